#!/bin/bash
# per-kernel device time of one render (ncu --metrics gpu__time_duration.sum).  usage: tools/launch_list.sh <scene> <w> <spp> <tag>
REPS=1 GRT_VARIANT=2 python tools/render_scene.py $1 $2 $3 > /dev/null 2>&1 || exit 1
REPS=1 GRT_VARIANT=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$4.csv python tools/render_scene.py $1 $2 $3 > gpurun_out/ncu_$4.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_$4.csv") if l.startswith('"')))
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value"); u = h.index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    name = r[k].split("<")[0].split("(")[0]
    if "wf_shade" in r[k]: name += "<Q=" + r[k].split(",")[-1].split(">")[0].strip() + ">"
    x = float(r[v].replace(",", "")); x *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u].replace("second", "s").replace("usecond","us"), 1e-6)
    t[name] += x; n[name] += 1
tot = sum(t.values())
for name, x in t.most_common(): print(f"{name:40s} {n[name]:6d} launches {x:10.2f} ms {100*x/tot:5.1f} %")
PY
