#!/bin/bash
# launch list of the default bench command (per-launch device times), as profiles/README.md describes.  Run under gpurun.
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ll_bench.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ll.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r1_launches.csv") if l.startswith('"')))
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value"); u = h.index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    x = float(r[v].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(r[u], 1e-6)
    t[r[k][:90]] += x; n[r[k][:90]] += 1
tot = sum(t.values())
with open("gpurun_out/r1_launches_summary.txt", "w") as f:
    f.write("# per-kernel device time from profiles/r1_launches.csv (ncu --metrics gpu__time_duration.sum, bench.py --steps 2 --warmup 3 --no-cpu: full size 1024^2 x 4096 spp)\n# cold-cache, serialised: compare SHARES, not absolutes\n")
    for name, x in t.most_common(): f.write(f"{100*x/tot:6.2f}%  n={n[name]:3d} {x:10.3f} ms  {name}\n")
print(open("gpurun_out/r1_launches_summary.txt").read())
PY
