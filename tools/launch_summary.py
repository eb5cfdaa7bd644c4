#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
usage: tools/launch_summary.py x.csv  ->  launches, total ms, share of the listed device time, per kernel name."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {n: i for i, n in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if r is hdr or len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])
    name = re.sub(r"<.*", "", name).replace("void ", "")
    if "wf_shade" in r[ix["Kernel Name"]]:
        m = re.search(r"\(int\)(\d)>", r[ix["Kernel Name"]]); name += f"<Q={m.group(1)}>" if m else ""
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else (v if unit in ("ms", "msecond") else v * 1e3))
    tot[name] += ms; cnt[name] += 1
allms = sum(tot.values())
print(f"# {sys.argv[1].split('/')[-1]}: {sum(cnt.values())} launches, {allms:.2f} ms of device time listed")
for k, v in tot.most_common():
    print(f"{k:40s} {cnt[k]:6d} launches {v:10.2f} ms {100 * v / allms:6.1f} %")
