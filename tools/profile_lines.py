#!/usr/bin/env python
"""Per-source-line view of an ncu report: joins the report's per-SASS-instruction counters (page `source`, sass view)
with the line table of the library that was profiled (nvdisasm -g on the cubin inside the .so, built with -lineinfo).
usage: tools/profile_lines.py <prof.ncu-rep> <lib.so> [top_n]
Columns: share of warp instructions, average active threads, share of stall samples, top stall reason, file:line."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernel = rows[0][1]
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
insts = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    insts.append((int(r[ix["Address"]], 16), r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]),
                  int(r[ix["Warp Stall Sampling (All Samples)"]] or 0), {c: int(r[ix[c]] or 0) for c in stall_cols}))
base = insts[0][0]
# line table from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
mangled_pat = re.sub(r"[^A-Za-z0-9_]", "", kernel.split("<")[0].split()[-1])   # e.g. wf_extend_dyn
targs = re.findall(r"\)(\d+)", kernel)                                            # template integers, in order
line_of = {}
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, want = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            dem = subprocess.run(["c++filt", cur_fn], capture_output=True, text=True).stdout.strip()
            want = dem.replace("(unsigned int)", "").replace("(int)", "").replace("(bool)", "").replace(" ", "") == \
                kernel.replace("(unsigned int)", "").replace("(int)", "").replace("(bool)", "").replace(" ", "") or \
                (mangled_pat in dem and all(t in re.sub(r"u\b", "", dem) for t in targs) and dem.split("<")[0].split()[-1] == kernel.split("<")[0].split()[-1])
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = os.path.basename(m.group(1)) + ":" + m.group(2)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and want:
            line_of[int(m.group(1), 16)] = cur_line
    if line_of:
        break
inst = collections.Counter(); thr = collections.Counter(); samp = collections.Counter(); reasons = collections.defaultdict(collections.Counter)
for addr, src, w, t, s, st in insts:
    loc = line_of.get(addr - base, "?")
    inst[loc] += w; thr[loc] += t; samp[loc] += s
    for k, v in st.items():
        reasons[loc][k] += v
tw = sum(inst.values()); ts = max(1, sum(samp.values()))
print(f"# {kernel}\n# {tw} warp instructions, {sum(thr.values()) / max(tw, 1):.2f} active threads per instruction; {len(line_of)} SASS instructions mapped to lines")
tot_reason = collections.Counter()
for loc in reasons:
    tot_reason.update(reasons[loc])
print("# stall samples by reason: " + ", ".join(f"{k[6:]} {100 * v / max(1, sum(tot_reason.values())):.0f}%" for k, v in tot_reason.most_common(6)))
for loc, w in inst.most_common(top):
    top_r = reasons[loc].most_common(1)[0][0][6:] if reasons[loc] else ""
    print(f"{100 * w / tw:5.1f} % inst  thr {thr[loc] / max(w, 1):5.1f}  stall {100 * samp[loc] / ts:5.1f} % ({top_r})  {loc}")
