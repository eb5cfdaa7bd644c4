#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep) into the small text summary kept under profiles/.
usage: tools/summarize_profile.py gpurun_out/prof.ncu-rep > profiles/name.txt"""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout

raw = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]
print(f"# ncu summary of {rep.split('/')[-1]}")
kname = ncu("--page", "source", "--csv").splitlines()[0]
print("kernel:", kname)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {vals[i]:>16s} {units[i]}")
print("\n## warp stall reasons (pct of warp-active cycles)")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h:
        try:
            if float(vals[i]) >= 0.5:
                print(f"{h:75s} {vals[i]:>16s}")
        except ValueError:
            pass
src = list(csv.reader(ncu("--page", "source", "--csv").splitlines()))
h2 = src[1]; ix = {h: i for i, h in enumerate(h2)}
tot = collections.Counter(); thr = collections.Counter(); allw = allt = 0
for r in src[2:]:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    op = m.group(2) if m else r[ix["Source"]]
    base = op.split(".")[0]
    key = op if base in ("F2F", "I2F", "MUFU", "DFMA", "DADD", "DMUL") else base
    w = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]])
    tot[key] += w; thr[key] += t; allw += w; allt += t
print(f"\n## SASS mix: {allw} warp instructions, {allt / max(allw, 1):.2f} active threads per instruction (warp execution efficiency {100 * allt / max(allw, 1) / 32:.1f} %)")
for k, v in tot.most_common(22):
    print(f"{k:16s} {100 * v / allw:5.1f} %   avg active threads {thr[k] / max(v, 1):5.1f}")
