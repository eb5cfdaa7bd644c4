"""Registers / stack / spills per kernel from the Makefile's *.ptxas.log files (nvcc -Xptxas -v)."""
import re, subprocess, sys, glob, os
d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "go_raytracer_b200", "csrc")
pat = sys.argv[1] if len(sys.argv) > 1 else "wf_extend|render_mega|trace_batch|wf_shade"
for f in sorted(glob.glob(os.path.join(d, "*.ptxas.log"))):
    cur, stack = None, None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m and cur and re.search(pat, cur):
            print(f"{cur[:100]:100s} regs {m.group(1):>3s} stack {stack[0]:>5s} spill st/ld {stack[1]}/{stack[2]}")
