#!/usr/bin/env python
"""Print the SASS of one kernel annotated with source lines (from the -lineinfo tables), optionally only a window
around the first instruction of a given source line.

  tools/sass_lines.py LIB KERNEL_SUBSTRING [file.cuh:LINE [before [after]]]
"""
import os, re, subprocess, sys, tempfile


def annotated(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        want, cur, out = False, None, []
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                want = kernel in m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = m.group(1).split("/")[-1] + ":" + m.group(2)
                continue
            if want and (re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) or re.match(r"\s*\.L_x_\d+:", ln)):
                out.append((cur, ln.strip()[:110]))
        if out:
            return out
    return []


if __name__ == "__main__":
    out = annotated(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        before = int(sys.argv[4]) if len(sys.argv) > 4 else 40
        after = int(sys.argv[5]) if len(sys.argv) > 5 else 40
        idx = [i for i, (c, _) in enumerate(out) if c == sys.argv[3]]
        if not idx:
            sys.exit("no instruction carries " + sys.argv[3])
        out = out[max(0, idx[0] - before): idx[0] + after]
    for c, l in out:
        print((c or "").ljust(24), l)
