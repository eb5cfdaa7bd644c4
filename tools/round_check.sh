#!/bin/bash
# tests + smoke + default bench (+ reference arm) + other configs in one call (run under gpurun); logs under gpurun_out/
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json")); r = json.load(open("gpurun_out/bench_reference.json"))
print("ours", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
print("reference arm", r["value"], r["cpu_baseline"]["cores"], "cores; e2e ratio", round(d["e2e"]["value"] / r["value"]))
PY
python tools/fullsize_check.py 2>&1 | grep -v Warn
