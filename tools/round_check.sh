#!/bin/bash
# tests + default bench + other configs (run under gpurun); logs under gpurun_out/
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
python tools/fullsize_check.py 2>&1 | grep -v Warn
