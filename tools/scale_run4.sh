#!/bin/bash
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/scale_1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 bench.py --gpus 4 --steps 3 --warmup 3 2>gpurun_out/scale_4.err > gpurun_out/scale_4.out
grep -c . gpurun_out/scale_4.out; grep '^{' gpurun_out/scale_4.out > gpurun_out/scale_4.json
python - <<'PY'
import json
a = json.load(open("gpurun_out/scale_1.json")); b = json.load(open("gpurun_out/scale_4.json"))
print(f"N=1 {a['value']:.0f}  N=4 {b['value']:.0f} ({b['ms_per_step']:.2f} ms/step, e2e {b['e2e']['value']:.0f})  efficiency {100*b['value']/(4*a['value']):.1f} %")
PY
