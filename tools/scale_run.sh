#!/bin/bash
# 1 -> 8 GPU runs of bench.py back to back on one box (run under `gpurun --gpus 8`); lines land in gpurun_out/scale_N.json
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/scale_1.json
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/scale_$n.err | grep '^{' > gpurun_out/scale_$n.json
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_{n}.json"))
    except Exception as e:
        print(n, "failed", e); continue
    base = base or d["value"]
    print(f"N={n}: {d['value']:.0f} Mpaths/s  e2e {d['e2e']['value']:.0f}  {d['ms_per_step']:.2f} ms/step  efficiency {100 * d['value'] / (n * base):.1f} %")
PY
