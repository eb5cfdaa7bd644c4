#!/bin/bash
# throughput at the per-GPU strata counts of 1 / 2 / 4 / 8-GPU sharding (4096 / 2048 / 1024 / 512 spp on one GPU)
for spp in 4096 2048 1024 512; do
  python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($spp, round(d['value']), round(d['ms_per_step'],2))"
done
