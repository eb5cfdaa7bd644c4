#!/bin/bash
# quick throughput check of the megakernel on the Cornell box (4096 / 484 spp) and the smoke scene
for spp in 4096 512; do
  python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cornell', $spp, round(d['value']), round(d['ms_per_step'],2))"
done
REPS=3 GRT_VARIANT=0 python tools/render_scene.py 7 1024 1024 2>&1 | grep "^variant"
