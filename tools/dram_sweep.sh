#!/bin/bash
# DRAM traffic of render_mega_kernel vs image size / spp (ncu metrics only).  Run under gpurun from the repo root.
python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1 || exit 1
for cfg in "1024 256" "1024 1024" "512 4096" "256 4096"; do
  set -- $cfg
  ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:render_mega -s 0 -c 1 --csv \
      python bench.py --width $1 --spp $2 --steps 1 --warmup 1 --no-cpu 2>/dev/null | grep render_mega | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print('$1 x $2', r[-3], r[-2], r[-1])"
done
