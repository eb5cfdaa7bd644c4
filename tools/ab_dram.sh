#!/bin/bash
# A/B of two builds: speed and DRAM write traffic.  usage (under gpurun): tools/ab_dram.sh a.so b.so
for lib in "$@"; do
  export GRT_CUDA_LIB=$PWD/$lib
  echo "== $lib"
  python bench.py --spp 1024 --steps 3 --warmup 2 --no-cpu 2>/dev/null | tail -1 | cut -c1-60
  ncu --metrics dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:render_mega -s 0 -c 1 --csv \
      python bench.py --spp 1024 --steps 1 --warmup 1 --no-cpu 2>/dev/null | grep render_mega | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print(r[-3], r[-2], r[-1])"
done
