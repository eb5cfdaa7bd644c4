#!/bin/bash
# Closing launch lists (run under gpurun): `ncu --metrics gpu__time_duration.sum` over the wavefront configs' bench commands
# on the final build, each after a plain run of the same command; summarised on the box (tools/launch_summary.py), csv gzipped.
set -x
: > gpurun_out/r2c_launches_summary.txt
for cs in "C1 100" "C5 4" "C4 16"; do
  set -- $cs; c=$1; spp=$2
  Bc="python bench.py --config $c --spp $spp --steps 1 --warmup 1 --no-cpu --no-configs"
  $Bc > gpurun_out/r2c_ll_plain_$c.json 2> gpurun_out/r2c_ll_plain_$c.err || continue
  timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches_$c.csv $Bc > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2c_launches_$c.csv >> gpurun_out/r2c_launches_summary.txt 2>&1
  gzip -f gpurun_out/r2c_launches_$c.csv
done
du -sh gpurun_out
