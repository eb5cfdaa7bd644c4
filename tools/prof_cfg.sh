#!/bin/bash
# usage (under gpurun): tools/prof_cfg.sh <scene> <width> <spp> <tag>   -- timing + stats, then one ncu --set full capture
set -e
python tools/render_scene.py $1 $2 $3 | tee gpurun_out/cfg_$4.log
ncu --set full --clock-control none --import-source on -k regex:render_mega -s 1 -c 1 -f -o gpurun_out/prof_$4 python tools/render_scene.py $1 $2 $3 > gpurun_out/ncu_$4.log 2>&1
