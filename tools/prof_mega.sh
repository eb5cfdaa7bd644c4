#!/bin/bash
# usage (under gpurun): tools/prof_mega.sh <scene> <width> <spp> <tag>  -- one `ncu --set full` capture of the megakernel after a plain run
set -e
export GRT_VARIANT=0
python tools/render_scene.py $1 $2 $3 > gpurun_out/cfg_$4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_mega -s 1 -c 1 -f -o gpurun_out/prof_$4 python tools/render_scene.py $1 $2 $3 > gpurun_out/ncu_$4.log 2>&1
