#!/bin/bash
# usage (under gpurun): tools/prof_wf.sh <scene> <width> <spp> <kernel regex> <skip> <tag>
# one `ncu --set full` capture of a wavefront kernel in steady state (after <skip> launches), after a plain run
set -e
export GRT_VARIANT=2
python tools/render_scene.py $1 $2 $3 > gpurun_out/cfg_$6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/prof_$6 python tools/render_scene.py $1 $2 $3 > gpurun_out/ncu_$6.log 2>&1
