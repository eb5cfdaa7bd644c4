#!/bin/bash
# A/B of several builds on the BVH configs.  usage (under gpurun): tools/ab_cfg.sh a.so b.so ...
export REPS=3 GRT_VARIANT=2
for lib in "$@"; do
  export GRT_CUDA_LIB=$PWD/$lib
  echo "== $lib"
  python tools/render_scene.py 8 480 1024 2>&1 | grep "^variant"
  python tools/render_scene.py 2 480 1024 2>&1 | grep "^variant"
  python tools/render_scene.py 1 1200 100 2>&1 | grep "^variant"
done
