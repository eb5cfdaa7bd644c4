#!/bin/bash
# round 2 A/B of extend-kernel builds on the BVH configs: usage (under gpurun) tools/ab_r2.sh a.so b.so ...
# each library is timed with the default extend choice and with the persistent extend forced on (GRT_WF_DYN=2)
export REPS=2 GRT_VARIANT=2
for lib in "$@"; do
  export GRT_CUDA_LIB=$PWD/$lib
  for dyn in ${DYNS:-default 2 0}; do
    echo "== $lib GRT_WF_DYN=$dyn"
    if [ "$dyn" = default ]; then unset GRT_WF_DYN; else export GRT_WF_DYN=$dyn; fi
    python tools/render_scene.py 8 480 1024 2>&1 | grep "^variant"
    python tools/render_scene.py 2 480 1024 2>&1 | grep "^variant"
    python tools/render_scene.py 1 1200 100 2>&1 | grep "^variant"
    python tools/render_scene.py 7 512 1024 2>&1 | grep "^variant"
  done
done
