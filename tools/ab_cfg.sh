#!/bin/bash
# A/B of several builds on the BVH configs.  usage (under gpurun): [EXITS="4 8 12"] tools/ab_cfg.sh a.so b.so ...
for lib in "$@"; do
  export GRT_CUDA_LIB=$PWD/$lib
  for b in ${EXITS:-8}; do
    echo "== $lib  exit16 $b"
    GRT_TRAV_EXIT16=$b python tools/render_scene.py 8 480 1024 2>&1 | grep "scene"
    GRT_TRAV_EXIT16=$b python tools/render_scene.py 2 480 1024 2>&1 | grep "scene"
    GRT_TRAV_EXIT16=$b python tools/render_scene.py 1 400 100 2>&1 | grep "scene"
  done
done
