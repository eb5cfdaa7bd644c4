#!/bin/bash
# traversal-slice exit divisor sweep on the BVH configs (run under gpurun)
for b in 1 2 4 8 1000; do
  echo "== exit divisor $b"
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 8 480 1024 2>&1 | grep "^scene"
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 2 480 1024 2>&1 | grep "^scene"
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 1 400 100 2>&1 | grep "^scene"
done
