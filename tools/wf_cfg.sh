#!/bin/bash
# wavefront path with and without the CUDA graph of the bounce loop
export REPS=3 GRT_VARIANT=2
for gph in 1 0; do
  echo "== GRT_WF_GRAPH=$gph"
  GRT_WF_GRAPH=$gph python tools/render_scene.py 1 400 100 2>&1 | grep "^variant"
  GRT_WF_GRAPH=$gph python tools/render_scene.py 1 1200 100 2>&1 | grep "^variant"
  GRT_WF_GRAPH=$gph python tools/render_scene.py 2 480 1024 2>&1 | grep "^variant"
  GRT_WF_GRAPH=$gph python tools/render_scene.py 8 480 1024 2>&1 | grep "^variant"
done
