#!/bin/bash
export REPS=3 GRT_VARIANT=1
for d in 0 2; do for e in ${EXITS:-14}; do
echo "== dyn $d exit16 $e"
GRT_WF_DYN=$d GRT_WF_EXIT16=$e python tools/render_scene.py 8 480 1024 2>&1 | grep "^variant"
GRT_WF_DYN=$d GRT_WF_EXIT16=$e python tools/render_scene.py 2 480 1024 2>&1 | grep "^variant"
GRT_WF_DYN=$d GRT_WF_EXIT16=$e python tools/render_scene.py 1 1200 100 2>&1 | grep "^variant"
[ $d = 0 ] && break
done; done
