#!/bin/bash
for b in 64 100000; do
  echo "== budget $b"
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 8 480 256 2>&1 | grep -v Trace
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 8 480 1024 2>&1 | grep -v Trace
  GRT_TRAV_BUDGET=$b python tools/render_scene.py 2 480 1024 2>&1 | grep -v Trace
done
