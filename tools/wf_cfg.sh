#!/bin/bash
for v in 0 1; do
  GRT_VARIANT=$v python tools/render_scene.py 8 480 256 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 2 480 256 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 1 400 100 2>&1 | grep "scene"
done
