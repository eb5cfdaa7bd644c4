#!/bin/bash
for v in 2 0; do
  GRT_VARIANT=$v python tools/render_scene.py 8 480 1024 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 2 480 1024 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 1 1200 100 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 7 1024 256 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 3 1024 256 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 4 1024 256 2>&1 | grep "scene"
  GRT_VARIANT=$v python tools/render_scene.py 5 1024 256 2>&1 | grep "scene"
done
