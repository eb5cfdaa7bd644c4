#!/bin/bash
export GRT_CUDA_LIB=$PWD/go_raytracer_b200/csrc/ab/prof.so
for b in 2 4 1000; do echo "== div $b"; GRT_TRAV_BUDGET=$b python tools/render_scene.py 8 480 256 2>&1 | grep -v Trace; done
