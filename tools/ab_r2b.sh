#!/bin/bash
# usage (under gpurun): tools/ab_r2b.sh "<scene w spp [env...]>;..." lib.so...   — each case on each library
IFS=';' read -ra CASES <<< "$1"; shift
export REPS=2 GRT_VARIANT=2
for lib in "$@"; do
  export GRT_CUDA_LIB=$PWD/$lib
  for c in "${CASES[@]}"; do
    set -- $c
    sid=$1; w=$2; spp=$3; shift 3
    echo -n "$lib [$*] "
    env "$@" python tools/render_scene.py $sid $w $spp 2>&1 | grep "^variant"
  done
done
