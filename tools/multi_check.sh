#!/bin/bash
# in-process multi-GPU path: tests, then fused peer-memory accumulation vs ncclReduce at the headline size (needs >= 2 GPUs)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export GRT_VERBOSE=1
for p in 1 0; do
  for rep in 1 2; do GRT_MULTI_P2P=$p ./go_raytracer_b200/csrc/grt_main -S 6 -gpus 2 -width 1024 -spp 4096 -variant mega -o gpurun_out/c2_p2p$p.pnm 2>&1 | grep grt_main | sed "s/^/P2P=$p /"; done
done
cmp gpurun_out/c2_p2p1.pnm gpurun_out/c2_p2p0.pnm && echo "identical images" || echo "images differ (fp32 summation order)"
rm -f gpurun_out/c2_p2p*.pnm
