#!/usr/bin/env python
"""Surface-area cost of the device's 4-wide BVH for a builtin scene (no GPU needed): expected node visits and leaf
tests of a random ray that hits the root box, from the tree grt_debug_repack returns.
usage: [GRT_WIDE_SAH=1] python tools/wide_cost.py <scene> [mesh_segments]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import go_raytracer_b200 as g
from test_wide_bvh import repack

sid = int(sys.argv[1])
kw = {"mesh_segments": int(sys.argv[2])} if len(sys.argv) > 2 else {}
if sid in (2, 5):
    kw["image"] = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "earthmap_rgb8.npz"))["rgb"]
s, cfg = g.builtin_scene(sid, width=64, spp=1, **kw)
flat = s.flatten()
t0 = time.time()
R = repack(flat)
dt = time.time() - t0
wn = R["wnodes"]
lo = wn[:, 0:12].reshape(-1, 3, 4).astype(np.float64)
hi = wn[:, 12:24].reshape(-1, 3, 4).astype(np.float64)
ref = wn[:, 24:28].copy().view(np.uint32)
cnt = wn[:, 29].copy().view(np.uint32)
e = np.clip(hi - lo, 0, None)
area = e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]      # [n, 4]
valid = np.arange(4)[None, :] < cnt[:, None]
area = np.where(valid & np.isfinite(area), area, 0.0)
run = valid & ((ref >> 31) == 1)
big = area.max()
# exclude the handful of huge children (ground sphere, world list): they are common to both builds
small = area < 0.01 * big
print(f"scene {sid}: {len(wn)} wide nodes, stack need {R['need_main']}, repack {dt:.2f} s; children per node {np.bincount(cnt, minlength=5)[1:5]}")
print(f"sum of child-box areas: inner children {area[valid & ~run & small].sum():.4g}, leaf children {area[run & small].sum():.4g}  (excluded {int((valid & ~small).sum())} huge)")
