#!/usr/bin/env python
"""Traversal counters (GRT_OPT_STATS) of one builtin scene per path segment: box tests, primitive tests, shading.
usage (under gpurun): python tools/scene_stats.py <scene> [width [spp [mesh_segments]]]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import go_raytracer_b200 as g
sid = int(sys.argv[1])
w = int(sys.argv[2]) if len(sys.argv) > 2 else 480
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kw = {"mesh_segments": int(sys.argv[4])} if len(sys.argv) > 4 else {}
if sid in (2, 5):
    kw["image"] = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "earthmap_rgb8.npz"))["rgb"]
s, cfg = g.builtin_scene(sid, width=w, spp=spp, **kw)
cam = g.derive_camera(cfg)
_, _, st = g.DeviceScene(s).render(cam, want_stats=True)
seg = max(1, st["segments"])
print({k: v for k, v in st.items()})
print(f"scene {sid}: {st['paths']} paths, {seg / max(1, st['paths']):.2f} segments/path; per segment: "
      + ", ".join(f"{k} {st[k] / seg:.2f}" for k in st if k.endswith("_tests")))
