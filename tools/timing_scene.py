#!/usr/bin/env python
"""Device time per kernel class (GRT_OPT_TIMING: CUDA events around every launch) of one builtin scene's wavefront render.
usage (under gpurun): python tools/timing_scene.py <scene> [width [spp]]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N
sid = int(sys.argv[1]); w = int(sys.argv[2]) if len(sys.argv) > 2 else 480; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kw = {}
if sid in (2, 5):
    kw["image"] = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "earthmap_rgb8.npz"))["rgb"]
s, cfg = g.builtin_scene(sid, width=w, spp=spp, **kw)
cam = g.derive_camera(cfg)
scene = g.DeviceScene(s)
acc = torch.zeros(cam.height * cam.width * 3, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
for flags in (0, N.GRT_OPT_TIMING):
    acc.zero_()
    scene.render_device(cam, acc.data_ptr(), st.cuda_stream, variant=N.GRT_VARIANT_AUTO, flags=flags)
    torch.cuda.synchronize()
t = N.GrtTiming()
N.check(N.lib().grt_last_timing(C.byref(t)))
paths = cam.width * cam.height * cam.spp_sqrt ** 2
print(f"scene {sid} {cam.width}x{cam.height}x{cam.spp_sqrt ** 2}: total {t.total_ms:.1f} ms = generate {t.generate_ms:.1f} + extend {t.extend_ms:.1f} + shade {t.shade_ms:.1f} "
      f"(+ {t.total_ms - t.generate_ms - t.extend_ms - t.shade_ms:.1f} other); {t.launches} launches, {t.extend_launches} extend; {paths / t.total_ms / 1e3:.0f} Mpaths/s under timing")
