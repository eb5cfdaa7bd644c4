"""Render one built-in scene once on the GPU (profiling helper): python tools/render_scene.py <scene> [width] [spp]"""
import sys, os
sys.path.insert(0, ".")
import numpy as np
import go_raytracer_b200 as g
import torch
VAR = int(os.environ.get("GRT_VARIANT", "0"))
sid = int(sys.argv[1]); width = int(sys.argv[2]) if len(sys.argv) > 2 else 0; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cw = int(sys.argv[4]) if len(sys.argv) > 4 else None; cl = int(sys.argv[5]) if len(sys.argv) > 5 else None
if os.environ.get("CL"):   # flatten options from the environment (A/B scripts): collapse_whole, collapse_leaf
    cw, cl = int(os.environ.get("CW", "32")), int(os.environ["CL"])
kw = {}
if sid in (2, 5):
    kw["image"] = np.load("tests/golden/earthmap_rgb8.npz")["rgb"]
if sid == 2 and width:
    kw["aspect"] = 16 / 9
if sid == 8 and os.environ.get("MESH_SEGMENTS"):   # smaller meshes than config C5's 708 x 708 segments
    kw["mesh_segments"] = int(os.environ["MESH_SEGMENTS"])
s, cfg = g.builtin_scene(sid, width=width, spp=spp, **kw)
dev = g.DeviceScene(s, 0, cw, cl)
cam = g.derive_camera(cfg)
acc = torch.zeros(cam.width * cam.height * 3, dtype=torch.float32, device="cuda")
import time
for _ in range(int(os.environ.get("REPS", "2"))):
    acc.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.render_device(cam, acc.data_ptr(), variant=VAR); e1.record(); torch.cuda.synchronize()
    if os.environ.get("REPS"): print("  rep: %.1f ms (events)  %.1f ms (wall)" % (e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
paths = cam.width * cam.height * cam.spp_sqrt ** 2
print(f"variant {VAR} scene {sid} collapse=({cw},{cl}) {cam.width}x{cam.height}x{cam.spp_sqrt**2}: {e0.elapsed_time(e1):.1f} ms, {paths / e0.elapsed_time(e1) / 1e3:.1f} Mpaths/s")
_, _, st = dev.render(cam, want_stats=True)
print({k: round(v / st["paths"], 2) for k, v in st.items() if k != "paths"}, "lanes/iter %.1f" % (st["lane_iterations"] / max(1, st["warp_iterations"])))
