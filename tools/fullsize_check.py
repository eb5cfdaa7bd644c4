"""Full-size sanity of the non-headline configs (C1, C3, C4, C5) at reduced spp: runs, finite, Mpaths/s."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import go_raytracer_b200 as g
import torch, os
VAR = int(os.environ.get("GRT_VARIANT", "2"))   # AUTO
earth = np.load("tests/golden/earthmap_rgb8.npz")["rgb"]
# full-size scenes; spp as in BASELINE.json where affordable, otherwise the resolution is reduced (NOT the spp: the
# megakernel keeps a warp full only when a pixel has >= 64 strata)
cases = [("C1 book1 400x225 x 100", dict(scene_id=1)),
         ("C3 smoke 1024^2 x 1024", dict(scene_id=7, width=1024, spp=1024)),
         ("C4 book2 1920x1080 x 64", dict(scene_id=2, width=1920, aspect=16 / 9, spp=64, image=earth)),
         ("C4 book2 480x270 x 1024", dict(scene_id=2, width=480, aspect=16 / 9, spp=1024, image=earth)),
         ("C5 mesh 1M tris 3840x2160 x 16", dict(scene_id=8, width=3840, spp=16)),
         ("C5 mesh 1M tris 480x270 x 1024", dict(scene_id=8, width=480, spp=1024))]
for name, kw in cases:
    t0 = time.time()
    s, cfg = g.builtin_scene(**kw)
    t1 = time.time()
    dev = g.DeviceScene(s)
    t2 = time.time()
    cam = g.derive_camera(cfg)
    n = cam.width * cam.height * 3
    acc = torch.zeros(n, dtype=torch.float32, device="cuda")
    dev.render_device(cam, acc.data_ptr(), variant=VAR)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc.zero_(); e0.record(); dev.render_device(cam, acc.data_ptr(), variant=VAR); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    a = acc.cpu().numpy()
    paths = cam.width * cam.height * cam.spp_sqrt ** 2
    fl = s.flatten()
    print(f"{name}: build {t1-t0:.1f}s flatten+upload {t2-t1:.1f}s  nodes {fl.n_nodes} tris {fl.n_tris} quads {fl.n_quads} boxes {fl.n_boxes} spheres {fl.n_spheres} hint {fl.max_depth_hint} | "
          f"{ms:.1f} ms  {paths/ms/1e3:.1f} Mpaths/s  mean {np.nanmean(a)/cam.spp_sqrt**2:.4f} nan {np.isnan(a).sum()} ", flush=True)
