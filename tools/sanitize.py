"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck): megakernel (list, BVH, mesh variants),
wavefront kernels incl. the dynamic extend, ray queries, tonemap, GPU BuildBVH order."""
import sys
sys.path.insert(0, ".")
import numpy as np
import go_raytracer_b200 as g
for sid, kw, w, spp in ((6, {}, 24, 16), (7, {}, 16, 16), (1, {}, 24, 4), (2, {}, 24, 4), (8, {"mesh_segments": 48}, 24, 4)):
    s, cfg = g.builtin_scene(sid, width=w, spp=spp, **kw)
    cam = g.derive_camera(cfg)
    dev = g.DeviceScene(s)
    for v in (g.GRT_VARIANT_MEGAKERNEL, g.GRT_VARIANT_WAVEFRONT):
        sums, rgb8, _ = dev.render(cam, variant=v, want_rgb8=True)
        assert np.isfinite(sums).mean() > 0.99
    _, _, st = dev.render(cam, want_stats=True)
    dev.close()
    print("scene", sid, "ok", st["paths"])
b = np.random.default_rng(1).uniform(-5, 5, size=(40000, 3))
g.bvh_order(np.concatenate([b, b + 0.5], axis=1))
print("sanitize pass ok")
