import sys; sys.path.insert(0, ".")
import go_raytracer_b200 as g
for spp in (16, 64, 256, 1024, 4096):
    s, cfg = g.builtin_scene(6, width=128, spp=spp)
    cam = g.derive_camera(cfg)
    _, _, st = g.DeviceScene(s).render(cam, want_stats=True)
    print(spp, "segments/path %.3f" % (st["segments"] / st["paths"]), "lanes per warp-iteration %.2f" % (st["lane_iterations"] / st["warp_iterations"]),
          "quad tests/segment %.2f box/segment %.2f" % (st["quad_tests"] / st["segments"], st["box_tests"] / st["segments"]))
