"""Time-to-first-pixel pieces of the 1M-triangle config: scene build, flatten (BuildBVH on GPU or host), upload."""
import os, sys, time
sys.path.insert(0, ".")
import go_raytracer_b200 as g
import torch
torch.zeros(1, device="cuda")
s, cfg = g.builtin_scene(8, width=480, spp=16)
for mode in ("gpu", "cpu", "gpu"):
    os.environ["GRT_BVH_BUILD"] = mode
    t0 = time.time(); f = s.flatten(); t1 = time.time()
    dev = g.DeviceScene(s); t2 = time.time()
    print(f"BVH build on {mode}: flatten {t1 - t0:.3f} s, flatten+upload (DeviceScene) {t2 - t1:.3f} s", flush=True)
    dev.close()
