"""In-process 2-GPU render of the headline frame: fused peer-memory accumulation vs ncclReduce (device time per call)."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
s, cfg = g.builtin_scene(6, width=1024, spp=4096)
cam = g.derive_camera(cfg)
out = np.zeros(cam.width * cam.height * 3, dtype=np.float32)
for p2p in ("1", "0", "1", "0"):
    os.environ["GRT_MULTI_P2P"] = p2p
    t = []
    for rep in range(4):
        out[:] = 0
        ms = C.c_double(0)
        N.check(N.lib().grt_host_camera_render(s._h, C.byref(cfg), 0xC0FFEE, 0, n, out.ctypes.data, None, 0, None, C.byref(ms)))
        t.append(round(ms.value, 1))
    print(f"{n} GPUs, P2P={p2p}: device ms per call {t}  mean radiance {out.mean() / 4096:.5f}", flush=True)
