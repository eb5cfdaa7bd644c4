import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU
np.set_printoptions(precision=9, linewidth=200)
for sid in [int(x) for x in sys.argv[1:]]:
    kw = {"mesh_segments": 96} if sid == 8 else {}
    s, cfg = g.builtin_scene(sid, width=160, spp=4, **kw)
    ow, dev = O.OracleWorld(s), g.DeviceScene(s)
    cam = O.derived_camera(cfg)
    prim = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    oh = ow.trace_batch(prim)
    sec = PU.secondary_batch(oh, np.random.default_rng(sid), time=prim["time"])
    for name, rays, excl in (("plain", sec.copy(), False), ("self", sec, True)):
        if not excl:
            rays["self_id"] = PU.NO_ID
        o2 = ow.trace_batch(rays, audit_eps=1e-5, use_exclusion=excl)
        g2 = dev.trace_batch(rays)
        r = PU.compare_hits(g2, o2)
        print(sid, name, {k: v for k, v in r.items() if not k.endswith("idx")})
        for i in list(r["bad_id_idx"][:4]) + list(r["bad_t_idx"][:4]):
            print("  ray", i, "o", rays["o"][i], "d", rays["d"][i], "self", sec["self_id"][i])
            print("     gpu id", g2["id"][i], "t", g2["t"][i], " orc id", o2["id"][i], "t", o2["t"][i], "flags", o2["flags"][i], "second_t", o2["second_t"][i])
