"""Generates tests/golden/hits_golden.npz: frozen oracle outputs (hit ids, t, front face) for the primary rays of
small views of scenes 6 (Cornell box), 1 (book-1 cover), 7 (smoke, no RNG on primary hits of surfaces) and the
oracle's 16x16 x 16 spp Cornell render.  Run from the repo root:  python tests/golden/make_hit_golden.py
The fixture pins the oracle against drift; the oracle itself is pinned as described in oracle/oracle.cpp's header."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import go_raytracer_b200 as g          # noqa: E402
from oracle import oracle_py as O      # noqa: E402
import parity_util as PU               # noqa: E402

out = {}
for sid, w in ((6, 48), (1, 64), (4, 48)):
    s, cfg = g.builtin_scene(sid, width=w, spp=1)
    cam = O.derived_camera(cfg)
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    h = O.OracleWorld(s).trace_batch(rays, audit_eps=1e-5)
    out[f"s{sid}_w"] = np.int32(w)
    out[f"s{sid}_id"] = h["id"].astype(np.int64)
    out[f"s{sid}_t"] = h["t"].astype(np.float64)
    out[f"s{sid}_front"] = h["front_face"].astype(np.uint8)
    out[f"s{sid}_flags"] = h["flags"].astype(np.uint32)
s, cfg = g.builtin_scene(6, width=16, spp=16)
sums, _, _, _ = O.OracleWorld(s).render(cfg, seed=0xC0FFEE, use_exclusion=True)
out["render6_sum"] = sums
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hits_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
