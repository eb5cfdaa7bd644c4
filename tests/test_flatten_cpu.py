"""CPU checks of the product's host layer: the flattener (BuildBVH order, instance baking, outward-rounded fp32
boxes) is validated against the oracle with a small fp64 interpreter of the FLAT arrays written here in numpy."""
import ctypes as C
import numpy as np
import pytest
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N
from oracle import oracle_py as O
import parity_util as PU


def as_np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class FlatInterp:
    """Closest hit over the flat scene in fp64, following the traversal contract of include/grt.h
    (left child first, lists in order, closed quad / open sphere intervals).  No media."""

    def __init__(self, flat):
        self.nodes = as_np(flat.nodes, flat.n_nodes, N.NODE_DTYPE)
        self.spheres = as_np(flat.spheres, flat.n_spheres, N.SPHERE_DTYPE)
        self.quads = as_np(flat.quads, flat.n_quads, N.QUAD_DTYPE)
        self.tris = as_np(flat.tris, flat.n_tris, N.TRI_DTYPE)
        self.boxes = as_np(flat.boxes, flat.n_boxes, N.BOX_DTYPE)
        self.items = as_np(flat.items, flat.n_items, np.dtype("<u4"))
        self.root = flat.root

    def hit(self, o, d, time, tmin, tmax):
        best = (-1, np.inf)
        stack = [self.root]
        while stack:
            ref = stack.pop()
            t, i = (ref >> 28) & 7, ref & N.REF_MASK
            if t == N.REF_LIST:
                item = int(self.items[i])
                if not item & N.LIST_LAST:
                    stack.append((N.REF_LIST << 28) | (i + 1))
                stack.append(item & ~N.LIST_LAST & 0xFFFFFFFF)
            elif t == N.REF_NODE:
                n = self.nodes[i]
                lo, hi = tmin, tmax
                ok = True
                for a in range(3):
                    if d[a] == 0:
                        if not (n["bmin"][a] <= o[a] <= n["bmax"][a]):
                            ok = False
                        continue
                    t0, t1 = (n["bmin"][a] - o[a]) / d[a], (n["bmax"][a] - o[a]) / d[a]
                    if t0 > t1:
                        t0, t1 = t1, t0
                    lo, hi = max(lo, t0), min(hi, t1)
                if ok and hi > lo:
                    stack.append(int(n["right"]) & 0x7FFFFFFF); stack.append(int(n["left"]) & 0x7FFFFFFF)   # bit 31: order hint
            elif t == N.REF_BOX:      # the six quads of a NewBox, in order
                fq = int(self.boxes[i]["first_quad"])
                for f in range(5, -1, -1):
                    stack.append((N.REF_QUAD << 28) | (fq + f))
            elif t == N.REF_QUAD:
                q = self.quads[i]
                den = q["n64"] @ d
                if abs(den) < 1e-8:
                    continue
                tt = (q["D64"] - q["n64"] @ o) / den
                if not (tmin <= tt <= tmax):
                    continue
                p = o + tt * d - q["Q"].astype(np.float64)
                al, be = q["A"].astype(np.float64) @ p, q["B"].astype(np.float64) @ p
                if 0 <= al <= 1 and 0 <= be <= 1:
                    best = (int(q["id"]), tt); tmax = tt
            elif t == N.REF_SPHERE:
                s = self.spheres[i]
                c = s["c0"] + time * s["dc"].astype(np.float64)
                oc = c - o
                a, h, cc = d @ d, d @ oc, oc @ oc - s["r"] ** 2
                disc = h * h - a * cc
                if disc < 0:
                    continue
                sq = np.sqrt(disc)
                root = (h - sq) / a
                if not (tmin < root < tmax):
                    root = (h + sq) / a
                    if not (tmin < root < tmax):
                        continue
                best = (int(s["id"]), root); tmax = root
            elif t == N.REF_TRI:
                tr = self.tris[i]
                e0, e1 = tr["e0"].astype(np.float64), tr["e1"].astype(np.float64)
                pv = np.cross(d, e1)
                det = e0 @ pv
                if abs(det) < 1e-8:
                    continue
                tv = o - tr["v0"].astype(np.float64)
                u = (tv @ pv) / det
                if u < 0 or u > 1:
                    continue
                qv = np.cross(tv, e0)
                v = (d @ qv) / det
                if v < 0 or u + v > 1:
                    continue
                tl = (e1 @ qv) / det
                if tl < tmin or tl > tmax:
                    continue
                best = (int(tr["id"]), tl); tmax = tl
        return best


@pytest.mark.parametrize("collapse", [None, (0, 0)])
@pytest.mark.parametrize("sid,kw", [(6, {}), (3, {}), (1, {}), (5, {}), (4, {}), (8, {"mesh_segments": 24})])
def test_flat_scene_agrees_with_oracle(sid, kw, collapse):
    s, cfg = g.builtin_scene(sid, width=24, spp=4, **kw)
    flat = s.flatten() if collapse is None else s.flatten(*collapse)
    fi = FlatInterp(flat)
    ow = O.OracleWorld(s)
    cam = O.derived_camera(cfg)
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    oh = ow.trace_batch(rays, audit_eps=1e-6)
    sec = PU.secondary_batch(oh, np.random.default_rng(3), time=rays["time"])
    sec["self_id"] = PU.NO_ID
    rays = np.concatenate([rays, sec[:300]])
    oh = ow.trace_batch(rays, audit_eps=1e-6)
    bad = 0
    for r, h in zip(rays, oh):
        fid, ft = fi.hit(r["o"].astype(np.float64), r["d"].astype(np.float64), float(r["time"]), float(r["tmin"]), float(r["tmax"]))
        if h["flags"]:
            continue
        if fid != h["id"] or (fid >= 0 and abs(ft - h["t"]) > 1e-5 * abs(h["t"])):
            bad += 1
    assert bad == 0


def test_collapse_options_do_not_change_topology_semantics():
    """Cornell: the whole world BVH (18 leaves) is emitted as one ordered run; boxes' span-1 duplicates appear once."""
    s, _ = g.builtin_scene(6)
    flat = s.flatten()
    assert flat.n_nodes == 0 and flat.n_quads == 18 and flat.n_boxes == 2
    items = as_np(flat.items, flat.n_items, np.dtype("<u4"))
    assert (flat.root >> 28) & 7 == N.REF_LIST
    run = items[flat.root & N.REF_MASK:]
    assert len(run) == 8 and run[-1] & N.LIST_LAST            # 5 walls + light + 2 boxes
    kinds = sorted((int(x) >> 28) & 7 for x in run)
    assert kinds == [N.REF_QUAD] * 6 + [N.REF_BOX] * 2
    noprim = s.flatten(32, 4)                                    # explicit options keep box primitives too
    assert noprim.n_boxes == 2
    # a larger tree keeps its nodes, in depth-first left-first order (left child = next node when it is a node)
    s1, _ = g.builtin_scene(1)
    f1 = s1.flatten()
    nodes = as_np(f1.nodes, f1.n_nodes, N.NODE_DTYPE)
    assert f1.n_nodes > 50
    for i, n in enumerate(nodes):
        if (n["left"] >> 28) & 7 == N.REF_NODE:
            assert (n["left"] & N.REF_MASK) == i + 1          # (bit 31 is the traversal hint)
        assert (n["bmin"] < n["bmax"]).all()


def test_boxes_are_rounded_outward():
    s, _ = g.builtin_scene(1)
    flat = s.flatten()
    nodes = as_np(flat.nodes, flat.n_nodes, N.NODE_DTYPE)
    wb = O.OracleWorld(s).world_bbox()
    assert (nodes[0]["bmin"].astype(np.float64) < wb[:3]).all() and (nodes[0]["bmax"].astype(np.float64) > wb[3:]).all()


def test_host_camera_matches_oracle_initialize():
    for sid in (1, 2, 6, 8):
        _, cfg = g.builtin_scene(sid)
        a, b = g.derive_camera(cfg), O.derived_camera(cfg)
        for f in ("width", "height", "spp_sqrt", "max_depth", "defocus_angle", "max_contribution"):
            assert getattr(a, f) == getattr(b, f)
        for f in ("center", "pixel00", "delta_u", "delta_v", "defocus_u", "defocus_v", "background"):
            assert np.allclose(list(getattr(a, f)), list(getattr(b, f)), rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("collapse", [None, (0, 0), (8, 2)])
def test_random_scenes_flatten_to_the_same_closest_hits(seed, collapse):
    """The flattener (instance baking, BuildBVH order, leaf runs, box records, hints) against the oracle, which walks
    the un-flattened description with the reference's transform-at-traversal-time recursion."""
    sc = PU.random_scene(seed)
    flat = sc.flatten() if collapse is None else sc.flatten(*collapse)
    fi = FlatInterp(flat)
    ow = O.OracleWorld(sc)
    rng = np.random.default_rng(1000 + seed)
    n = 500
    o = rng.uniform(-14, 14, size=(n, 3))
    tgt = rng.uniform(-6, 6, size=(n, 3))
    rays = PU.make_rays(o, tgt - o, time=rng.uniform(0, 1, size=n))
    oh = ow.trace_batch(rays, audit_eps=1e-6)
    assert (oh["id"] >= 0).mean() > 0.08
    bad = []
    for k, (r, h) in enumerate(zip(rays, oh)):
        if h["flags"]:
            continue
        fid, ft = fi.hit(r["o"].astype(np.float64), r["d"].astype(np.float64), float(r["time"]), float(r["tmin"]), float(r["tmax"]))
        if fid != h["id"] or (fid >= 0 and abs(ft - h["t"]) > 1e-5 * abs(h["t"])):
            bad.append((k, fid, int(h["id"]), ft, float(h["t"])))
    assert not bad, bad[:5]
