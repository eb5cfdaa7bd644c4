"""Host logic of the device-internal 4-wide BVH (csrc/wide_bvh.hpp), checked without a GPU through the
grt_debug_repack hook: the structure the upload would put into HBM, and a Python emulation of the device
traversal (dev_trace.cuh: slab tests on child boxes, nearest-first order, leaf runs) against the oracle.

The closest hit must not depend on the tree shape: every primitive the oracle hits has to be among the primitives
the wide traversal still visits after culling, for every ray."""
import ctypes as C
import numpy as np
import pytest
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N
from oracle import oracle_py as O
import parity_util as PU

RUN_BIT, RUN_MASK = 0x80000000, 0x01FFFFFF


def repack(flat):
    L = N.lib()
    fn = L.grt_debug_repack
    fn.restype = C.c_int
    vp, u32p, ip = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    fn.argtypes = [C.POINTER(N.GrtScene), vp, C.c_uint32, u32p, vp, C.c_uint32, u32p, u32p, vp, ip, ip]
    cap_n, cap_e = max(1, flat.n_nodes), max(1, flat.n_items)
    wn = np.zeros((cap_n, 32), dtype=np.float32)
    en = np.zeros((cap_e, 2), dtype=np.uint32)
    mb = np.zeros(max(1, flat.n_media), dtype=np.uint32)
    nw, ne, root, nm, nb = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_int()
    N.check(fn(C.byref(flat), wn.ctypes.data, cap_n, C.byref(nw), en.ctypes.data, cap_e, C.byref(ne), C.byref(root), mb.ctypes.data,
               C.byref(nm), C.byref(nb)))
    return {"wnodes": wn[:nw.value], "entries": en[:ne.value], "root": root.value, "media": mb[:flat.n_media],
            "need_main": nm.value, "need_boundary": nb.value}


def arrays(flat):
    def view(ptr, n, dt):
        if not ptr or n == 0:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dtype=dt).copy()
    return {"nodes": view(flat.nodes, flat.n_nodes, N.NODE_DTYPE), "spheres": view(flat.spheres, flat.n_spheres, N.SPHERE_DTYPE),
            "quads": view(flat.quads, flat.n_quads, N.QUAD_DTYPE), "boxes": view(flat.boxes, flat.n_boxes, N.BOX_DTYPE),
            "tris": view(flat.tris, flat.n_tris, N.TRI_DTYPE), "items": view(flat.items, flat.n_items, np.dtype("<u4")),
            "media": view(flat.media, flat.n_media, np.dtype([("boundary", "<u4"), ("nid", "<f4"), ("mat", "<u4"), ("id", "<u4")]))}


def rtype(r):
    return (int(r) >> 28) & 7


def node_fields(w):
    lo = np.stack([w[0:4], w[4:8], w[8:12]], axis=1).astype(np.float64)      # [child, axis]
    hi = np.stack([w[12:16], w[16:20], w[20:24]], axis=1).astype(np.float64)
    refs = w[24:28].view(np.uint32)
    meta = w[28:32].view(np.uint32)
    return lo, hi, refs, meta


def prims_of_ref(ref, R, out):
    """All primitive refs (type << 28 | index) reachable from a device ref, runs and lists expanded; media as refs."""
    ref = int(ref)
    if ref & RUN_BIT:
        t, cnt, first = (ref >> 28) & 7, ((ref >> 25) & 7) + 1, ref & RUN_MASK
        out.extend((t << 28) | (first + k) for k in range(cnt))
        return
    t, i = rtype(ref), ref & N.REF_MASK
    if t == N.REF_NODE:
        lo, hi, refs, meta = node_fields(R["wnodes"][i])
        for k in range(int(meta[1])):
            prims_of_ref(refs[k], R, out)
    elif t == N.REF_LIST:
        k = i
        while True:
            e, cnt = int(R["entries"][k][0]), int(R["entries"][k][1])
            first = e & ~N.LIST_LAST
            if rtype(first) in (N.REF_SPHERE, N.REF_QUAD, N.REF_TRI, N.REF_BOX):
                out.extend(first + j for j in range(cnt))
            else:
                prims_of_ref(first, R, out)
            if e & N.LIST_LAST:
                break
            k += 1
    elif t != N.REF_NONE:
        out.append(ref)


def binary_prims(ref, A, out):
    """The same walk over the ABI's binary tree (GrtNode / items)."""
    ref = int(ref) & ~0x80000000
    t, i = rtype(ref), ref & N.REF_MASK
    if t == N.REF_NODE:
        binary_prims(A["nodes"][i]["left"], A, out)
        binary_prims(A["nodes"][i]["right"], A, out)
    elif t == N.REF_LIST:
        k = i
        while True:
            e = int(A["items"][k])
            binary_prims(e & ~N.LIST_LAST, A, out)
            if e & N.LIST_LAST:
                break
            k += 1
    elif t != N.REF_NONE:
        out.append(ref)


SCENES = [(1, {}), (2, {}), (3, {}), (6, {}), (7, {}), (8, {"mesh_segments": 24})]


@pytest.fixture(autouse=True, params=["sah", "plain"])
def tree_build(request, monkeypatch):
    """Every test runs on both trees the upload can build: medium-free subtrees regrouped by the surface area heuristic
    (the default) and BuildBVH's own topology, only widened (GRT_WIDE_SAH=0)."""
    monkeypatch.setenv("GRT_WIDE_SAH", "1" if request.param == "sah" else "0")
    return request.param


@pytest.mark.parametrize("sid,kw", SCENES)
def test_wide_tree_holds_exactly_the_binary_trees_primitives(sid, kw):
    s, cfg = g.builtin_scene(sid, width=64, spp=1, **kw)
    flat = s.flatten()
    A, R = arrays(flat), repack(flat)
    wide, binary = [], []
    prims_of_ref(R["root"], R, wide)
    binary_prims(flat.root, A, binary)
    assert set(wide) == set(binary)
    # nothing is visited more often than the reference visits it (span-1 duplicates of surfaces are dropped)
    from collections import Counter
    cw, cb = Counter(wide), Counter(binary)
    assert all(cw[k] <= cb[k] for k in cw)
    media = [k for k in cb if rtype(k) == N.REF_MEDIUM]
    assert all(cw[k] == cb[k] for k in media)          # a medium draws per test (medium.go:47): same multiplicity
    # breadth-first numbering: children come after their parent, so a prefix of the array is the top of the tree
    for i, w in enumerate(R["wnodes"]):
        lo, hi, refs, meta = node_fields(w)
        assert 1 <= meta[1] <= 4
        for k in range(4):
            r = int(refs[k])
            if k >= meta[1]:
                assert rtype(r) == N.REF_NONE and lo[k, 0] > hi[k, 0]        # empty slot: empty box
            elif not (r & RUN_BIT) and rtype(r) == N.REF_NODE:
                assert (r & N.REF_MASK) > i
    assert R["need_main"] <= 48 and R["need_boundary"] <= 32


def _prim_points(ref, A):
    """Points that must lie inside any box claiming to bound this primitive."""
    t, i = rtype(ref), ref & N.REF_MASK
    if t == N.REF_SPHERE:
        s = A["spheres"][i]
        pts = []
        for time in (0.0, 1.0):
            c = s["c0"] + time * s["dc"].astype(np.float64)
            pts += [c - s["r"], c + s["r"]]
        return np.array(pts)
    if t == N.REF_TRI:
        tr = A["tris"][i]
        v0 = tr["v0"].astype(np.float64)
        return np.array([v0, v0 + tr["e0"], v0 + tr["e1"]])
    if t == N.REF_BOX:
        b = A["boxes"][i]
        pts = []
        for k in range(8):
            o = np.array([b["mx"][0] if k & 1 else b["mn"][0], b["mx"][1] if k & 2 else b["mn"][1], b["mx"][2] if k & 4 else b["mn"][2]], dtype=np.float64)
            pts.append(np.array([b["rc"] * o[0] + b["rs"] * o[2], o[1], -b["rs"] * o[0] + b["rc"] * o[2]]) + b["T"])
        return np.array(pts)
    return None


@pytest.mark.parametrize("sid,kw", SCENES)
def test_every_child_box_bounds_what_is_below_it(sid, kw):
    s, cfg = g.builtin_scene(sid, width=64, spp=1, **kw)
    flat = s.flatten()
    A, R = arrays(flat), repack(flat)
    for w in R["wnodes"]:
        lo, hi, refs, meta = node_fields(w)
        for k in range(int(meta[1])):
            below = []
            prims_of_ref(refs[k], R, below)
            for p in below:
                if rtype(p) == N.REF_MEDIUM:
                    continue
                pts = _prim_points(p, A)
                if pts is None:
                    continue
                assert (pts >= lo[k] - 1e-9).all() and (pts <= hi[k] + 1e-9).all(), (p, lo[k], hi[k], pts)


def _emulate(R, ray_o, ray_d, tmin=0.001):
    """dev_trace.cuh's traversal without primitive tests: returns (set of primitive refs whose leaf is reached when no
    hit ever shrinks tmax — a superset of what the device tests —, the deepest stack seen)."""
    inv = 1.0 / ray_d
    stack, deepest, seen = [R["root"]], 1, []
    while stack:
        ref = int(stack.pop())
        if not (ref & RUN_BIT) and rtype(ref) == N.REF_NODE:
            lo, hi, refs, meta = node_fields(R["wnodes"][ref & N.REF_MASK])
            hits = []
            for k in range(int(meta[1])):
                with np.errstate(invalid="ignore", over="ignore"):
                    t0, t1 = (lo[k] - ray_o) * inv, (hi[k] - ray_o) * inv
                tn = np.fmax(np.fmax.reduce(np.fmin(t0, t1)), tmin)
                tf = np.fmin.reduce(np.fmax(t0, t1))
                if not (tf <= tn):
                    hits.append((tn if meta[0] & 1 else k, int(refs[k])))
            hits.sort(key=lambda h: h[0])
            for _, r in reversed(hits):
                stack.append(r)
            deepest = max(deepest, len(stack))
        elif not (ref & RUN_BIT) and rtype(ref) == N.REF_LIST:
            k = ref & N.REF_MASK
            while True:
                e, cnt = int(R["entries"][k][0]), int(R["entries"][k][1])
                first = e & ~N.LIST_LAST
                last = bool(e & N.LIST_LAST)
                if rtype(first) in (N.REF_SPHERE, N.REF_QUAD, N.REF_TRI, N.REF_BOX):
                    seen.extend(first + j for j in range(cnt))
                else:
                    if not last:
                        stack.append((N.REF_LIST << 28) | (k + 1))
                    stack.append(first)
                    deepest = max(deepest, len(stack))
                    break
                if last:
                    break
                k += 1
        else:
            prims_of_ref(ref, R, seen)
    return seen, deepest


@pytest.mark.parametrize("sid,kw", [(1, {}), (2, {}), (8, {"mesh_segments": 24})])
def test_culling_never_loses_the_oracles_hit(sid, kw):
    s, cfg = g.builtin_scene(sid, width=24, spp=1, **kw)
    flat = s.flatten()
    A, R = arrays(flat), repack(flat)
    ow = O.OracleWorld(s)
    cam = O.derived_camera(cfg)
    prim = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    oh = ow.trace_batch(prim)
    sec = PU.secondary_batch(oh, np.random.default_rng(sid))
    sec["self_id"] = PU.NO_ID
    rays = np.concatenate([prim, sec])
    oh = ow.trace_batch(rays)
    ids = {N.REF_SPHERE: A["spheres"]["id"], N.REF_TRI: A["tris"]["id"], N.REF_QUAD: A["quads"]["id"]}
    deepest_all, checked = 0, 0
    for ray, h in zip(rays[::3], oh[::3]):
        seen, deepest = _emulate(R, ray["o"].astype(np.float64), ray["d"].astype(np.float64))
        deepest_all = max(deepest_all, deepest)
        if h["id"] < 0:
            continue
        seen_ids = set()
        for p in seen:
            t, i = rtype(p), p & N.REF_MASK
            if t == N.REF_BOX:
                fq = int(A["boxes"][i]["first_quad"])
                seen_ids.update(int(x) for x in A["quads"]["id"][fq:fq + 6])
            elif t == N.REF_MEDIUM:
                seen_ids.add(int(A["media"][i]["id"]))
            elif t in ids:
                seen_ids.add(int(ids[t][i]))
        assert int(h["id"]) in seen_ids, f"the oracle hits object {h['id']} but the wide traversal culled it"
        checked += 1
    assert checked > 50
    assert deepest_all <= R["need_main"], (deepest_all, R["need_main"])


def _inner_area(R):
    """Sum of the box areas of all inner-node children below the huge ones (ground sphere, world list): proportional to
    the node visits of a random ray."""
    wn = R["wnodes"]
    lo = wn[:, 0:12].reshape(-1, 3, 4).astype(np.float64)
    hi = wn[:, 12:24].reshape(-1, 3, 4).astype(np.float64)
    ref = wn[:, 24:28].copy().view(np.uint32)
    cnt = wn[:, 29].copy().view(np.uint32)
    e = np.clip(hi - lo, 0, None)
    area = e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]
    valid = (np.arange(4)[None, :] < cnt[:, None]) & np.isfinite(area)
    area = np.where(valid, area, 0.0)
    inner = valid & ((ref >> 31) == 0) & (area < 0.01 * area.max())
    return area[inner].sum()


@pytest.mark.parametrize("sid,kw", [(1, {}), (8, {"mesh_segments": 64})])
def test_sah_regrouping_lowers_the_expected_node_visits(sid, kw, monkeypatch, tree_build):
    if tree_build != "sah":
        pytest.skip("compares the two builds itself")
    s, cfg = g.builtin_scene(sid, width=64, spp=1, **kw)
    flat = s.flatten()
    monkeypatch.setenv("GRT_WIDE_SAH", "0")
    plain = repack(flat)
    monkeypatch.setenv("GRT_WIDE_SAH", "1")
    sah = repack(flat)
    assert len(sah["wnodes"]) <= len(plain["wnodes"])
    assert _inner_area(sah) < 0.9 * _inner_area(plain), (_inner_area(sah), _inner_area(plain))
    assert sah["need_main"] <= 64


def test_subtrees_with_a_medium_keep_the_references_order(monkeypatch, tree_build):
    """Book 2 holds two constant media in its world list.  A medium test depends on the closest hit among everything
    BEFORE it in the reference's order and draws a random number (hittable.go:129-136, medium.go:38-47), so: every node
    on a path from the root to a medium is flagged in-order (meta bit 0 clear), an in-order walk meets the media in the
    reference's order, and the SET of surfaces between two consecutive media is the reference's — how each such set is
    boxed (BuildBVH's topology or the SAH regrouping) is free."""
    if tree_build != "sah":
        pytest.skip("compares the two builds itself")
    s, cfg = g.builtin_scene(2, width=64, spp=1, image=np.zeros((2, 2, 3), np.uint8))
    flat = s.flatten()
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GRT_WIDE_SAH", mode)
        R = repack(flat)

        def has_medium(ref):
            ref = int(ref)
            if ref & RUN_BIT:
                return False
            t, i = rtype(ref), ref & N.REF_MASK
            if t == N.REF_MEDIUM:
                return True
            if t == N.REF_NODE:
                lo, hi, refs, meta = node_fields(R["wnodes"][i])
                return any(has_medium(refs[k]) for k in range(int(meta[1])))
            return False

        segments, media = [set()], []

        def walk(ref):
            lo, hi, refs, meta = node_fields(R["wnodes"][int(ref) & N.REF_MASK])
            assert (int(meta[0]) & 1) == 0, "a node above a medium is flagged nearest-first"
            for k in range(int(meta[1])):
                c = int(refs[k])
                if not (c & RUN_BIT) and rtype(c) == N.REF_MEDIUM:
                    media.append(c)
                    segments.append(set())
                elif has_medium(c):
                    walk(c)
                else:
                    prims = []
                    prims_of_ref(c, R, prims)
                    segments[-1].update(prims)

        assert has_medium(R["root"])
        walk(R["root"])
        out[mode] = (media, segments)
    assert out["0"][0] == out["1"][0] and len(out["0"][0]) == 2
    assert out["0"][1] == out["1"][1]
    assert sum(len(x) for x in out["1"][1]) > 1000
