"""Shared helpers for the parity tests: ray batches and comparison against the oracle."""
import numpy as np
import go_raytracer_b200 as g
from oracle import oracle_py as O

NO_ID = 0xFFFFFFFF
INF = np.float32(np.inf)


def make_rays(o, d, time=None, tmin=0.001, tmax=np.inf, self_id=None):
    n = len(o)
    r = np.zeros(n, dtype=g.RAY_DTYPE)
    r["o"] = np.asarray(o, dtype=np.float32)
    r["d"] = np.asarray(d, dtype=np.float32)
    r["tmin"] = np.float32(tmin)
    r["tmax"] = np.float32(tmax)
    r["time"] = 0 if time is None else np.asarray(time, dtype=np.float32)
    r["self_id"] = NO_ID if self_id is None else np.asarray(self_id, dtype=np.uint32)
    return r


def primary_batch(cfg, window, sample=0, seed=0xC0FFEE):
    """Camera rays of getRay (camera.go:256-270) for stratum `sample`, rounded to fp32 (the common input)."""
    o, d, t = O.primary_rays(cfg, seed, window, sample)
    return make_rays(o, d, t)


def secondary_batch(orc_hits, rng, time=None, around_normal=True):
    """Rays leaving the oracle's first-bounce hit points: origin on a primitive (self_id set)."""
    ok = orc_hits["id"] >= 0
    p = orc_hits["p"][ok]
    n = orc_hits["n"][ok]
    v = rng.normal(size=p.shape)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if around_normal:
        # hemisphere around the face-forwarded normal, like a diffuse bounce; unnormalised lengths like light samples
        flip = np.sum(v * n, axis=1) < 0
        v[flip] *= -1
    v *= rng.uniform(0.3, 300.0, size=(len(v), 1))
    tm = None if time is None else np.asarray(time)[ok]
    return make_rays(p, v, tm, self_id=orc_hits["id"][ok].astype(np.uint32))


def compare_hits(gpu, orc, audit=True, t_rel=1e-5):
    """Parity checks 1 and 2: ids bit-exact except documented ties/edges (flags), t within t_rel."""
    gid = gpu["id"].astype(np.int64)
    gid[gpu["id"] == NO_ID] = -1
    oid = orc["id"].astype(np.int64)
    flagged = orc["flags"] != 0 if audit else np.zeros(len(orc), bool)
    id_mismatch = (gid != oid)
    bad_id = id_mismatch & ~flagged
    both = (gid >= 0) & (oid >= 0) & ~id_mismatch
    rel = np.zeros(len(orc))
    rel[both] = np.abs(gpu["t"][both].astype(np.float64) - orc["t"][both]) / np.maximum(np.abs(orc["t"][both]), 1e-30)
    bad_t = both & (rel > t_rel) & ~flagged
    return {
        "n": len(orc), "hits": int((oid >= 0).sum()), "flagged": int(flagged.sum()),
        "id_mismatch_total": int(id_mismatch.sum()), "id_mismatch_unflagged": int(bad_id.sum()),
        "t_max_rel_unflagged": float(rel[both & ~flagged].max()) if (both & ~flagged).any() else 0.0,
        "t_bad": int(bad_t.sum()), "bad_id_idx": np.nonzero(bad_id)[0][:10], "bad_t_idx": np.nonzero(bad_t)[0][:10],
    }


def ref_bvh_order(boxes):
    """CPU checker for grt_bvh_order: the object order bvhHelper (bvh.go:35-61) ends with — every span of >= 3
    objects sorted (stably) by boxCompare (bvh.go:25-32) along LongestAxis (aabb.go:73-87) of the span's padded
    union box (aabb.go:54-59,118-129), split at the median, recursively."""
    import numpy as np
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 6)
    order = np.arange(b.shape[0])
    stack = [(0, b.shape[0])]
    while stack:
        s, e = stack.pop()
        if e - s < 3:
            continue
        idx = order[s:e]
        lo, hi = b[idx, :3].min(axis=0), b[idx, 3:].max(axis=0)
        size = []
        for a in range(3):
            l, h = lo[a], hi[a]
            if h - l < 0.0001:
                l, h = l - 0.0001 / 2, h + 0.0001 / 2
            size.append(h - l)
        if size[0] > size[1]:
            axis = 0 if size[0] > size[2] else 2
        else:
            axis = 1 if size[1] > size[2] else 2
        k = np.lexsort((b[idx, 3 + axis], b[idx, axis]))    # stable; primary key = box min, then box max
        order[s:e] = idx[k]
        mid = s + (e - s) // 2
        stack.append((s, mid)); stack.append((mid, e))
    return order


def random_scene(seed):
    """Random instancing / grouping of every flat-able surface type: spheres (static, moving), quads, boxes, triangles,
    under random Translate / RotateY chains, grouped into lists and BVHs of random size, nested up to three deep."""
    import numpy as np
    import go_raytracer_b200 as g
    rng = np.random.default_rng(seed)
    sc = g.Scene()
    mat = sc.NewLambertian((.5, .5, .5))

    def prim():
        k = rng.integers(0, 5)
        c = rng.uniform(-8, 8, size=3)
        if k == 0:
            return sc.NewSphere(tuple(c), float(rng.uniform(0.2, 1.5)), mat)
        if k == 1:
            return sc.NewMotionSphere(tuple(c), tuple(c + rng.uniform(-1, 1, size=3)), float(rng.uniform(0.2, 1.0)), mat)
        if k == 2:
            return sc.NewQuad(tuple(c), tuple(rng.uniform(-2, 2, size=3)), tuple(rng.uniform(-2, 2, size=3)), mat)
        if k == 3:
            return sc.NewBox(tuple(c), tuple(c + rng.uniform(0.3, 2.0, size=3)), mat)
        v = [tuple(c + rng.uniform(-1.5, 1.5, size=3)) for _ in range(3)]
        return sc.NewTriangle(v, mat)

    def instance(obj):
        for _ in range(rng.integers(0, 3)):
            if rng.random() < 0.5:
                obj = sc.Translate(obj, tuple(rng.uniform(-3, 3, size=3)))
            else:
                obj = sc.RotateY(obj, float(rng.uniform(-180, 180)))
        return obj

    def group(depth):
        n = int(rng.integers(1, 12 if depth else 40))
        objs = []
        for _ in range(n):
            if depth < 2 and rng.random() < 0.15:
                objs.append(instance(group(depth + 1)))
            else:
                objs.append(instance(prim()))
        lst = sc.NewHittableList(objs)
        return sc.BuildBVH(lst) if rng.random() < 0.7 else lst

    light = sc.NewQuad((-2, 20, -2), (4, 0, 0), (0, 0, 4), sc.NewDiffuseLight((5, 5, 5)))
    world = sc.NewHittableList([group(0), light])
    sc.set_world(world)
    sc.set_lights(sc.NewHittableList([light]))
    return sc


def random_shaded_scene(seed):
    """random_scene's geometry with every material class (Lambertian with solid / checker / noise textures, metal,
    dielectric, isotropic) and one or two constant media whose boundaries are instanced spheres or boxes."""
    import numpy as np
    import go_raytracer_b200 as g
    rng = np.random.default_rng(5000 + seed)
    sc = g.Scene()

    def material():
        k = rng.integers(0, 6)
        c = tuple(rng.uniform(0.2, 0.9, size=3))
        if k == 0:
            return sc.NewLambertian(c)
        if k == 1:
            return sc.NewTexturedLambertian(sc.NewCheckerboardColors(float(rng.uniform(0.3, 2.0)), c, (0.9, 0.9, 0.9)))
        if k == 2:
            return sc.NewTexturedLambertian(sc.NewNoiseTextureWithType(float(rng.uniform(0.5, 4.0)), int(rng.integers(1, 4)), int(seed + 1)))
        if k == 3:
            return sc.NewMetal(c, float(rng.uniform(0.0, 0.6)))
        if k == 4:
            return sc.NewDielectric(float(rng.uniform(1.2, 1.8)))
        return sc.NewLambertian(c)

    def instance(obj):
        for _ in range(rng.integers(0, 3)):
            obj = sc.Translate(obj, tuple(rng.uniform(-2, 2, size=3))) if rng.random() < 0.5 else sc.RotateY(obj, float(rng.uniform(-180, 180)))
        return obj

    objs = [sc.NewQuad((-12, -3, -12), (24, 0, 0), (0, 0, 24), sc.NewLambertian((.6, .6, .6)))]      # a floor
    for _ in range(int(rng.integers(12, 70))):     # below ~32 leaves the flattener emits one run, above it a BVH
        c = rng.uniform(-6, 6, size=3)
        k = rng.integers(0, 4)
        m = material()
        if k == 0:
            o = sc.NewSphere(tuple(c), float(rng.uniform(0.4, 1.6)), m)
        elif k == 1:
            o = sc.NewBox(tuple(c), tuple(c + rng.uniform(0.5, 2.0, size=3)), m)
        elif k == 2:
            o = sc.NewQuad(tuple(c), tuple(rng.uniform(-2, 2, size=3)), tuple(rng.uniform(-2, 2, size=3)), m)
        else:
            o = sc.NewTriangle([tuple(c + rng.uniform(-1.5, 1.5, size=3)) for _ in range(3)], m)
        objs.append(instance(o))
    for _ in range(int(rng.integers(1, 3))):
        c = rng.uniform(-5, 5, size=3)
        white = sc.NewLambertian((1, 1, 1))
        if rng.random() < 0.5:
            b = sc.NewSphere(tuple(c), float(rng.uniform(1.0, 2.5)), white)
        else:
            b = sc.NewBox(tuple(c), tuple(c + rng.uniform(1.0, 3.0, size=3)), white)
        objs.append(sc.ConstantMedium(instance(b), float(rng.uniform(0.05, 0.6)), tuple(rng.uniform(0.1, 0.9, size=3))))
    light = sc.NewQuad((-3, 9, -3), (6, 0, 0), (0, 0, 6), sc.NewDiffuseLight((6, 6, 6)))
    objs.append(light)
    lst = sc.NewHittableList(objs)
    sc.set_world(sc.BuildBVH(lst) if seed % 2 == 1 else lst)      # odd seeds: a BVH over everything, media included
    sc.set_lights(sc.NewHittableList([light]))
    cam = g.Camera()
    cam.AspectRatio, cam.Width, cam.SamplesPerPixel, cam.MaxDepth = 1.0, 40, 16, 20
    cam.VerticalFOV, cam.Background = 50.0, (0.3, 0.4, 0.6)
    cam.PositionCamera((0, 3, -16), (0, 0, 0), (0, 1, 0))
    return sc, cam.config()
