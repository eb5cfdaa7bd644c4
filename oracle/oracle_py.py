"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under go_raytracer_b200/ does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

ORC_RAY = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4"), ("time", "<f4"),
                    ("self_id", "<u4"), ("pad", "<u4", 2)])
ORC_HIT = np.dtype([("t", "<f8"), ("id", "<i4"), ("front_face", "<i4"), ("p", "<f8", 3), ("n", "<f8", 3),
                    ("u", "<f8"), ("v", "<f8"), ("flags", "<i4"), ("pad", "<i4"), ("second_t", "<f8")])
assert ORC_RAY.itemsize == 48 and ORC_HIT.itemsize == 96


class OrcDerivedCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp_sqrt", C.c_int32), ("max_depth", C.c_int32),
                ("center", C.c_double * 3), ("pixel00", C.c_double * 3), ("delta_u", C.c_double * 3),
                ("delta_v", C.c_double * 3), ("defocus_u", C.c_double * 3), ("defocus_v", C.c_double * 3),
                ("defocus_angle", C.c_double), ("background", C.c_double * 3), ("max_contribution", C.c_double)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "box_tests", "sphere_tests", "quad_tests", "tri_tests",
                                          "medium_tests", "shade_diffuse", "shade_specular", "light_pdf_evals",
                                          "nan_samples")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, dbl, i32, u64, u32 = C.c_void_p, C.c_double, C.c_int, C.c_uint64, C.c_uint32
        P = C.POINTER
        L.orc_build.restype = vp; L.orc_build.argtypes = [vp]
        L.orc_free.restype = None; L.orc_free.argtypes = [vp]
        L.orc_trace_batch.restype = i32; L.orc_trace_batch.argtypes = [vp, vp, u64, vp, dbl, i32, i32]
        L.orc_camera_derive.restype = i32; L.orc_camera_derive.argtypes = [vp, P(OrcDerivedCamera)]
        L.orc_primary_rays.restype = i32; L.orc_primary_rays.argtypes = [vp, u64, i32, i32, i32, i32, i32, vp, vp, vp]
        L.orc_render.restype = i32
        L.orc_render.argtypes = [vp, vp, u64, u32, u32, i32, i32, i32, i32, i32, i32, vp, vp, P(OrcStats), P(dbl)]
        L.orc_write_ppm.restype = C.c_long; L.orc_write_ppm.argtypes = [vp, i32, i32, dbl, vp, C.c_long]
        L.orc_vec_op.restype = None; L.orc_vec_op.argtypes = [i32, P(dbl), P(dbl), dbl, P(dbl)]
        L.orc_vec_scalar.restype = dbl; L.orc_vec_scalar.argtypes = [i32, P(dbl), P(dbl)]
        L.orc_color_bytes.restype = None; L.orc_color_bytes.argtypes = [P(dbl), P(i32)]
        L.orc_interval.restype = dbl; L.orc_interval.argtypes = [i32, dbl, dbl, dbl]
        L.orc_ray_at.restype = None; L.orc_ray_at.argtypes = [P(dbl), P(dbl), dbl, P(dbl)]
        L.orc_aabb_hit.restype = i32; L.orc_aabb_hit.argtypes = [P(dbl), P(dbl), P(dbl), P(dbl), dbl, dbl]
        L.orc_philox.restype = None; L.orc_philox.argtypes = [P(u32), P(u32), P(u32)]
        L.orc_uniform.restype = dbl; L.orc_uniform.argtypes = [u64, u32, u32, u32, i32, u32]
        L.orc_texture_value.restype = None; L.orc_texture_value.argtypes = [vp, i32, dbl, dbl, P(dbl), P(dbl)]
        L.orc_world_bbox.restype = None; L.orc_world_bbox.argtypes = [vp, P(dbl)]
        _lib = L
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class OracleWorld:
    """The oracle's own object tree built from a scene DESCRIPTION (not from the flat arrays)."""

    def __init__(self, scene):
        self._L = lib()
        self._scene = scene   # keep the description alive
        self._w = self._L.orc_build(scene.description_ptr())
        if not self._w:
            raise RuntimeError("orc_build failed")

    def __del__(self):
        try:
            if self._w:
                self._L.orc_free(self._w)
                self._w = None
        except Exception:
            pass

    def trace_batch(self, rays, audit_eps=0.0, use_exclusion=False, nthreads=0):
        rays = np.ascontiguousarray(rays, dtype=ORC_RAY)
        hits = np.zeros(len(rays), dtype=ORC_HIT)
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        self._L.orc_trace_batch(self._w, rays.ctypes.data, len(rays), hits.ctypes.data, float(audit_eps),
                                int(use_exclusion), int(nthreads))
        return hits

    def render(self, cfg, seed=0xC0FFEE, sample_first=0, sample_stride=1, window=None, use_exclusion=False,
               nthreads=0, want_sumsq=False, want_stats=False):
        """Restated render loop.  Returns (sum[H,W,3] f64, sumsq or None, stats or None, seconds)."""
        cam = derived_camera(cfg)
        x0, y0, x1, y1 = window if window is not None else (0, 0, 0, 0)
        sums = np.zeros((cam.height, cam.width, 3), dtype=np.float64)
        sq = np.zeros_like(sums) if want_sumsq else None
        st = OrcStats()
        sec = C.c_double(0)
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        rc = self._L.orc_render(self._w, C.addressof(cfg), int(seed), int(sample_first), int(sample_stride), x0, y0, x1, y1,
                                int(use_exclusion), int(nthreads), sums.ctypes.data, sq.ctypes.data if want_sumsq else None,
                                C.byref(st), C.byref(sec))
        if rc != 0:
            raise RuntimeError("oracle: hit an invalid PDF function (hittable.go:69-72)")
        return sums, sq, (st.as_dict() if want_stats else None), sec.value

    def texture_value(self, tex, u, v, p):
        out = (C.c_double * 3)()
        self._L.orc_texture_value(self._w, int(tex), float(u), float(v), _d3(p), out)
        return np.array(out[:])

    def world_bbox(self):
        out = (C.c_double * 6)()
        self._L.orc_world_bbox(self._w, out)
        return np.array(out[:])


def derived_camera(cfg):
    d = OrcDerivedCamera()
    lib().orc_camera_derive(C.addressof(cfg), C.byref(d))
    return d


def primary_rays(cfg, seed, window, sample):
    """getRay (camera.go:256-270) in fp64 for the pixels of `window` and stratum `sample`."""
    x0, y0, x1, y1 = window
    n = (x1 - x0) * (y1 - y0)
    o = np.zeros((n, 3)); d = np.zeros((n, 3)); t = np.zeros(n)
    lib().orc_primary_rays(C.addressof(cfg), int(seed), x0, y0, x1, y1, int(sample), o.ctypes.data, d.ctypes.data, t.ctypes.data)
    return o, d, t


def write_ppm(sums, scale):
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    h, w, _ = sums.shape
    cap = 32 + w * h * 12
    buf = C.create_string_buffer(cap)
    n = lib().orc_write_ppm(sums.ctypes.data, w, h, float(scale), buf, cap)
    return buf.raw[:n]
