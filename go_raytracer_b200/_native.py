"""ctypes binding of libgrt_cuda.so (include/grt.h, include/grt_host.h).

The library is the product: hand-written sm_100a kernels behind a C ABI.  There
is no Python or CPU fallback — if the shared library is missing this module
raises, and every compute entry point fails with GRT_E_NO_DEVICE on a machine
without a CUDA device.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRT_CUDA_LIB") or os.path.join(_HERE, "csrc", "libgrt_cuda.so")   # env override: A/B builds

GRT_OK, GRT_E_INVALID, GRT_E_NO_DEVICE, GRT_E_CUDA, GRT_E_UNSUPPORTED, GRT_E_NCCL = 0, -1, -2, -3, -4, -5
GRT_VARIANT_MEGAKERNEL, GRT_VARIANT_WAVEFRONT, GRT_VARIANT_AUTO = 0, 1, 2
GRT_OPT_STATS, GRT_OPT_ATOMIC_SUM, GRT_OPT_TIMING = 1, 2, 4
GRT_NO_ID = 0xFFFFFFFF
REF_SHIFT, REF_MASK = 28, 0x0FFFFFFF
REF_NODE, REF_SPHERE, REF_QUAD, REF_TRI, REF_LIST, REF_MEDIUM, REF_BOX, REF_NONE = 0, 1, 2, 3, 4, 5, 6, 7
LIST_LAST = 0x80000000


class GrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgrt_cuda error {code}: {msg}")
        self.code = code


class GrtCameraConfig(C.Structure):
    """Public fields of camera.Camera (camera.go:24-36) + PositionCamera (:65)."""
    _fields_ = [("AspectRatio", C.c_double), ("Width", C.c_int32), ("SamplesPerPixel", C.c_int32),
                ("MaxDepth", C.c_int32), ("MaxThreads", C.c_int32), ("VerticalFOV", C.c_double),
                ("DefocusAngle", C.c_double), ("FocusDistance", C.c_double), ("Background", C.c_double * 3),
                ("MaxContribution", C.c_double), ("lookFrom", C.c_double * 3), ("lookAt", C.c_double * 3),
                ("vup", C.c_double * 3)]


class GrtSceneOptions(C.Structure):
    _fields_ = [("width", C.c_int32), ("spp", C.c_int32), ("aspect", C.c_double), ("seed", C.c_uint64),
                ("mesh_segments", C.c_int32), ("image_w", C.c_int32), ("image_h", C.c_int32),
                ("image_rgb", C.c_void_p)]


class GrtObjOptions(C.Structure):
    """objLoader.LoadObjOptions (objLoader.go:17-29)."""
    _fields_ = [("ScaleFactor", C.c_double), ("FlipYZ", C.c_int32), ("IgnoreNormals", C.c_int32), ("Center", C.c_int32),
                ("FlipFaces", C.c_int32), ("IgnoreMtl", C.c_int32), ("FindWindows", C.c_int32),
                ("Position", C.c_double * 3), ("DefaultMaterial", C.c_int32)]


class GrtCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp_sqrt", C.c_int32), ("max_depth", C.c_int32),
                ("center", C.c_double * 3), ("pixel00", C.c_double * 3), ("delta_u", C.c_double * 3),
                ("delta_v", C.c_double * 3), ("defocus_u", C.c_double * 3), ("defocus_v", C.c_double * 3),
                ("defocus_angle", C.c_double), ("background", C.c_double * 3), ("max_contribution", C.c_double)]


class GrtOptions(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("variant", C.c_int32), ("device", C.c_int32), ("sample_first", C.c_uint32),
                ("sample_stride", C.c_uint32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32),
                ("y1", C.c_int32), ("flags", C.c_uint32), ("pad", C.c_uint32)]


class GrtStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "box_tests", "sphere_tests", "quad_tests", "tri_tests",
                                          "medium_tests", "shade_diffuse", "shade_specular", "light_pdf_evals",
                                          "nan_samples", "warp_iterations", "lane_iterations")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class GrtTiming(C.Structure):
    """grt_last_timing: device time per kernel class of the last GRT_OPT_TIMING render (CUDA events)."""
    _fields_ = [("total_ms", C.c_double), ("generate_ms", C.c_double), ("extend_ms", C.c_double), ("shade_ms", C.c_double),
                ("extend_launches", C.c_uint64), ("launches", C.c_uint64), ("extend_kernel", C.c_uint64)]


class GrtScene(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("root", C.c_uint32),
                ("nodes", C.c_void_p), ("n_nodes", C.c_uint32),
                ("spheres", C.c_void_p), ("n_spheres", C.c_uint32),
                ("quads", C.c_void_p), ("n_quads", C.c_uint32),
                ("boxes", C.c_void_p), ("n_boxes", C.c_uint32),
                ("tris", C.c_void_p), ("n_tris", C.c_uint32),
                ("tri_shade", C.c_void_p), ("tri_v64", C.c_void_p),
                ("items", C.c_void_p), ("n_items", C.c_uint32),
                ("media", C.c_void_p), ("n_media", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32),
                ("textures", C.c_void_p), ("n_textures", C.c_uint32),
                ("images", C.c_void_p), ("n_images", C.c_uint32),
                ("texels", C.c_void_p), ("n_texel_bytes", C.c_uint64),
                ("perlins", C.c_void_p), ("n_perlins", C.c_uint32),
                ("lights", C.c_void_p), ("n_lights", C.c_uint32),
                ("lights_mode", C.c_uint32), ("max_depth_hint", C.c_uint32)]


# numpy views of the flat records (for tests that inspect the flattened scene)
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<u4"), ("right", "<u4")])
SPHERE_DTYPE = np.dtype([("c0", "<f8", 3), ("r", "<f8"), ("dc", "<f4", 3), ("mat", "<u4"), ("id", "<u4"),
                         ("flags", "<u4"), ("uvrot", "<f4", 2)])
QUAD_DTYPE = np.dtype([("n", "<f4", 3), ("D", "<f4"), ("Q", "<f4", 3), ("flags", "<u4"), ("A", "<f4", 3),
                       ("mat", "<u4"), ("B", "<f4", 3), ("id", "<u4"), ("n64", "<f8", 3), ("D64", "<f8")])
BOX_DTYPE = np.dtype([("mn", "<f4", 3), ("first_quad", "<u4"), ("mx", "<f4", 3), ("flags", "<u4"), ("T", "<f4", 3),
                      ("rc", "<f4"), ("rs", "<f4"), ("pad", "<f4", 3)])
TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("mat", "<u4"), ("e0", "<f4", 3), ("id", "<u4"), ("e1", "<f4", 3),
                      ("flags", "<u4")])
MATERIAL_DTYPE = np.dtype([("type", "<u4"), ("tex", "<u4"), ("albedo", "<f4", 3), ("fuzz", "<f4"), ("ior", "<f4"), ("pad", "<u4")])
TEXTURE_DTYPE = np.dtype([("type", "<u4"), ("color", "<f4", 3), ("scale", "<f4"), ("even", "<u4"), ("odd", "<u4"), ("aux", "<u4")])
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4"), ("time", "<f4"),
                      ("self_id", "<u4"), ("pad", "<u4", 2)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("id", "<u4"), ("ref", "<u4"), ("front_face", "<u4"), ("p", "<f4", 3),
                      ("u", "<f4"), ("n", "<f4", 3), ("v", "<f4")])
assert NODE_DTYPE.itemsize == 32 and SPHERE_DTYPE.itemsize == 64 and QUAD_DTYPE.itemsize == 96
assert TRI_DTYPE.itemsize == 48 and RAY_DTYPE.itemsize == 48 and HIT_DTYPE.itemsize == 48
assert MATERIAL_DTYPE.itemsize == 32 and TEXTURE_DTYPE.itemsize == 32 and BOX_DTYPE.itemsize == 64

_lib = None


def bvh_order(boxes, device=0):
    """grt_bvh_order: the object order BuildBVH (bvh.go:21-61) ends with, computed on the GPU.
    boxes: (n, 6) float64 {lo.xyz, hi.xyz} in list order; returns order[p] = list index at position p."""
    b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 6)
    out = np.empty(b.shape[0], dtype=np.uint32)
    check(lib().grt_bvh_order(b.ctypes.data, b.shape[0], int(device), out.ctypes.data))
    return out


def lib():
    """Load libgrt_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C go_raytracer_b200/csrc`). There is no fallback implementation.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, u64, dbl = C.c_void_p, C.c_int, C.c_uint64, C.c_double
    P = C.POINTER
    sig = {
        "grt_abi_version": (i32, []), "grt_device_count": (i32, []), "grt_last_error": (C.c_char_p, []),
        "grt_launch_count": (u64, []), "grt_last_timing": (i32, [P(GrtTiming)]), "grt_bvh_order": (i32, [vp, C.c_uint32, i32, vp]),
        "grt_scene_upload": (i32, [P(GrtScene), i32, P(vp)]), "grt_scene_free": (i32, [vp]),
        "grt_trace_batch": (i32, [vp, vp, u64, vp]), "grt_trace_batch_device": (i32, [vp, vp, u64, vp, vp]),
        "grt_render": (i32, [vp, P(GrtCamera), P(GrtOptions), vp, vp, P(GrtStats)]),
        "grt_render_device": (i32, [vp, P(GrtCamera), P(GrtOptions), vp, vp, vp]),
        "grt_tonemap_device": (i32, [vp, vp, u64, C.c_float, vp]),
        "grt_render_multi": (i32, [P(GrtScene), P(GrtCamera), P(GrtOptions), P(i32), i32, vp, vp, P(dbl)]),
        "grt_host_last_error": (C.c_char_p, []), "grt_host_scene_new": (vp, []), "grt_host_scene_free": (None, [vp]),
        "grt_host_solid_color": (i32, [vp, dbl, dbl, dbl]), "grt_host_checkerboard": (i32, [vp, dbl, i32, i32]),
        "grt_host_image": (i32, [vp, i32, i32, vp]),
        "grt_host_decode_jpeg": (i32, [vp, C.c_size_t, P(i32), P(i32), vp, C.c_size_t]), "grt_host_image_texture": (i32, [vp, i32]),
        "grt_host_noise_texture": (i32, [vp, dbl, i32, u64]),
        "grt_host_lambertian": (i32, [vp, i32]), "grt_host_metal": (i32, [vp, dbl, dbl, dbl, dbl]),
        "grt_host_dielectric": (i32, [vp, dbl]), "grt_host_diffuse_light": (i32, [vp, i32]),
        "grt_host_isotropic": (i32, [vp, i32]),
        "grt_host_sphere": (i32, [vp, P(dbl), dbl, i32]), "grt_host_motion_sphere": (i32, [vp, P(dbl), P(dbl), dbl, i32]),
        "grt_host_quad": (i32, [vp, P(dbl), P(dbl), P(dbl), i32]), "grt_host_box": (i32, [vp, P(dbl), P(dbl), i32]),
        "grt_host_triangle": (i32, [vp, P(dbl), P(dbl), P(dbl), i32]),
        "grt_host_list": (i32, [vp]), "grt_host_list_add": (i32, [vp, i32, i32]), "grt_host_bvh": (i32, [vp, i32]),
        "grt_host_translate": (i32, [vp, i32, P(dbl)]), "grt_host_rotate_y": (i32, [vp, i32, dbl]),
        "grt_host_constant_medium": (i32, [vp, i32, dbl, i32]),
        "grt_host_set_world": (i32, [vp, i32]), "grt_host_set_lights": (i32, [vp, i32]),
        "grt_host_load_obj": (i32, [vp, C.c_char_p, C.c_char_p, P(GrtObjOptions), P(i32), P(i32), P(i32)]),
        "grt_host_load_obj_file": (i32, [vp, C.c_char_p, P(GrtObjOptions), P(i32), P(i32), P(i32)]),
        "grt_host_builtin_scene": (i32, [vp, i32, P(GrtSceneOptions), P(GrtCameraConfig)]),
        "grt_host_flatten": (i32, [vp, P(GrtScene)]), "grt_host_flatten_opts": (i32, [vp, i32, i32, P(GrtScene)]), "grt_host_camera_derive": (i32, [P(GrtCameraConfig), P(GrtCamera)]),
        "grt_host_write_ppm": (C.c_long, [vp, i32, i32, vp, C.c_long]),
        "grt_host_write_p6": (C.c_long, [vp, i32, i32, vp, C.c_long]), "grt_host_write_png": (C.c_long, [vp, i32, i32, vp, C.c_long]),
        "grt_host_camera_render_rgb8": (i32, [vp, P(GrtCameraConfig), u64, i32, i32, vp, vp, vp, C.c_long, P(C.c_long), P(dbl)]),
        "grt_host_camera_render": (i32, [vp, P(GrtCameraConfig), u64, i32, i32, vp, vp, C.c_long, P(C.c_long), P(dbl)]),
        "grt_host_scene_description": (vp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = a symbol include/*.h declares is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    # include/grt.h
    "grt_abi_version", "grt_device_count", "grt_last_error", "grt_scene_upload", "grt_scene_free", "grt_trace_batch",
    "grt_trace_batch_device", "grt_render", "grt_render_device", "grt_tonemap_device", "grt_render_multi",
    "grt_launch_count", "grt_last_timing", "grt_bvh_order",
    # include/grt_host.h
    "grt_host_last_error", "grt_host_scene_new", "grt_host_scene_free", "grt_host_solid_color", "grt_host_checkerboard",
    "grt_host_image", "grt_host_decode_jpeg", "grt_host_image_texture", "grt_host_noise_texture", "grt_host_lambertian", "grt_host_metal",
    "grt_host_dielectric", "grt_host_diffuse_light", "grt_host_isotropic", "grt_host_sphere", "grt_host_motion_sphere",
    "grt_host_quad", "grt_host_box", "grt_host_triangle", "grt_host_list", "grt_host_list_add", "grt_host_bvh",
    "grt_host_translate", "grt_host_rotate_y", "grt_host_constant_medium", "grt_host_set_world", "grt_host_set_lights",
    "grt_host_load_obj", "grt_host_load_obj_file", "grt_host_builtin_scene", "grt_host_flatten", "grt_host_flatten_opts", "grt_host_camera_derive", "grt_host_write_ppm", "grt_host_write_p6", "grt_host_write_png", "grt_host_camera_render_rgb8",
    "grt_host_camera_render", "grt_host_scene_description",
]


def check(rc):
    if rc != 0:
        L = lib()
        msg = (L.grt_last_error() or b"").decode() or (L.grt_host_last_error() or b"").decode()
        raise GrtError(rc, msg)


def host_check(rc):
    if rc < 0:
        raise GrtError(rc, (lib().grt_host_last_error() or b"").decode())
    return rc
