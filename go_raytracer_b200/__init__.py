"""go_raytracer_b200 — B200-native backend for go_raytracer's render hot path.

The product is `csrc/libgrt_cuda.so` (hand-written sm_100a CUDA behind the C ABI
of include/grt.h).  This package is the thin host-side mirror of the reference's
Go API used by tests and bench: `Scene` (package hittable's constructors),
`Camera` (camera.Camera) and `DeviceScene` (upload / trace_batch / render).
"""
from ._native import (GrtError, GrtCameraConfig, GrtCamera, GrtOptions, GrtStats, GrtScene, RAY_DTYPE, HIT_DTYPE,
                      GRT_VARIANT_MEGAKERNEL, GRT_VARIANT_WAVEFRONT, GRT_VARIANT_AUTO, GRT_NO_ID, lib, LIB_PATH, bvh_order)
from .scene import Scene, builtin_scene, decode_jpeg, load_image, SCENE_NAMES, PERLIN, MARBLE, TURBULENT
from .camera import Camera, DeviceScene, derive_camera, write_ppm, write_p6, write_png

__all__ = ["GrtError", "GrtCameraConfig", "GrtCamera", "GrtOptions", "GrtStats", "GrtScene", "RAY_DTYPE", "HIT_DTYPE",
           "GRT_VARIANT_MEGAKERNEL", "GRT_VARIANT_WAVEFRONT", "GRT_VARIANT_AUTO", "GRT_NO_ID", "lib", "LIB_PATH", "bvh_order", "Scene", "builtin_scene", "decode_jpeg", "load_image",
           "SCENE_NAMES", "PERLIN", "MARBLE", "TURBULENT", "Camera", "DeviceScene", "derive_camera", "write_ppm", "write_p6", "write_png"]
