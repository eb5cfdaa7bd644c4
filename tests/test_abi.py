"""The C-ABI library loads, exports every symbol include/*.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("grt.h", "grt_host.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(grt_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported():
    L = C.CDLL(N.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 40
    for name in sorted(decl):
        assert hasattr(L, name), f"{name} is declared in include/*.h but not exported by libgrt_cuda.so"
    assert decl == set(N.EXPORTED_SYMBOLS)


def test_struct_sizes_match_header():
    # sizes the device code relies on (include/grt.h)
    assert C.sizeof(N.GrtCamera) == 16 + 6 * 24 + 8 + 24 + 8
    assert C.sizeof(N.GrtOptions) == 48
    assert C.sizeof(N.GrtStats) == 104
    assert g.lib().grt_abi_version() == 1


def test_host_layer_errors_are_codes_not_aborts():
    s = g.Scene()
    with pytest.raises(g.GrtError):
        s.NewSphere((0, 0, 0), 1, 99)                    # bad material id
    with pytest.raises(g.GrtError):
        s.flatten()                                      # no world / lights
    m = s.NewLambertian((1, 1, 1))
    sp = s.NewSphere((0, 0, 0), 1, m)
    s.set_world(s.NewHittableList([sp]))
    s.set_lights(s.BuildBVH(s.NewHittableList([sp])))    # BVHNode as lights: "hit an invalid PDF function" (hittable.go:69-72)
    with pytest.raises(g.GrtError, match="invalid PDF"):
        s.flatten()
    with pytest.raises(g.GrtError):
        g.builtin_scene(0)                               # defaultScene is empty (main.go:412)


@pytest.mark.skipif(g.lib().grt_device_count() > 0, reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_device():
    s, cfg = g.builtin_scene(6, width=8, spp=1)
    with pytest.raises(g.GrtError) as e:
        g.DeviceScene(s)
    assert e.value.code == N.GRT_E_NO_DEVICE


def test_ppm_writer_format():
    rgb = np.array([[[0, 255, 255], [1, 20, 100]]], dtype=np.uint8)
    assert g.write_ppm(rgb) == b"P3\n2 1\n255\n0 255 255\n1 20 100\n"      # camera.go:160, color.go:45
