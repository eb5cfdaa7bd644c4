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


C_CLIENT = r"""
/* A plain C99 client of the boundary: what a cgo binding compiles against (INTEGRATION.md). */
#include <stdio.h>
#include <string.h>
#include "grt.h"
#include "grt_host.h"
int main(void) {
    if (grt_abi_version() != 1) return 10;
    GrtHostScene* s = grt_host_scene_new();
    GrtCameraConfig cam;
    GrtSceneOptions so;
    memset(&so, 0, sizeof so);
    so.width = 8; so.spp = 4;
    if (grt_host_builtin_scene(s, 6, &so, &cam)) return 11;          /* cornellBox, main.go:278 */
    GrtScene flat;
    if (grt_host_flatten(s, &flat)) return 12;
    if (flat.n_quads != 18 || flat.n_boxes != 2 || flat.n_lights != 1) return 13;
    GrtCamera dc;
    if (grt_host_camera_derive(&cam, &dc)) return 14;
    if (dc.width != 8 || dc.height != 8 || dc.spp_sqrt != 2 || dc.max_depth != 50) return 15;
    int ndev = grt_device_count();
    GrtSceneHandle h = 0;
    int rc = grt_scene_upload(&flat, 0, &h);
    if (ndev == 0) {                                                   /* no CPU fallback: a status code and a message */
        if (rc != GRT_E_NO_DEVICE || !grt_last_error()[0]) return 16;
    } else {
        if (rc) return 17;
        float sum[8 * 8 * 3]; unsigned char rgb[8 * 8 * 3];
        GrtOptions opt;
        memset(&opt, 0, sizeof opt); memset(sum, 0, sizeof sum);
        opt.seed = 0xC0FFEEull; opt.sample_stride = 1; opt.variant = GRT_VARIANT_AUTO;
        if (grt_render(h, &dc, &opt, sum, rgb, 0)) return 18;
        float m = 0; for (int i = 0; i < 8 * 8 * 3; i++) m += sum[i];
        if (!(m > 0)) return 19;
        grt_scene_free(h);
    }
    grt_host_scene_free(s);
    printf("c client ok (%d device%s)\n", ndev, ndev == 1 ? "" : "s");
    return 0;
}
"""


def _run_c_client(tmp_path):
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lgrt_cuda", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, f"C client failed with code {r.returncode}: {r.stdout} {r.stderr}"
    assert "c client ok" in r.stdout
    return r.stdout


def test_plain_c99_client_compiles_links_and_runs(tmp_path):
    """include/*.h are C headers (no C++ in the signatures) and the library is usable from plain C: builds a C99
    client with gcc, links it against libgrt_cuda.so and runs it.  Without a device the client checks that upload fails
    with GRT_E_NO_DEVICE and a message (no CPU fallback); test_plain_c99_client_renders_on_the_gpu is its GPU twin."""
    _run_c_client(tmp_path)


@pytest.mark.gpu
def test_plain_c99_client_renders_on_the_gpu(tmp_path):
    """The same C99 client under `-m gpu`: grt_scene_upload + grt_render with host buffers from plain C — the calls a
    cgo binding makes."""
    out = _run_c_client(tmp_path)
    assert "0 devices" not in out
