"""Parity pinned to OUTPUT OF THE REFERENCE ITSELF: the Go renderer's own images (README.md:39-49), see
tests/ref_render_util.py for the cases and bars.

  not gpu: the oracle (oracle/oracle.cpp, the fp64 restatement) against the Go renders — this is what pins the oracle's
           integrator, recursive clamp, media and light/cosine mixture sampling to the reference (camera.go:293-341);
  gpu:     the CUDA path through the C ABI (grt_render with host buffers, both variants, the device tonemap) against
           the same JPEGs with the same bars.
"""
import numpy as np
import pytest
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU
import ref_render_util as RU


def _earth():
    """earthmap.jpg as Go decodes it (tests/golden/make_earthmap_fixture.py, csrc/jpeg_go.hpp)."""
    import os
    return np.load(os.path.join(RU.HERE, "golden", "earthmap_rgb8.npz"))["rgb"]


def _quads_mask(cfg_like_width):
    """Pixels of the `quads` scene (main.go:219-246) whose colour does not depend on Go's unseeded math/rand (the Perlin
    tables of the right quad, which the metal quad also reflects): the sky, the light quad (z = 0), the teal quad
    (y = -3) and the earth-textured quad (x = -3; its texels are Go's, test_jpeg_go.py), found by tracing one ray
    through every pixel and eroded by two pixels."""
    s, cfg = g.builtin_scene(5, width=cfg_like_width, spp=1, image=_earth())
    cam = O.derived_camera(cfg)
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    oh = O.OracleWorld(s).trace_batch(rays)
    p = oh["p"]
    ok = (oh["id"] < 0) | (np.abs(p[:, 2]) < 1e-6) | (np.abs(p[:, 1] + 3.0) < 1e-6) | (np.abs(p[:, 0] + 3.0) < 1e-6)
    m = ok.reshape(cam.height, cam.width)
    er = m.copy()
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            er &= np.roll(np.roll(m, dy, axis=0), dx, axis=1)
    er[:2] = er[-2:] = False
    er[:, :2] = er[:, -2:] = False
    return er


@pytest.mark.parametrize("name", sorted(RU.CASES))
def test_oracle_reproduces_the_go_renders(name):
    sid, width, spp = RU.CASES[name]
    s, cfg = g.builtin_scene(sid, width=width, spp=spp)
    cam = O.derived_camera(cfg)
    sums, _, _, _ = O.OracleWorld(s).render(cfg)           # reference semantics: no self-exclusion, fp64
    img = RU.print_color(sums, cam.spp_sqrt ** 2)
    dmean, dblock, jm, jb = RU.assert_matches_reference(img, name, "oracle")
    # the sample count really is the one the readme image was made with (see ref_render_util's table)
    ref = RU.load_reference(name)
    assert 0.8 <= RU.roughness(img) / RU.roughness(ref) <= 1.25, (RU.roughness(img), RU.roughness(ref))
    print(f"{name}: channel-mean diff {dmean.round(3)}, block diff {dblock:.3f}; through the same JPEG tables {jm.round(3)}, {jb:.3f} (/255)")


def test_oracle_reproduces_the_go_quads_render_on_its_deterministic_pixels():
    s, cfg = g.builtin_scene(5, image=_earth())             # main.go:237-240 as shipped: 400x400, 100 spp
    cam = O.derived_camera(cfg)
    assert (cam.width, cam.height, cam.spp_sqrt) == (400, 400, 10)
    sums, _, _, _ = O.OracleWorld(s).render(cfg)
    mask = _quads_mask(400)
    assert mask.mean() > 0.5
    RU.assert_matches_reference(RU.print_color(sums, 100), "quads", "oracle", mask)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["mega", "wavefront"])
@pytest.mark.parametrize("name", sorted(RU.CASES))
def test_cuda_render_reproduces_the_go_renders(name, variant):
    sid, width, spp = RU.CASES[name]
    s, cfg = g.builtin_scene(sid, width=width, spp=spp)
    cam = g.derive_camera(cfg)
    v = g.GRT_VARIANT_MEGAKERNEL if variant == "mega" else g.GRT_VARIANT_WAVEFRONT
    sums, rgb8, _ = g.DeviceScene(s).render(cam, seed=0xBEEF, variant=v, want_rgb8=True)
    # the device tonemap (color.go:14-46, fp32) is what a caller writes to the PPM: the host formula (fp64) on the same
    # sums, up to one code value where sqrt(x) * 256 lands within fp32 rounding of an integer
    d = np.abs(rgb8.astype(np.float64) - RU.print_color(sums, cam.spp_sqrt ** 2))
    assert d.max() <= 1 and (d == 0).mean() > 0.999, (d.max(), (d == 0).mean())
    dmean, dblock, jm, jb = RU.assert_matches_reference(rgb8.astype(np.float64), name, f"CUDA {variant}")
    print(f"{name}/{variant}: channel-mean diff {dmean.round(3)}, block diff {dblock:.3f}; through the same JPEG tables {jm.round(3)}, {jb:.3f} (/255)")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["mega", "wavefront"])
def test_cuda_render_reproduces_the_go_quads_render_on_its_deterministic_pixels(variant):
    s, cfg = g.builtin_scene(5, image=_earth())
    cam = g.derive_camera(cfg)
    v = g.GRT_VARIANT_MEGAKERNEL if variant == "mega" else g.GRT_VARIANT_WAVEFRONT
    _, rgb8, _ = g.DeviceScene(s).render(cam, seed=0xBEEF, variant=v, want_rgb8=True)
    RU.assert_matches_reference(rgb8.astype(np.float64), "quads", f"CUDA {variant}", _quads_mask(400))
