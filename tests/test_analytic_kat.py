"""Analytic known-answer tests of the integrator: closed-form radiance that no restatement can share an error with.

Each case is rendered by the oracle (not gpu) and by the CUDA path, both variants (gpu), and compared with the
ANALYTIC answer — never with each other.  MaxContribution is set huge so the firefly clamp (camera.go:334-341)
is the identity and the estimator of camera.go:319-330 is an unbiased mixture importance sampler.

 1. White furnace.  A closed room whose six walls are DiffuseLights of radiance E facing inwards (materials.go:142-155
    emits on the front face only).  Inside: a Lambertian sphere of albedo 1, a Dielectric sphere, a fuzzy Metal sphere
    of albedo 1, an Isotropic ConstantMedium of albedo 1 in a rotated box.  Every material conserves energy, so the
    radiance arriving from ANY direction is E and every pixel's expectation is exactly E — whatever the light /
    cosine mixture (pdf.go:65-74), the re-intersecting light pdf (objects.go:152-160), Schlick / refraction
    (materials.go:94-130), the medium's free-flight sampling (medium.go:27-58) or the 50-bounce recursion do, as
    long as they are unbiased.
 2. Grey furnace.  One convex Lambertian object of albedo rho in the same room: it cannot see itself, so its
    radiance is exactly rho * E per channel (sphere, and a box under RotateY + Translate).
 3. Direct irradiance.  The Cornell light (main.go:293: 130 x 105 quad, radiance 15, facing down) over a Lambertian
    floor, nothing else, black background.  Radiance leaving floor point x is rho * Le * F(x) with F the closed-form
    configuration factor from a surface element to a parallel rectangle; F is also integrated numerically here so a
    wrong formula cannot pass.
"""
import numpy as np
import pytest
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU

E = 0.5
BIG = 1e30


def _room(sc, light_mat, lo=-10.0, hi=10.0):
    """Six quads with u x v pointing INTO the room (quad normal = unit(u x v), objects.go:129-141)."""
    s = hi - lo
    walls = [
        sc.NewQuad((lo, lo, lo), (0, 0, s), (s, 0, 0), light_mat),    # floor   y = lo, normal +y
        sc.NewQuad((lo, hi, lo), (s, 0, 0), (0, 0, s), light_mat),    # ceiling y = hi, normal -y
        sc.NewQuad((lo, lo, lo), (0, s, 0), (0, 0, s), light_mat),    # x = lo, normal +x
        sc.NewQuad((hi, lo, lo), (0, 0, s), (0, s, 0), light_mat),    # x = hi, normal -x
        sc.NewQuad((lo, lo, lo), (s, 0, 0), (0, s, 0), light_mat),    # z = lo, normal +z
        sc.NewQuad((lo, lo, hi), (0, s, 0), (s, 0, 0), light_mat),    # z = hi, normal -z
    ]
    return walls


def _camera(width, spp, frm, at, fov=60.0, depth=50):
    cam = g.Camera()
    cam.AspectRatio, cam.Width, cam.SamplesPerPixel, cam.MaxDepth = 1.0, width, spp, depth
    cam.VerticalFOV, cam.Background, cam.MaxContribution = fov, (0, 0, 0), BIG
    cam.PositionCamera(frm, at, (0, 1, 0))
    return cam.config()


def white_furnace(as_bvh):
    sc = g.Scene()
    walls = _room(sc, sc.NewDiffuseLight((E, E, E)))
    objs = list(walls)
    objs.append(sc.NewSphere((-4, -2, 2), 2.5, sc.NewLambertian((1, 1, 1))))
    objs.append(sc.NewSphere((3, -1, 0), 2.0, sc.NewDielectric(1.5)))
    objs.append(sc.NewSphere((0, 4, 3), 2.0, sc.NewMetal((1, 1, 1), 0.3)))
    box = sc.Translate(sc.RotateY(sc.NewBox((0, 0, 0), (3, 4, 3), sc.NewLambertian((1, 1, 1))), 25.0), (-1.5, -7, -4))
    objs.append(sc.ConstantMedium(box, 0.25, (1, 1, 1)))
    lst = sc.NewHittableList(objs)
    sc.set_world(sc.BuildBVH(lst) if as_bvh else lst)
    sc.set_lights(sc.NewHittableList([walls[1]]))            # only the ceiling is importance-sampled; all walls emit
    return sc, _camera(48, 256, (0, 0, -9.5), (0, 0, 0))


def grey_furnace(kind):
    rho = (0.2, 0.5, 0.8)
    sc = g.Scene()
    walls = _room(sc, sc.NewDiffuseLight((E, E, E)))
    if kind == "sphere":
        obj = sc.NewSphere((0, 0, 0), 3.0, sc.NewLambertian(rho))
    else:
        obj = sc.Translate(sc.RotateY(sc.NewBox((-2, -2, -2), (2, 3, 2), sc.NewLambertian(rho)), 33.0), (0.5, -0.5, 0))
    sc.set_world(sc.NewHittableList(walls + [obj]))
    sc.set_lights(sc.NewHittableList([walls[1], walls[3]]))   # two lights: HittableList.PdfValue / Random (hittable.go:89-103)
    return sc, _camera(48, 256, (0, 1, -9.5), (0, 0, 0)), np.array(rho)


def _object_mask(trace, cfg, cam):
    """True where the ray through the pixel (stratum 0 of 1) hits something nearer than the walls."""
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    h = trace(rays)
    p = h["p"].astype(np.float64)
    on_wall = (np.abs(np.abs(p).max(axis=1) - 10.0) < 1e-3)
    m = (~on_wall).reshape(cam.height, cam.width)
    return m


def _erode(m, k=1):
    out = m.copy()
    for dy in range(-k, k + 1):
        for dx in range(-k, k + 1):
            out &= np.roll(np.roll(m, dy, axis=0), dx, axis=1)
    return out


def _render_oracle(sc, cfg, seed):
    cam = O.derived_camera(cfg)
    sums, _, _, _ = O.OracleWorld(sc).render(cfg, seed=seed)
    return sums / cam.spp_sqrt ** 2


def _render_gpu(variant):
    def f(sc, cfg, seed):
        cam = g.derive_camera(cfg)
        sums, _, _ = g.DeviceScene(sc).render(cam, seed=seed, variant=variant)
        return sums.astype(np.float64) / cam.spp_sqrt ** 2
    return f


def _check_white_furnace(render, as_bvh):
    sc, cfg = white_furnace(as_bvh)
    img = render(sc, cfg, 11)
    assert np.isfinite(img).all()
    # every pixel's expectation is E: the image mean over 48*48*256 samples is far inside 0.5 %
    assert abs(img.mean() / E - 1.0) < 5e-3, img.mean() / E
    # and no region is off: 8x8-pixel block means (16384 samples each) within 4 %
    b = img.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3))
    assert np.abs(b / E - 1.0).max() < 0.04, np.abs(b / E - 1.0).max()


def _check_grey_furnace(render, trace_of, kind):
    sc, cfg, rho = grey_furnace(kind)
    cam = O.derived_camera(cfg)
    img = render(sc, cfg, 12)
    m = _object_mask(trace_of(sc), cfg, cam)
    inner, outer = _erode(m, 1), _erode(~m, 1)
    assert inner.sum() > 200 and outer.sum() > 200
    got = img[inner].mean(axis=0)
    assert np.abs(got / (rho * E) - 1.0).max() < 0.01, (got, rho * E)      # object radiance = rho * E, per channel
    assert np.abs(img[outer] / E - 1.0).max() < 1e-5                        # a wall seen directly: exactly E


# ---- direct irradiance under the Cornell light ---------------------------------------------------------------
LQ, LU, LV, LE = np.array([343.0, 550.0, 332.0]), np.array([-130.0, 0, 0]), np.array([0, 0, -105.0]), 15.0
RHO = 0.73


def _F_corner(a, b, h):
    """Configuration factor from a surface element to a parallel rectangle [0,a] x [0,b] at height h, one corner on
    the element's normal (Siegel & Howell, configuration B-4); odd in a and in b."""
    A, B = np.sqrt(a * a + h * h), np.sqrt(b * b + h * h)
    return (a / A * np.arctan(b / A) + b / B * np.arctan(a / B)) / (2 * np.pi)


def form_factor(px, pz):
    x1, x2 = LQ[0] + LU[0] - px, LQ[0] - px
    z1, z2 = LQ[2] + LV[2] - pz, LQ[2] - pz
    h = LQ[1]
    return _F_corner(x2, z2, h) - _F_corner(x1, z2, h) - _F_corner(x2, z1, h) + _F_corner(x1, z1, h)


def form_factor_numeric(px, pz, n=400):
    u = (np.arange(n) + 0.5) / n
    X = LQ[0] + LU[0] * u[:, None] - px
    Z = LQ[2] + LV[2] * u[None, :] - pz
    h = LQ[1]
    r2 = X * X + Z * Z + h * h
    return float((h * h / (np.pi * r2 * r2)).sum() * (130.0 * 105.0) / (n * n))


def direct_light_scene():
    sc = g.Scene()
    floor = sc.NewQuad((0, 0, 0), (555, 0, 0), (0, 0, 555), sc.NewLambertian((RHO, RHO, RHO)))
    light = sc.NewQuad(tuple(LQ), tuple(LU), tuple(LV), sc.NewDiffuseLight((LE, LE, LE)))
    sc.set_world(sc.NewHittableList([floor, light]))
    sc.set_lights(sc.NewHittableList([light]))
    cfg = _camera(60, 4096, (278, 420, -250), (278, 0, 278), fov=60.0)
    return sc, cfg


def test_configuration_factor_formula_against_quadrature():
    for px, pz in [(278.0, 279.5), (0.0, 0.0), (500.0, 100.0), (343.0, 332.0), (150.0, 450.0)]:
        assert abs(form_factor(px, pz) / form_factor_numeric(px, pz) - 1.0) < 1e-4


def _check_direct_irradiance(render, trace_of):
    sc, cfg = direct_light_scene()
    cam = O.derived_camera(cfg)
    img = render(sc, cfg, 13)
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    h = trace_of(sc)(rays)
    p = h["p"].astype(np.float64)
    on_floor = (np.abs(p[:, 1]) < 1e-3) & np.isfinite(h["t"]) & (h["t"] < 1e30)      # the floor is y = 0 (fp32 hit points on the GPU side)
    expect = (RHO * LE * form_factor(p[:, 0], p[:, 2])).reshape(cam.height, cam.width)
    m = _erode(on_floor.reshape(cam.height, cam.width), 1)
    assert m.sum() > 500
    got = img[..., 0]
    assert np.allclose(img[..., 0], img[..., 1]) and np.allclose(img[..., 0], img[..., 2])
    # image-level: total floor radiance within 0.5 %; block-level (6x6 pixels x 4096 spp, fully on the floor; sigma ~ 0.8 %): within 4 %
    assert abs(got[m].sum() / expect[m].sum() - 1.0) < 5e-3, got[m].sum() / expect[m].sum()
    mb = m.reshape(10, 6, 10, 6).all(axis=(1, 3))
    gb, eb = got.reshape(10, 6, 10, 6).mean(axis=(1, 3)), expect.reshape(10, 6, 10, 6).mean(axis=(1, 3))
    assert mb.sum() >= 10
    assert np.abs(gb[mb] / eb[mb] - 1.0).max() < 0.04, np.abs(gb[mb] / eb[mb] - 1.0).max()


# ---- the oracle ------------------------------------------------------------------------------------------------
def _oracle_trace(sc):
    ow = O.OracleWorld(sc)
    return lambda rays: ow.trace_batch(rays)


@pytest.mark.parametrize("as_bvh", [False, True])
def test_oracle_white_furnace(as_bvh):
    _check_white_furnace(_render_oracle, as_bvh)


@pytest.mark.parametrize("kind", ["sphere", "box"])
def test_oracle_grey_furnace(kind):
    _check_grey_furnace(_render_oracle, _oracle_trace, kind)


def test_oracle_direct_irradiance_under_the_cornell_light():
    _check_direct_irradiance(_render_oracle, _oracle_trace)


# ---- the CUDA path ----------------------------------------------------------------------------------------------
def _gpu_trace(sc):
    dev = g.DeviceScene(sc)
    return lambda rays: dev.trace_batch(rays)


VARIANTS = [g.GRT_VARIANT_MEGAKERNEL, g.GRT_VARIANT_WAVEFRONT]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("as_bvh", [False, True])
def test_cuda_white_furnace(as_bvh, variant):
    _check_white_furnace(_render_gpu(variant), as_bvh)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kind", ["sphere", "box"])
def test_cuda_grey_furnace(kind, variant):
    _check_grey_furnace(_render_gpu(variant), _gpu_trace, kind)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_cuda_direct_irradiance_under_the_cornell_light(variant):
    _check_direct_irradiance(_render_gpu(variant), _gpu_trace)
