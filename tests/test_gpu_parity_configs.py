"""Parity of the CUDA path on EVERY BASELINE.json config class, not only the Cornell box (round-1 verdict, items 4):

  * image parity proper (north star check 3) for the smoke, book-1, book-2 and mesh scenes: INDEPENDENT random streams on
    the two sides, the oracle with the reference's semantics (no self-exclusion, fp64), per-pixel |diff| <= 3 sigma of
    the Monte Carlo error for the stated fraction of pixel-channels and mean RGB error <= 1e-3, at the config's own
    per-pixel sample count and a reduced resolution;
  * same-seed renders with the reference's semantics on the oracle side (use_exclusion=False): how many pixel-channels
    the fp32 kernel follows sample by sample, per scene, with a measured bar each;
  * the SURVEY 8d ray batches (>= 2^20 primary rays + as many secondary rays) for book 2 and for the FULL 708-segment,
    1.0 M-triangle mesh of config C5.
"""
import os
import numpy as np
import pytest
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
T_REL = 1e-5


def _earth():
    return np.load(os.path.join(HERE, "golden", "earthmap_rgb8.npz"))["rgb"]


def _scene(sid, **kw):
    if sid in (2, 5):
        kw["image"] = _earth()
    return g.builtin_scene(sid, **kw)


# scene id -> (width, spp of its BASELINE.json config, builtin_scene kwargs)
# The bar is self-calibrating.  "Within 3 sigma for 99.7 %" holds for Gaussian errors with a KNOWN sigma; these
# estimators are heavy-tailed (light paths through glass, metal and media, clamped at MaxContribution) and sigma is
# itself estimated from the samples, so two runs of the ORACLE with different seeds agree on only 99.1-99.6 % of
# pixel-channels of the book scenes and the mesh (99.7-99.9 % on the smoke box).  The test therefore renders the oracle
# twice and requires the CUDA image to be as close to the oracle as the oracle is to itself (less 0.4 % of noise on
# the fraction), and never below 98.5 %.
STAT_CASES = {
    7: (32, 4096, {}),                              # C3 Cornell smoke, 4096 spp
    1: (64, 100, {}),                               # C1 book-1 cover, 100 spp
    2: (48, 1024, {"aspect": 16 / 9}),              # C4 book-2 cover, 1024 spp
    8: (40, 1024, {"mesh_segments": 128}),          # C5 mesh scene, 1024 spp (32 k triangles: the oracle renders it in seconds)
}
_ORACLE_CACHE = {}


def _oracle_pair(sid):
    """(scene, cfg, oracle mean image, per-pixel sigma of a difference of two such images, oracle-vs-oracle fraction)."""
    if sid not in _ORACLE_CACHE:
        width, spp, kw = STAT_CASES[sid]
        s, cfg = _scene(sid, width=width, spp=spp, **kw)
        cam = g.derive_camera(cfg)
        S2 = cam.spp_sqrt ** 2
        ow = O.OracleWorld(s)
        os_, osq, _, _ = ow.render(cfg, seed=0xC0FFEE, want_sumsq=True, use_exclusion=False)
        other, _, _, _ = ow.render(cfg, seed=0xFACADE, use_exclusion=False)
        om = os_ / S2
        var = np.maximum(osq / S2 - om ** 2, 1e-12)
        sigma = np.sqrt(2 * var / S2)                 # both estimators have (about) this variance
        self_frac = (np.abs(other / S2 - om) <= 3 * sigma + 1e-6).mean()
        _ORACLE_CACHE[sid] = (s, cfg, om, var, sigma, self_frac)
    return _ORACLE_CACHE[sid]


@pytest.mark.parametrize("variant", ["mega", "wavefront"])
@pytest.mark.parametrize("sid", sorted(STAT_CASES))
def test_statistical_image_parity_per_config(sid, variant):
    s, cfg, om, var, sigma, self_frac = _oracle_pair(sid)
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    v = g.GRT_VARIANT_MEGAKERNEL if variant == "mega" else g.GRT_VARIANT_WAVEFRONT
    gs, _, _ = g.DeviceScene(s).render(cam, seed=424242, variant=v)
    gm = gs.astype(np.float64) / S2
    fin = np.isfinite(gm) & np.isfinite(om)
    assert fin.mean() > 0.999
    within = (np.abs(gm - om)[fin] <= 3 * sigma[fin] + 1e-6).mean()
    need = max(0.985, min(0.997, self_frac - 0.004))
    print(f"scene {sid} {variant}: {within:.4f} of pixel-channels within 3 sigma of the oracle (oracle vs oracle, other seed: {self_frac:.4f})")
    assert within >= need, f"scene {sid} {variant}: only {within:.4f} of pixel-channels within 3 sigma (oracle vs itself: {self_frac:.4f}, need {need:.4f})"
    # mean RGB error: the north star's 1e-3 is quoted at 4096 spp; allow the Monte Carlo error of THIS image's mean on top
    sigma_mean = np.sqrt((2 * var[fin] / S2).sum()) / fin.sum()
    err = abs(gm[fin].mean() - om[fin].mean())
    assert err <= 1e-3 + 3 * sigma_mean, f"scene {sid} {variant}: mean RGB error {err:.2e} (sigma of the mean {sigma_mean:.1e})"


# scene id -> (width, spp, kwargs, bar).  Measured on B200 (round 2): the fp32 kernel follows the fp64 reference-semantics
# oracle sample by sample (per-pixel means within 1e-4) on 1.0000 of the pixel-channels of scenes 3-7, 0.9967 of book 1,
# 0.9942 of book 2 and 0.955 of the mesh scene; the bars sit a little below.  What makes a sample leave the oracle's path:
#   1, 2, 8  metal fuzz and refraction are chaotic: an fp32 direction error grows with the curvature at every bounce, and
#            on the mesh a hit within fp32 error of a shared edge lands on the neighbouring triangle (other vertex normals);
#   7, 2     a medium's free-flight distance -ln(u)/rho against a boundary distance can flip within 1e-6 relative.
# The oracle's self-exclusion extension (use_exclusion=True, what the other same-seed tests use) moved NO pixel of
# these renders: the measured fractions with and without it are identical.
SAME_SEED = {6: (48, 64, {}, 0.995), 3: (40, 36, {}, 0.995), 4: (48, 36, {}, 0.995), 5: (40, 36, {}, 0.995),
             7: (48, 64, {}, 0.99), 1: (48, 36, {}, 0.985), 2: (40, 16, {}, 0.98), 8: (40, 16, {"mesh_segments": 64}, 0.92)}


@pytest.mark.parametrize("sid", sorted(SAME_SEED))
def test_same_seed_render_against_reference_semantics(sid):
    """The oracle WITHOUT its self-exclusion extension (tmin = 0.001 only, camera.go:300): exactly the reference's
    semantics on the same random streams.  Reports how much of the image the fp32 kernel reproduces sample by sample."""
    w, spp, kw, bar = SAME_SEED[sid]
    s, cfg = _scene(sid, width=w, spp=spp, **kw)
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    gs, _, _ = g.DeviceScene(s).render(cam)
    ref, _, _, _ = O.OracleWorld(s).render(cfg, use_exclusion=False)
    ext, _, _, _ = O.OracleWorld(s).render(cfg, use_exclusion=True)
    gm, rm, em = gs.astype(np.float64) / S2, ref / S2, ext / S2
    fin = np.isfinite(gm) & np.isfinite(rm)
    same_ref = (np.abs(gm - rm)[fin] < 1e-4).mean()
    same_ext = (np.abs(gm - em)[fin] < 1e-4).mean()
    excl_effect = (np.abs(rm - em)[fin] >= 1e-4).mean()        # what the oracle-side extension changes at all
    print(f"scene {sid}: follows the reference-semantics oracle on {same_ref:.4f} of pixel-channels "
          f"(with the self-exclusion extension {same_ext:.4f}; the extension itself moves {excl_effect:.4f})")
    assert same_ref > bar, f"scene {sid}: only {same_ref:.3f} of pixel-channels follow the reference-semantics oracle (bar {bar})"
    assert abs(gm[fin].mean() - rm[fin].mean()) <= 1e-3 * max(1.0, rm[fin].mean())


def _check(r, name, max_flag_frac=0.08):
    assert r["id_mismatch_unflagged"] == 0, f"{name}: {r['id_mismatch_unflagged']} hit-id mismatches outside documented ties, e.g. rays {r['bad_id_idx']}"
    assert r["t_bad"] == 0, f"{name}: t off by up to {r['t_max_rel_unflagged']:.2e} relative at rays {r['bad_t_idx']}"
    assert r["flagged"] <= max_flag_frac * r["n"], f"{name}: too many rays excluded as ties/edges ({r['flagged']}/{r['n']})"
    assert r["hits"] > 0.05 * r["n"]


@pytest.mark.parametrize("sid,kw", [(2, {"aspect": 1.0}), (8, {"aspect": 1.0})])
def test_million_ray_batches_book2_and_full_mesh(sid, kw):
    """SURVEY 8d's fixed ray batch on the two BVH configs: the S = 1 primary rays of a 1024 x 1024 view (2^20 rays) and
    as many secondary rays from the oracle's first hits.  Scene 8 is the FULL config-C5 mesh: 708 x 708 segments,
    1 002 528 triangles through the OBJ loader and BuildBVH (objects.go:408-461, bvh.go:35-61)."""
    s, cfg = _scene(sid, width=1024, spp=1, **kw)
    flat = s.flatten()
    if sid == 8:
        assert flat.n_tris > 1_000_000
    ow, dev = O.OracleWorld(s), g.DeviceScene(s)
    cam = O.derived_camera(cfg)
    prim = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    assert len(prim) >= 1 << 20
    oh = ow.trace_batch(prim, audit_eps=1e-5)
    _check(PU.compare_hits(dev.trace_batch(prim), oh, t_rel=T_REL), f"scene {sid} 1M primary")
    sec = PU.secondary_batch(oh, np.random.default_rng(200 + sid), time=prim["time"])
    plain = sec.copy(); plain["self_id"] = PU.NO_ID
    _check(PU.compare_hits(dev.trace_batch(plain), ow.trace_batch(plain, audit_eps=1e-5), t_rel=T_REL), f"scene {sid} 1M secondary")
    _check(PU.compare_hits(dev.trace_batch(sec), ow.trace_batch(sec, audit_eps=1e-5, use_exclusion=True), t_rel=T_REL),
           f"scene {sid} 1M secondary+self")


@pytest.mark.parametrize("sid,kw", [(1, {}), (2, {"aspect": 16 / 9}), (8, {"mesh_segments": 96})])
def test_traversal_scheduling_does_not_change_the_image(sid, kw, monkeypatch):
    """The wavefront extend step is scheduled per scene (persistent warps or one thread per slot, node-step quorum,
    slice exit, treelet, tree build: grt_wavefront.cu / wide_bvh.hpp).  None of that may change WHAT a ray hits: every
    combination must give the same image up to the order of the float additions into a pixel."""
    s, cfg = _scene(sid, width=96, spp=16, **kw)
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2

    def render(env):
        for k in ("GRT_WF_DYN", "GRT_WF_VOTE16", "GRT_WF_EXIT16", "GRT_WF_TREELET", "GRT_WIDE_SAH", "GRT_WF_SLOTS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sums, _, _ = g.DeviceScene(s).render(cam, variant=g.GRT_VARIANT_WAVEFRONT)     # (the tree is built at upload: a fresh handle per setting)
        return sums.astype(np.float64) / S2

    base = render({})
    assert np.isfinite(base).mean() > 0.999
    for env in ({"GRT_WF_DYN": "0"}, {"GRT_WF_DYN": "2"}, {"GRT_WF_DYN": "2", "GRT_WF_VOTE16": "8", "GRT_WF_EXIT16": "12"},
                {"GRT_WF_DYN": "2", "GRT_WF_VOTE16": "1", "GRT_WF_EXIT16": "0"}, {"GRT_WF_DYN": "2", "GRT_WF_TREELET": "16"},
                {"GRT_WF_SLOTS": "4096"}):
        img = render(env)
        fin = np.isfinite(base) & np.isfinite(img)
        assert (np.isfinite(base) == np.isfinite(img)).all(), env
        assert np.allclose(img[fin], base[fin], rtol=1e-4, atol=1e-5), (env, float(np.abs(img - base)[fin].max()))
    # the two tree builds may break EXACT ties between equidistant primitives differently: a handful of samples
    plain = render({"GRT_WIDE_SAH": "0"})
    fin = np.isfinite(base) & np.isfinite(plain)
    same = np.isclose(plain[fin], base[fin], rtol=1e-4, atol=1e-5).mean()
    assert same > 0.995, same
    assert abs(plain[fin].mean() - base[fin].mean()) <= 2e-3 * max(1.0, base[fin].mean())
