"""Developer script: quick GPU-vs-oracle checks (run under gpurun).  Not a test."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU


def trace_check(scene_id, W=128, **kw):
    s, cfg = g.builtin_scene(scene_id, width=W, spp=4, **kw)
    ow = O.OracleWorld(s)
    dev = g.DeviceScene(s)
    cam = O.derived_camera(cfg)
    rays = PU.primary_batch(cfg, (0, 0, cam.width, cam.height))
    oh = ow.trace_batch(rays, audit_eps=1e-5)
    gh = dev.trace_batch(rays)
    r1 = PU.compare_hits(gh, oh)
    rng = np.random.default_rng(1)
    sec = PU.secondary_batch(oh, rng, time=rays["time"])
    sec_noself = sec.copy(); sec_noself["self_id"] = PU.NO_ID
    oh2 = ow.trace_batch(sec_noself, audit_eps=1e-5)
    gh2 = dev.trace_batch(sec_noself)
    r2 = PU.compare_hits(gh2, oh2)
    oh3 = ow.trace_batch(sec, audit_eps=1e-5, use_exclusion=True)
    gh3 = dev.trace_batch(sec)
    r3 = PU.compare_hits(gh3, oh3)
    for name, r in (("primary", r1), ("secondary", r2), ("secondary+self", r3)):
        print(f"scene {scene_id} {name}: n={r['n']} hits={r['hits']} flagged={r['flagged']} id_mis={r['id_mismatch_total']} "
              f"id_bad={r['id_mismatch_unflagged']} t_max_rel={r['t_max_rel_unflagged']:.2e} t_bad={r['t_bad']}")
        if r['id_mismatch_unflagged'] or r['t_bad']:
            print("   bad idx", r['bad_id_idx'], r['bad_t_idx'])
    return dev, ow, cfg, s


def render_check(scene_id, W=32, spp=64, **kw):
    s, cfg = g.builtin_scene(scene_id, width=W, spp=spp, **kw)
    ow = O.OracleWorld(s)
    dev = g.DeviceScene(s)
    cam = g.derive_camera(cfg)
    t0 = time.time()
    gs, _, st = dev.render(cam, want_stats=True)
    t1 = time.time()
    os_, osq, ost, sec = ow.render(cfg, use_exclusion=True, want_sumsq=True, want_stats=True)
    S2 = cam.spp_sqrt ** 2
    gm = gs.astype(np.float64) / S2
    om = os_ / S2
    diff = np.abs(gm - om)
    print(f"scene {scene_id} render {cam.width}x{cam.height}x{S2}: gpu mean {np.nanmean(gm):.5f} oracle mean {np.nanmean(om):.5f} "
          f"mean|d| {np.nanmean(diff):.2e} max|d| {np.nanmax(diff):.3f} exact-ish px {(diff < 1e-4).mean():.3f} gpu {t1-t0:.2f}s oracle {sec:.2f}s")
    print("   gpu stats", st)
    print("   orc stats", ost)
    return gm, om


if __name__ == "__main__":
    which = sys.argv[1:] or ["6"]
    for w in which:
        sid = int(w)
        kw = {"mesh_segments": 64} if sid == 8 else {}
        try:
            trace_check(sid, **kw)
            render_check(sid, **kw)
        except Exception as e:
            import traceback; traceback.print_exc()
