#!/usr/bin/env python
"""bench.py — Mpaths/s of the render hot path on BASELINE.json's configs (headline: C2, Cornell box 1024^2 x 4096 spp).

  python bench.py --gpus N --steps K --warmup W              our arm (CUDA backend through the C ABI), config C2
  python bench.py --config C1|C2|C3|C4|C5 ...                another config as the headline line
  python bench.py --impl reference --gpus N --steps K ...    the reference's CPU algorithm (oracle port) on host cores
  python bench.py --inproc --gpus N                          one process, N GPUs through grt_render_multi (the cgo path)

One "step" = one complete render of the workload (one pass of the hot path over all W*H*spp paths).
Multi-GPU (torchrun, one rank per GPU): the strata set is split s = rank (mod N) (strong scaling: the
workload is fixed), every rank renders its shard into a private fp32 sum buffer and ONE NCCL reduce to rank 0
combines them (camera.go has no analogue).

The default (N = 1) line also carries `configs`: one short measurement per BASELINE.json config (full resolution, the
spp named in its `workload`) with its own roofline fractions (FP32 issue and L2/HBM bytes) and CPU baseline.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Frozen algorithmic-cost table: flops and bytes per EVENT, counted from
# the cited reference lines / the flat fp32 layout.  Events are counted by the oracle (the reference algorithm).
COSTS = {
    "box_tests": (21, 32),        # aabb.go:94-110, one 32-byte node
    "quad_tests": (30, 64),       # objects.go:168-194: 12 (plane reject) .. 54 (hit); 30 = documented average
    "sphere_tests": (40, 64),     # objects.go:84-113: 29 (reject) .. 55 (hit)
    "tri_tests": (35, 48),        # objects.go:409-456: 20 .. 63
    "medium_tests": (21, 16),     # medium.go:39-51 (the two boundary traversals are counted as box/quad/sphere tests)
    "shade_diffuse": (125, 32),   # onb.go:13-25 + vec.go:177-186 + pdf.go:33-74 + camera.go:325-330
    "shade_specular": (50, 32),   # materials.go:70-130
    "light_pdf_evals": (80, 0),   # objects.go:152-160 re-intersection + pdf
    "paths": (23, 0),             # camera.go:257-269
}

# BASELINE.json configs.  `spp` is the config's sample count, `side_spp` the reduced one used for the
# short per-config lines of the default run and `cpu` the bounded CPU sample (width, spp) of the same scene.
CONFIGS = {
    "C1": dict(scene=6 - 5, what="book-1 cover (main.go:19-91), MaxDepth 50", width=400, aspect=0.0, spp=100, side_spp=100, cpu=(400, 100)),
    "C2": dict(scene=6, what="cornellBox (main.go:278-320), MaxDepth 50", width=1024, aspect=0.0, spp=4096, side_spp=256, cpu=(1024, 256)),
    "C3": dict(scene=7, what="cornellSmoke (main.go:323-367), MaxDepth 50", width=1024, aspect=0.0, spp=4096, side_spp=256, cpu=(1024, 16)),
    "C4": dict(scene=2, what="book-2 cover (main.go:94-174) at 16:9, MaxDepth 40", width=1920, aspect=16 / 9, spp=1024, side_spp=64, cpu=(1920, 4)),
    "C5": dict(scene=8, what="1.0 M-triangle displaced sphere through objLoader + BuildBVH, modelExample's camera and materials (main.go:371-409), MaxDepth 50",
               width=3840, aspect=0.0, spp=1024, side_spp=16, cpu=(1920, 4)),
}

# Event counts per path of the REFERENCE algorithm on C2 (the oracle's counters: reference BVH, every continuation
# traced).  Used when the CPU leg is skipped; the CPU leg re-measures them.
ORACLE_EVENTS_C2 = {"paths": 1.0, "box_tests": 27.06, "quad_tests": 17.55, "sphere_tests": 0.0, "tri_tests": 0.0,
                    "medium_tests": 0.0, "shade_diffuse": 1.925, "shade_specular": 0.0, "light_pdf_evals": 1.925}


def algorithmic_cost(events_per_path):
    fl = sum(events_per_path.get(k, 0.0) * c[0] for k, c in COSTS.items())
    by = sum(events_per_path.get(k, 0.0) * c[1] for k, c in COSTS.items())
    return fl, by


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _earth():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "earthmap_rgb8.npz"))["rgb"]


def build_config(name, width=None, spp=None):
    """(Scene, GrtCameraConfig) of a BASELINE.json config at the given (or the config's own) width / spp."""
    import go_raytracer_b200 as g
    c = CONFIGS[name]
    kw = {}
    if c["scene"] in (2, 5):
        kw["image"] = _earth()          # earthmap.jpg as Go decodes it (tests/golden/make_earthmap_fixture.py)
    return g.builtin_scene(c["scene"], width=width or c["width"], spp=spp or c["spp"], aspect=c["aspect"], **kw)


def workload(name, cam_width, cam_height, spp_used, full_spp):
    c = CONFIGS[name]
    return {"workload": f"{name}: {c['what']}, {cam_width}x{cam_height} x {spp_used} spp", "name": name, "scene": c["scene"],
            "width": cam_width, "height": cam_height, "spp": spp_used, "config_spp": full_spp,
            "paths_per_step": cam_width * cam_height * spp_used}


def cpu_sample(name, cores, width=None, spp=None, steps=1, warm=False):
    """The oracle port (the reference's algorithm, fp64, one task per image row like threadedRenderer, camera.go:111-132)
    on a bounded sample of config `name`.  Returns (Mpaths/s, seconds per step, events per path, description)."""
    from oracle import oracle_py as O
    w, s_ = CONFIGS[name]["cpu"]
    s, cfg = build_config(name, width or w, spp or s_)
    ow = O.OracleWorld(s)
    cam = O.derived_camera(cfg)
    paths = cam.width * cam.height * cam.spp_sqrt ** 2
    if warm:
        ow.render(cfg, nthreads=cores, window=(0, 0, cam.width, max(1, cam.height // 8)))
    t, ost = 0.0, None
    for _ in range(steps):
        _, _, ost, sec = ow.render(cfg, nthreads=cores, want_stats=True)
        t += sec
    ev = {k: ost[k] / max(1, ost["paths"]) for k in COSTS}
    what = (f"{name} at {cam.width}x{cam.height} x {cam.spp_sqrt ** 2} spp ({paths / 1e6:.1f} M paths, {t / steps:.1f} s per step), {cores} threads, one task per row; "
            "C++ fp64 restatement of the Go renderer (no Go toolchain here): no per-Vec3 heap allocation, so faster than the Go binary would be")
    return paths * steps / t / 1e6, t / steps, ev, what, (cam.width, cam.height, cam.spp_sqrt ** 2, paths)


def run_reference(args):
    """The reference's own CPU algorithm (fp64 oracle port; the Go binary cannot be built here: no Go
    toolchain) on all host cores.  Each step is a BOUNDED sample of the workload: full resolution,
    reduced spp (Mpaths/s is spp-independent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    c = CONFIGS[args.config]
    width = args.width or c["width"]
    spp = args.ref_spp or (64 if args.config in ("C2", "C3") else c["cpu"][1])
    val, sec, _, what, (w, h, s2, paths) = cpu_sample(args.config, cores, width=width if args.config != "C5" else c["cpu"][0], spp=spp, steps=args.steps, warm=args.warmup > 0)
    full = build_config(args.config, width, args.spp or c["spp"])[1]
    import go_raytracer_b200 as g
    fc = g.derive_camera(full)
    cfg = workload(args.config, fc.width, fc.height, fc.spp_sqrt ** 2, c["spp"])
    # what each step of THIS arm really rendered (a bounded sample of the workload above)
    cfg["sample"] = {"width": w, "height": h, "spp": s2, "paths_per_step": paths}
    line = {"impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec, "ms_per_step_is": "one bounded sample (config.sample), not the full workload",
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def scene_bytes(flat):
    return (flat.n_nodes * 32 + flat.n_spheres * 64 + flat.n_quads * 96 + flat.n_boxes * 64 + flat.n_tris * (48 + 64 + 72) + flat.n_items * 4 +
            flat.n_media * 16 + flat.n_materials * 32 + flat.n_textures * 32 + flat.n_lights * 304 + int(flat.n_texel_bytes))


class stdout_to_stderr:
    """File-descriptor level: whatever native libraries print to stdout inside (NCCL's version banner when NCCL_DEBUG
    is set in the environment) goes to stderr, so that stdout stays the ONE JSON line of the contract."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def last_timing(L):
    from go_raytracer_b200 import _native as N
    t = N.GrtTiming()
    N.check(L.grt_last_timing(C.byref(t)))
    return t


def measure_config(name, spp, steps, warmup, local, seed, cores, want_cpu, peaks, sm_mhz, props):
    """One short line for the `configs` object: device-resident throughput of config `name` at `spp`, the kernel class
    that dominates it (CUDA events per class, GRT_OPT_TIMING) and both roofline fractions."""
    import torch
    import go_raytracer_b200 as g
    from go_raytracer_b200 import _native as N
    L = g.lib()
    s, cfg = build_config(name, None, spp)
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    paths = cam.width * cam.height * S2
    t0 = time.perf_counter()
    scene = g.DeviceScene(s, local)
    upload_s = time.perf_counter() - t0
    flat = s.flatten()
    variant = "wavefront" if flat.n_nodes > 0 else "megakernel"      # what GRT_VARIANT_AUTO picks (grt_render_device)
    dev = torch.device("cuda", local)
    acc = torch.zeros(cam.width * cam.height * 3, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        acc.zero_()
        scene.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=seed, variant=N.GRT_VARIANT_AUTO)
    torch.cuda.synchronize()
    l0 = L.grt_launch_count()
    ms = 0.0
    for _ in range(steps):
        flush.fill_(1)
        acc.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        scene.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=seed, variant=N.GRT_VARIANT_AUTO)
        e1.record(stream)
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    launches = L.grt_launch_count() - l0
    ms /= steps
    value = paths / (ms / 1e3) / 1e6
    finite = bool(torch.isfinite(acc).all().item())
    # per-kernel-class device time of one more (untimed) render: which kernel dominates, and by how much
    acc.zero_()
    scene.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=seed, variant=N.GRT_VARIANT_AUTO, flags=N.GRT_OPT_TIMING)
    torch.cuda.synchronize()
    T = last_timing(L)
    share = T.extend_ms / T.total_ms if T.total_ms > 0 else None
    dominant = ("render_mega_kernel", "wf_extend", "wf_extend_dyn")[int(T.extend_kernel)]    # the library picks per scene (grt_wavefront.cu)
    scene.close()
    cpu, ev, ev_src = None, None, None
    if want_cpu:
        v, sec, ev, what, _ = cpu_sample(name, cores)
        cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": what}
        ev_src = "oracle counters of this config's cpu_baseline run"
    line = {"workload": workload(name, cam.width, cam.height, S2, CONFIGS[name]["spp"])["workload"], "value": value, "unit": "Mpaths/s",
            "ms_per_step": ms, "steps": steps, "warmup": warmup, "variant": variant, "gpu_launches_per_step": launches / steps,
            "all_pixels_finite": finite, "upload_s": upload_s,
            "dominant_kernel": {"name": dominant, "ms_per_step": T.extend_ms, "launches_per_step": int(T.extend_launches),
                                "share_of_step": share, "timed": "CUDA events around every launch of the class on the launching stream (GRT_OPT_TIMING pass, plain launches)"},
            "cpu_baseline": cpu}
    if ev:
        line["roofline"] = roofline_pair(ev, ev_src, value * 1e6, S2, T.extend_ms / T.total_ms if T.total_ms else 1.0, peaks, sm_mhz, props, flat)
    return line


def roofline_pair(ev, ev_src, paths_per_s, S2, dom_share, peaks, sm_mhz, props, flat):
    """Both fractions SURVEY.md §8d asks for: FP32 issue (flops/path x paths/s over SMs x 128 x 2 x clock) and L2/HBM
    bytes (bytes/path x paths/s over the measured HBM copy bandwidth), for the dominant kernel: its algorithmic work is
    the traversal + intersection + (megakernel) shading events, and its time is its share of the step."""
    flops_pp, bytes_pp = algorithmic_cost(ev)
    bytes_pp += 12.0 / S2
    fp32_peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    share = dom_share if dom_share and dom_share > 0 else 1.0
    ach_tf = paths_per_s * flops_pp / 1e12 / share           # the dominant kernel does this work in `share` of the step
    ach_gbs = paths_per_s * bytes_pp / 1e9 / share
    f32, fby = ach_tf / fp32_peak, ach_gbs / hbm_peak
    resident = "shared memory" if flat.n_nodes == 0 else ("L1/L2" if flat.n_nodes < 100000 else "L2 (126 MB)")
    # list-only scenes (C2, C3) are staged whole in shared memory, their node/primitive bytes never
    # leave the SM and FP32 issue is the binding roofline; BVH scenes (C1, C4, C5) are judged on both, the larger binds
    bound = "fp32_issue" if (flat.n_nodes == 0 or f32 >= fby) else "l2_hbm_bytes"
    return {"events_per_path": ev, "events_source": ev_src, "flops_per_path": flops_pp, "bytes_per_path": bytes_pp,
            "fp32": {"achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": f32,
                     "peak_source": f"SMs({props.multi_processor_count}) x 128 lanes x 2 x {sm_mhz:.0f} MHz observed during the run"},
            "bytes": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fby,
                      "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                      "note": f"algorithmic node + primitive bytes over the measured HBM copy bandwidth; the scene is {resident}-resident, so these bytes "
                              "are served on chip (a fraction above 1 is possible and means exactly that)"},
            "bound": bound, "frac": f32 if bound == "fp32_issue" else fby, "dominant_kernel_share": share}


def run_inproc(args):
    """One process, N GPUs: grt_render_multi (what a cgo caller or `grt_main -gpus N` runs) — fused NVLink peer-atomic
    accumulation when the devices have native peer atomics, else one ncclReduce."""
    import numpy as np
    import go_raytracer_b200 as g
    from go_raytracer_b200 import _native as N
    L = g.lib()
    c = CONFIGS[args.config]
    s, cfg = build_config(args.config, args.width or c["width"], args.spp or c["spp"])
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    paths = cam.width * cam.height * S2
    flat = s.flatten()
    n = args.gpus
    devs = (C.c_int * n)(*range(n))
    opt = N.GrtOptions()
    opt.seed, opt.variant = args.seed, N.GRT_VARIANT_AUTO
    sums = np.zeros((cam.height, cam.width, 3), dtype=np.float32)
    rgb8 = np.zeros((cam.height, cam.width, 3), dtype=np.uint8)
    kms = C.c_double(0)
    out = {}
    for mode in (["1", "0"] if n > 1 else ["1"]):
        os.environ["GRT_MULTI_P2P"] = mode
        with stdout_to_stderr():
            for _ in range(max(1, args.warmup // 2)):
                sums[:] = 0
                N.check(L.grt_render_multi(C.byref(flat), C.byref(cam), C.byref(opt), devs, n, sums.ctypes.data, rgb8.ctypes.data, C.byref(kms)))
            t_dev, t_wall = 0.0, 0.0
            l0 = L.grt_launch_count()
            for _ in range(args.steps):
                sums[:] = 0
                t0 = time.perf_counter()
                N.check(L.grt_render_multi(C.byref(flat), C.byref(cam), C.byref(opt), devs, n, sums.ctypes.data, rgb8.ctypes.data, C.byref(kms)))
                t_wall += time.perf_counter() - t0
                t_dev += kms.value
        out["fused_peer_atomics" if mode == "1" else "nccl_reduce"] = {
            "value": paths * args.steps / (t_dev / 1e3) / 1e6, "device_ms_per_step": t_dev / args.steps,
            "e2e_value": paths * args.steps / t_wall / 1e6, "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "gpu_launches": int(L.grt_launch_count() - l0), "mean": float(sums.mean() / S2)}
    best = max(out.values(), key=lambda r: r["value"])
    line = {"metric": "Mpaths/s", "value": best["value"], "unit": "Mpaths/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": best["device_ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload(args.config, cam.width, cam.height, S2, c["spp"]), parallelism=f"in-process grt_render_multi x{n} (one host thread per device)"),
            "e2e": {"value": best["e2e_value"], "unit": "Mpaths/s", "h2d_bytes_per_step": n * scene_bytes(flat) + sums.nbytes, "d2h_bytes_per_step": sums.nbytes + rgb8.nbytes,
                    "note": "wall clock around grt_render_multi: scene upload to every device, render, reduce, tonemap, read-back"},
            "gpu_launches": best["gpu_launches"], "inproc": out}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import go_raytracer_b200 as g
    from go_raytracer_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries ONE JSON line: whatever libraries print on fd 1 (NCCL's version banner does, whatever
    # NCCL_DEBUG_FILE says) is sent to stderr; the JSON line goes out through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = g.lib()

    name = args.config
    c = CONFIGS[name]
    s, cfg = build_config(name, args.width or c["width"], args.spp or c["spp"])
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    paths = cam.width * cam.height * S2
    nval = cam.width * cam.height * 3
    scene = g.DeviceScene(s, local)
    flat = s.flatten()
    auto = "wavefront" if flat.n_nodes > 0 else "mega"
    vname = args.variant if args.variant != "auto" else auto
    variant = N.GRT_VARIANT_WAVEFRONT if vname == "wavefront" else N.GRT_VARIANT_MEGAKERNEL

    acc = torch.zeros(nval, dtype=torch.float32, device=dev)
    rgb8 = torch.zeros(nval, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step():
        acc.zero_()
        scene.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=args.seed, variant=variant,
                            sample_first=rank, sample_stride=world)
        if world > 1:
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.grt_launch_count()
    evs = []
    sync()
    for _ in range(args.steps):
        flush.fill_(1)                     # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    sync()
    launches = L.grt_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = paths * args.steps / (ms / 1e3) / 1e6

    # ---- end to end ("e2e") ------------------------------------------------------------------------
    # N = 1: the call a client makes, with HOST buffers: grt_scene_upload (host scene -> HBM), then grt_render (the
    # caller's zeroed rgb_sum goes up with a pageable cudaMemcpy, render, tonemap, sums + rgb8 come back down), free.
    # N > 1 (torchrun): every rank uploads + renders its shard, one NCCL reduce, rank 0 tonemaps and reads back.
    sbytes = scene_bytes(flat)
    h_sum = torch.zeros(nval, dtype=torch.float32).pin_memory()
    h_rgb8 = torch.zeros(nval, dtype=torch.uint8).pin_memory()
    h_zero = torch.zeros(nval, dtype=torch.float32).pin_memory()

    def e2e_step():
        if world == 1:
            sc = g.DeviceScene(s, local)
            sc.render(cam, seed=args.seed, variant=variant, want_rgb8=True)      # grt_render, host buffers
            sc.close()
            return
        sc = g.DeviceScene(s, local)                        # grt_scene_upload: host scene -> HBM
        acc.copy_(h_zero, non_blocking=True)                # caller's (zeroed) rgb_sum buffer up
        sc.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=args.seed, variant=variant,
                         sample_first=rank, sample_stride=world)
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            sc.tonemap_device(acc.data_ptr(), rgb8.data_ptr(), nval, 1.0 / S2, stream.cuda_stream)
            h_sum.copy_(acc, non_blocking=True)
            h_rgb8.copy_(rgb8, non_blocking=True)
        torch.cuda.synchronize()
        sc.close()

    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = paths * args.steps / float(t.item()) / 1e6

    # ---- (N > 1) the in-process path on the same GPUs, while the other ranks wait at the barrier ----------------
    inproc = None
    if world > 1 and not args.no_inproc:
        sync()
        # the other ranks wait on a CPU (gloo) barrier: an NCCL barrier is a kernel spinning on their GPUs, and rank 0 is
        # about to use those GPUs from its own process
        cpu_group = dist.new_group(backend="gloo")
        if rank == 0:
            try:
                devs = (C.c_int * world)(*range(world))
                opt = N.GrtOptions()
                opt.seed, opt.variant = args.seed, variant
                flat = s.flatten()          # (the view handed out earlier died with the e2e steps' own flatten calls)
                hs = np.zeros(nval, dtype=np.float32)
                kms = C.c_double(0)
                os.environ["GRT_MULTI_P2P"] = "1"
                with stdout_to_stderr():
                    N.check(L.grt_render_multi(C.byref(flat), C.byref(cam), C.byref(opt), devs, world, hs.ctypes.data, None, C.byref(kms)))
                    hs[:] = 0
                    tw = time.perf_counter()
                    N.check(L.grt_render_multi(C.byref(flat), C.byref(cam), C.byref(opt), devs, world, hs.ctypes.data, None, C.byref(kms)))
                    tw = time.perf_counter() - tw
                inproc = {"what": f"grt_render_multi on {world} GPUs from ONE process (rank 0; fused NVLink peer-atomic accumulation when available), one call",
                          "value": paths / (kms.value / 1e3) / 1e6, "device_ms": kms.value, "e2e_value": paths / tw / 1e6, "wall_ms": 1e3 * tw,
                          "mean": float(hs.mean() / S2)}
            except Exception as e:      # measurement extra: never fail the bench line over it
                inproc = {"error": str(e)}
        dist.barrier(group=cpu_group)
        sync()

    if rank == 0:
        props = torch.cuda.get_device_properties(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        cores = os.cpu_count() or 1
        # ---- roofline for the dominant kernel -----------------------------------------------------------
        acc.zero_()
        scene.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=args.seed, variant=variant, sample_first=rank, sample_stride=world,
                            flags=N.GRT_OPT_TIMING) if args.steps and paths / world <= 6e9 else None
        torch.cuda.synchronize()
        T = last_timing(L)
        dom_share = (T.extend_ms / T.total_ms) if T.total_ms > 0 else 1.0
        executed, lanes = None, None
        if name != "C5":
            # what THIS build executes (ordered runs, box slabs, wide nodes with leaf boxes): the megakernel's event counters
            # on an untimed full-resolution pass at low spp; lane occupancy at the REAL per-pixel sample count (centre window)
            s2, cfg2 = build_config(name, args.width or c["width"], 16 if S2 >= 16 else S2)
            sc2 = g.DeviceScene(s2, local)
            _, _, st = sc2.render(g.derive_camera(cfg2), seed=args.seed, variant=N.GRT_VARIANT_MEGAKERNEL, want_stats=True)
            sc2.close()
            executed = {k: st[k] / max(1, st["paths"]) for k in COSTS}
            if vname == "mega":
                cx, cy = max(0, cam.width // 2 - 64), max(0, cam.height // 2 - 64)
                _, _, so = scene.render(cam, seed=args.seed, variant=N.GRT_VARIANT_MEGAKERNEL, sample_first=rank, sample_stride=world,
                                        window=(cx, cy, min(cam.width, cx + 128), min(cam.height, cy + 128)), want_stats=True)
                lanes = so["lane_iterations"] / max(1, so["warp_iterations"])
        ref_events = dict(ORACLE_EVENTS_C2) if name == "C2" else None
        ref_src = "frozen (bench.py:ORACLE_EVENTS_C2)" if name == "C2" else None
        cpu = None
        if world == 1 and not args.no_cpu:
            # ---- CPU baseline (oracle port) on a bounded sample; it also counts the reference algorithm's events
            w_, s_ = c["cpu"]
            v, sec, ref_events, what, _ = cpu_sample(name, cores, width=(args.width or w_) if name != "C5" else w_, spp=args.cpu_spp or s_)
            cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": what}
            ref_src = "oracle counters of the cpu_baseline run"
        roofline = None
        if ref_events:
            rp = roofline_pair(ref_events, ref_src, value * 1e6 / world, S2, dom_share, peaks, sm_mhz, props, flat)
            traffic = None
            try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this size
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{name}:{cam.width}x{S2}")
            except Exception:
                pass
            bind = rp["fp32"] if rp["bound"] == "fp32_issue" else rp["bytes"]
            roofline = {"kernel": ("render_mega_kernel", "wf_extend (the wavefront variant's traversal + intersection kernel, one thread per slot)",
                                   "wf_extend_dyn (the wavefront variant's traversal + intersection kernel, persistent warps)")[int(T.extend_kernel)],
                        "bound": rp["bound"], "achieved": bind["achieved"], "peak": bind["peak"], "unit": bind["unit"], "frac": bind["frac"],
                        "peak_source": bind["peak_source"], "kernel_ms_per_step": T.extend_ms, "kernel_share_of_step": dom_share,
                        "kernel_launches_per_step": int(T.extend_launches),
                        "flops_per_path": rp["flops_per_path"], "bytes_per_path": rp["bytes_per_path"], "hbm_bytes_per_path": 12.0 / S2,
                        "events_per_path": ref_events, "events_source": ref_src, "executed_events_per_path": executed,
                        "lanes_per_warp_iteration": lanes, "traffic": traffic,
                        "fp32": rp["fp32"], "bytes": rp["bytes"],
                        "note": "not tensor-core work (no dense contraction). achieved = algorithmic flops (bytes) per launch / the kernel's launch time, "
                                "both from the frozen event-cost table (SURVEY.md 8d) and CUDA events inside this run; `fp32` and `bytes` give both fractions, "
                                "`bound` names the larger"}
        line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (f64 only for the winning hit's t, sphere quadratics and decisions within fp32 error of a boundary)", "data": "synthetic",
                "config": dict(workload(name, cam.width, cam.height, S2, c["spp"]), variant=vname, parallelism=f"spp-shard x{world} + ncclReduce",
                               l2="flushed between steps (256 MiB write); per-step CUDA events summed"),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": sbytes + nval * 4, "d2h_bytes_per_step": nval * 4 + nval,
                        "through": "grt_scene_upload + grt_render with host buffers (pageable), per step" if world == 1 else
                                   "per rank: grt_scene_upload + grt_render_device, one NCCL reduce, rank 0 tonemap + pinned read-back"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        if inproc is not None:
            line["inproc"] = inproc
        # ---- the other BASELINE.json configs, one short line each (N = 1 only) ------------------------------------
        if world == 1 and not args.no_configs and name == "C2" and not args.width and not args.spp:
            del acc, rgb8
            scene.close()
            torch.cuda.empty_cache()
            cfgs = {}
            for cn in CONFIGS:
                if cn == name:
                    cfgs[cn] = {"workload": line["config"]["workload"], "value": value, "unit": "Mpaths/s", "ms_per_step": ms / args.steps, "variant": vname,
                                "roofline": {k: roofline[k] for k in ("fp32", "bytes", "bound", "frac", "flops_per_path", "bytes_per_path")} if roofline else None,
                                "cpu_baseline": cpu, "see": "the top-level fields of this line"}
                    continue
                try:
                    cfgs[cn] = measure_config(cn, CONFIGS[cn]["side_spp"], 2, 1, local, args.seed, cores, not args.no_cpu, peaks, sm_mhz, props)
                except Exception as e:          # a side line must not cost the headline
                    cfgs[cn] = {"error": f"{type(e).__name__}: {e}"}
            line["configs"] = cfgs
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--variant", default="auto", choices=["auto", "mega", "wavefront"])
    ap.add_argument("--width", type=int, default=0, help="override the config's image width (not a BASELINE.json config then)")
    ap.add_argument("--spp", type=int, default=0, help="override the config's samples per pixel")
    ap.add_argument("--ref-spp", type=int, default=0, help="--impl reference: spp of each step's bounded sample")
    ap.add_argument("--cpu-spp", type=int, default=0, help="spp of the cpu_baseline sample (about 10-30 s of CPU work)")
    ap.add_argument("--seed", type=int, default=0xC0FFEE)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config lines of the default run")
    ap.add_argument("--no-inproc", action="store_true", help="N > 1: skip the extra grt_render_multi measurement on rank 0")
    ap.add_argument("--inproc", action="store_true", help="one process, N GPUs through grt_render_multi (no torchrun)")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build(quiet=True)
    if args.impl == "reference":
        run_reference(args)
    elif args.inproc:
        run_inproc(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
