#!/usr/bin/env python
"""bench.py — Mpaths/s of the render hot path on the Cornell box 1024^2 x 4096 spp (BASELINE.json config C2).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA backend through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on host cores

One "step" = one complete render of the workload (one pass of the hot path over all W*H*spp paths).
Multi-GPU (torchrun, one rank per GPU): the strata set is split s = rank (mod N) (strong scaling: the
workload is fixed), every rank renders its shard into a private fp32 sum buffer and ONE NCCL reduce to rank 0
combines them (camera.go has no analogue).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Frozen algorithmic-cost table: flops and bytes per EVENT, counted from
# the cited reference lines / the flat fp32 layout.  Events are counted by the kernel's STATS build.
COSTS = {
    "box_tests": (21, 32),        # aabb.go:94-110, one 32-byte node
    "quad_tests": (30, 64),       # objects.go:168-194: 12 (plane reject) .. 54 (hit); 30 = documented average
    "sphere_tests": (40, 64),     # objects.go:84-113: 29 (reject) .. 55 (hit)
    "tri_tests": (35, 48),        # objects.go:409-456: 20 .. 63
    "medium_tests": (21, 16),     # medium.go:39-51 (the two boundary traversals are counted as box/quad tests)
    "shade_diffuse": (125, 32),   # onb.go:13-25 + vec.go:177-186 + pdf.go:33-74 + camera.go:325-330
    "shade_specular": (50, 32),   # materials.go:70-130
    "light_pdf_evals": (80, 0),   # objects.go:152-160 re-intersection + pdf
    "paths": (23, 0),             # camera.go:257-269
}


# Event counts per path of the REFERENCE algorithm on this workload (the oracle's counters: reference BVH, every
# continuation traced, SURVEY.md 8d "events counted by the oracle").  Used when the CPU leg is skipped; the CPU leg
# re-measures them.  Source: oracle, cornellBox 1024x1024 x 16 spp, seed 0xC0FFEE.
ORACLE_EVENTS_PER_PATH = {"paths": 1.0, "box_tests": 27.06, "quad_tests": 17.55, "sphere_tests": 0.0, "tri_tests": 0.0,
                          "medium_tests": 0.0, "shade_diffuse": 1.925, "shade_specular": 0.0, "light_pdf_evals": 1.925}


def algorithmic_cost(stats):
    paths = max(1, stats["paths"])
    fl = sum(stats[k] * c[0] for k, c in COSTS.items())
    by = sum(stats[k] * c[1] for k, c in COSTS.items())
    return fl / paths, by / paths


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


def workload(args):
    return {"workload": f"cornellBox (main.go:278-320) {args.width}x{args.width} x {args.spp} spp, MaxDepth 50",
            "scene": 6, "width": args.width, "spp": args.spp, "paths_per_step": args.width * args.width * (int(args.spp ** 0.5) ** 2)}


def run_reference(args):
    """The reference's own CPU algorithm (fp64 oracle port; the Go binary cannot be built here: no Go
    toolchain) on all host cores.  Each step is a BOUNDED sample of the workload: full resolution,
    reduced spp (Mpaths/s is spp-independent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import go_raytracer_b200 as g
    from oracle import oracle_py as O
    cores = os.cpu_count() or 1
    s, cfg = g.builtin_scene(6, width=args.width, spp=args.ref_spp)
    ow = O.OracleWorld(s)
    cam = O.derived_camera(cfg)
    paths = cam.width * cam.height * cam.spp_sqrt ** 2
    for _ in range(args.warmup):
        ow.render(cfg, nthreads=cores, window=(0, 0, cam.width, max(1, cam.height // 8)))
    t = 0.0
    for _ in range(args.steps):
        _, _, _, sec = ow.render(cfg, nthreads=cores)
        t += sec
    val = paths * args.steps / t / 1e6
    sample = f"{cam.width}x{cam.height} x {cam.spp_sqrt ** 2} spp per step (full resolution, reduced spp), {cores} threads, one task per row"
    line = {"impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import ctypes as C
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

    s, cfg = g.builtin_scene(6, width=args.width, spp=args.spp)
    cam = g.derive_camera(cfg)
    S2 = cam.spp_sqrt ** 2
    paths = cam.width * cam.height * S2
    nval = cam.width * cam.height * 3
    scene = g.DeviceScene(s, local)
    variant = N.GRT_VARIANT_WAVEFRONT if args.variant == "wavefront" else N.GRT_VARIANT_MEGAKERNEL

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

    # ---- end to end through the C ABI with HOST buffers ("e2e") ------------------------------------
    # per step: scene upload (H2D), zeroed host accumulation buffer up, render, reduce, tonemap, sums + rgb8 down
    flat = s.flatten()
    scene_bytes = (flat.n_nodes * 32 + flat.n_spheres * 64 + flat.n_quads * 96 + flat.n_tris * 48 + flat.n_items * 4 +
                   flat.n_media * 16 + flat.n_materials * 32 + flat.n_textures * 32 + flat.n_lights * 304)
    h_sum = torch.zeros(nval, dtype=torch.float32).pin_memory()
    h_rgb8 = torch.zeros(nval, dtype=torch.uint8).pin_memory()
    h_zero = torch.zeros(nval, dtype=torch.float32).pin_memory()

    def e2e_step():
        sc = g.DeviceScene(s, local)                        # grt_scene_upload: host scene -> HBM
        acc.copy_(h_zero, non_blocking=True)                # caller's (zeroed) rgb_sum buffer up
        sc.render_device(cam, acc.data_ptr(), stream.cuda_stream, seed=args.seed, variant=variant,
                         sample_first=rank, sample_stride=world)
        if world > 1:
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

    if rank == 0:
        # ---- roofline for the dominant kernel (render_mega_kernel) ------------------------------------
        # event counts from the STATS build on an untimed full-resolution pass at low spp
        s2, cfg2 = g.builtin_scene(6, width=args.width, spp=16)
        cam2 = g.derive_camera(cfg2)
        sc2 = g.DeviceScene(s2, local)
        _, _, st = sc2.render(cam2, seed=args.seed, variant=N.GRT_VARIANT_MEGAKERNEL, want_stats=True)   # event counts are variant-independent
        sc2.close()
        # lane occupancy of the kernel's trace phase at the REAL per-pixel sample count (a centre window is enough)
        c0 = max(0, cam.width // 2 - 64)
        _, _, st_occ = scene.render(cam, seed=args.seed, variant=N.GRT_VARIANT_MEGAKERNEL, sample_first=rank, sample_stride=world,
                                    window=(c0, c0, min(cam.width, c0 + 128), min(cam.height, c0 + 128)), want_stats=True)
        executed = {k: st[k] / st["paths"] for k in COSTS}            # what THIS kernel executes (ordered runs, box slabs, zero-weight cut-off)
        ref_events = dict(ORACLE_EVENTS_PER_PATH)
        ref_src = "frozen (bench.py:ORACLE_EVENTS_PER_PATH)"
        hbm_bytes_pp = 12.0 / S2                                      # what must reach HBM: one fp32 RGB store per pixel
        cpu = None
        if world == 1 and not args.no_cpu:
            # ---- CPU baseline (oracle port) on a bounded sample; it also counts the reference algorithm's events
            from oracle import oracle_py as O
            cores = os.cpu_count() or 1
            sb, cfgb = g.builtin_scene(6, width=args.width, spp=args.cpu_spp)
            ow = O.OracleWorld(sb)
            camb = O.derived_camera(cfgb)
            _, _, ost, sec = ow.render(cfgb, nthreads=cores, want_stats=True)
            pb = camb.width * camb.height * camb.spp_sqrt ** 2
            cpu = {"value": pb / sec / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "port",
                   "sample": f"{camb.width}x{camb.height} x {camb.spp_sqrt ** 2} spp (full resolution, reduced spp), {sec:.1f} s, "
                             "C++ fp64 restatement of the Go renderer (no Go toolchain here); faster than the Go binary would be"}
            ref_events = {k: ost[k] / ost["paths"] for k in COSTS}
            ref_src = "oracle counters of the cpu_baseline run"
        flops_pp, onchip_bytes_pp = algorithmic_cost(dict({k: v for k, v in ref_events.items()}, paths=1.0))
        kernel_ms = ms / args.steps                                  # the megakernel is >99.9 % of the step
        props = torch.cuda.get_device_properties(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12        # TFLOP/s at the clock seen
        per_gpu_paths = paths / world
        ach_tf = per_gpu_paths * flops_pp / (kernel_ms / 1e3) / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_gbs = per_gpu_paths * hbm_bytes_pp / (kernel_ms / 1e3) / 1e9
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this size
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.width}x{args.spp}")
        except Exception:
            pass
        roofline = {"kernel": "render_mega_kernel" if variant == 0 else "wavefront kernels",
                    "bound": "fp32_issue", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                    "peak_source": f"SMs({props.multi_processor_count}) x 128 lanes x 2 x {sm_mhz:.0f} MHz observed during the run",
                    "flops_per_path": flops_pp, "onchip_bytes_per_path": onchip_bytes_pp, "hbm_bytes_per_path": hbm_bytes_pp,
                    "events_per_path": ref_events, "events_source": ref_src, "executed_events_per_path": executed,
                    "lanes_per_warp_iteration": st_occ["lane_iterations"] / max(1, st_occ["warp_iterations"]),
                    "traffic": traffic,
                    "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                    "note": "not tensor-core work (no dense contraction); the 4 KB scene is shared-memory resident, so the binding "
                            "roofline is FP32 issue (SURVEY.md 8d); the hbm object is the HBM roofline on the bytes that must reach "
                            "HBM (one RGB store per pixel), reported to show the kernel is nowhere near it"}
        line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (f64 only for the winning hit's t and for decisions within fp32 error of a boundary)", "data": "synthetic",
                "config": dict(workload(args), variant=args.variant, parallelism=f"spp-shard x{world} + ncclReduce",
                               l2="flushed between steps (256 MiB write); per-step CUDA events summed"),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": scene_bytes + nval * 4,
                        "d2h_bytes_per_step": nval * 4 + nval},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
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
    ap.add_argument("--variant", default="mega", choices=["mega", "wavefront"])
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--ref-spp", type=int, default=64, help="--impl reference: spp of each step's bounded sample")
    ap.add_argument("--cpu-spp", type=int, default=256, help="spp of the cpu_baseline sample (about 10-30 s of CPU work)")
    ap.add_argument("--seed", type=int, default=0xC0FFEE)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build(quiet=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
