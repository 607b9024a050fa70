#!/usr/bin/env python
"""bench.py — Mrays/s and Mpaths/s of the render path on the RTOW final scene (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the CPU arm: C++ restatement of the F# algorithm

A step is one frame: Scene.render of the C2 scene (1201x801, 500 spp, depth 50, adaptive early-out as the
reference does it).  `value` times the device-resident frame (probe + compact + main [+ all-reduces] +
finalize) with CUDA events, L2 flushed between steps, max over ranks.  `e2e` times the public API with host
buffers: Scene.make (BVH build + host->device upload) + Scene.render + Image.render (device->host image).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# FLOPs of the reference algorithm per ray: 3 (inverse directions) + 12 per slab test +
# 20 per sphere test + 14 per plane test + 1 per candidate + 6 (strike point)
def flops_per_ray(c):
    rays = max(1, c["rays"])
    return 3.0 + 12.0 * c["box_tests"] / rays + 20.0 * c["sphere_tests"] / rays + 14.0 * c["plane_tests"] / rays + c["candidates"] / rays + 6.0


# the same figure per config, measured by the oracle; used when the CPU sample does not run (N > 1)
FLOPS_PER_RAY = {"C1": 96.0, "C2": 414.0, "C3": 75.0, "C4": 131.0, "C5": 1927.0}


def metric_name(spec, config):
    if config == "C2":
        return "Mrays/s, RTOW final scene 1201x801 500spp depth 50 (Mpaths/s alongside)"
    return f"Mrays/s, {spec.name}, {spec.cols}x{spec.rows} {spec.spp}spp depth {spec.bounce_depth} (Mpaths/s alongside)"


def build_spec(args):
    from ray_tracing_fsharp_b200 import sample_images
    fn = sample_images.CONFIGS[args.config]
    spec = fn()
    if args.spp:
        spec.spp = args.spp
    if args.half_extents:
        spec.max_width_coord, spec.max_height_coord = args.half_extents
    return spec


def workload_name(spec, adaptive):
    return (f"{spec.name}, {spec.cols}x{spec.rows}, {spec.spp} spp, depth {spec.bounce_depth}, "
            f"adaptive early-out {'on (Scene.fs:157-194)' if adaptive else 'off'}")


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the F# algorithm), all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_sample(spec, target_seconds, threads, seed=7):
    """Renders every `row_step`-th row of the frame at full spp with the oracle, sized for ~target_seconds."""
    import oracle
    from ray_tracing_fsharp_b200.domain import marshal
    hs, ts, _keep = marshal(spec.objects)
    scene = oracle.Scene(hs, ts)
    cam = oracle.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    # calibrate on a few rows spread over the frame, at no more than 64 spp (a row of the 100 k-sphere frame at 4096 spp
    # takes the host cores 20 s); the time of a row is at most linear in spp beyond that
    step0 = max(1, spec.rows // 8)
    spp_cal = min(spec.spp, 64)
    cam_cal = oracle.camera_make_basic(spp_cal, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam_cal.bounce_depth = spec.bounce_depth
    t0 = time.perf_counter()
    _, _, c0, rows0 = scene.render(cam_cal, spec.max_width_coord, spec.max_height_coord, seed=seed, rng_mode=0, adaptive=True, threads=threads,
                                   row_begin=step0 // 2, row_step=step0)
    dt0 = time.perf_counter() - t0
    per_row = dt0 / max(1, rows0) * (spec.spp / spp_cal)
    n_rows = int(min(spec.rows, max(1, target_seconds / max(per_row, 1e-9))))
    row_step = max(1, spec.rows // max(1, n_rows))
    t0 = time.perf_counter()
    _, _, c, rows = scene.render(cam, spec.max_width_coord, spec.max_height_coord, seed=seed + 1, rng_mode=0, adaptive=True, threads=threads,
                                 row_begin=row_step // 2, row_step=row_step)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "rows": rows, "row_step": row_step, "counters": c}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec = build_spec(args)
    threads = os.cpu_count() or 1
    per_step = []
    counters = None
    rows = row_step = 0
    # every step is a bounded sample of the frame; size it so that the whole run stays within a couple of minutes
    per_step_seconds = min(args.cpu_seconds, 120.0 / max(1, args.warmup + args.steps))
    for i in range(args.warmup + args.steps):
        s = cpu_sample(spec, per_step_seconds, threads, seed=100 + i)
        if i >= args.warmup:
            per_step.append(s)
        counters, rows, row_step = s["counters"], s["rows"], s["row_step"]
    secs = sum(s["seconds"] for s in per_step)
    rays = sum(s["counters"]["rays"] for s in per_step)
    paths = sum(s["counters"]["paths"] for s in per_step)
    mrays = rays / secs / 1e6
    sample = f"every {row_step}th row of the frame ({rows} of {spec.rows} rows) at full spp per step"
    line = {
        "impl": "reference",
        "metric": metric_name(spec, args.config),
        "note": "CPU arm = the oracle, a C++ restatement of the F# CPU renderer, row-parallel on all host threads (.NET is absent from this image)",
        "value": mrays, "unit": "Mrays/s", "mpaths_per_s": paths / secs / 1e6,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(per_step)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(spec, True), "sample": sample},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "flops_per_ray_reference_algorithm": flops_per_ray(counters),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.samples:
            if t < t_begin or t > t_end + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ray_tracing_fsharp_b200 import abi, native
    from ray_tracing_fsharp_b200.distributed import DeviceBackend, render_split_frame
    from ray_tracing_fsharp_b200.domain import marshal
    from ray_tracing_fsharp_b200.scene import Camera, Image, Scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if native.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    adaptive = not args.no_adaptive
    spec = build_spec(args)
    cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    max_w, max_h = spec.max_width_coord, spec.max_height_coord
    hs, ts, keep = marshal(spec.objects)
    scene = native.SceneHandle(hs, ts, local_rank, keepalive=keep)
    flags = (abi.RT_FLAG_NO_SMEM if args.no_smem else 0)
    backend = DeviceBackend(scene, cam, max_w, max_h, seed=1, adaptive=adaptive, flags=flags)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def red_max(t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX)

    def red_sum(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    main_events = []

    def frame(seed, time_main=False):
        """probe -> [all-reduce flags] -> compact + main -> [all-reduce sums] -> finalize on rank 0 (distributed.py)."""
        backend.opts.seed = seed
        stats, flags = backend.alloc()
        backend.probe(rank, world, stats, flags)
        if world > 1:
            red_max(flags)
        if time_main:  # the dominant kernel, for the roofline: events on the stream the kernels are launched on
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
        backend.main(rank, world, stats, flags)
        if time_main:
            m1.record()
            main_events.append((m0, m1))
        if world > 1:
            red_sum(stats)
        return backend.finalize(stats) if rank == 0 else None

    fp32_peak = native.measure_fp32_peak(local_rank) if rank == 0 else 0.0

    for i in range(args.warmup):
        frame(1000 + i)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t_begin = time.perf_counter()
    ms_total = 0.0
    rays = paths = 0
    launches0 = backend.launches
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frame(2000 + i, time_main=True)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        c = backend.counters()
        work = torch.tensor([c.rays, c.paths], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(work, op=dist.ReduceOp.SUM)
        ms_total += float(ms.item())
        rays += int(work[0].item())
        paths += int(work[1].item())
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = backend.launches - launches0
    launch_t = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(launch_t, op=dist.ReduceOp.SUM)
    secs = ms_total / 1e3
    value = rays / secs / 1e6
    # the main-phase kernel alone (rank 0's launches): its share of the rays comes from an untimed probe-only frame
    main_ms = sum(a.elapsed_time(b) for a, b in main_events) / max(1, len(main_events))
    backend.opts.seed = 2000 + args.steps - 1
    ps_, pf_ = backend.alloc()
    backend.probe(rank, world, ps_, pf_)
    probe_rays_rank0 = backend.counters().rays
    if world > 1:
        red_max(pf_)
    backend.main(rank, world, ps_, pf_)
    main_rays_rank0 = backend.counters().rays - probe_rays_rank0

    # ---- end to end through the public API with host buffers ----
    e2e_secs = 0.0
    e2e_rays = 0
    h2d = d2h = 0
    n_e2e = max(1, min(args.steps, 3))
    for i in range(-2, n_e2e):  # two untimed warm-up iterations (first-use allocations: handle pool, the two recycled output frames)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            sc = Scene.make(spec.objects, device=local_rank)              # host BVH build + H2D upload of the scene
            _, image = Scene.render(lambda _p: None, lambda _s: None, max_w, max_h, cam, sc, seed=3000 + i, adaptive=adaptive, flags=flags)
            pixels = Image.render(image)                                  # kernels + D2H of the image
            st = sc.last_stats
            r_rays = st.rays
            h2d = sc.handle.device_bytes()
            d2h = pixels.nbytes + 64
            sc.handle.close()
        else:
            hs2, ts2, keep2 = marshal(spec.objects)
            sc2 = native.SceneHandle(hs2, ts2, local_rank, keepalive=keep2)
            be = DeviceBackend(sc2, cam, max_w, max_h, seed=3000 + i, adaptive=adaptive, flags=flags)
            stats_t, _ = render_split_frame(be, rank, world, red_max, red_sum)
            if rank == 0:
                pixels = be.finalize_to_host(stats_t)  # device->host copy of the frame into pinned host memory
                d2h = pixels.nbytes + 64
            torch.cuda.synchronize(dev)
            r_rays = be.counters().rays
            h2d = sc2.device_bytes()
            sc2.close()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        rr = torch.tensor([r_rays], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        if i >= 0:
            e2e_secs += float(dt.item())
            e2e_rays += int(rr.item())

    # one extra, untimed frame with the traversal counters on (a slower kernel variant): tests per ray of OUR traversal
    traversal = None
    if args.counters:
        be = DeviceBackend(scene, cam, max_w, max_h, seed=4000, adaptive=adaptive, flags=flags | abi.RT_FLAG_COUNTERS)
        render_split_frame(be, rank, world, red_max, red_sum)
        c = be.counters()
        traversal = {"box_tests_per_ray": c.box_tests / max(1, c.rays), "prim_tests_per_ray": c.prim_tests / max(1, c.rays),
                     "rays_per_path": c.rays / max(1, c.paths)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            s = cpu_sample(spec, args.cpu_seconds, threads)
            cpu = {"value": s["counters"]["rays"] / s["seconds"] / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"every {s['row_step']}th row of the frame ({s['rows']} of {spec.rows} rows) at full spp, {s['seconds']:.1f} s",
                   "mpaths_per_s": s["counters"]["paths"] / s["seconds"] / 1e6, "counters": s["counters"]}
        # algorithmic FLOPs per ray of the REFERENCE traversal on this scene: measured by the oracle on the CPU
        # sample when it ran, else the figure recorded in DESIGN.md for C2
        f_ray = flops_per_ray(cpu["counters"]) if cpu else (args.flops_per_ray or FLOPS_PER_RAY[args.config])
        achieved = main_rays_rank0 / (main_ms / 1e3) * f_ray / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp) and args.config == "C2" and not args.spp and not args.half_extents:  # the ncu capture is of the C2 main kernel
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": metric_name(spec, args.config),
            "value": value, "unit": "Mrays/s", "mpaths_per_s": paths / secs / 1e6,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(spec, adaptive), "l2": "flushed between steps (256 MiB fill)",
                       "parallelism": f"sample-split x{world}" + (", NCCL all-reduce of flags (max) and sums (int32 sum)" if world > 1 else ""),
                       "scene_bytes_staged_in_shared_memory": 0 if args.no_smem else scene.shared_memory_bytes()},
            "rays_per_step": rays / args.steps, "paths_per_step": paths / args.steps,
            "e2e": {"value": e2e_rays / e2e_secs / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_secs / n_e2e, "steps": n_e2e,
                    "what": ("Scene.make (BVH build + upload) + Scene.render + Image.render with host buffers" if world == 1 else
                             "per rank: marshal + rt_scene_create (BVH build + upload) + rt_device_probe / all-reduce / rt_device_main / all-reduce; "
                             "rank 0: rt_device_finalize + device->host copy of the frame into pinned host memory")},
            "gpu_launches": int(launch_t.item()),
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                         "traffic": traffic, "flops_per_ray": f_ray,
                         "kernel": "render_kernel<PROBE=0> (main phase), rank 0", "kernel_ms_per_launch": main_ms, "rays_per_launch": main_rays_rank0,
                         "note": "achieved = rays per launch x FLOPs the REFERENCE traversal spends per ray (exhaustive DFS, SURVEY 8d) / CUDA-event "
                                 "duration of the launch; peak = FP32 FMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 "
                                 "figure); traffic = DRAM bytes per launch, ncu, profiles/roofline_traffic.json"},
        }
        if traversal:
            line["traversal"] = traversal
        if cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu.items() if k != "counters"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (a reduced-size run; not the headline)")
    ap.add_argument("--half-extents", type=int, nargs=2, default=None, help="override maxWidthCoord maxHeightCoord")
    ap.add_argument("--no-adaptive", action="store_true")
    ap.add_argument("--no-smem", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--counters", action="store_true", help="also report box / primitive tests per ray of the device traversal")
    ap.add_argument("--cpu-seconds", type=float, default=25.0)
    ap.add_argument("--flops-per-ray", type=float, default=0.0, help="FLOPs of the reference traversal per ray, used when the CPU sample does not run (default: the config's figure from DESIGN.md)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
