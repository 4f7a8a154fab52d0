#!/usr/bin/env python
"""Headline benchmark: Mpaths/s at 1080p, max depth 8 (BASELINE.json metric), config C2
(floor + 1000 mixed-material spheres, 1920x1080, 64 spp, depth 8), synthetic seeded scene and sky.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference --gpus N ...           # the reference's own CPU code (oracle/_ref host build)

A step = one frame of the workload (W*H*spp paths) through the wavefront path.  `value` is device-timed
(CUDA events, max over ranks) with the scene resident in HBM; `e2e` goes through the C ABI with host buffers
(scene upload + LBVH build + render + read-back of the mean image into pinned host memory) by wall clock.
N > 1 is weak scaling by samples-per-pixel: every rank renders its own 64-sample Philox range of the same
frame and the SUM buffers are reduced to rank 0 with one NCCL reduce per frame (inside the timed region).

Beside the headline the line carries a `configs` block with the other BASELINE.json configurations, measured in
the same process: C1 and C3 on one GPU (rank 0), C4 (3840x2160, a FIXED 1024 spp split over the ranks: strong
scaling, one reduce) and C5 (120 frames round-robin over the ranks, per-frame rebuild, wall clock) on all ranks,
and for N > 1 `multi_gpu_check`: the reduced SUM buffer against rank 0 rendering the same samples alone.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpaths/s at 1080p, max depth 8"
WORKLOAD = {"workload": "C2: floor + 1000 mixed-material spheres, 1920x1080, 64 spp, depth 8",
            "width": 1920, "height": 1080, "spp": 64, "max_depth": 8, "n_spheres": 1000, "scene_seed": 12345,
            "l2": "hit queues (about 30 GB streamed per 64-spp frame, written by one launch and read by the next) exceed the 126 MB L2; no explicit flush"}
# HBM byte models of the dominant kernel k_bounce.
#   a wavefront design that keeps rays in HBM moves 144 B per path segment + 16 B per path.  That is the
#     contract's algorithmic figure and what `roofline.frac` is computed from.
#   This kernel's own traffic: rays never leave registers, a segment that leaves the scene is shaded where it ends, so
#     only a segment THAT HITS costs one 48-byte hit record written by the launch that traces it and read by the launch
#     that shades it (96 B), plus 16 B of radiance per path.  Reported as `roofline_own_bytes`; ncu's measured DRAM
#     bytes agree with this model to 4 %.
SURVEY_BYTES_PER_SEGMENT = 144
OWN_BYTES_PER_HIT_RECORD = 96
BYTES_PER_PATH = 16


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which made the reference
    arm single-threaded at N > 1 in round 1: the thread count is passed to the OpenMP region explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []   # lines: (arrival time, text)
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # nvidia-smi needs a few hundred ms to start, so it is launched before the warm-up; only the samples that
        # arrived inside the marked (timed) region count
        inside = [ln for (t, ln) in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload():
    from cpppathtracer_b200 import scenes
    sc = scenes.sphere_field(WORKLOAD["n_spheres"], WORKLOAD["width"], WORKLOAD["height"], WORKLOAD["scene_seed"])
    sky = scenes.synthetic_sky(1280, 1280)
    return sc, sky


def cpu_baseline_leg(sc, sky, budget_s: float = 15.0):
    """The reference's own tracing code on the host cores (oracle/_ref), bounded sample of the workload."""
    from oracle import ref_host
    orc = ref_host.best_oracle()
    orc.load_scene(sc)
    orc.set_sky(sky)
    cores = host_threads()
    t0 = time.perf_counter()
    _, st = orc.render(sc.camera, 1, WORKLOAD["max_depth"], seed=1, nthreads=cores)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(WORKLOAD["spp"], budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    _, st = orc.render(sc.camera, spp, WORKLOAD["max_depth"], seed=1, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": st["paths"] / dt / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": orc.kind,
            "sample": f"{sc.camera.width}x{sc.camera.height} x {spp} spp of the workload scene, depth 8 "
                      f"({st['paths'] / 1e6:.1f} M paths, {dt:.1f} s, OpenMP over rows, {cores} threads)",
            "segments_per_path": st["segments"] / st["paths"]}


def reference_arm(args, out):
    """`--impl reference`: the reference's CPU implementation of the path, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)     # before the OpenMP runtime of the oracle library initialises
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import ref_host
    sc, sky = build_workload()
    orc = ref_host.best_oracle()
    orc.load_scene(sc)
    orc.set_sky(sky)
    # bounded sample per step: ~1.5 s of CPU work
    t0 = time.perf_counter()
    orc.render(sc.camera, 1, WORKLOAD["max_depth"], seed=1, nthreads=cores)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(WORKLOAD["spp"], 1.5 / max(t1, 1e-3))))
    for _ in range(args.warmup):
        orc.render(sc.camera, spp, WORKLOAD["max_depth"], seed=2, nthreads=cores)
    paths = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, st = orc.render(sc.camera, spp, WORKLOAD["max_depth"], seed=3, first_sample=k * spp, nthreads=cores)
        paths += st["paths"]
    dt = time.perf_counter() - t0
    v = paths / dt / 1e6
    sample = f"{sc.camera.width}x{sc.camera.height} x {spp} spp per step of the workload scene, depth 8, {cores} OpenMP threads"
    print(file=out, flush=True, *[json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(WORKLOAD, sample=sample),
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": orc.kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})])


def visit_counts(r, cam, depth, spp=2):
    """Node visits / primitive tests per segment of THIS BVH (instrumented kernel, a few spp)."""
    from cpppathtracer_b200 import api
    r.set_option(api.PT_OPT_COUNT_STEPS, 1)
    r.reset_accum()
    r.render(cam, spp, depth, seed=7)
    st = r.stats()
    r.set_option(api.PT_OPT_COUNT_STEPS, 0)
    return st["node_steps"] / st["segments"], st["prim_tests"] / st["segments"], st["paths"] / st["segments"]


def fp32_view(r, cam, depth, mrays_per_s, sm_mhz):
    """SURVEY.md 8(d)'s second ceiling: algorithmic FP32 work per segment from THIS BVH's measured visit counts
    (instrumented kernel, 2 spp) against 148 SMs x 128 lanes x 2 flop x clock."""
    nodes, prims, paths_per_seg = visit_counts(r, cam, depth)
    # per node visit two child boxes of 18 flops (slab test with precomputed 1/d), ~30 per sphere test, 10 for the
    # platform, ~150 per shade, 40 per camera ray
    flops_seg = nodes * 2 * 18 + prims * 30 + 10 + 150 + 40 * paths_per_seg
    peak = 148 * 128 * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12
    achieved = mrays_per_s * 1e6 * flops_seg / 1e12
    return {"bound": "fp32 issue", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "algorithmic_flops_per_segment": flops_seg, "node_visits_per_segment": nodes, "prim_tests_per_segment": prims,
            "note": "algorithmic flops only (min/max, compares, address and stack arithmetic, Philox and IEEE division "
                    "sequences are not counted): the kernel issues 2.4-3.2 of 4 warp-instructions per clock per SM"}


def fma_variant_leg(sc, sky):
    """Same workload through libpt_b200_fma.so (reported beside the headline, which is the default library's)."""
    from cpppathtracer_b200 import api
    if not os.path.exists(api.LIB_PATH_FMA):
        return None
    r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_FMA)
    for _ in range(3):
        r.reset_accum()
        r.render(sc.camera, WORKLOAD["spp"], WORKLOAD["max_depth"], seed=2026)
    st = r.stats()
    r.close()
    return {"value": st["paths"] / st["render_ms"] / 1e3, "unit": "Mpaths/s", "how": "device time of one 64-spp frame (CUDA events), third of three"}


def ref_cuda_leg(sc, sky):
    """Reported baseline (north_star): the reference's own CUDA kernels rebuilt for sm_100, same scene."""
    try:
        from oracle import ref_cuda
        if not ref_cuda.available():
            return None
        R = ref_cuda.RefCuda()
        R.load_scene(sc)
        R.set_sky(sky)
        out = R.render(sc.camera, 8, warmup=2)
        return {"mpaths_kernel_only": out["mpaths_kernel"], "mpaths_service_loop": out["mpaths_service"],
                "traced_pixels": out["traced_pixels"], "spp": 8,
                "note": "SamplePixel megakernel (path_tracer.cu:124-175), XORWOW RNG, grid W/16 x H/16 (1072 of 1080 rows)"}
    except Exception as e:  # the baseline is optional evidence, never a reason to fail the bench
        return {"unavailable": str(e)[:200]}


def graze_leg(r, sc, segments=(0, 1, 2, 3)):
    """Closest-hit primitive ids of this path against the reference's own intersection code on real path rays of the
    config's scene (camera at 480x270, segments 0-3): north_star excludes grazing / tie rays from the bit-exact id
    requirement and wants them COUNTED.  The oracle is the checker here, nothing it computes is timed or reported as
    a result of the product."""
    from oracle import ref_host
    port = ref_host.PortOracle()
    port.load_scene(sc)
    ref = ref_host.RefHost() if ref_host.available() else port
    if ref is not port:
        ref.load_scene(sc)
    cam = copy.deepcopy(sc.camera)
    cam.width, cam.height = 480, 270
    n = mism = t_diff = hits = 0
    for seg in segments:
        rays, valid = port.segment_rays(cam, seg, sample=0, max_depth=8, seed=1)
        rays = rays[valid]
        if len(rays) == 0:
            continue
        pg, tg, _, _ = r.trace(rays)
        pr, tr, _, _ = ref.trace(rays)
        bad = pg != pr
        n += len(rays); mism += int(bad.sum())
        both = (~bad) & (pg >= 0)
        hits += int(both.sum()); t_diff += int((tg[both] != tr[both]).sum())
    return {"rays": n, "prim_id_mismatches": mism, "hits_compared": hits, "hit_distance_not_bit_identical": t_diff,
            "oracle": ref.kind, "what": "graze / tie class (DESIGN.md section 4): excluded from the bit-exact id bar and counted"}


def config_single_gpu(name, what, sc, sky, spp, frames, depth, stream, with_graze):
    """One BASELINE config on this rank's GPU: device-timed frames, statistics of the last one, visit counts."""
    import torch
    from cpppathtracer_b200 import api
    r = api.Renderer(sc, sky)
    r.set_stream(stream)
    cam = sc.camera
    for _ in range(3):
        r.reset_accum()
        r.render(cam, spp, depth, seed=2026, sync=False)
    r.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        r.reset_accum()
        r.render(cam, spp, depth, seed=2026, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    r.synchronize()
    st = r.stats()
    nodes, prims, _ = visit_counts(r, cam, depth, spp=1 if sc.n_prims > 10000 else 2)
    out = {"workload": what, "mpaths_per_s": cam.width * cam.height * spp / ms / 1e3, "ms_per_frame": ms, "frames_timed": frames,
           "n_gpus": 1, "segments_per_path": st["segments"] / st["paths"], "mrays_per_s": st["segments"] / ms / 1e3,
           "node_visits_per_segment": nodes, "prim_tests_per_segment": prims, "n_prims": sc.n_prims,
           "bvh_sah": st["bvh_sah"], "entry_grid_levels": st["entry_levels"], "wave_paths": st["wave_paths"]}
    if with_graze:
        try:
            out["graze"] = graze_leg(r, sc)
        except Exception as e:
            out["graze"] = {"unavailable": str(e)[:200]}
    r.close()
    return out


def _quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the driver wants exactly ONE JSON line there.
    Point fd 1 at stderr for the duration of the run and return a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu / reference-CUDA baseline legs and the graze counts")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 legs (profiling runs)")
    ap.add_argument("--wave-paths", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args, real_stdout)
        return

    import torch
    import torch.distributed as dist
    from cpppathtracer_b200 import api, scenes, video, distributed as ptd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(args.warmup, 3)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sc, sky = build_workload()
    cam = sc.camera
    spp, depth = WORKLOAD["spp"], WORKLOAD["max_depth"]
    r = api.Renderer(sc, sky)
    r.set_stream(stream)
    r.set_option(api.PT_OPT_STAGE_TIMING, 1)
    if args.wave_paths:
        r.set_option(api.PT_OPT_WAVE_PATHS, args.wave_paths)
    first = rank * spp          # weak scaling: rank g renders samples [64 g, 64 (g+1)) of the same frame

    def step():
        r.reset_accum()
        r.render(cam, spp, depth, seed=2026, first_sample=first, sync=False)
        if world > 1:
            ptd.reduce_sum_to_root(ptd.accum_tensor(r), renderer=r)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step()
    r.synchronize()
    launches0 = r.stats()["launches"]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = None
    r.synchronize()
    st = r.stats()
    launches = st["launches"] - launches0
    ms = max_over_ranks(ms)
    paths_per_step = cam.width * cam.height * spp * world
    value = paths_per_step * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers (wall clock) --------------------------------
    host_imgs = [torch.empty((cam.height, cam.width, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    host_img = host_imgs[0]
    e2e_steps = max(5, min(args.steps, 20))
    h2d = sc.n_prims * (32 + 4 + 4) + len(sc.materials) * 48
    d2h = cam.width * cam.height * 16

    def e2e_step(k, first_of_run):
        # the caller edits an object, rebuilds (upload + LBVH), renders and reads the mean image back.  The read-back is the
        # pipelined pair pt_read_accum_begin / _end: the copy of frame k runs under the build and render of frame k + 1, like
        # the reference's render thread hands a finished frame to its callback while it goes on (path_tracer.cu:256-306);
        # every frame is complete in host memory before the timed region ends
        r.update_prim(1, sc.center[1], float(sc.radius[1]))
        r.build()
        r.reset_accum()
        r.render(cam, spp, depth, seed=2026, first_sample=first, sync=False)
        if world > 1:
            ptd.reduce_sum_to_root(ptd.accum_tensor(r), renderer=r)
        if rank == 0:
            r.read_accum_begin(host_imgs[k & 1])
            if not first_of_run:
                r.read_accum_end()
        else:
            r.synchronize()

    e2e_step(0, True)
    if rank == 0:
        r.read_accum_end()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k, k == 0)
    if rank == 0:
        r.read_accum_end()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = paths_per_step * e2e_steps / e2e_s / 1e6
    if rank == 0:
        clocks = sampler.stop()
        if not clocks.get("samples"):
            # the timed region was shorter than one sampling period: use the samples of the e2e leg (same workload,
            # same kernels, back to back) and say so
            sampler.t1 = time.perf_counter()
            clocks = sampler.stop()
            clocks["window"] = "timed region + e2e leg"
        else:
            clocks["window"] = "timed region"

    # ---- N > 1: the reduced image against rank 0 rendering the same samples alone ----------------------
    multi_check = None
    if world > 1:
        sc_s = scenes.sphere_field(1000, 960, 540)
        rs = api.Renderer(sc_s, sky)
        rs.set_stream(stream)
        total_spp = 8 * world + 3               # not divisible by the world size on purpose
        t = ptd.render_sharded(rs, sc_s.camera, total_spp, depth, seed=11, rank=rank, world=world)
        if rank == 0:
            sharded = t.cpu().numpy().reshape(540, 960, 4).copy()
            rs.reset_accum(); rs.render(sc_s.camera, total_spp, depth, seed=11)
            alone = ptd.accum_tensor(rs).cpu().numpy().reshape(540, 960, 4)
            d = np.abs(sharded[..., :3] - alone[..., :3])
            multi_check = {"what": f"960x540 x {total_spp} spp sharded over {world} ranks + one NCCL reduce, vs the same samples on rank 0 alone",
                           "samples_ok": bool((sharded[..., 3] == total_spp).all()),
                           "max_rel_diff": float((d / np.maximum(alone[..., :3], 1e-3)).max()),
                           "max_abs_diff_of_sums": float(d.max()),
                           "bitwise_equal_fraction": float((sharded[..., :3] == alone[..., :3]).mean())}
        rs.close()
        barrier()

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        segs = st["segments"]                      # rays traced in the last frame on this rank
        kern_ms = st["traverse_ms"] + st["shade_ms"]   # all k_bounce launches of the last frame (CUDA events per launch)
        # k_bounce launches of the last frame: per wave `depth` launches that trace + one that only shades, which is
        # not enqueued in a scene without emitters (every launch of the timed region minus the two bookkeeping kernels
        # per wave, per step; minus nothing else: the reduce is NCCL's)
        n_launch = max(int(round(launches / max(args.steps, 1))) - 2 * st["waves"], 1)
        hits = st["hit_records"]                   # segments that hit something: the only ones with a queue record
        survey_bytes = SURVEY_BYTES_PER_SEGMENT * segs + BYTES_PER_PATH * st["paths"]
        own_bytes = OWN_BYTES_PER_HIT_RECORD * hits + BYTES_PER_PATH * st["paths"]
        achieved = survey_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
        achieved_own = own_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
        # measured DRAM bytes per launch: the ncu capture is taken on a 16-spp wave (tools/profile_target.py; ncu replays
        # every launch ~40 times), the step here is one 64-spp wave of the same kernels, so the captured bytes of the
        # wave are scaled by the ratio of the wave sizes and divided by this step's launches
        traffic, tj = None, {}
        tp = os.path.join(ROOT, "profiles", "traverse_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("dram_bytes_per_wave") and tj.get("wave_paths"):
                traffic = tj["dram_bytes_per_wave"] * (st["paths"] / tj["wave_paths"]) / n_launch
            else:
                traffic = tj.get("dram_bytes_per_launch")
        dram_gbs = traffic * n_launch / (kern_ms * 1e-3) / 1e9 if (traffic and kern_ms > 0) else None
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(WORKLOAD, parallelism=f"spp-sharded x{world}, one NCCL reduce per frame" if world > 1 else "single GPU",
                           spp_total=spp * world, wave_paths=st["wave_paths"]),
            "roofline": {"bound": "hbm", "kernel": "k_bounce (shade previous hits + trace the segment, fused)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "byte_model": "SURVEY.md 8(d): 144 B per path segment + 16 B per path (a wavefront design with ray queues in HBM)",
                         "traffic_source": "ncu --set full of a 16-spp wave (profiles/traverse_traffic.json), scaled to this step's paths, per launch",
                         "algorithmic_bytes_per_launch": survey_bytes / n_launch, "avg_launch_ms": kern_ms / n_launch,
                         "launches_per_step": n_launch,
                         "share_of_step": kern_ms / st["render_ms"] if st["render_ms"] > 0 else None,
                         "segments_per_step": segs,
                         "dram_gbs_measured": dram_gbs, "dram_frac_measured": (dram_gbs / peak) if dram_gbs else None,
                         "note": "bound by instruction issue x SIMD efficiency (IPC 2.5-3.2 of 4, 17-30 of 32 lanes active, "
                                 "long-scoreboard on L1/L2 node fetches; profiles/), not by HBM: the kernel moves a quarter of "
                                 "the bytes this model budgets (see roofline_own_bytes) because rays stay in registers"},
            "roofline_own_bytes": {"bound": "hbm", "achieved": achieved_own, "peak": peak, "unit": "GB/s",
                                   "frac": (achieved_own / peak) if achieved_own else None,
                                   "byte_model": "this kernel's own traffic: 96 B per segment that hits (one 48-byte hit record written + read) + 16 B per path",
                                   "hit_records_per_step": hits, "bytes_per_launch": own_bytes / n_launch,
                                   "measured_over_model": (traffic * n_launch / own_bytes) if traffic else None},
            "path": {"mrays_per_s": segs / (st["render_ms"] * 1e-3) / 1e6 if st["render_ms"] > 0 else None,
                     "segments_per_path": segs / st["paths"], "render_ms_last_frame": st["render_ms"],
                     "bounce_kernel_ms": kern_ms, "entry_grid_levels": st["entry_levels"], "entry_grid_cells": st["entry_cells"]},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "what": "pt_update_prim + pt_build (upload, LBVH) + pt_render_frame_async + pt_read_accum_begin / _end into pinned host memory (the copy of one frame runs under the next frame's build and render; all frames are on the host before the clock stops)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if multi_check is not None:
            line["multi_gpu_check"] = multi_check
        line["roofline_fp32"] = fp32_view(r, cam, depth, line["path"]["mrays_per_s"], (clocks or {}).get("sm_mhz"))
        if tj.get("ipc_time_weighted"):
            # third view, from the same ncu capture: what the kernel is actually bound by.  4 warp-instructions can
            # issue per SM per clock; the capture says how many do and how many of an instruction's 32 lanes are alive.
            line["roofline_issue"] = {"bound": "warp-instruction issue", "achieved": tj["ipc_time_weighted"], "peak": 4.0,
                                      "unit": "warp-instructions / clock / SM", "frac": tj["ipc_time_weighted"] / 4.0,
                                      "active_lanes_of_32": tj["lanes_per_instruction_time_weighted"],
                                      "source": "ncu --set full of a 16-spp wave (profiles/traverse_traffic.json), time-weighted over its k_bounce launches"}
        if world == 1 and not args.no_baselines:
            line["variants"] = {"libpt_b200_fma.so (nvcc-contracted arithmetic of the reference's CUDA build)": fma_variant_leg(sc, sky)}
            line["cpu_baseline"] = cpu_baseline_leg(sc, sky)
            rc = ref_cuda_leg(sc, sky)
            if rc is not None:
                line["ref_cuda_baseline"] = rc
            try:
                line["graze_c2"] = graze_leg(r, sc)
            except Exception as e:
                line["graze_c2"] = {"unavailable": str(e)[:200]}
    r.close()

    # ---- the other BASELINE.json configs ---------------------------------------------------------------------
    if not args.no_configs:
        configs = {}
        graze = world == 1 and not args.no_baselines
        barrier()
        if rank == 0:
            # C1 and C3 on ONE GPU (this rank's); the other ranks wait at the barrier below
            c1 = scenes.demo_scene(640, 360)
            configs["C1"] = config_single_gpu("C1", "reference demo scene (floor + 74 spheres / cylinders, 4 materials, sky), 640x360, 16 spp, depth 8",
                                              c1, sky, 16, 20, depth, stream, graze)
            c3 = scenes.stress_field(100000, 1920, 1080)
            configs["C3"] = config_single_gpu("C3", "BVH stress: floor + 100 k spheres / cylinders (128 836 primitives with the hollow shells), glass-heavy, 1920x1080, 32 spp, depth 8",
                                              c3, sky, 32, 3, depth, stream, graze)
            del c3
        barrier()
        # C4: 3840x2160, a FIXED 1024 spp split over the ranks (strong scaling), one reduce of the 133 MB SUM buffers
        sc4 = scenes.sphere_field(1000, 3840, 2160)
        r4 = api.Renderer(sc4, sky)
        r4.set_stream(stream)
        ptd.render_sharded(r4, sc4.camera, 16 * world, depth, seed=5, rank=rank, world=world)      # warm-up: allocations, NCCL
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ptd.render_sharded(r4, sc4.camera, 1024, depth, seed=5, rank=rank, world=world, sync=False)
        e1.record()
        torch.cuda.synchronize()
        ms4 = max_over_ranks(e0.elapsed_time(e1))
        r4.synchronize()
        st4 = r4.stats()
        if rank == 0:
            configs["C4"] = {"workload": "C2 scene at 3840x2160, 1024 spp progressive (16-spp waves), depth 8; the sample range is split over the ranks, one NCCL reduce of the SUM buffers at the end",
                             "mpaths_per_s": 3840 * 2160 * 1024 / ms4 / 1e3, "ms": ms4, "n_gpus": world, "scaling": "strong",
                             "spp_total": 1024, "spp_per_rank": ptd.shard_samples(1024, world, 0)[1], "waves_per_rank": st4["waves"],
                             "segments_per_path": st4["segments"] / st4["paths"], "timing": "CUDA events, max over ranks, reduce inside"}
        r4.close()
        barrier()
        # C5: 120 frames at 1080p x 64 spp, camera orbit, 10 % of the spheres move: every frame is pt_update_prim x 100 +
        # pt_build (upload + LBVH on the device) + render + read-back of the mean image; frame f belongs to rank f % N
        r5 = api.Renderer(sc, sky)
        r5.set_stream(stream)
        n_frames = 120
        cams = video.camera_path(cam, n_frames)
        mine = ptd.frames_of_rank(n_frames, world, rank)
        sc5 = copy.deepcopy(sc)

        def frame(f, k, first_of_run):
            ids, centres = video.moving_spheres(sc5, f)
            for p, c in zip(ids, centres):
                r5.update_prim(int(p), c, float(sc5.radius[p]))
            r5.build()
            r5.reset_accum()
            r5.render(cams[f], spp, depth, seed=2026, sync=False)
            r5.read_accum_begin(host_imgs[k & 1])
            if not first_of_run:
                r5.read_accum_end()

        frame(mine[0] if mine else 0, 0, True)
        r5.read_accum_end()
        barrier()
        t0 = time.perf_counter()
        for k, f in enumerate(mine):
            frame(f, k, k == 0)
        if mine:
            r5.read_accum_end()
        torch.cuda.synchronize()
        wall5 = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            configs["C5"] = {"workload": "MotionalCamera video: 120 frames, 1920x1080 x 64 spp, depth 8, camera orbit + 10 % of the spheres moving, per-frame upload + LBVH rebuild, frames round-robin over the ranks, every frame read back to pinned host memory (pipelined: the copy of a frame runs under the next one)",
                             "mpaths_per_s": n_frames * cam.width * cam.height * spp / wall5 / 1e6, "wall_s": wall5, "n_gpus": world,
                             "frames": n_frames, "frames_per_s": n_frames / wall5, "timing": "wall clock, max over ranks"}
        r5.close()
        if rank == 0:
            line["configs"] = configs
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
