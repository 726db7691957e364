#!/usr/bin/env python
"""Benchmark of the ray-propagation hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- the default full-colour RGB design of
couplers_coor_full_color() at the runner's sizes (gpu_ray_tracing_pro_fullColor.py:16-17, 60-61):
100 x 75 FoV cells x 3 wavelengths x 5000 rays per cell = 112.5 M rays per launch, bins
3 x 75 x 100 x 80 x 120 float32 (864 MB), synthetic complex128 LUTs of the real shapes (the real
LUT files are not reachable).  One "step" = one launch of the walk over the whole ray set, which is
what the runner repeats num_iter = 4 times with continuing RNG streams (RUN:169-177).

Printed JSON line (rank 0):
  value    ray-bounces / s, all ranks, inputs resident in HBM, CUDA events on the launch stream
           (a bounce = one position advance x += gap; counted exactly for the
           timed launches by replaying them with device counters from the saved RNG states)
  e2e      the same metric through the C ABI host entry wgrt_trace_fullcolor_host with pinned HOST
           buffers, copies inside the timed region, every step.  The headline uses the runner
           layout (the engine derives the rays from the start points: uploads LUTs + geometry +
           points, downloads the bins); e2e_dropin is the literal 33-array call (uploads the 12
           materialised ray arrays too), e2e_job does the runner's K launches in one call, and
           e2e_eval adds the evaluation's pupil sums with the bins never leaving the device
           ("full-colour eval wall time").  At N > 1: per rank upload + walk into device bins, ONE
           NCCL reduce-scatter, D2H of the rank's 1/N slice of the reduced bins
  reference_numba_cuda   the reference's OWN kernel (Numba -> PTX built from /root/reference in the build
           container, oracle/_ref) on the same device-resident inputs: the second baseline of
           BASELINE.json, and the full-size parity check (bins and RNG states bit-equal)
  roofline dominant kernel (walk_warp_kernel) against the FP64 FMA peak measured live on this GPU
           (the path is FP64-issue bound, not HBM bound: SURVEY.md section 8d); roofline_hbm gives
           the algorithmic-bytes view against MEASURED_PEAKS.json
  cpu_baseline   the CPU oracle port (oracle/) on the host cores, bounded sample (rank 0, N = 1)
Multi-GPU (torchrun): weak scaling -- every rank walks the full workload with its own RNG streams
(more Monte-Carlo samples per FoV), then ONE NCCL all-reduce of the bin tensor inside the timed
region.  --workload c3_dense_fov_41x41x3x10000 is the PARTITIONED (strong-scaling) arm on BASELINE
configs[2].  --impl reference times the CPU oracle port on all host threads (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray_bounces_per_s"
UNIT = "ray-bounces/s"
WORKLOADS = {
    # name: (num_FOV_x, num_FOV_y, rays_per_FoV, (EBy, EBx))
    "c2_default_fullcolor_100x75x3x5000": (100, 75, 5000, (80, 120)),
    "small_20x15x3x2000": (20, 15, 2000, (80, 120)),
    # BASELINE.json configs[2]: dense FoV sweep, PARTITIONED over the ranks (strong scaling): contiguous
    # cell ranges per rank, one NCCL all-reduce of the bins, result checked against the 1-GPU job
    "c3_dense_fov_41x41x3x10000": (41, 41, 10000, (80, 120)),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("WGRT_BENCH_WORKLOAD", "c2_default_fullcolor_100x75x3x5000"),
                    choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-partitioned", action="store_true", help="skip the partitioned config-3 sub-record")
    ap.add_argument("--no-legacy", action="store_true", help="skip the legacy energy-splitting tracer sub-record")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the bounded baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t_enter, self.t_exit = index, [], None, 0.0, float("inf")

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()                     # nvidia-smi needs ~0.1 s to start: the timed region is shorter than that
            while not self.rows and time.time() - t0 < 2.0:
                time.sleep(0.005)
            self.t_enter = time.time()           # rows that arrived before now were sampled before the timed region
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def __exit__(self, *exc):
        self.t_exit = time.time()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        # samples taken DURING the timed region (arrival time within it, 25 ms slack for the pipe); a region
        # shorter than the 20 ms sampling period may hold none: then the first sample after it started
        rows = [r for r in self.rows if self.t_enter <= r[-1] <= self.t_exit + 0.025]
        if not rows:
            rows = [r for r in self.rows if r[-1] >= self.t_enter][:1] or self.rows[-1:]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_flops(c):
    """F = 40 N_E + 30 N_draw2 + 40 N_draw3 + 8 N_straddle + 9 N_cross + 6 N_bounce."""
    return (40 * c["efield"] + 30 * c["draw2"] + 40 * c["draw3"] + 8 * c["straddle"] + 9 * c["cross"]
            + 6 * c["bounces"])


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def pinned_like(a: np.ndarray):
    """A pinned host copy of ``a`` (as a torch tensor and a NumPy view of the same memory)."""
    import torch
    v = a.view(np.float64) if a.dtype == np.complex128 else a
    v = v.view(np.int32) if v.dtype == np.uint32 else v
    t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)


# ------------------------------------------------------------------------------------------------
def sample_scene(scene, target_rays: int):
    """Bounded sample of the workload for CPU timing: every stride-th cell, whole cells."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    rpc = scene.meta["num_rays_per_FoV"]
    n_cells = scene.rays.num_rays // rpc
    stride = max(1, int(round(n_cells * rpc / max(target_rays, 1))))
    idx = np.arange(0, n_cells, stride)
    sel = (idx[:, None] * rpc + np.arange(rpc)[None, :]).ravel()
    rays = scene.rays
    sub = si.RaySet(*(a[sel].copy() for a in rays.arrays()), rays.rng_states[sel].copy())
    return sub, f"every {stride}th FoV-wavelength cell of the workload ({len(idx)} cells x {rpc} rays = {len(sel)} rays), one launch"


def run_cpu_oracle(scene, sub, threads: int):
    """Time one launch of the oracle port over the sample. Returns (seconds, counters)."""
    from oracle import oracle
    EB = scene.new_matrix_EB()
    rng = sub.rng_states.copy()
    full = scene.rays
    scene.rays = sub
    try:
        args = scene.kernel_args(EB, rng)
        t0 = time.perf_counter()
        cnt = oracle.trace(*args, num_threads=threads, counters=True)
        dt = time.perf_counter() - t0
    finally:
        scene.rays = full
    return dt, cnt


def reference_arm(args, nx, ny, rpc, eb):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the
    reference itself is Python/Numba and cannot travel to this box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    from oracle import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    scene = si.make_scene(nx, ny, rpc, eb=eb, seed=2024)
    # size the per-step sample so that (steps + warmup) launches stay within a few minutes
    probe, _ = sample_scene(scene, 200_000)
    dt, cnt = run_cpu_oracle(scene, probe, threads)
    rate = cnt["rays"] / dt
    budget_s = min(args.cpu_seconds, 150.0 / max(args.steps + args.warmup, 1))
    sub, desc = sample_scene(scene, int(rate * budget_s))
    for _ in range(args.warmup):
        run_cpu_oracle(scene, sub, threads)
    tot_t, tot_b, tot_r = 0.0, 0, 0
    for _ in range(args.steps):
        dt, cnt = run_cpu_oracle(scene, sub, threads)
        tot_t += dt; tot_b += cnt["bounces"]; tot_r += cnt["rays"]
    value = tot_b / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": desc},
        "rays_per_s": tot_r / tot_t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def partitioned_record(args, nx, ny, rpc, eb, world, rank, local_rank, workload, sample_clocks=True):
    """Strong scaling on BASELINE.json configs[2]: the job's FoV-wavelength cells are split into contiguous
    ranges (multi_gpu.cell_range) and every rank walks its range with the runner layout and the GLOBAL RNG
    seeds (RUN:158).  A cell's deposits land in its own tile of matrix_EB, so the ranks' outputs are disjoint:
    NO collective is needed -- `value` times K launches of every rank's range, device
    resident; `e2e` is multi_gpu.trace_partitioned from pinned host memory to pinned host memory (each rank
    uploads only its table columns, downloads only its matrix_EB columns).  Outside the timed regions the
    ranks' bins are summed onto rank 0 and compared with the same job walked on rank 0's GPU alone: they
    must be bit-equal.  Returns the record (rank 0) or None."""
    import contextlib
    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, multi_gpu, synthetic_inputs as si
    scene = si.make_scene(nx, ny, 2, eb=eb, seed=2024)             # tables; the rays are implicit
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 2025)
    L = scene.eb_shape[0]
    n_cells = nx * ny * L
    c0, c1 = multi_gpu.cell_range(n_cells, world, rank)
    stream = torch.cuda.current_stream()
    steps, warm = args.steps, max(args.warmup, 3)

    def to_dev(a):
        v = a.view(np.float64) if a.dtype == np.complex128 else a
        t = torch.from_numpy(np.ascontiguousarray(v.view(np.int32) if v.dtype == np.uint32 else v)).cuda()
        return GRTF._TorchAlias(t, a.shape, a.dtype)

    g = {k: to_dev(v) for k, v in scene.geom.items()}
    lt = {k: to_dev(v) for k, v in scene.luts.items()}
    d_px, d_py = to_dev(pts[:, 0].astype(np.float32)), to_dev(pts[:, 1].astype(np.float32))

    def job(first_cell, cells, counters=False):
        n = cells * rpc
        rng = to_dev(si.initial_rng_states(n, offset=first_cell * rpc))
        ebt = to_dev(scene.new_matrix_EB())
        k = GRTF.process_rays_kernel_pro_fullColor.configured(counters=counters, ray_index_base=first_cell * rpc)
        k = k.runner_layout(rpc // 2, n, first_cell)
        a = [d_px, d_py] + [None] * 10 + [rng, g["IC"], g["FC"], g["FC_offset"], g["OC"], g["OC_offset"], scene.n_g,
                                           g["eff_reg1"], g["eff_reg2"], g["eff_reg_FOV"], g["eff_reg_FOV_range"],
                                           lt["lut_ic1"], lt["lut_ic2"], lt["lut_ic3"], lt["lut_fc1"], lt["lut_fc2"],
                                           lt["lut_oc1"], lt["lut_oc2"], g["lut_TIR"], g["lut_gap"], ebt]
        return k[1, 256, stream], a, rng, ebt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launch, a, rng, ebt = job(c0, c1 - c0)
    rng0 = rng._t.clone()
    for _ in range(warm):
        launch(*a)
    torch.cuda.synchronize()
    rng._t.copy_(rng0); ebt._t.zero_()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the clock sampler (rank 0 only) starts BEFORE the barrier: its start-up must not skew the ranks
    with (ClockSampler(local_rank) if (rank == 0 and sample_clocks) else contextlib.nullcontext()) as clocks:
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            launch(*a)
        ev1.record(stream)
        barrier()
    mine_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([mine_ms], device="cuda", dtype=torch.float64)
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank = [t.clone()]
    ms = float(t.item())
    # bounces of the timed launches: replay with counters
    claunch, ca, crng, cebt = job(c0, c1 - c0, counters=True)
    _capi.reset_counters()
    for _ in range(steps):
        claunch(*ca)
    cnt = _capi.read_counters()
    bt = torch.tensor([cnt["bounces"], cnt["rays"], cnt["near_tie"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bt)
    del crng, cebt, claunch, ca

    # ---- end to end: pinned host -> pinned host through multi_gpu.trace_partitioned, no collective -----
    pin_keep, geom_p, luts_p = [], {}, {}
    for src, dst in ((scene.geom, geom_p), (scene.luts, luts_p)):
        for k, v in src.items():
            tpin, view = pinned_like(v)
            pin_keep.append(tpin); dst[k] = view
    tpin, eb_view = pinned_like(scene.new_matrix_EB())
    pin_keep.append(tpin)
    m0 = m1 = 0
    for _ in range(2):
        _, (m0, m1) = multi_gpu.trace_partitioned(pts, geom_p, scene.n_g, luts_p, rpc, world, rank, num_iter=1,
                                                  eb=eb, matrix_EB=eb_view)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        multi_gpu.trace_partitioned(pts, geom_p, scene.n_g, luts_p, rpc, world, rank, num_iter=1, eb=eb,
                                    matrix_EB=eb_view)
    tw = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    cols = m1 - m0
    frac_cols = cols / nx
    h2d = sum(v.nbytes for k, v in list(geom_p.items()) + list(luts_p.items())
              if k.startswith("lut_") or k.startswith("eff_reg_FOV")) * frac_cols + pts.shape[0] * 8
    d2h = eb_view.nbytes * frac_cols
    # one launch of this rank's range from the same seeds on the device == the columns that came back
    l1, a1, r1, e1 = job(c0, c1 - c0)
    l1(*a1)
    torch.cuda.synchronize()
    own = e1._t.cpu().numpy().reshape(scene.eb_shape)
    e2e_same = bool(np.array_equal(own[:, :, m0:m1], eb_view[:, :, m0:m1]))
    io = torch.tensor([h2d, d2h, 1.0 if e2e_same else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(io)
    del l1, a1, r1, e1, own, pin_keep, geom_p, luts_p, eb_view

    # ---- correctness, outside the timed regions: sum of the ranks' (disjoint) bins == the 1-GPU job ----
    total = ebt._t.clone()
    if world > 1:
        dist.reduce(total, dst=0)
    rec = None
    if rank == 0:
        flaunch, fa, frng, febt = job(0, n_cells)
        for _ in range(steps):
            flaunch(*fa)
        torch.cuda.synchronize()
        same = bool(torch.equal(febt._t, total))
        bounces_per_launch = float(bt[0].item()) / max(steps, 1)
        rec = {"metric": METRIC, "value": float(bt[0].item()) / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
               "steps": steps, "warmup": warm, "ms_per_step": ms / max(steps, 1),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "num_FOV_x": nx, "num_FOV_y": ny, "wavelengths": L,
                          "rays_per_FoV": rpc, "rays_per_launch_all_gpus": n_cells * rpc, "eyebox_bins": list(eb),
                          "partition": f"contiguous FoV-wavelength cell ranges, {n_cells} cells over {world} ranks, global RNG "
                                       "seeds; the ranks' bins are disjoint: no collective (each rank keeps / downloads "
                                       "the FoV-x columns of its cells)",
                          "l2_policy": "runner layout: per-launch inputs are the RNG states (2 GB over all ranks) > L2"},
               "rays_per_s": float(bt[1].item()) / (ms * 1e-3), "full_colour_wall_ms": ms,
               "per_rank_walk_ms_per_step": [float(x.item()) / max(steps, 1) for x in per_rank],
               "collective_ms": 0.0, "near_tie_rays": int(bt[2].item()),
               "deposits": float(total.sum(dtype=torch.float64).item()),
               "summed_bins_bit_equal_to_single_gpu_job": same,
               "e2e": {"value": bounces_per_launch * steps / wall, "unit": UNIT, "ms_per_step": wall / max(steps, 1) * 1e3,
                       "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
                       "api": "multi_gpu.trace_partitioned -> wgrt_trace_fullcolor_host (runner layout, WGRT_FLAG_BINS_COLUMNS): "
                              "per rank H2D of its table columns, walk of its cell range, D2H of its matrix_EB columns; "
                              "one launch per call, pinned host buffers, no collective",
                       "columns_bit_equal_to_device_launch": bool(io[2].item() == world)},
               "gpu_launches": steps * 17}
        if clocks is not None:
            rec["clocks"] = clocks.summary()
    del total
    torch.cuda.empty_cache()
    return rec


def partitioned_arm(args, nx, ny, rpc, eb):
    """--workload c3_...: the partitioned arm alone, printed as the run's JSON line."""
    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _capi.load_library()
    rec = partitioned_record(args, nx, ny, rpc, eb, world, rank, local_rank, args.workload)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    nx, ny, rpc, eb = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, nx, ny, rpc, eb)
        return
    if args.workload.startswith("c3_"):
        partitioned_arm(args, nx, ny, rpc, eb)
        return

    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, multi_gpu, synthetic_inputs as si

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _capi.load_library()

    # ---- inputs: identical scene on every rank; rank-specific RNG streams -----------------------
    scene = si.make_scene(nx, ny, rpc, eb=eb, seed=2024)
    N = scene.rays.num_rays
    scene.rays.rng_states = si.initial_rng_states(N, offset=rank * N)
    host_args = list(scene.kernel_args(scene.new_matrix_EB()))

    def to_dev(a):
        if not isinstance(a, np.ndarray):
            return a
        v = a.view(np.float64) if a.dtype == np.complex128 else a
        t = torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda()
        return GRTF._TorchAlias(t, a.shape, a.dtype)

    dev_args = [None if i in (2, 3, 4, 5) else to_dev(a) for i, a in enumerate(host_args)]
    rng_t, eb_t = dev_args[12]._t, dev_args[32]._t
    stream = torch.cuda.current_stream()
    kern = GRTF.process_rays_kernel_pro_fullColor
    launch = kern[(N + 255) // 256, 256, stream]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import contextlib
    reducer = multi_gpu.BinReducer(eb_t.numel(), eb_t.device) if world > 1 else None
    for _ in range(max(args.warmup, 3)):
        launch(*dev_args)
    if world > 1:
        for _ in range(3):   # NCCL warm-up at the timed message size (the first uses set up channels and pick the algorithm)
            reducer.reduce_scatter(eb_t)
    torch.cuda.synchronize()
    eb_t.zero_()
    rng_saved = rng_t.clone()

    # ---- timed region: K launches (+ ONE reduce-scatter of the bins when N > 1) -------------------
    # CUDA events on the launch stream; the clock sampler runs on rank 0 only and starts BEFORE the barrier,
    # so that its start-up (~0.1 s of nvidia-smi) cannot skew the ranks against each other -- in round 1 every
    # rank started its own sampler between the barrier and its first launch, and the skew (6 ms at N = 2,
    # 52 ms at N = 8) was charged to the collective that waits for the slowest rank.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 3)]
    part = None
    with (ClockSampler(local_rank) if rank == 0 else contextlib.nullcontext()) as clocks:
        barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            launch(*dev_args)
            ev[k + 1].record(stream)
        if world > 1:
            part = reducer.reduce_scatter(eb_t)   # as uint8 when that is exact (multi_gpu.BinReducer); no host sync
        ev[args.steps + 1].record(stream)
        barrier()
    rng_final = rng_t.clone()
    total_ms = ev[0].elapsed_time(ev[args.steps + 1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    collective_ms = ev[args.steps].elapsed_time(ev[args.steps + 1]) if world > 1 else 0.0
    t = torch.tensor([total_ms, sum(step_ms), collective_ms], device="cuda", dtype=torch.float64)
    per_rank = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
    else:
        per_rank = [t]
    total_ms_max = max(float(x[0].item()) for x in per_rank)
    narrow_exact = None
    if world > 1:
        narrow_exact = reducer.narrow_ok()
        if not narrow_exact:                      # some count exceeded 255 // world: the float32 form (untimed here)
            part = reducer.reduce_scatter(eb_t, wide=True)
        dsum = part.sum(dtype=torch.float64).reshape(1)
        dist.all_reduce(dsum)
        deposits_total = float(dsum.item())
    else:
        deposits_total = float(eb_t.sum(dtype=torch.float64).item())

    # ---- exact event counts of the timed launches: replay them with device counters -------------
    rng_t.copy_(rng_saved)
    scratch_eb = torch.zeros_like(eb_t)
    count_args = list(dev_args)
    count_args[32] = GRTF._TorchAlias(scratch_eb, host_args[32].shape, np.float32)
    _capi.reset_counters()
    counted = kern.configured(counters=True)[1, 256, stream]
    for _ in range(args.steps):
        counted(*count_args)
    cnt = _capi.read_counters()
    del scratch_eb
    bt = torch.tensor([cnt["bounces"], cnt["rays"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bt)
    bounces_all, rays_all = float(bt[0].item()), float(bt[1].item())
    value = bounces_all / (total_ms_max * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "num_FOV_x": nx, "num_FOV_y": ny, "wavelengths": 3,
                   "rays_per_FoV": rpc, "rays_per_launch_per_gpu": N, "eyebox_bins": list(eb),
                   "l2_policy": "inputs (4.1 GB of ray state per launch) exceed the 126 MB L2; no flush needed",
                   "partition": "replicated design, rank-specific RNG streams, one NCCL reduce-scatter of the bins (as uint8 "
                                "when exact): every rank ends with 1/N of the summed tensor" if world > 1 else "single GPU"},
        "rays_per_s": rays_all / (total_ms_max * 1e-3),
        "full_colour_wall_ms": total_ms_max,
        "bounces_per_ray": bounces_all / max(rays_all, 1),
        "deposits": deposits_total,
        "step_ms_rank0": step_ms,
        "per_rank_ms": {"walk_total": [float(x[1].item()) for x in per_rank],
                        "collective": [float(x[2].item()) for x in per_rank],
                        "timed_region": [float(x[0].item()) for x in per_rank]},
        "collective_ms": max(float(x[2].item()) for x in per_rank),
        "collective": None if world == 1 else
        ("one NCCL reduce-scatter of the bins as uint8 (exact: every count <= 255 // N), 216 MB per rank in"
         if narrow_exact else "narrow form not exact for these counts: float32 reduce-scatter needed (untimed re-run)"),
        # per launch (profiles/r2_launch_shares_final.txt): geometry hash, 6 region-index / atlas kernels and 7 zone-table
        # kernels (no-ops when the geometry is unchanged), tile pick, walk, near-tie redo = 17 kernels
        "gpu_launches": args.steps * 17 + (3 if world > 1 else 0),
        "near_tie_rays": cnt["near_tie"],
        "clocks": clocks.summary() if clocks is not None else None,
    }

    if rank == 0:
        # ---- roofline of the dominant kernel (walk_warp_kernel) ---------------------------------
        # literal-algorithm work per launch (straddling edges / cross products only exist in the
        # literal scan): one strict launch with counters from the same RNG states
        rng_t.copy_(rng_saved)
        scratch_eb = torch.zeros_like(eb_t)
        count_args[32] = GRTF._TorchAlias(scratch_eb, host_args[32].shape, np.float32)
        _capi.reset_counters()
        kern.configured(strict=True, counters=True)[1, 256, stream](*count_args)
        cs = _capi.read_counters()
        del scratch_eb
        flops_per_launch = algorithmic_flops(cs)
        p64, p32 = C.c_double(), C.c_double()
        _capi.check(lib.wgrt_debug_fma_peak(C.byref(p64), C.byref(p32)), lib)
        walk_ms = float(np.mean(step_ms))         # region-index + tile-pick kernels are < 0.1 % of it
        ach = flops_per_launch / (walk_ms * 1e-3) / 1e12
        line["roofline"] = {
            "kernel": "walk_warp_kernel", "bound": "fp64", "achieved": ach, "peak": p64.value, "unit": "TFLOP/s",
            "frac": ach / p64.value, "traffic": load_traffic(),
            "peak_source": "DFMA chain micro-benchmark run live in this process (wgrt_debug_fma_peak)",
            "algorithmic_flops_per_ray": flops_per_launch / max(cs["rays"], 1),
            "fp32_peak_tflops": p32.value,
        }
        peaks, src = measured_peaks()
        cells = nx * ny * 3
        alg_bytes = 40.0 * N + cells * 93e3            # 40 B/ray + ~93 KB per cell
        gbs = alg_bytes / (walk_ms * 1e-3) / 1e9
        line["roofline_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": gbs / peaks["hbm_gbs"], "peak_source": src,
                                "algorithmic_bytes_per_launch": alg_bytes}
        rng_t.copy_(rng_saved)

    # ---- the reference's own Numba-CUDA kernel on this GPU (rank 0, N = 1 only) -----------------
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        line["reference_numba_cuda"] = reference_gpu_leg(args, dev_args, host_args, rng_saved, rng_final, eb_t,
                                                         N, stream, value, bounces_all)
    del rng_final

    # ---- end to end through the C ABI with pinned host buffers -----------------------------------
    if not args.no_e2e and world > 1:
        del dev_args, count_args
        torch.cuda.empty_cache()
        line["e2e"] = e2e_multi_gpu(args, scene, rpc, N, world, rank, stream)
    if not args.no_e2e and world == 1:
        from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner
        del dev_args, count_args
        torch.cuda.empty_cache()

        def wall_max(seconds):
            tw = torch.tensor([seconds], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            return float(tw.item())

        def all_sum(v):
            tv = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tv)
            return float(tv.item())

        # (1) e2e -- the runner-level entry: what the reference runner does
        # between having its inputs and having its bins (RUN:59-185), with the ray set derived on the
        # device from the start points.  Every step is one complete job of ONE launch: upload LUTs,
        # geometry and start points from pinned host memory, seed the RNG as RUN:158, walk, download the
        # bins.  (Steps are independent jobs, so each uploads and downloads everything it needs.)
        pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 2024 + 1)
        pin_keep, geom_p, luts_p = [], {}, {}
        for src, dst in ((scene.geom, geom_p), (scene.luts, luts_p)):
            for k, a in src.items():
                tpin, view = pinned_like(a)
                pin_keep.append(tpin); dst[k] = view
        tpin, eb_view = pinned_like(scene.new_matrix_EB())
        pin_keep.append(tpin)
        first_cell = 0
        h2d_r = sum(a.nbytes for a in list(geom_p.values()) + list(luts_p.values())) + pts.shape[0] * 8
        d2h_r = eb_view.nbytes
        for _ in range(2):
            runner.trace_full_color(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=1, matrix_EB=eb_view,
                                    bins_start_zero=True)
        barrier()
        parts = np.zeros(3)
        tms_r = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            runner.trace_full_color(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=1, matrix_EB=eb_view,
                                    bins_start_zero=True, timings=tms_r)
            parts += np.array(tms_r)
        wall_r = wall_max(time.perf_counter() - t0)
        # bounces of that job: replay launch #1 from the runner's seeds on the device with counters
        _capi.reset_counters()
        ck = kern.configured(counters=True).runner_layout(rpc // 2, N)
        d_geom = {k: to_dev(v) for k, v in scene.geom.items()}
        d_luts = {k: to_dev(v) for k, v in scene.luts.items()}
        d_rng = to_dev(si.initial_rng_states(N))
        d_eb = to_dev(scene.new_matrix_EB())
        d_px, d_py = to_dev(pts[:, 0].astype(np.float32)), to_dev(pts[:, 1].astype(np.float32))
        rargs = [d_px, d_py, None, None, None, None, None, None, None, None, None, None, d_rng,
                 d_geom["IC"], d_geom["FC"], d_geom["FC_offset"], d_geom["OC"], d_geom["OC_offset"], scene.n_g,
                 d_geom["eff_reg1"], d_geom["eff_reg2"], d_geom["eff_reg_FOV"], d_geom["eff_reg_FOV_range"],
                 d_luts["lut_ic1"], d_luts["lut_ic2"], d_luts["lut_ic3"], d_luts["lut_fc1"], d_luts["lut_fc2"],
                 d_luts["lut_oc1"], d_luts["lut_oc2"], d_geom["lut_TIR"], d_geom["lut_gap"], d_eb]
        ck[1, 256, stream](*rargs)
        c1 = _capi.read_counters()
        same1 = bool(np.array_equal(d_eb._t.cpu().numpy(), eb_view))
        line["e2e"] = {
            "value": all_sum(c1["bounces"]) * args.steps / wall_r, "unit": UNIT,
            "h2d_bytes_per_step": int(h2d_r), "d2h_bytes_per_step": int(d2h_r),
            "ms_per_step": wall_r / max(args.steps, 1) * 1e3,
            "stream_spans_ms_per_step_rank0": {"h2d": parts[0] / args.steps, "trace": parts[1] / args.steps,
                                               "d2h": parts[2] / args.steps,
                                               "note": "the three streams run as a pipeline over FoV-x column chunks: spans overlap"},
            "api": "runner.trace_full_color -> wgrt_trace_fullcolor_host, runner layout (start points instead of "
                   "the 12 materialised ray arrays), one launch per call, pinned host buffers",
            "bins_bit_equal_to_device_launch": same1}

        # (2) e2e_job -- the same entry doing the runner's whole job in one call: K launches with
        # continuing RNG streams (RUN:169-177), RNG states returned to the host as well
        seeds_t, seeds = pinned_like(si.initial_rng_states(N))
        pin_keep.append(seeds_t)
        tms_j = []
        barrier()
        t0 = time.perf_counter()
        runner.trace_full_color(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=args.steps, matrix_EB=eb_view,
                                bins_start_zero=True, rng_states=seeds, timings=tms_j)
        wall_j = wall_max(time.perf_counter() - t0)
        for _ in range(args.steps - 1):
            ck[1, 256, stream](*rargs)
        cj = _capi.read_counters()
        samej = bool(np.array_equal(d_eb._t.cpu().numpy(), eb_view)) and \
            bool(np.array_equal(d_rng._t.cpu().numpy().view(np.uint32), seeds))
        line["e2e_job"] = {
            "value": all_sum(cj["bounces"]) / wall_j, "unit": UNIT, "ms_per_step": wall_j / max(args.steps, 1) * 1e3,
            "launches_per_call": args.steps, "h2d_bytes_per_call": int(h2d_r + seeds.nbytes),
            "d2h_bytes_per_call": int(d2h_r + seeds.nbytes),
            "stream_spans_ms_per_call_rank0": {"h2d": tms_j[0], "trace": tms_j[1], "d2h": tms_j[2]},
            "bins_and_rng_bit_equal_to_device_launches": samej}
        # (2b) e2e_eval -- "full-colour eval wall time": the runner from its inputs to its evaluation results
        # (RUN:59-198) in one call: K launches, the pupil-mask sums and per-cell totals (EVAL:68-109, RUN:186) AND the
        # rest of evaluation() (EVAL:110-160: display model, Lab / CIEDE2000, luminance statistics) on the device; the
        # bin tensor never leaves it and 8 doubles per eye position + the per-cell totals come back.  The host only
        # averages those over the 56 eye positions (`host_evaluation_ms`).  For comparison the same call with the
        # round-1 host-finished evaluation (NumPy on the downloaded pupil sums) is timed too.
        from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import AR_system_evaluation_functions as EV
        runner.trace_and_evaluate(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=1)
        tms_e = []
        barrier()
        t0 = time.perf_counter()
        res = runner.trace_and_evaluate(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=args.steps, timings=tms_e)
        wall_e = wall_max(time.perf_counter() - t0)
        t0 = time.perf_counter()
        EV.finish_metrics(res["eval_metrics"], ny * nx, *(((eb[0] - 30) // 8 + 1), ((eb[1] - 30) // 12 + 1)))
        host_eval_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        res_h = runner.trace_and_evaluate(pts, geom_p, scene.n_g, luts_p, rpc, num_iter=args.steps, host_evaluation=True)
        wall_h = time.perf_counter() - t0
        cells_ref = eb_view.reshape(eb_view.shape[0], eb_view.shape[1], eb_view.shape[2], -1).sum(axis=-1, dtype=np.float64)
        rel = max(abs(res[k] - res_h[k]) / max(abs(res_h[k]), 1e-300) for k in ("delta_e", "U_fov", "U_EB"))
        line["e2e_eval"] = {
            "value": all_sum(cj["bounces"]) / wall_e, "unit": UNIT, "wall_ms_per_call": wall_e * 1e3,
            "launches_per_call": args.steps, "h2d_bytes_per_call": int(h2d_r),
            "d2h_bytes_per_call": int(res["eval_metrics"].nbytes + res["cell_sums"].nbytes),
            "stream_spans_ms_per_call_rank0": {"h2d": tms_e[0], "trace": tms_e[1], "d2h_and_reductions": tms_e[2]},
            "host_evaluation_ms": host_eval_s * 1e3,
            "maps": {"efficiency_per_colour": [float(v) for v in res["efficiency"]], "U_fov": float(res["U_fov"]),
                     "U_EB": float(res["U_EB"]), "delta_e_2000": float(res["delta_e"])},
            "host_finished_variant": {"wall_ms_per_call": wall_h * 1e3, "max_rel_diff_of_maps": rel,
                                      "note": "round-1 path: pupil sums downloaded, evaluation() finished in NumPy"},
            "cell_sums_equal_to_downloaded_bins": bool(np.array_equal(res["cell_sums"].astype(np.float64), cells_ref)),
            "api": "runner.trace_and_evaluate -> wgrt_trace_evaluate_metrics_host: K launches + pupil-mask sums + per-cell "
                   "totals + evaluation() lines 110-160 on the device, bins never downloaded"}
        del d_geom, d_luts, d_rng, d_eb, rargs, pin_keep, geom_p, luts_p, eb_view, seeds
        torch.cuda.empty_cache()

        # (3) e2e_dropin -- the literal 33-argument call on the runner's materialised arrays: H2D of all
        # ray arrays, LUTs, geometry and bins plus D2H of bins and RNG states, every step
        pinned, e2e_args = [], []
        for i, a in enumerate(host_args):
            if isinstance(a, np.ndarray) and i not in (2, 3, 4, 5):
                tpin, view = pinned_like(a)
                pinned.append(tpin)
                e2e_args.append(view)
            elif i in (2, 3, 4, 5):
                e2e_args.append(None)
            else:
                e2e_args.append(a)
        rng_host0 = rng_saved.cpu().numpy().view(np.uint32).copy()   # same streams as the timed launches
        prob, keep = GRTF.pack_problem(e2e_args, host=True)
        h2d = sum(a.nbytes for i, a in enumerate(e2e_args) if isinstance(a, np.ndarray))
        d2h = e2e_args[12].nbytes + e2e_args[32].nbytes
        tms = (C.c_float * 3)()
        for _ in range(2):
            _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 1, tms), lib)
        e2e_args[12][...] = rng_host0
        e2e_args[32][...] = 0
        barrier()
        t0 = time.perf_counter()
        parts = np.zeros(3)
        for _ in range(args.steps):
            _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 1, tms), lib)
            parts += np.array(list(tms))
        wall = wall_max(time.perf_counter() - t0)
        # same RNG streams as the timed device-resident launches -> same bounce count
        line["e2e_dropin"] = {"value": bounces_all / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                              "d2h_bytes_per_step": int(d2h), "ms_per_step": wall / max(args.steps, 1) * 1e3,
                              "stream_spans_ms_per_step_rank0": {"h2d": parts[0] / args.steps, "trace": parts[1] / args.steps,
                                                                 "d2h": parts[2] / args.steps},
                              "api": "wgrt_trace_fullcolor_host on the 12 materialised ray arrays (pinned), 1 launch per call"}
        e2e_deposits = float(e2e_args[32].sum(dtype=np.float64))
        line["e2e_dropin"]["deposits_match_device_run"] = bool(world > 1 or e2e_deposits == deposits_total)
        del prob, keep, e2e_args, pinned

    # ---- BASELINE configs[2], partitioned over the ranks (strong scaling), as a sub-record ----------
    if not args.no_partitioned:
        wl = "c3_dense_fov_41x41x3x10000"
        torch.cuda.empty_cache()
        rec = partitioned_record(args, *WORKLOADS[wl], world, rank, local_rank, wl, sample_clocks=False)
        if rank == 0:
            line["partitioned_c3"] = rec

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        threads = os.cpu_count() or 1
        probe, _ = sample_scene(scene, 200_000)
        dt, c0 = run_cpu_oracle(scene, probe, threads)
        sub, desc = sample_scene(scene, int(c0["rays"] / dt * args.cpu_seconds))
        dt, c1 = run_cpu_oracle(scene, sub, threads)
        line["cpu_baseline"] = {"value": c1["bounces"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": desc, "seconds": dt, "rays_per_s": c1["rays"] / dt}

    if rank == 0 and world == 1 and not args.no_legacy:   # (the sub-records that are no part of the step)
        line["small_launch"] = small_launch_record(stream)
    if rank == 0 and world == 1 and not args.no_legacy:
        line["legacy_split_tracer"] = legacy_record(args)
    if rank == 0 and world == 1:
        line["reference_cudasim"] = reference_cudasim_record()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def small_launch_record(stream):
    """Per-launch overhead: BASELINE configs[0] (532 nm, 5 x 5 FoV cells, 64 rays per cell = 1600
    rays) device resident through the reference-shaped kernel object, 200 launches back to back -- host time per call
    (argument packing in Python + the C ABI + 17 kernel launches, 14 of them no-ops of the cached region index) and device
    time per launch."""
    import torch
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    scene = si.make_scene(5, 5, 64, seed=11, lmd_subset=[1])

    def to_dev(a):
        if not isinstance(a, np.ndarray):
            return a
        v = a.view(np.float64) if a.dtype == np.complex128 else a
        return GRTF._TorchAlias(torch.from_numpy(np.ascontiguousarray(v.view(np.int32) if v.dtype == np.uint32 else v)).cuda(),
                                a.shape, a.dtype)
    dev = [to_dev(a) for a in scene.kernel_args(scene.new_matrix_EB())]
    launch = GRTF.process_rays_kernel_pro_fullColor[(scene.rays.num_rays + 255) // 256, 256, stream]
    for _ in range(20):
        launch(*dev)
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(n):
        launch(*dev)
    host = time.perf_counter() - t0
    e1.record(stream)
    torch.cuda.synchronize()
    return {"rays": scene.rays.num_rays, "launches": n, "host_us_per_call": host / n * 1e6,
            "device_us_per_launch": e0.elapsed_time(e1) / n * 1e3,
            "note": "device time is bounded below by the host rate when the host is the slower side"}


def legacy_record(args):
    """the legacy deterministic energy-splitting tracer (GRTF:192-417 + the compaction
    kernel GRTF:178-190) as a whole job -- generation after generation on the device, SoA queues, ballot / prefix-sum
    compaction -- against the CPU oracle restatement of the same kernels (one host thread, the same job)."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import legacy, synthetic_inputs as si
    from oracle import oracle
    oracle.build()
    nx, ny, npts, gens, cap = 24, 18, 64, 6, 1 << 21
    scene = si.make_scene(nx, ny, 2, seed=9, build_rays=False)
    geom, luts = legacy.make_legacy_luts(scene, 1, seed=4)
    pts = si.points_in_disc(scene.geom["IC"], npts, 10)
    rows0 = legacy.initial_rows(pts, nx, ny)
    kw = dict(max_steps=300, max_generations=gens, capacity=cap)
    legacy.trace(rows0, geom, luts, **kw)                       # warm-up (index build, arena)
    t0 = time.perf_counter()
    EB, live, st = legacy.trace(rows0, geom, luts, **kw)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    EB_o, live_o, st_o = oracle.legacy_trace(rows0, geom, luts, **kw)
    dt_o = time.perf_counter() - t0
    return {"job": f"{nx}x{ny} FoV cells x {npts} start points x TE/TM = {len(rows0)} initial rows, {gens} generations, "
                   "single wavelength, synthetic 3-order tables",
            "rows_processed": st["rows_processed"], "children": st["children"], "children_dropped": st["children_dropped"],
            "max_live_rows": st["max_live_rows"], "generations": st["generations"],
            "wall_ms": dt * 1e3, "rows_per_s": st["rows_processed"] / dt,
            "cpu_oracle": {"wall_ms": dt_o * 1e3, "rows_per_s": st_o["rows_processed"] / dt_o, "threads": 1},
            "counts_equal_to_oracle": st == st_o,
            "deposited_energy": float(EB.sum(dtype=np.float64)),
            "bins_max_rel_diff_to_oracle": float(np.max(np.abs(EB - EB_o) / np.maximum(np.abs(EB_o), 1e-30))) if EB_o.any() else 0.0,
            "api": "legacy.trace -> wgrt_legacy_trace_host (host arrays in and out, all generations on the device)"}


def reference_cudasim_record():
    """BASELINE.json: "the reference's only CPU path (its Numba kernels under NUMBA_ENABLE_CUDASIM on the box's
    host cores, core count stated)".  The reference is Python source and does not travel to the GPU box, so
    this is the rate measured where it exists -- the build container, when oracle/make_golden.py ran the
    unmodified reference kernel under the simulator to produce tests/golden/walk_c1.npz (BASELINE
    configs[0]: 532 nm, 5 x 5 FoV cells, 64 rays per cell) -- stored in the fixture with its process count."""
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "walk_c1.npz"))
        rays = int(g["rng_states"].shape[0]) * int(g["num_iter"])
        secs, procs = float(g["sim_seconds"]), int(g["sim_procs"])
        return {"rays_per_s": rays / secs, "seconds": secs, "rays": rays, "processes": procs,
                "config": "BASELINE configs[0]: single wavelength 532 nm, 5x5 FoV grid, 64 rays per FoV, x2 launches",
                "kind": "reference (GPU_ray_tracing_functions.py unmodified, NUMBA_ENABLE_CUDASIM=1, ray-chunked over "
                        "processes); measured in the build container at fixture generation, not on this box"}
    except (OSError, KeyError) as e:
        return {"unavailable": str(e)}


def e2e_multi_gpu(args, scene, rpc, N, world, rank, stream):
    """End to end on N GPUs, the way a replicated multi-GPU job runs (multi_gpu.ReplicatedJob; north_star:
    "one NCCL reduce of the bin tensors over NVLink at the end").  Every step, from pinned host memory to
    pinned host memory: each rank uploads 1/N of every table over its own PCIe link + NCCL all-gather;
    device-side seeding of rank-specific streams; one launch of the full C2 ray set through the
    reference-shaped kernel object (runner layout); ONE reduce-scatter of the bins; D2H of the rank's 1/N
    slice of the sum.  (N ranks each pushing full tables and full bin tensors through one host were measured
    host-bound in round 1: 57 ms per step at N = 4.)"""
    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, multi_gpu, synthetic_inputs as si
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 2024 + 1)
    job = multi_gpu.ReplicatedJob(pts, scene.geom, scene.n_g, scene.luts, rpc, eb=scene.eb_shape[3:])
    for _ in range(2):
        job.step(1)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        part, (lo, hi) = job.step(1)
    tw = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    # check: the same job walked from explicitly seeded host states, counters on, then summed over the ranks
    d_rng = torch.from_numpy(si.initial_rng_states(N, offset=rank * N).view(np.int32)).cuda()
    numel = int(np.prod(scene.eb_shape))
    chk = torch.zeros(numel, dtype=torch.float32, device="cuda")
    _capi.reset_counters()
    kern = job.kernel.configured(counters=True)
    kern[1, 256, stream](*job.launch_args(GRTF._TorchAlias(d_rng, (N,), np.uint32),
                                          GRTF._TorchAlias(chk, scene.eb_shape, np.float32)))
    c1 = _capi.read_counters()
    dist.all_reduce(chk)
    same = bool(torch.equal(chk[lo:hi].cpu(), torch.from_numpy(part)))
    tv = torch.tensor([float(c1["bounces"]), 1.0 if same else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(tv)
    out = {"value": float(tv[0].item()) * args.steps / wall, "unit": UNIT, "ms_per_step": wall / max(args.steps, 1) * 1e3,
           "h2d_bytes_per_step": int(job.h2d_bytes) * world, "d2h_bytes_per_step": int(numel * 4),
           "api": "multi_gpu.ReplicatedJob.step: per rank H2D of 1/N of every table + NCCL all-gather, device-side seeding, "
                  "GRTF.process_rays_kernel_pro_fullColor (runner layout) into device bins, one NCCL reduce-scatter "
                  "(uint8 when exact), D2H of the rank's 1/N slice of the summed bins",
           "reduced_slices_bit_equal_to_device_run": bool(tv[1].item() == world)}
    del job, chk, d_rng
    torch.cuda.empty_cache()
    return out


def reference_gpu_leg(args, dev_args, host_args, rng_saved, rng_final, eb_engine, N, stream, value, bounces_all):
    """BASELINE.json's second baseline: the reference's own kernel (GPU_ray_tracing_functions.py:833-1246,
    compiled by Numba to PTX in the build container, oracle/build_ref_ptx.py; JITed here by the driver as
    at a Numba launch) on the SAME device-resident inputs and RNG states as the timed launches, with the
    runner's launch shape (256 threads, RUN:160-177).  Also the full-size parity check: its bins and final
    RNG states must be bit-equal to the engine's."""
    import torch
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    from oracle import ref_numba_cuda as ref
    if not ref.available():
        return {"unavailable": "oracle/_ref PTX not built (needs /root/reference at build time)"}
    name = "process_rays_kernel_pro_fullColor"
    dead = torch.zeros(N, dtype=torch.float32, device="cuda")       # gap_x, gap_y, pol, azi: overwritten before use
    r_rng = rng_saved.clone()
    r_eb = torch.zeros_like(eb_engine)
    r_args = list(dev_args)
    for i in (2, 3, 4, 5):
        r_args[i] = GRTF._TorchAlias(dead, (N,), np.float32)
    r_args[12] = GRTF._TorchAlias(r_rng, (N,), np.uint32)
    r_args[32] = GRTF._TorchAlias(r_eb, host_args[32].shape, np.float32)
    t0 = time.perf_counter()
    attrs = ref.function_attributes(name)                           # module load = driver JIT of the PTX
    jit_s = time.perf_counter() - t0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for k in range(args.steps):
        ref.launch(name, r_args, stream=stream.cuda_stream)
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_s = ev[0].elapsed_time(ev[args.steps]) * 1e-3
    ref_value = bounces_all / total_s
    out = {"kernel": "GPU_ray_tracing_functions.process_rays_kernel_pro_fullColor (Numba 0.65 -> PTX sm_90, driver JIT)",
           "launch_shape": [(N + 255) // 256, 256], "ms_per_launch": ms, "value": ref_value, "unit": UNIT,
           "rays_per_s": N * args.steps / total_s, "driver_jit_s": jit_s, "registers": attrs["registers"],
           "local_bytes": attrs["local_bytes"],
           "bins_bit_equal_to_engine": bool(torch.equal(r_eb, eb_engine)),
           "rng_states_bit_equal_to_engine": bool(torch.equal(r_rng, rng_final)),
           "engine_over_reference_device_resident": value / ref_value,
           "note": "same inputs, same RNG states, same number of launches as the timed engine launches; warm "
                   "(the reference as shipped also pays Numba's Python->PTX compile, ~10 s, in its first launch)"}
    del dead, r_rng, r_eb
    torch.cuda.empty_cache()
    return out


def load_traffic():
    """DRAM bytes per launch of walk_warp_kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "walk_warp_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except OSError:
        return None


if __name__ == "__main__":
    main()
