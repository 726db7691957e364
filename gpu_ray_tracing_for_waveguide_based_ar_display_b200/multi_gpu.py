"""One process per GPU: partition the runner's work over ranks, reduce the bins once at the end.

The reference is single GPU.  Rays are independent, RNG state is per ray and
the only shared output -- the bin tensor -- is additive (GPU_ray_tracing_functions.py:33, 164), so
the path shards without any collective inside the walk:

* ``cell_range(n_cells, world, rank)`` gives each rank a contiguous range of the runner's cell
  sequence (FoV-x outer, FoV-y, wavelength inner; gpu_ray_tracing_pro_fullColor.py:82-84);
* ``shard_rays`` builds that rank's ray set with the GLOBAL per-ray RNG seeds, so every ray draws
  exactly the numbers it would draw in a single-GPU run;
* ``reduce_bins`` is the one collective: a sum all-reduce (NCCL over NVLink on GPUs, gloo in the CPU
  tests) of the float32 bin tensor after the last launch.  Bins are integer counts far below 2^24,
  so float32 summation is exact and the result is bit-identical for any number of ranks.

``run_partitioned`` strings these together around a caller-supplied ``trace`` callable (the engine's
kernel object on GPUs).  PyTorch is plumbing here: device memory and ``torch.distributed``.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import synthetic_inputs as si

__all__ = ["cell_range", "shard_rays", "reduce_bins", "run_partitioned"]


def cell_range(n_cells: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of the runner's cell sequence for ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_cells, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_rays(points: np.ndarray, num_FOV_x: int, num_FOV_y: int, n_lmd: int, num_rays_per_FoV: int,
               world_size: int, rank: int) -> Tuple[si.RaySet, Tuple[int, int]]:
    """This rank's rays (whole cells) with the RNG seeds of the unpartitioned launch (RUN:158)."""
    ii, jj, ll = np.meshgrid(np.arange(num_FOV_x), np.arange(num_FOV_y), np.arange(n_lmd), indexing="ij")
    cells = np.stack((ii.ravel(), jj.ravel(), ll.ravel()), axis=1)
    c0, c1 = cell_range(len(cells), world_size, rank)
    rays = si.build_ray_set(points, num_FOV_x, num_FOV_y, n_lmd, num_rays_per_FoV, cells=cells[c0:c1])
    rays.rng_states = si.initial_rng_states(rays.num_rays, offset=c0 * num_rays_per_FoV)
    return rays, (c0 * num_rays_per_FoV, c1 * num_rays_per_FoV)


def reduce_bins(matrix_EB, group=None, narrow: bool = True):
    """Sum ``matrix_EB`` over all ranks, in place.  Accepts a torch tensor (CUDA -> NCCL, CPU -> gloo)
    or a NumPy array (wrapped without copying).

    The bins are small integer counts stored in float32 (one deposit adds exactly 1.0,
    GPU_ray_tracing_functions.py:1168).  With ``narrow`` they travel as uint8, four to an int32 word: if
    every entry of every rank is an integer in [0, 255 // world_size], no byte sum can exceed 255, no
    carry crosses a byte, and the int32 SUM all-reduce IS the sum of the counts -- a quarter of the bytes
    over NVLink (864 MB -> 216 MB at the default size), bit-identical result.  A trailing word of the same
    all-reduce carries every rank's "my entries do not qualify" flag; if it comes back non-zero the bins
    (still untouched) are reduced as float32 instead.  ONE collective in the common case."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return matrix_EB
    t = torch.from_numpy(matrix_EB) if isinstance(matrix_EB, np.ndarray) else matrix_EB
    world = dist.get_world_size(group)
    limit = 255 // world
    if narrow and limit >= 1 and t.dtype == torch.float32 and t.numel() > 0 and t.numel() % 4 == 0 and t.is_contiguous():
        n = t.numel()
        # packed counts + the flag word, padded to a multiple of 4 KB (collectives like round sizes)
        nw = (n // 4 + 1 + 1023) // 1024 * 1024
        words = torch.empty(nw, dtype=torch.int32, device=t.device)
        words[n // 4:].zero_()
        q = words[:n // 4].view(torch.uint8)
        on_gpu = t.is_cuda and t.data_ptr() % 16 == 0
        if on_gpu:
            # one fused pass of the engine: convert and flag entries that are not integers in [0, limit]
            import ctypes as C
            from . import _capi
            lib = _capi.load_library()
            st = torch.empty(2, dtype=torch.int32, device=t.device)
            stream = torch.cuda.current_stream(t.device).cuda_stream
            _capi.check(lib.wgrt_bins_pack_u8(C.c_void_p(t.data_ptr()), n, C.c_void_p(q.data_ptr()),
                                              C.c_void_p(st.data_ptr()), C.c_float(float(limit)), C.c_void_p(stream)), lib)
            words[-1:] = st[1:2]
        else:
            flat = t.reshape(-1)
            clipped = flat.clamp(0, limit)
            q.copy_(clipped.to(torch.uint8))
            words[-1] = int(not bool((q.to(torch.float32) == flat).all()))
        dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
        if int(words[-1].item()) == 0:                       # every rank qualified: q holds the exact sums
            if on_gpu:
                _capi.check(lib.wgrt_bins_unpack_u8(C.c_void_p(q.data_ptr()), n, C.c_void_p(t.data_ptr()),
                                                    C.c_void_p(stream)), lib)
            else:
                t.reshape(-1).copy_(q)
            return matrix_EB
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return matrix_EB


def run_partitioned(scene: si.Scene, points: np.ndarray, trace: Callable, num_iter: int = 1,
                    world_size: Optional[int] = None, rank: Optional[int] = None, group=None):
    """Trace this rank's share of ``scene`` for ``num_iter`` launches and all-reduce the bins.

    ``trace(*args33)`` is the launch callable (e.g. ``kernel[grid, block]``); it must mutate
    ``rng_states`` and ``matrix_EB`` in place like the reference kernel.  A ``RayWalkKernel`` may be
    passed instead: it is then launched with ``ray_index_base`` = this rank's first ray, so that even a
    zero RNG state reseeds (GRTF:28-29) as in the single launch over the whole job.  Returns
    ``(matrix_EB, rng_states, (first_ray, last_ray))`` -- bins summed over all ranks, and this
    rank's slice of the global RNG state array.
    """
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    m = scene.meta
    L = scene.eb_shape[0]
    rays, span = shard_rays(points, m["num_FOV_x"], m["num_FOV_y"], L, m["num_rays_per_FoV"], world_size, rank)
    if hasattr(trace, "configured"):
        trace = trace.configured(ray_index_base=span[0])[(rays.num_rays + 255) // 256, 256]
    full = scene.rays
    scene.rays = rays
    try:
        EB = scene.new_matrix_EB()
        rng = rays.rng_states
        for _ in range(num_iter):
            trace(*scene.kernel_args(EB, rng))
    finally:
        scene.rays = full
    reduce_bins(EB, group)
    return EB, rng, span
