"""One process per GPU: the two ways the runner's job spreads over the GPUs of a box.

The reference is single GPU.  Rays are independent, RNG state is per ray and the
only shared output -- the bin tensor -- is additive (GPU_ray_tracing_functions.py:33, 164), so the path
shards without any collective inside the walk.

PARTITIONED (strong scaling, ``trace_partitioned``): the job's FoV-wavelength cells -- the runner's cell
sequence, FoV-x outer, FoV-y, wavelength inner (gpu_ray_tracing_pro_fullColor.py:82-84) -- are cut into
contiguous, balanced ranges (``cell_range``), one per rank, with the GLOBAL per-ray RNG seeds (RUN:158), so
every ray draws exactly the numbers it draws in a single-GPU run.  A cell's deposits land in that cell's
tile of ``matrix_EB`` only, so the ranks' outputs are disjoint: NO collective is needed.  Each rank uploads
only the table columns of its cells over its own PCIe link, walks them, and downloads only the FoV-x
columns of ``matrix_EB`` its cells touch (the pipelined host entry with ``WGRT_FLAG_BINS_COLUMNS``).  A
boundary column shared by two ranks comes back from both, each holding its own cells' counts and zeros
elsewhere: ``merge_columns`` adds them.

REPLICATED (weak scaling, ``ReplicatedJob``): every rank walks the whole design with its own RNG streams
(more Monte-Carlo samples per FoV); the bins must then be SUMMED over the ranks: one NCCL reduce-scatter
(``BinReducer``), after which every rank holds -- and downloads -- 1/N of the summed tensor.  The job's tables
enter the node once: each rank uploads 1/N of every table over its own PCIe link and an NCCL all-gather
over NVLink completes them on every GPU.

``BinReducer`` moves the bins as uint8 whenever that is exact (small integer counts; four to an int32
word, a quarter of the bytes) and as float32 otherwise; the result is bit-identical for any number of ranks
because the counts are integers far below 2^24.

PyTorch is plumbing here: device memory and ``torch.distributed``.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import _capi, synthetic_inputs as si

__all__ = ["cell_range", "column_range", "shard_rays", "reduce_bins", "run_partitioned", "trace_partitioned",
           "merge_columns", "BinReducer", "ReplicatedJob"]


def cell_range(n_cells: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of the runner's cell sequence for ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_cells, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def column_range(c0: int, c1: int, num_FOV_y: int, n_lmd: int) -> Tuple[int, int]:
    """FoV-x columns [m_lo, m_hi) touched by the cells [c0, c1) of the runner's sequence."""
    cpc = num_FOV_y * n_lmd
    if c1 <= c0:
        return 0, 0
    return c0 // cpc, (c1 - 1) // cpc + 1


def shard_rays(points: np.ndarray, num_FOV_x: int, num_FOV_y: int, n_lmd: int, num_rays_per_FoV: int,
               world_size: int, rank: int) -> Tuple[si.RaySet, Tuple[int, int]]:
    """This rank's rays (whole cells) with the RNG seeds of the unpartitioned launch (RUN:158)."""
    ii, jj, ll = np.meshgrid(np.arange(num_FOV_x), np.arange(num_FOV_y), np.arange(n_lmd), indexing="ij")
    cells = np.stack((ii.ravel(), jj.ravel(), ll.ravel()), axis=1)
    c0, c1 = cell_range(len(cells), world_size, rank)
    rays = si.build_ray_set(points, num_FOV_x, num_FOV_y, n_lmd, num_rays_per_FoV, cells=cells[c0:c1])
    rays.rng_states = si.initial_rng_states(rays.num_rays, offset=c0 * num_rays_per_FoV)
    return rays, (c0 * num_rays_per_FoV, c1 * num_rays_per_FoV)


# ------------------------------------------------------------------------------------------------
# partitioned job: no collective
# ------------------------------------------------------------------------------------------------
def trace_partitioned(points: np.ndarray, geom: Dict[str, np.ndarray], n_g: float, luts: Dict[str, np.ndarray],
                      num_rays_per_FoV: int, world_size: int, rank: int, num_iter: int = 4,
                      eb: Tuple[int, int] = (80, 120), matrix_EB: Optional[np.ndarray] = None,
                      timings: Optional[list] = None):
    """This rank's share of the runner's job: cells ``cell_range(L * X * Y, world_size, rank)``, walked
    ``num_iter`` times with the global RNG seeds, through the pipelined host entry
    (``runner.trace_full_color``): only this rank's table columns go up, only its ``matrix_EB`` columns come
    down.  ``matrix_EB`` is a full-shape float32 host array (created when omitted; pass a pinned one for
    speed) of which ONLY the returned column range is written (starting from zero).

    Returns ``(matrix_EB, (m_lo, m_hi))``.  Needs no process group: ``world_size`` / ``rank`` are plain
    integers (``torch.distributed`` ranks, MPI ranks, or a loop over devices in one process).
    """
    from . import runner
    L, X, Y, _ = geom["lut_TIR"].shape
    c0, c1 = cell_range(L * X * Y, world_size, rank)
    if matrix_EB is None:
        matrix_EB = np.zeros((L, Y, X, eb[0], eb[1]), dtype=np.float32)
    runner.trace_full_color(points, geom, n_g, luts, num_rays_per_FoV, num_iter=num_iter, eb=eb, first_cell=c0,
                            num_cells=c1 - c0, matrix_EB=matrix_EB, bins_start_zero=True,
                            flags=_capi.WGRT_FLAG_BINS_COLUMNS, timings=timings)
    return matrix_EB, column_range(c0, c1, Y, L)


def merge_columns(total: np.ndarray, part: np.ndarray, columns: Tuple[int, int]) -> np.ndarray:
    """Add the columns ``[m_lo, m_hi)`` of a rank's full-shape ``part`` into ``total`` (in place).  Columns
    shared by two ranks hold disjoint cells' counts, so the sum is the single-GPU result bit for bit."""
    m0, m1 = columns
    total[:, :, m0:m1] += part[:, :, m0:m1]
    return total


# ------------------------------------------------------------------------------------------------
# the one collective of a replicated job
# ------------------------------------------------------------------------------------------------
class BinReducer:
    """Sum of the float32 bin tensor over the ranks, with persistent buffers.

    The bins are small integer counts stored in float32 (one deposit adds exactly 1.0,
    GPU_ray_tracing_functions.py:1168).  They travel as uint8, four to an int32 word: if every entry of
    every rank is an integer in [0, 255 // world_size], no byte sum can exceed 255, no carry crosses a byte,
    and the int32 SUM IS the sum of the counts -- a quarter of the bytes over NVLink (864 MB -> 216 MB at
    the default size), bit-identical result.  Whether every rank qualified is itself reduced on the device
    (one 4-byte all-reduce); nothing on the critical path synchronises with the host.  ``narrow_ok()`` reads
    that flag afterwards; if it is False the narrow result must be discarded and ``wide=True`` used (the
    bins are never modified by a narrow reduce-scatter).
    """

    def __init__(self, numel: int, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.numel = int(numel)
        self.device = torch.device(device)
        self.limit = 255 // max(self.world, 1)
        # elements per rank: a multiple of 4 (uint8 packing) -- the tensor is padded up to world * per
        self.per = -(-self.numel // (4 * self.world)) * 4
        self.padded = self.per * self.world
        self.words = torch.zeros(self.padded // 4, dtype=torch.int32, device=self.device)
        self.part_words = torch.empty(self.per // 4, dtype=torch.int32, device=self.device)
        self.part = torch.empty(self.per, dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._wide_in = None
        self.nccl = self.device.type == "cuda"
        self.lib = _capi.load_library() if self.nccl else None

    def slice_of(self, rank: Optional[int] = None) -> Tuple[int, int]:
        """Flat element range [lo, hi) of the summed tensor that ``rank`` holds after ``reduce_scatter``."""
        r = self.rank if rank is None else rank
        return min(r * self.per, self.numel), min((r + 1) * self.per, self.numel)

    def _pack(self, t):
        torch = self.torch
        n4 = self.numel // 4 * 4
        if self.nccl and t.data_ptr() % 16 == 0:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            q = self.words.view(torch.uint8)
            _capi.check(self.lib.wgrt_bins_pack_u8(C.c_void_p(t.data_ptr()), n4, C.c_void_p(q.data_ptr()),
                                                   C.c_void_p(self.stats.data_ptr()), C.c_float(float(self.limit)),
                                                   C.c_void_p(stream)), self.lib)
            self.flag.copy_(self.stats[1:2])
            if n4 != self.numel:                                    # ragged tail (never the case for whole tiles)
                tail = t.reshape(-1)[n4:]
                q[n4:self.numel].copy_(tail.clamp(0, self.limit).to(torch.uint8))
                self.flag.add_((q[n4:self.numel].to(torch.float32) != tail).any().to(torch.int32))
        else:
            flat = t.reshape(-1)
            q = self.words.view(torch.uint8)[:self.numel]
            q.copy_(flat.clamp(0, self.limit).to(torch.uint8))
            self.flag.fill_(int(not bool((q.to(torch.float32) == flat).all())))

    def reduce_scatter(self, matrix_EB, wide: bool = False):
        """Sum over the ranks; returns this rank's slice ``slice_of()`` of the result as a float32 tensor on
        the bins' device (a view into a persistent buffer: consume it before the next call).  Asynchronous
        on the current stream.  ``matrix_EB`` itself is left untouched."""
        torch, dist = self.torch, self.dist
        t = matrix_EB
        lo, hi = self.slice_of()
        if self.world == 1:
            return t.reshape(-1)
        if wide or self.limit < 1:
            if self._wide_in is None or self._wide_in.numel() != self.padded:
                self._wide_in = torch.zeros(self.padded, dtype=torch.float32, device=self.device)
            self._wide_in[:self.numel].copy_(t.reshape(-1))
            if self.nccl:
                dist.reduce_scatter_tensor(self.part, self._wide_in, group=self.group)
            else:                                                    # gloo has no reduce-scatter
                dist.all_reduce(self._wide_in, group=self.group)
                self.part.copy_(self._wide_in[self.rank * self.per:(self.rank + 1) * self.per])
            self.flag.zero_()
            return self.part[:hi - lo]
        self._pack(t)
        dist.all_reduce(self.flag, group=self.group)                 # 4 bytes: did every rank qualify?
        if self.nccl:
            dist.reduce_scatter_tensor(self.part_words, self.words, group=self.group)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _capi.check(self.lib.wgrt_bins_unpack_u8(C.c_void_p(self.part_words.data_ptr()), self.per,
                                                     C.c_void_p(self.part.data_ptr()), C.c_void_p(stream)), self.lib)
        else:
            dist.all_reduce(self.words, group=self.group)
            mine = self.words[self.rank * (self.per // 4):(self.rank + 1) * (self.per // 4)]
            self.part.copy_(mine.view(torch.uint8))
        return self.part[:hi - lo]

    def narrow_ok(self) -> bool:
        """True when the last narrow ``reduce_scatter`` was exact on every rank (synchronises with the host)."""
        return int(self.flag.item()) == 0

    def reduce_scatter_checked(self, matrix_EB):
        """``reduce_scatter`` with the fallback: narrow first, float32 if some rank's bins did not qualify."""
        part = self.reduce_scatter(matrix_EB)
        if self.world > 1 and not self.narrow_ok():
            part = self.reduce_scatter(matrix_EB, wide=True)
        return part


def reduce_bins(matrix_EB, group=None, narrow: bool = True):
    """Sum ``matrix_EB`` over all ranks, in place (an all-reduce: every rank ends with the full sum).
    Accepts a torch tensor (CUDA -> NCCL, CPU -> gloo) or a NumPy array (wrapped without copying).  With
    ``narrow`` the bins travel as uint8 when that is exact (see ``BinReducer``)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return matrix_EB
    t = torch.from_numpy(matrix_EB) if isinstance(matrix_EB, np.ndarray) else matrix_EB
    if narrow and t.dtype == torch.float32 and t.numel() > 0 and t.is_contiguous():
        red = BinReducer(t.numel(), t.device, group)
        if red.limit >= 1:
            red._pack(t)
            dist.all_reduce(red.flag, group=group)
            if red.narrow_ok():
                dist.all_reduce(red.words, group=group)
                q = red.words.view(torch.uint8)
                n4 = t.numel() // 4 * 4
                if red.nccl and t.data_ptr() % 16 == 0:
                    stream = torch.cuda.current_stream(t.device).cuda_stream
                    _capi.check(red.lib.wgrt_bins_unpack_u8(C.c_void_p(q.data_ptr()), n4, C.c_void_p(t.data_ptr()),
                                                            C.c_void_p(stream)), red.lib)
                    if n4 != t.numel():
                        t.reshape(-1)[n4:].copy_(q[n4:t.numel()])
                else:
                    t.reshape(-1).copy_(q[:t.numel()])
                return matrix_EB
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return matrix_EB


# ------------------------------------------------------------------------------------------------
# replicated job: every rank walks the whole design, one reduce-scatter
# ------------------------------------------------------------------------------------------------
class ReplicatedJob:
    """The runner's job replicated over the ranks of a process group (weak scaling: world x the Monte-Carlo
    samples per FoV), from host memory to host memory.

    ``step(num_iter)`` does, every call:
      1. the job's inputs enter the node ONCE: each rank uploads 1/N of every large table from pinned host
         memory over its own PCIe link, and an NCCL all-gather over NVLink completes the tables on every GPU;
      2. each rank seeds its own RNG streams on the device (RUN:158 with the ray index shifted by
         rank * num_rays: independent streams) and walks the full ray set ``num_iter`` times through the
         reference-shaped kernel object (runner layout) into a device bin tensor;
      3. ONE reduce-scatter (``BinReducer``) sums the bins over the ranks; each rank downloads its 1/N slice
         of the sum into pinned host memory.
    Returns ``(part, (lo, hi))``: the flat element range of the summed ``matrix_EB`` this rank holds.
    """

    def __init__(self, points: np.ndarray, geom: Dict[str, np.ndarray], n_g: float, luts: Dict[str, np.ndarray],
                 num_rays_per_FoV: int, eb: Tuple[int, int] = (80, 120), group=None, counters: bool = False):
        import torch
        import torch.distributed as dist
        from . import GPU_ray_tracing_functions as GRTF
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lib = _capi.load_library()
        self.n_g = float(n_g)
        L, X, Y, _ = geom["lut_TIR"].shape
        self.shape = (L, Y, X, int(eb[0]), int(eb[1]))
        self.rpc = int(num_rays_per_FoV)
        self.N = L * X * Y * self.rpc
        tables = dict(geom)
        tables.update(luts)
        tables["px"] = np.ascontiguousarray(points[:, 0], dtype=np.float32)
        tables["py"] = np.ascontiguousarray(points[:, 1], dtype=np.float32)
        self.host, self.dev, self.alias, self.sharded = {}, {}, {}, {}
        self.h2d_bytes = 0
        for k, a in tables.items():
            v = a.view(np.float64) if a.dtype == np.complex128 else a
            tpin = torch.from_numpy(np.ascontiguousarray(v)).pin_memory().view(-1)
            self.host[k] = tpin
            self.dev[k] = torch.empty_like(tpin, device="cuda")
            self.alias[k] = GRTF._TorchAlias(self.dev[k], a.shape, a.dtype)
            self.sharded[k] = self.world > 1 and tpin.numel() % self.world == 0 and a.nbytes >= (1 << 20)
            self.h2d_bytes += a.nbytes // self.world if self.sharded[k] else a.nbytes
        numel = int(np.prod(self.shape))
        self.eb_dev = torch.zeros(numel, dtype=torch.float32, device="cuda")
        self.eb_alias = GRTF._TorchAlias(self.eb_dev, self.shape, np.float32)
        self.rng_dev = torch.empty(self.N, dtype=torch.int32, device="cuda")
        self.rng_alias = GRTF._TorchAlias(self.rng_dev, (self.N,), np.uint32)
        self.reducer = BinReducer(numel, "cuda", group)
        lo, hi = self.reducer.slice_of()
        self.part_host = torch.empty(hi - lo, dtype=torch.float32).pin_memory()
        self.kernel = GRTF.process_rays_kernel_pro_fullColor.configured(
            counters=counters, ray_index_base=self.rank * self.N).runner_layout(self.rpc // 2, self.N)

    def launch_args(self, rng_alias=None, eb_alias=None):
        g = self.alias
        return [g["px"], g["py"]] + [None] * 10 + [
            rng_alias or self.rng_alias, g["IC"], g["FC"], g["FC_offset"], g["OC"], g["OC_offset"], self.n_g,
            g["eff_reg1"], g["eff_reg2"], g["eff_reg_FOV"], g["eff_reg_FOV_range"], g["lut_ic1"], g["lut_ic2"],
            g["lut_ic3"], g["lut_fc1"], g["lut_fc2"], g["lut_oc1"], g["lut_oc2"], g["lut_TIR"], g["lut_gap"],
            eb_alias or self.eb_alias]

    def upload(self):
        torch, dist = self.torch, self.dist
        for k, h in self.host.items():
            if self.sharded[k]:
                n = h.numel() // self.world
                mine = self.dev[k][self.rank * n:(self.rank + 1) * n]
                mine.copy_(h[self.rank * n:(self.rank + 1) * n], non_blocking=True)
                dist.all_gather_into_tensor(self.dev[k], mine, group=self.group)
            else:
                self.dev[k].copy_(h, non_blocking=True)

    def seed(self):
        stream = self.torch.cuda.current_stream().cuda_stream
        _capi.check(self.lib.wgrt_seed_rng(C.c_void_p(self.rng_dev.data_ptr()), self.N, self.rank * self.N,
                                           C.c_void_p(stream)), self.lib)

    def step(self, num_iter: int = 1):
        torch = self.torch
        stream = torch.cuda.current_stream()
        self.upload()
        self.seed()
        self.eb_dev.zero_()
        launch = self.kernel[1, 256, stream]
        args = self.launch_args()
        for _ in range(num_iter):
            launch(*args)
        part = self.reducer.reduce_scatter(self.eb_dev) if self.world > 1 else self.eb_dev
        self.part_host.copy_(part, non_blocking=True)
        torch.cuda.synchronize()
        if self.world > 1 and not self.reducer.narrow_ok():          # some count exceeded 255 // world: float32
            part = self.reducer.reduce_scatter(self.eb_dev, wide=True)
            self.part_host.copy_(part, non_blocking=True)
            torch.cuda.synchronize()
        return self.part_host.numpy(), self.reducer.slice_of()


def run_partitioned(scene: si.Scene, points: np.ndarray, trace: Callable, num_iter: int = 1,
                    world_size: Optional[int] = None, rank: Optional[int] = None, group=None):
    """Host-array form of a partitioned job around a caller-supplied ``trace`` callable (used by the CPU
    multi-process tests with the oracle as ``trace``; GPU jobs use ``trace_partitioned``): trace this
    rank's share of ``scene`` for ``num_iter`` launches and all-reduce the bins.

    ``trace(*args33)`` must mutate ``rng_states`` and ``matrix_EB`` in place like the reference kernel.  A
    ``RayWalkKernel`` may be passed instead: it is then launched with ``ray_index_base`` = this rank's first
    ray, so that even a zero RNG state reseeds (GRTF:28-29) as in the single launch over the whole job.
    Returns ``(matrix_EB, rng_states, (first_ray, last_ray))`` -- bins summed over all ranks, and this
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
