"""Multi-GPU entries of the package (multi_gpu.py) on real devices.

* ``trace_partitioned`` needs no process group, so the partition / column logic and the
  WGRT_FLAG_BINS_COLUMNS host entry are covered on ONE GPU by walking every rank's share in turn and merging.
* With >= 2 GPUs, two NCCL ranks run the partitioned and the replicated job (``ReplicatedJob``: sharded table
  upload + all-gather, device-side seeding, reduce-scatter, per-rank download) and the parent checks them
  against the CPU oracle.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu, synthetic_inputs as si

NX, NY, RPC, ITERS = 9, 4, 120, 2


def _scene():
    scene = si.make_scene(NX, NY, RPC, seed=41)
    pts = si.points_in_disc(scene.geom["IC"], RPC // 2, 42)
    scene.rays = si.build_ray_set(pts, NX, NY, 3, RPC)
    return scene, pts


def _oracle_bins(oracle, scene, num_iter, seed_offset=0):
    EB = scene.new_matrix_EB()
    rng = si.initial_rng_states(scene.rays.num_rays, offset=seed_offset)
    for _ in range(num_iter):
        oracle.trace(*scene.kernel_args(EB, rng))
    return EB


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_partitioned_job_on_one_gpu(world, oracle):
    scene, pts = _scene()
    want = _oracle_bins(oracle, scene, ITERS)
    total = scene.new_matrix_EB()
    sentinel = np.float32(-7.0)
    for rank in range(world):
        part = np.full(scene.eb_shape, sentinel, dtype=np.float32)
        out, (m0, m1) = multi_gpu.trace_partitioned(pts, scene.geom, scene.n_g, scene.luts, RPC, world, rank,
                                                    num_iter=ITERS, matrix_EB=part)
        assert out is part
        # only the rank's own columns are written (from zero); the rest of the host array is untouched
        assert np.all(part[:, :, :m0] == sentinel) and np.all(part[:, :, m1:] == sentinel)
        assert np.all(part[:, :, m0:m1] >= 0)
        multi_gpu.merge_columns(total, part, (m0, m1))
    assert np.array_equal(total, want) and want.sum() > 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu as mg
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    scene, pts = _scene()
    part, cols = mg.trace_partitioned(pts, scene.geom, scene.n_g, scene.luts, RPC, world, rank, num_iter=ITERS)
    job = mg.ReplicatedJob(pts, scene.geom, scene.n_g, scene.luts, RPC)
    sl, (lo, hi) = job.step(ITERS)
    sl = sl.copy()
    sl2, _ = job.step(ITERS)                       # persistent buffers: a second step gives the same result
    # the all-reduce form and the reducer on a tensor whose counts do not qualify for the narrow path
    big = torch.full((1000,), 200.0, device="cuda")
    red = mg.BinReducer(big.numel(), "cuda")
    p1 = red.reduce_scatter_checked(big).clone()
    a, b = red.slice_of()
    ok_wide = bool(torch.all(p1 == 200.0 * world)) and p1.numel() == b - a
    full = torch.full((1000,), 3.0, device="cuda")
    mg.reduce_bins(full)
    ok_all = bool(torch.all(full == 3.0 * world))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), part=part[:, :, cols[0]:cols[1]], cols=np.array(cols),
             sl=sl, same=np.array(np.array_equal(sl, sl2)), lohi=np.array([lo, hi]), ok=np.array([ok_wide, ok_all]))
    dist.destroy_process_group()


def test_two_ranks_partitioned_and_replicated(tmp_path, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    scene, pts = _scene()
    want = _oracle_bins(oracle, scene, ITERS)
    total = scene.new_matrix_EB()
    n = scene.rays.num_rays
    rep = sum(_oracle_bins(oracle, scene, ITERS, seed_offset=r * n) for r in range(world))
    got = np.zeros(rep.size, dtype=np.float32)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        m0, m1 = d["cols"]
        total[:, :, m0:m1] += d["part"]
        lo, hi = d["lohi"]
        got[lo:hi] = d["sl"]
        assert bool(d["same"]) and bool(d["ok"].all())
    assert np.array_equal(total, want), "partitioned job differs from the single-GPU result"
    assert np.array_equal(got.reshape(rep.shape), rep), "replicated job: summed bins differ"
