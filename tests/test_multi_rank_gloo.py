"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): partition, per-ray RNG seeding, and the
single bin all-reduce must reproduce the single-process result bit for bit.  The trace callable is
the CPU oracle here (tests may use it); on GPUs it is the engine's kernel object."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu, synthetic_inputs as si
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = si.make_scene(5, 3, 40, seed=21)
    pts = si.points_in_disc(scene.geom["IC"], 20, 22)
    EB, rng, span = multi_gpu.run_partitioned(scene, pts, lambda *a: oracle.trace(*a, num_threads=2), num_iter=2)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), EB=EB, rng=rng, span=np.array(span))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_run_equals_single_process(world, tmp_path, oracle):
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu, synthetic_inputs as si
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    scene = si.make_scene(5, 3, 40, seed=21)
    pts = si.points_in_disc(scene.geom["IC"], 20, 22)
    # single process = world size 1 through the same code path
    EB1, rng1, span1 = multi_gpu.run_partitioned(scene, pts, oracle.trace, num_iter=2, world_size=1, rank=0)
    assert span1 == (0, 5 * 3 * 3 * 40)
    rng_all = np.zeros_like(rng1)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(d["EB"], EB1), "reduced bins differ from the single-process run"
        a, b = d["span"]
        rng_all[a:b] = d["rng"]
    assert np.array_equal(rng_all, rng1)
    assert EB1.sum() > 0


def _reduce_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rs = np.random.default_rng(100 + rank)
    cases = {"counts": rs.integers(0, 40, size=(3, 4, 5, 8, 12)).astype(np.float32),        # narrow path (3 * 39 <= 255)
             "large": rs.integers(0, 120, size=(2, 3, 3, 8, 12)).astype(np.float32),         # 3 * 119 > 255: float32 path
             "fractional": rs.random((2, 2, 2, 8, 12)).astype(np.float32)}                   # not counts: float32 path
    if rank == 1:
        cases["counts"][0, 0, 0, 0, 0] = 85.0                                                # 3 * 85 = 255: still narrow
    out = {}
    for k, a in cases.items():
        ref = torch.from_numpy(a.copy()); dist.all_reduce(ref)
        got = a.copy(); multi_gpu.reduce_bins(got)
        plain = a.copy(); multi_gpu.reduce_bins(plain, narrow=False)
        out[k] = bool(np.array_equal(got, ref.numpy()) and np.array_equal(plain, ref.numpy()))
    # BinReducer: persistent buffers, reduce-scatter form, flag reduced with the data (no host sync on the path)
    for k, a in cases.items():
        ref = torch.from_numpy(a.copy()); dist.all_reduce(ref)
        red = multi_gpu.BinReducer(a.size, "cpu")
        t = torch.from_numpy(a.copy())
        keep = t.clone()
        part = red.reduce_scatter(t)
        narrow = red.narrow_ok()
        lo, hi = red.slice_of()
        ok = torch.equal(t, keep) and narrow == (k == "counts")          # the bins are never modified
        if narrow:
            ok = ok and torch.equal(part, ref.reshape(-1)[lo:hi])
        part = red.reduce_scatter_checked(t).clone()
        ok = ok and torch.equal(part, ref.reshape(-1)[lo:hi])
        spans = [red.slice_of(r) for r in range(world)]
        ok = ok and spans[0][0] == 0 and spans[-1][1] == a.size and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        out["reducer_" + k] = bool(ok)
    np.savez(os.path.join(out_dir, f"reduce{rank}.npz"), **out)
    dist.destroy_process_group()


def test_reduce_bins_narrow_path_is_exact(tmp_path):
    """The uint8 all-reduce of the bins is used only when it is exact, and then gives the float32 result."""
    world = 3
    mp.spawn(_reduce_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = np.load(tmp_path / f"reduce{r}.npz")
        assert all(bool(d[k]) for k in d.files) and len(d.files) == 6, {k: bool(d[k]) for k in d.files}


def test_cell_range_partition():
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.multi_gpu import cell_range
    for n in (0, 1, 7, 22500):
        for w in (1, 2, 3, 8):
            spans = [cell_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cell_range(10, 2, 2)


def test_shard_rays_seeds_match_global_layout():
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import multi_gpu, synthetic_inputs as si
    pts = np.random.default_rng(0).uniform(size=(4, 2))
    full = si.build_ray_set(pts, 3, 2, 3, 8)
    for w in (2, 4):
        for r in range(w):
            rays, (a, b) = multi_gpu.shard_rays(pts, 3, 2, 3, 8, w, r)
            for f in si.RAY_FIELDS:
                assert np.array_equal(getattr(rays, f), getattr(full, f)[a:b]), f
            assert np.array_equal(rays.rng_states, full.rng_states[a:b])


def test_column_ranges_and_merge():
    """Partitioned jobs need no collective: every rank returns the FoV-x columns its cells touch; shared
    boundary columns add up (disjoint cells)."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.multi_gpu import cell_range, column_range, merge_columns
    L, X, Y = 3, 41, 41
    n = L * X * Y
    rs = np.random.default_rng(3)
    # bins of the whole job: cell (m, n, l) owns tile [l, n, m]
    full = rs.integers(0, 5, size=(L, Y, X, 2, 3)).astype(np.float32)
    for world in (1, 2, 5, 8):
        total = np.zeros_like(full)
        covered = np.zeros(X, dtype=int)
        for r in range(world):
            c0, c1 = cell_range(n, world, r)
            m0, m1 = column_range(c0, c1, Y, L)
            assert 0 <= m0 < m1 <= X and m0 == c0 // (Y * L) and (m1 - 1) == (c1 - 1) // (Y * L)
            covered[m0:m1] += 1
            part = np.zeros_like(full)                       # what the rank's walk produces: its own cells only
            cells = np.arange(c0, c1)
            mm, nn, ll = cells // (Y * L), (cells // L) % Y, cells % L
            part[ll, nn, mm] = full[ll, nn, mm]
            merge_columns(total, part, (m0, m1))
        assert np.array_equal(total, full) and covered.min() >= 1 and covered.max() <= 2
    assert column_range(5, 5, Y, L) == (0, 0)
