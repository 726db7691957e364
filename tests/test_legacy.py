"""Row f4: the legacy deterministic energy-splitting tracer (GRTF:178-190, 192-417).

CPU: the oracle restatement against the reference kernels themselves (tests/golden/legacy.npz, made by running
the unmodified ``process_rays_kernel`` / ``pack_active_to_front`` under Numba's simulator).  GPU: the CUDA kernels
(drop-in per-launch form on the reference's AoS rows, and the SoA generation loop) against the oracle and against
the reference's own PTX.

Parity bar: ray / child COUNTS, region states, flags and cell indices exact; ray data (positions, field amplitudes,
phases) to 1e-12 relative -- the reference appends children in thread-scheduling order, so rows are compared as a
sorted multiset -- and the bins (float32 energy sums, accumulation-order dependent) to 1e-5 relative.
"""
import ast
import os

import numpy as np
import pytest

from conftest import GOLDEN

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import legacy


def load():
    from oracle import make_golden
    g = np.load(os.path.join(GOLDEN, "legacy.npz"))
    r = ast.literal_eval(str(g["recipe"]))
    geom, luts, rows0 = make_golden.legacy_inputs(r)
    return g, r, geom, luts, rows0


def step_args(rows, count, counter, r, geom, luts, EB):
    return (rows, count, counter, r["max_steps"], geom["IC"], geom["FC"], geom["FC_offset"], geom["OC"], geom["OC_offset"],
            geom["eff_reg1"], geom["eff_reg2"], geom["eff_reg_FOV"], geom["eff_reg_FOV_range"], luts["lut_ic1"],
            luts["lut_ic2"], luts["lut_fc1"], luts["lut_fc2"], luts["lut_oc"], geom["lut_TIR"], geom["lut_gap"], EB)


def sort_rows(rows):
    rows = np.asarray(rows)
    return rows[np.lexsort(rows.T[::-1])]


def sort_rows_robust(rows):
    """Order that survives last-bit differences: by the exact columns (state, flag, m, n) and rounded data."""
    rows = np.asarray(rows)
    key = np.round(rows[:, [0, 1, 8, 9, 10]], 7)
    return rows[np.lexsort((key[:, 4], key[:, 3], key[:, 2], key[:, 1], key[:, 0], rows[:, 7], rows[:, 6], rows[:, 12],
                            rows[:, 11]))]


def assert_rows_close(a, b, what):
    assert a.shape == b.shape, f"{what}: {a.shape[0]} rows where {b.shape[0]} are expected"
    a, b = sort_rows_robust(a), sort_rows_robust(b)
    for c in (6, 7, 11, 12):
        assert np.array_equal(a[:, c], b[:, c]), f"{what}: column {c} (exact) differs"
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13, err_msg=what)


def test_oracle_reproduces_reference_kernels(oracle):
    g, r, geom, luts, rows0 = load()
    cap = r["capacity"]
    a = np.zeros((cap, 13)); b = np.zeros((cap, 13))
    a[:len(rows0)] = rows0
    EB = np.zeros(tuple(g["eb_shape"]), dtype=np.float32)
    count = len(rows0)
    for gen in range(r["generations"]):
        counter = np.array([count], dtype=np.int32)
        assert oracle.legacy_step(*step_args(a, count, counter, r, geom, luts, EB)) == 0
        total = int(counter[0])
        assert total == int(g[f"launch{gen}_counter"]), "child count"
        assert np.array_equal(sort_rows(a[:total]), g[f"launch{gen}_rows"]), f"rows after launch {gen}"
        b[:] = 0
        count = oracle.legacy_pack(a, b, total)
        assert np.array_equal(sort_rows(b[:count]), g[f"pack{gen}_rows"]), f"rows after compaction {gen}"
        a, b = b, a
    want = np.zeros(EB.size, dtype=np.float32); want[g["eb_index"]] = g["eb_value"]
    np.testing.assert_allclose(EB.ravel(), want, rtol=1e-6, atol=0)
    assert EB.sum() > 0 and count > len(rows0)


def test_oracle_generation_loop_and_capacity(oracle):
    g, r, geom, luts, rows0 = load()
    EB, live, st = oracle.legacy_trace(rows0, geom, luts, max_steps=r["max_steps"], max_generations=r["generations"],
                                       capacity=r["capacity"])
    assert st["generations"] == r["generations"] and st["children_dropped"] == 0
    assert np.array_equal(sort_rows(live), g[f"pack{r['generations'] - 1}_rows"])
    # children beyond the capacity are dropped and counted, never written
    EB2, live2, st2 = oracle.legacy_trace(rows0, geom, luts, max_steps=r["max_steps"], max_generations=r["generations"],
                                          capacity=150)
    assert st2["children_dropped"] > 0 and st2["max_live_rows"] <= 150


def test_argument_checks():
    g, r, geom, luts, rows0 = load()
    EB = np.zeros(tuple(g["eb_shape"]), dtype=np.float32)
    good = step_args(rows0.copy(), len(rows0), np.zeros(1, np.int32), r, geom, luts, EB)
    legacy.pack_legacy_problem(good, host=True)
    for i, bad in ((0, rows0.astype(np.float32)), (2, np.zeros(1, np.int64)), (13, luts["lut_ic1"][..., :20]),
                   (20, EB.astype(np.float64)), (18, geom["lut_TIR"][..., :2].copy())):
        a = list(good); a[i] = bad
        with pytest.raises((TypeError, ValueError)):
            legacy.pack_legacy_problem(a, host=True)
    a = list(good); a[1] = len(rows0) + 1
    with pytest.raises(ValueError):
        legacy.pack_legacy_problem(a, host=True)
    with pytest.raises(TypeError):
        legacy.pack_legacy_problem(good[:-1], host=True)


# ------------------------------------------------------------------------------------------------ GPU
def _to_dev(a):
    import torch
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
    v = a.view(np.float64) if a.dtype == np.complex128 else a
    return GRTF._TorchAlias(torch.from_numpy(np.ascontiguousarray(v)).cuda(), a.shape, a.dtype)


@pytest.mark.gpu
def test_gpu_launch_matches_golden_and_oracle(oracle):
    """GRTF.process_rays_kernel / pack_active_to_front drop-ins (device AoS rows) generation by generation."""
    g, r, geom, luts, rows0 = load()
    cap = r["capacity"]
    a = np.zeros((cap, 13)); b = np.zeros((cap, 13))
    a[:len(rows0)] = rows0
    EB = np.zeros(tuple(g["eb_shape"]), dtype=np.float32)
    count = len(rows0)
    for gen in range(r["generations"]):
        counter = np.array([count], dtype=np.int32)
        legacy.process_rays_kernel[(count + 255) // 256, 256](*step_args(a, count, counter, r, geom, luts, EB))
        total = int(counter[0])
        assert total == int(g[f"launch{gen}_counter"]), "child count"
        assert_rows_close(a[:total], g[f"launch{gen}_rows"], f"rows after launch {gen}")
        oc = np.zeros(1, dtype=np.int32)
        b[:] = 0
        legacy.pack_active_to_front[(total + 255) // 256, 256](a, b, total, oc)
        count = int(oc[0])
        assert_rows_close(b[:count], g[f"pack{gen}_rows"], f"rows after compaction {gen}")
        a, b = b, a
    want = np.zeros(EB.size, dtype=np.float32); want[g["eb_index"]] = g["eb_value"]
    np.testing.assert_allclose(EB.ravel(), want, rtol=1e-5, atol=0)
    z = np.ones(7, dtype=np.float32)
    legacy.zero_out_kernel[1, 32](z)
    assert not z.any()


@pytest.mark.gpu
def test_gpu_generation_loop_matches_oracle(oracle):
    """legacy.trace: SoA queues, ballot / prefix-sum compaction, all generations on the device."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    scene = si.make_scene(6, 5, 2, seed=9, build_rays=False)
    geom, luts = legacy.make_legacy_luts(scene, 1, seed=4)
    pts = si.points_in_disc(scene.geom["IC"], 12, 10)
    rows0 = legacy.initial_rows(pts, 6, 5)
    kw = dict(max_steps=300, max_generations=7, capacity=1 << 17)
    EB_o, live_o, st_o = oracle.legacy_trace(rows0, geom, luts, **kw)
    EB, live, st = legacy.trace(rows0, geom, luts, **kw)
    assert st == st_o and st["children"] > 10 * len(rows0) and st["children_dropped"] == 0
    assert_rows_close(live, live_o, "live rows after the last generation")
    np.testing.assert_allclose(EB, EB_o, rtol=1e-5, atol=1e-12)
    assert EB.sum() > 0
    # a queue that is too small: children are dropped and counted, the job still terminates
    EB2, live2, st2 = legacy.trace(rows0, geom, luts, max_steps=300, max_generations=7, capacity=2000)
    assert st2["children_dropped"] > 0 and st2["max_live_rows"] <= 2000


@pytest.mark.gpu
def test_gpu_launch_matches_reference_ptx():
    """The reference's own process_rays_kernel / pack_active_to_front (Numba -> PTX, oracle/_ref) on the GPU:
    same counts, same rows (1e-12), same bins (1e-5), two generations on a job the simulator could not do."""
    import torch
    from oracle import ref_numba_cuda as ref
    if not ref.available() or "process_rays_kernel" not in ref.manifest()["kernels"]:
        pytest.skip("oracle/_ref PTX of the legacy kernels not built")
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    scene = si.make_scene(8, 6, 2, seed=19, build_rays=False)
    geom, luts = legacy.make_legacy_luts(scene, 1, seed=5)
    pts = si.points_in_disc(scene.geom["IC"], 40, 11)
    rows0 = legacy.initial_rows(pts, 8, 6)
    cap = 1 << 16
    r = dict(max_steps=300)
    res = {}
    for impl in ("engine", "reference"):
        a = np.zeros((cap, 13)); a[:len(rows0)] = rows0
        d_a, d_b = _to_dev(a), _to_dev(np.zeros((cap, 13)))
        d_EB = _to_dev(np.zeros((6, 8, 80, 120), dtype=np.float32))
        count = len(rows0)
        dev = {k: _to_dev(v) for k, v in {**geom, **luts}.items()}
        snaps = []
        for gen in range(3):
            d_cnt = _to_dev(np.array([count], dtype=np.int32)); d_oc = _to_dev(np.zeros(1, dtype=np.int32))
            args = (d_a, count, d_cnt, r["max_steps"], dev["IC"], dev["FC"], dev["FC_offset"], dev["OC"], dev["OC_offset"],
                    dev["eff_reg1"], dev["eff_reg2"], dev["eff_reg_FOV"], dev["eff_reg_FOV_range"], dev["lut_ic1"],
                    dev["lut_ic2"], dev["lut_fc1"], dev["lut_fc2"], dev["lut_oc"], dev["lut_TIR"], dev["lut_gap"], d_EB)
            if impl == "engine":
                legacy.process_rays_kernel[1, 256](*args)
            else:
                ref.launch("process_rays_kernel", args, n=count)
            torch.cuda.synchronize()
            total = int(d_cnt._t.cpu()[0])
            assert total <= cap
            d_b._t.zero_()
            if impl == "engine":
                legacy.pack_active_to_front[1, 256](d_a, d_b, total, d_oc)
            else:
                ref.launch("pack_active_to_front", (d_a, d_b, total, d_oc), n=total)
            torch.cuda.synchronize()
            count = int(d_oc._t.cpu()[0])
            snaps.append((total, d_a._t.cpu().numpy()[:total].copy(), count, d_b._t.cpu().numpy()[:count].copy()))
            d_a, d_b = d_b, d_a
        res[impl] = (snaps, d_EB._t.cpu().numpy())
    for gen, (e, f) in enumerate(zip(res["engine"][0], res["reference"][0])):
        assert e[0] == f[0] and e[2] == f[2], f"generation {gen}: counts {e[0], e[2]} vs reference {f[0], f[2]}"
        assert_rows_close(e[1], f[1], f"rows after launch {gen}")
        assert_rows_close(e[3], f[3], f"rows after compaction {gen}")
    np.testing.assert_allclose(res["engine"][1], res["reference"][1], rtol=1e-5, atol=1e-12)
    assert res["engine"][1].sum() > 0 and res["engine"][0][-1][2] > 4 * len(rows0)
