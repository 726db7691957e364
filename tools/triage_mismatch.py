#!/usr/bin/env python
"""Triage of parity mismatches ("report mismatching rays and show they are ties").

Given a scene and two final ``rng_states`` arrays of the same launch(es) -- e.g. the engine's and the
oracle's, or the engine's and the reference kernel's -- every ray whose states differ drew a different
number of times, i.e. took a different branch somewhere.  Each such ray is replayed ALONE through the
CPU oracle (oracle/wgrt_oracle.c, test infrastructure) with a per-decision trace: the uniform draw and
the cumulative efficiencies it was compared with.  A mismatch is a TIE when some draw of the ray lies
within ``--tol`` (default 1e-9) of one of its thresholds -- two correct implementations whose
efficiencies differ in the last bits may then legitimately decide differently; otherwise it is reported
as a LOGIC difference.

    python tools/triage_mismatch.py --golden walk_deep --rng-a a.npy --rng-b b.npy
    python tools/triage_mismatch.py --golden walk_deep --engine            # engine (GPU) vs oracle
    python tools/triage_mismatch.py --scene 5,4,300,61 --engine            # nx,ny,rays_per_FoV,seed

Exit code 0: no mismatch, or ties only.  1: at least one logic difference.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def closest_threshold(ev: np.ndarray):
    """(distance, event index, which threshold) of the decision closest to a tie."""
    best = (np.inf, -1, "")
    for k, e in enumerate(ev):
        for nm, thr in (("e1", e[2]), ("e1+e2", e[3]), ("e1+e2+e3", e[4])):
            if np.isfinite(thr) and abs(e[1] - thr) < best[0]:
                best = (abs(e[1] - thr), k, nm)
    return best


def triage(scene, rng_start, rng_a, rng_b, tol=1e-9, num_iter=1, out=sys.stdout, names=("A", "B")):
    """Replay the rays whose final states differ; returns (n_mismatch, n_ties, n_logic)."""
    from oracle import oracle
    bad = np.flatnonzero(rng_a != rng_b)
    ties = logic = 0
    for i in bad:
        rng = rng_start.copy()
        EB = scene.new_matrix_EB()
        worst = (np.inf, -1, "", 0)
        n_ev = 0
        for it in range(num_iter):
            ev = oracle.trace_events(*scene.kernel_args(EB, rng), idx=int(i))
            n_ev += len(ev)
            d = closest_threshold(ev)
            if d[0] < worst[0]:
                worst = d + (it,)
        r = scene.rays
        verdict = "TIE" if worst[0] < tol else "LOGIC"
        ties += verdict == "TIE"
        logic += verdict == "LOGIC"
        print(f"ray {i}: cell (m={int(r.m[i])}, n={int(r.n[i])}, lambda={int(r.lmd_num[i])}) "
              f"final rng {names[0]}={rng_a[i]:#010x} {names[1]}={rng_b[i]:#010x} oracle={rng[i]:#010x}; "
              f"{n_ev} decisions, closest |u - threshold| = {worst[0]:.3e} (launch {worst[3]}, decision {worst[1]}, "
              f"vs {worst[2]}) -> {verdict}", file=out)
    print(f"{bad.size} mismatching rays of {rng_a.size}: {ties} ties (|u - threshold| < {tol:g}), {logic} logic differences",
          file=out)
    return bad.size, ties, logic


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--golden", help="name of a tests/golden walk fixture (its scene and, as B, its rng_states)")
    ap.add_argument("--scene", help="nx,ny,rays_per_FoV,seed of a synthetic scene")
    ap.add_argument("--rng-a", help=".npy of final rng_states (A)")
    ap.add_argument("--rng-b", help=".npy of final rng_states (B); default: the golden's, else the oracle's")
    ap.add_argument("--engine", action="store_true", help="A = the CUDA engine run here (needs a GPU)")
    ap.add_argument("--num-iter", type=int, default=None)
    ap.add_argument("--tol", type=float, default=1e-9)
    a = ap.parse_args(argv)
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    from oracle import oracle
    oracle.build()
    golden = None
    if a.golden:
        from conftest import load_golden_walk
        scene, golden = load_golden_walk(a.golden)
    elif a.scene:
        nx, ny, rpc, seed = (int(v) for v in a.scene.split(","))
        scene = si.make_scene(nx, ny, rpc, seed=seed)
    else:
        ap.error("--golden or --scene is required")
    num_iter = a.num_iter or (int(golden["num_iter"]) if golden is not None else 1)
    start = scene.rays.rng_states.copy()
    names = ["A", "B"]
    if a.engine:
        from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
        rng_a = start.copy(); EB = scene.new_matrix_EB()
        for _ in range(num_iter):
            GRTF.process_rays_kernel_pro_fullColor[1, 256](*scene.kernel_args(EB, rng_a))
        names[0] = "engine"
    elif a.rng_a:
        rng_a = np.load(a.rng_a).astype(np.uint32)
    else:
        ap.error("--rng-a or --engine is required")
    if a.rng_b:
        rng_b = np.load(a.rng_b).astype(np.uint32)
    elif golden is not None:
        rng_b = golden["rng_states"]; names[1] = "reference"
    else:
        rng_b = start.copy(); EB = scene.new_matrix_EB()
        for _ in range(num_iter):
            oracle.trace(*scene.kernel_args(EB, rng_b))
        names[1] = "oracle"
    _, _, logic = triage(scene, start, rng_a, rng_b, a.tol, num_iter, names=tuple(names))
    return 1 if logic else 0


if __name__ == "__main__":
    raise SystemExit(main())
