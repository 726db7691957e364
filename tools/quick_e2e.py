"""Ad-hoc probe (not the bench): wall time of runner.trace_full_color on C2 vs number of pipeline chunks."""
import os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner, synthetic_inputs as si

nx, ny, rpc = 100, 75, 5000
scene = si.make_scene(nx, ny, 2, seed=2024)          # tables only; rays are implicit
pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 2025)

def pinned_like(a):
    v = a.view(np.float64) if a.dtype == np.complex128 else a
    t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)
keep, geom, luts = [], {}, {}
for src, dst in ((scene.geom, geom), (scene.luts, luts)):
    for k, a in src.items():
        t, v = pinned_like(a); keep.append(t); dst[k] = v
t, eb = pinned_like(np.zeros((3, ny, nx, 80, 120), np.float32)); keep.append(t)
chunk_list = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4,8,12,16,25,50").split(",")]
ref = None
for it in (1, 4):
    for ch in chunk_list:
        os.environ["WGRT_HOST_CHUNKS"] = str(ch)
        tm = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            spans = []
            runner.trace_full_color(pts, geom, scene.n_g, luts, rpc, num_iter=it, matrix_EB=eb, bins_start_zero=True,
                                    timings=spans)
            tm.append((time.perf_counter() - t0) * 1e3)
        s = float(eb.sum(dtype=np.float64))
        if it == 1:
            ref = s if ref is None else ref
            assert s == ref, (s, ref)
        print(json.dumps({"num_iter": it, "chunks": ch, "wall_ms": [round(x, 2) for x in tm],
                          "spans_ms": [round(x, 2) for x in spans], "deposits": s}), flush=True)
