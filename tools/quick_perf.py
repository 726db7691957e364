"""Ad-hoc timing probe (not the bench): fast vs strict walk on a runner-shaped scene."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, synthetic_inputs as si

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=100); ap.add_argument("--ny", type=int, default=75)
ap.add_argument("--rays", type=int, default=1000); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--strict", action="store_true"); ap.add_argument("--tile", type=int, default=0)
a = ap.parse_args()
t0 = time.time()
scene = si.make_scene(a.nx, a.ny, a.rays, seed=1)
print("scene built", time.time() - t0, "s; rays", scene.rays.num_rays, flush=True)

def to_dev(x):
    if isinstance(x, np.ndarray):
        v = x.view(np.float64) if x.dtype == np.complex128 else x
        t = torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda()
        return GRTF._TorchAlias(t, x.shape, x.dtype)
    return x
dev = [to_dev(x) for x in scene.kernel_args(scene.new_matrix_EB())]
rng0 = dev[12]._t.clone()
N = scene.rays.num_rays
res = {}
for name, kern in (("fast", GRTF.process_rays_kernel_pro_fullColor.configured(tile_hint=a.tile)),) + \
        ((("strict", GRTF.process_rays_kernel_pro_fullColor.configured(strict=True)),) if a.strict else ()):
    dev[12]._t.copy_(rng0); dev[32]._t.zero_()
    kc = kern.configured(counters=True)
    _capi.reset_counters()
    kc[1, 256](*dev); torch.cuda.synchronize()
    cnt = _capi.read_counters()
    times = []
    for it in range(a.iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); kern[1, 256](*dev); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    res[name] = dict(ms=times, rays_per_s=N / ms * 1e3, ray_bounces_per_s=cnt["bounces"] / ms * 1e3,
                     counters={k: round(v / max(cnt["rays"], 1), 3) for k, v in cnt.items()},
                     deposits=float(dev[32]._t.sum().item()))
    print(name, json.dumps(res[name]), flush=True)
