python -m pytest tests/test_evaluation.py -x -q -m gpu 2>&1 | tail -5
python - <<'PY'
import numpy as np, time, torch, ctypes as C
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi
lib=_capi.load_library()
for shape, mask, sy, sx in (((3,75,100,80,120),30,1,1), ((3,19,25,320,480),120,1,1), ((3,75,100,80,120),30,8,12)):
    n=int(np.prod(shape)); eb=torch.randint(0,3,(n,),device='cuda').float()
    ny=(shape[3]-mask)//sy+1; nx=(shape[4]-mask)//sx+1
    out=torch.empty(shape[0]*shape[1]*shape[2]*ny*nx,device='cuda'); cs=torch.empty(shape[0]*shape[1]*shape[2],device='cuda')
    for _ in range(2):
        _capi.check(lib.wgrt_eval_pupil_sums(C.c_void_p(eb.data_ptr()), *shape, mask, sy, sx, C.c_void_p(out.data_ptr()), C.c_void_p(cs.data_ptr()), None), lib)
    torch.cuda.synchronize(); t0=time.perf_counter()
    _capi.check(lib.wgrt_eval_pupil_sums(C.c_void_p(eb.data_ptr()), *shape, mask, sy, sx, C.c_void_p(out.data_ptr()), C.c_void_p(cs.data_ptr()), None), lib)
    torch.cuda.synchronize(); print(shape, mask, sy, sx, "outputs/tile", ny*nx, "ms", round((time.perf_counter()-t0)*1e3,3))
PY
