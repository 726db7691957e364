python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kernel.py tests/test_region_index.py -x -q -m gpu 2>&1 | tail -3
for w in 16 20 24 28 32; do echo "warps/SM cap $w"; WGRT_WARPS_PER_SM=$w python tools/quick_perf.py --rays 5000 --iters 3 2>&1 | tail -1 | cut -c1-100; done
