python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kernel.py -x -q -m gpu 2>&1 | tail -3
python tools/quick_perf.py --rays 5000 --iters 4 2>&1 | tail -1 | cut -c1-110
