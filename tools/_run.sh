python tools/quick_perf.py --rays 5000 --iters 4 2>&1 | tail -1 | cut -c1-120
