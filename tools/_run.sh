for w in 24 26 27 28; do echo "cap $w"; WGRT_WARPS_PER_SM=$w python tools/quick_perf.py --rays 5000 --iters 4 2>&1 | tail -1 | cut -c1-110; done
