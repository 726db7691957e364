set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; tail -c 600 gpurun_out/bench_r1_e.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref_e.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_warp.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_list_e.log 2>&1
for t in 2500 1250; do python tools/quick_perf.py --rays 5000 --iters 3 --tile $t 2>&1 | tail -1 | cut -c1-90; done
