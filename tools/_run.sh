python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -c 400 gpurun_out/bench_r1_g.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_list_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk_warp -s 3 -c 1 -o gpurun_out/prof_walk_warp_final python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full_g.log 2>&1
tail -2 gpurun_out/ncu_full_g.log
