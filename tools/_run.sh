python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; tail -c 800 gpurun_out/bench_r1_f.err
