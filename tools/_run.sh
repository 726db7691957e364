python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --workload c3_dense_fov_41x41x3x10000 > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; tail -c 300 gpurun_out/bench_c3_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c3_dense_fov_41x41x3x10000 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; tail -c 300 gpurun_out/bench_c3_2gpu.err
python tools/quick_e2e.py 8,16 2>&1 | tail -4 | cut -c1-120
python tools/quick_perf.py --rays 5000 --iters 3 2>&1 | tail -1 | cut -c1-100
