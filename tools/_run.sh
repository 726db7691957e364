python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/bench_r1_2gpu_c.json 2> gpurun_out/bench_r1_2gpu_c.err; tail -c 600 gpurun_out/bench_r1_2gpu_c.err
