#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/i_bench_2gpu.log 2> gpurun_out/i_bench_2gpu.err; echo "rc=$?"; tail -1 gpurun_out/i_bench_2gpu.log | cut -c1-400; tail -3 gpurun_out/i_bench_2gpu.err
