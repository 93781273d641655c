#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py > gpurun_out/h_bench.log 2> gpurun_out/h_bench.err; echo "rc=$?"; tail -1 gpurun_out/h_bench.log | cut -c1-300; tail -3 gpurun_out/h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.log 2> gpurun_out/h_bench_ref.err; echo "rc=$?"; tail -1 gpurun_out/h_bench_ref.log | cut -c1-900
timeout 600 python -m pytest tests/test_stage2_gpu.py -q -m gpu -s -k "gradients" 2>&1 | tail -15
