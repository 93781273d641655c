#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NUNERF_CHAIN_IMPL=ts timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "fused or sdf_network or cta_pair" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/f_bench_ss.log 2>&1; tail -1 gpurun_out/f_bench_ss.log | cut -c1-300
NUNERF_CHAIN_IMPL=ts timeout 600 python bench.py > gpurun_out/f_bench_ts.log 2>&1; tail -1 gpurun_out/f_bench_ts.log | cut -c1-300
NUNERF_CHAIN_IMPL=ts NUNERF_CHAIN_KORDER=1 timeout 600 python bench.py > gpurun_out/f_bench_ts_k.log 2>&1; tail -1 gpurun_out/f_bench_ts_k.log | cut -c1-300
