#!/bin/bash
# round-2 GPU call A: TS-mode chain kernel correctness + micro-benchmarks against the SS kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== TS correctness (chain tests)"
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "fused or sdf_network or cta_pair" > gpurun_out/a_ts_tests.log 2>&1
echo "rc=$?" >> gpurun_out/a_ts_tests.log
tail -5 gpurun_out/a_ts_tests.log
for v in "ts 1" "ss 1" "ss 0"; do
  set -- $v
  echo "== micro-benchmarks impl=$1 hiprio=$2"
  NUNERF_CHAIN_IMPL=$1 NUNERF_CHAIN_HIPRIO=$2 timeout 300 python tools/bench_chain.py > gpurun_out/a_chain_$1_$2.log 2>&1
  NUNERF_CHAIN_IMPL=$1 NUNERF_CHAIN_HIPRIO=$2 timeout 300 python tools/bench_pred.py > gpurun_out/a_pred_$1_$2.log 2>&1
  grep -h "fused" gpurun_out/a_chain_$1_$2.log gpurun_out/a_pred_$1_$2.log
done
echo "== full GPU tests (TS default)"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/a_all_tests.log
tail -4 gpurun_out/a_all_tests.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/a_bench_ts.log 2>&1; tail -1 gpurun_out/a_bench_ts.log | cut -c1-600
NUNERF_CHAIN_IMPL=ss NUNERF_CHAIN_HIPRIO=1 timeout 600 python bench.py > gpurun_out/a_bench_ss1.log 2>&1; tail -1 gpurun_out/a_bench_ss1.log | cut -c1-400
