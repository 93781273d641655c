#!/bin/bash
# round-2 GPU call D: cta_group::2 variant and skeleton ablations re-measured after the elect.sync issue fix
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in "1 0" "2 0" "1 16" "1 32" "1 48" "2 32"; do
  set -- $v
  echo "== pair=$1 debug=$2"
  NUNERF_CHAIN_PAIR=$1 NUNERF_CHAIN_DEBUG=$2 timeout 300 python tools/bench_chain.py 2>&1 | grep "fused"
  NUNERF_CHAIN_PAIR=$1 NUNERF_CHAIN_DEBUG=$2 timeout 300 python tools/bench_pred.py 2>&1 | grep -i "fused" | head -4
done
echo "== bench pair=2"
NUNERF_CHAIN_PAIR=2 timeout 600 python bench.py --no-subrecords > gpurun_out/d_bench_pair2.log 2>&1; tail -1 gpurun_out/d_bench_pair2.log | cut -c1-300
