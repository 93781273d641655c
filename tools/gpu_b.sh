#!/bin/bash
export PYTHONUNBUFFERED=1
for d in 0 64 2 66 1 67 16; do
  echo "== NUNERF_CHAIN_DEBUG=$d"; NUNERF_CHAIN_DEBUG=$d timeout 120 python tools/bench_chain.py 2>&1 | grep "fused" | tail -1
done
