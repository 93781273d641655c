#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NUNERF_CHAIN_IMPL=ts timeout 120 python tools/run_chain_once.py > gpurun_out/c_plain_ts.log 2>&1 && \
NUNERF_CHAIN_IMPL=ts timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 2 -c 1 -f -o gpurun_out/prof_r2_chain_ts python tools/run_chain_once.py > gpurun_out/c_ncu_ts.log 2>&1
NUNERF_CHAIN_IMPL=ss timeout 120 python tools/run_chain_once.py > gpurun_out/c_plain_ss.log 2>&1 && \
NUNERF_CHAIN_IMPL=ss timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 2 -c 1 -f -o gpurun_out/prof_r2_chain_ss python tools/run_chain_once.py > gpurun_out/c_ncu_ss.log 2>&1
tail -3 gpurun_out/c_ncu_ts.log gpurun_out/c_ncu_ss.log
ls -la gpurun_out/*.ncu-rep
