#!/bin/bash
# Final ncu --set full captures of the round (run on the GPU box: gpurun -- 'bash tools/gpu_e.sh').  The commands are
# run without ncu first (numbers are never taken under the profiler), then captured.
out=gpurun_out
python tools/bench_pred.py once > $out/r2u_pred_once.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|dw_tc" -c 12 -f -o $out/r2u_chain \
    python tools/bench_pred.py once > $out/r2u_ncu_chain.log 2>&1
python tools/profile_stage2.py --thick > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"shell_bounce|shade_encode_.*var|seg_composite|hit_interp" \
    -c 16 -f -o $out/r2u_nz python tools/profile_stage2.py --thick > $out/r2u_ncu_nz.log 2>&1
ncu -i $out/r2u_chain.ncu-rep --page raw --csv > $out/r2u_chain_raw.csv 2>/dev/null
ncu -i $out/r2u_nz.ncu-rep --page raw --csv > $out/r2u_nz_raw.csv 2>/dev/null
ls -la $out/r2u_chain.ncu-rep $out/r2u_nz.ncu-rep $out/r2u_chain_raw.csv $out/r2u_nz_raw.csv
tail -2 $out/r2u_ncu_chain.log $out/r2u_ncu_nz.log
