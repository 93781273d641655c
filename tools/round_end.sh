#!/bin/bash
# Round-end measurement set (run on the GPU box: gpurun -- 'bash tools/round_end.sh r1p').  Everything lands in
# gpurun_out/<tag>_*; numbers are taken WITHOUT a profiler attached, the ncu passes run afterwards on the same commands.
tag=${1:-rN}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/${tag}_tests.log
python bench.py > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_line.json 2> $out/${tag}_bench_reference.err
python bench.py --step 20000 --no-cpu-baseline --no-subrecords > $out/${tag}_bench_step20000.json 2>> $out/${tag}_bench.err
python bench.py --rays-per-gpu 32768 --chunk 8192 --no-cpu-baseline --no-subrecords > $out/${tag}_bench_config3_32768rays_1gpu.json 2>> $out/${tag}_bench.err
python tools/bench_hbm_kernels.py > $out/${tag}_hbm_kernels_32768rays.json 2> $out/${tag}_hbm.err
python tools/bench_stage2.py > $out/${tag}_stage2_config4.json 2> $out/${tag}_stage2.err
python tools/bench_stage2.py 4096 --thick > $out/${tag}_stage2_nonzero_thickness.json 2>> $out/${tag}_stage2.err
python tools/bench_sweep.py 512 > $out/${tag}_config5_sweep512_evalrender.json 2> $out/${tag}_sweep.err
# ncu: launch list of the bench step, then one --set full capture of the HBM-side kernels and the marching-cubes kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"composite|compact|geometry|upsample|ray_setup|merge_sdf|points_kernel" \
    -c 40 -o $out/${tag}_hbm python tools/bench_hbm_kernels.py once > $out/${tag}_ncu_hbm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mc_count|mc_emit" -c 4 -o $out/${tag}_mc \
    python tools/bench_sweep.py 256 > $out/${tag}_ncu_mc.log 2>&1
cat $out/${tag}_tests.log; cat $out/${tag}_bench_line.json | cut -c1-400
