#!/bin/bash
# round-2 ncu evidence: launch list of one bench step + full captures of the scan kernels (strict, mixed, multi-statistic)
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gram_i8|gt_finish|pack_planes|correct_hist|multi_stat|marg_sum|stat_kernel|reduce_cov' \
    -c 60 -o gpurun_out/r2_scan python tools/profile_scan.py ssu > gpurun_out/r2_scan_ncu.log 2>&1
tail -3 gpurun_out/r2_scan_ncu.log
