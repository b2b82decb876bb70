#!/bin/bash
# round-2 ncu evidence: launch list of one bench step + full captures of the scan kernels (strict, mixed, multi-statistic).
# Reports are exported to CSV on the box; only the contraction's report is kept whole (gpurun_out/ is limited to 64 MiB).
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gram_i8' -c 8 -o gpurun_out/r2_gram python tools/profile_scan.py ssu > gpurun_out/r2_scan_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'gt_finish|pack_planes|correct_hist|multi_stat|marg_sum|stat_kernel|reduce_cov' -c 30 -o /tmp/r2_aux python tools/profile_scan.py ssu >> gpurun_out/r2_scan_ncu.log 2>&1
ncu -i gpurun_out/r2_gram.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_gram_raw.csv 2>/dev/null
ncu -i /tmp/r2_aux.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_aux_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
