#!/bin/bash
# ncu evidence for generator B (null_simulate_level_kernel, the cov_GenerateAlignment kernel) and for generator A's kernels at the SSU shape:
# launch list with DRAM bytes, and a full capture of a few mid-tree levels of the largest chunk
cd "$(dirname "$0")/.."
GEN_ITERS=1 GEN_SIMULATE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'null_simulate|fitch|replay|permut' -c 1400 --csv \
    --log-file gpurun_out/r2_launches_generators.csv python tools/gen_time.py 100 > gpurun_out/r2_launches_generators.log 2>&1
GEN_ITERS=1 GEN_SIMULATE=1 ncu --set full --clock-control none --import-source on -k regex:'null_simulate' --launch-skip 158 --launch-count 4 -o gpurun_out/r2_sim python tools/gen_time.py 100 >> gpurun_out/r2_launches_generators.log 2>&1
ncu -i gpurun_out/r2_sim.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_simulate_raw.csv 2>/dev/null
rm -f gpurun_out/r2_sim.ncu-rep
tail -4 gpurun_out/r2_launches_generators.log; ls -la gpurun_out | grep -E "generators|simulate"
