#!/bin/bash
# One GPU call: smoke(), the whole GPU suite, timings and an ncu launch list of the stages after the scan.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1b_gpu.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1b_smoke.log
tail -n 3 gpurun_out/r1b_smoke.log
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r1b_gpu_all.log 2>&1
echo "all exit $?" >> gpurun_out/r1b_gpu_all.log
tail -n 4 gpurun_out/r1b_gpu_all.log
timeout 120 python tools/stage_time.py > gpurun_out/r1b_stage_time.jsonl 2> gpurun_out/r1b_stage_time.err
cat gpurun_out/r1b_stage_time.jsonl
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'evalue_hits|subs_tables|branch_rows|gram_i8|pack_planes' --csv --log-file gpurun_out/r1b_ncu_newkernels.csv \
  python tools/stage_time.py > /dev/null 2>&1
