#!/bin/bash
cd "$(dirname "$0")/.."
python tools/stage_time.py > gpurun_out/r2_stage_time_hits_treesubs.jsonl 2> gpurun_out/r2_stage_time.err; cut -c1-260 gpurun_out/r2_stage_time_hits_treesubs.jsonl
bash tools/r2_sanitize.sh
RSCAPE_B200_TRACE=1 python tools/loop_time.py 2>&1 | grep -E "100 nulls|rsb\] gram 0\.[1-9]"
RSCAPE_B200_LIB=$PWD/r-scape_b200/build/alt/fin4.so RSCAPE_B200_TRACE=1 python tools/loop_time.py 2>&1 | grep -E "100 nulls|rsb\] gram 0\.[1-9]"
bash tools/r2_profile.sh
