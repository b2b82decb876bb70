#!/bin/bash
# BASELINE config 5: statistic sweep on the SSU shape with 500 nulls (one bench line per statistic x correction).
# Usage (GPU box):  bash tools/sweep.sh > gpurun_out/sweep.jsonl
for stat in GT MI MIr MIg CHI OMES RAFS; do
  for act in APC ASC; do
    python bench.py --workload sweep --stat $stat --actype $act --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1
  done
done
