#!/bin/bash
# compute-sanitizer passes over the small-shape driver; logs land in gpurun_out/ (summaries are copied to profiles/)
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool exit=$?"; tail -3 gpurun_out/r2_sanitizer_$tool.log
done
