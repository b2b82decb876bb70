#!/bin/bash
# end-to-end overlap diagnostics + sanitizers (1 GPU)
cd "$(dirname "$0")/.."
python tools/gen_time.py 100 2>&1 | tail -3
for s in 0 2; do echo "null slices $s"; python tools/e2e_timeline.py $s 2>&1 | tail -6; done
for c in 2 8 16; do echo "GEN_CHUNK $c"; RSCAPE_B200_GEN_CHUNK=$c python tools/e2e_timeline.py 0 2>&1 | tail -2; done
BENCH_PHASES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt 2>&1 >/dev/null | grep "e2e phases" | tail -3
bash tools/r2_sanitize.sh
