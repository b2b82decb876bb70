#!/bin/bash
# slot groups of the pipelined loop on one GPU: bench with 2 / 3 / 4 slots, strict and mixed; then the GPU suite
cd "$(dirname "$0")/.."
for slots in 2 3 4; do
  python bench.py --slots $slots --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/groups_$slots.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/groups_$slots.json')); o=d['other_precision_mode']
print('slots $slots strict value ms %.2f e2e ms %.2f gram %.3f share %.2f | mixed value ms %.2f e2e ms %.2f gram %.3f share %.2f | clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'], d['roofline']['gram_share_of_step'], o['ms_per_step'], o['e2e']['ms_per_step'], o['roofline']['gram_ms'], o['roofline']['gram_share_of_step'], d['clocks']['sm_mhz']))"
done
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2_gputests5.log; tail -4 gpurun_out/r2_gputests5.log
