#!/bin/bash
# final 1-GPU evidence of the round: GPU suite, smoke, default bench, LSU, config 5, launch list + ncu captures
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2_gputests_final.log; tail -3 gpurun_out/r2_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_final.log
python bench.py > gpurun_out/r2_bench_ssu_1gpu.json 2> gpurun_out/r2_bench_ssu_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ssu_1gpu.json')); o=d['other_precision_mode']
print('strict value %.3g ms %.2f e2e %.3g ms %.2f frac %.3f gram %.3f share %.2f | mixed value %.3g ms %.2f e2e ms %.2f frac %.3f | clocks %s cpu %.3g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['gram_ms'], d['roofline']['gram_share_of_step'], o['value'], o['ms_per_step'], o['e2e']['ms_per_step'], o['roofline']['frac'], d['clocks']['sm_mhz'], d['cpu_baseline']['value']))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
python bench.py --workload lsu --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_lsu_1gpu.json 2> gpurun_out/r2_bench_lsu_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_lsu_1gpu.json')); print('lsu 1 gpu value %.3g ms %.2f e2e %.3g ms %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
for wl in rnasep trna; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/r2_bench_${wl}_1gpu.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${wl}_1gpu.json')); print('$wl value %.3g ms %.3f e2e %.3g ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"; done
BENCH_PHASES=1 python bench.py --workload sweep --stat all --steps 1 --warmup 1 > gpurun_out/r2_bench_sweep_all.json 2> gpurun_out/r2_bench_sweep_all.err; grep phases gpurun_out/r2_bench_sweep_all.err | tail -1
python bench.py --workload sweep --steps 1 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r2_bench_sweep_gt.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sweep_gt.json')); print('sweep GT alone ms', d['ms_per_step']); d=json.load(open('gpurun_out/r2_bench_sweep_all.json')); print('sweep all ms', d['ms_per_step'], d['contraction'])"
bash tools/r2_profile.sh
