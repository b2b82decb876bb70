#!/bin/bash
# ring depth 3 vs 4 of the contraction: kernel time, overlap with the generator, whole bench; then the GPU suite, config 5 and LSU on 1 GPU
cd "$(dirname "$0")/.."
A=$PWD/r-scape_b200/build/alt
python tools/gram_time.py ssu 2,4 2>&1 | grep gram
RSCAPE_B200_LIB=$A/ns3.so python tools/gram_time.py ssu 2,4 2>&1 | grep gram
for s in 0 2; do echo "ns3, null slices $s"; RSCAPE_B200_LIB=$A/ns3.so python tools/e2e_timeline.py $s 2>&1 | tail -4; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ns4_bench.json 2>/dev/null
RSCAPE_B200_LIB=$A/ns3.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ns3_bench.json 2>/dev/null
for f in ns4 ns3; do python -c "
import json; d=json.load(open('gpurun_out/r2_${f}_bench.json')); o=d['other_precision_mode']
print('$f strict value ms %.2f e2e ms %.2f gram %.3f | mixed value ms %.2f e2e ms %.2f gram %.3f | clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'], o['ms_per_step'], o['e2e']['ms_per_step'], o['roofline']['gram_ms'], d['clocks']['sm_mhz']))"; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_gputests3.log; cat gpurun_out/r2_gputests3.log
BENCH_PHASES=1 python bench.py --workload sweep --stat all --steps 1 --warmup 1 > gpurun_out/r2_bench_sweep_all.json 2> gpurun_out/r2_bench_sweep_all.err; grep phases gpurun_out/r2_bench_sweep_all.err | tail -1
python bench.py --workload sweep --steps 1 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r2_bench_sweep_gt.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sweep_gt.json')); print('sweep GT alone ms', d['ms_per_step']); d=json.load(open('gpurun_out/r2_bench_sweep_all.json')); print('sweep all ms', d['ms_per_step'], d['contraction'], d['config']['histogram'])"
python bench.py --workload lsu --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_lsu_1gpu.json 2> gpurun_out/r2_bench_lsu_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_lsu_1gpu.json')); print('lsu 1 gpu value %.3g ms %.2f e2e %.3g ms %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
