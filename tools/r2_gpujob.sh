cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_gputests2.log; tail -12 gpurun_out/r2_gputests2.log
python tools/gen_time.py 100 2>&1 | tail -4
RSCAPE_B200_REPLAY=rank python tools/gen_time.py 100 2>&1 | tail -3
python tools/gram_time.py ssu 2,4 2>&1 | tail -4
python bench.py > gpurun_out/r2_bench_ssu_1gpu.json 2> gpurun_out/r2_bench_ssu_1gpu.err; tail -2 gpurun_out/r2_bench_ssu_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ssu_1gpu.json')); print('strict', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['gram_ms'], d['roofline']['gram_share_of_step']); o=d['other_precision_mode']; print('mixed', o['ms_per_step'], o['e2e']['ms_per_step'], o['roofline']['frac'], o['roofline']['gram_ms'], o['roofline']['gram_share_of_step'])"
BENCH_PHASES=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-alt --nulls 13 > gpurun_out/r2_bench_13nulls.json 2> gpurun_out/r2_bench_13nulls.err; grep phases gpurun_out/r2_bench_13nulls.err | tail -2
BENCH_PHASES=1 python bench.py --workload sweep --stat all --steps 1 --warmup 1 > gpurun_out/r2_bench_sweep_all.json 2> gpurun_out/r2_bench_sweep_all.err; grep phases gpurun_out/r2_bench_sweep_all.err | tail -1
python bench.py --workload sweep --steps 1 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r2_bench_sweep_gt.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sweep_gt.json')); print('sweep GT alone ms', d['ms_per_step']); d=json.load(open('gpurun_out/r2_bench_sweep_all.json')); print('sweep all ms', d['ms_per_step'], d['contraction'], d['config']['histogram'])"
