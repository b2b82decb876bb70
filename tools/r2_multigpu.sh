#!/bin/bash
# multi-GPU lines of round 2: tools/r2_multigpu.sh N  (run under gpurun --gpus N)
cd "$(dirname "$0")/.."
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
one() { python bench.py --gpus 1 "$@"; }
L=$( [ "$N" = 1 ] && echo one || echo run )
BENCH_PHASES=1 $L --steps 3 --warmup 3 --no-alt > gpurun_out/r2_bench_ssu_${N}gpu.json 2> gpurun_out/r2_bench_ssu_${N}gpu.err
$L --steps 3 --warmup 3 --no-alt --null-slices 2 > gpurun_out/r2_bench_ssu_mixed_${N}gpu.json 2>> gpurun_out/r2_bench_ssu_${N}gpu.err
$L --workload lsu --steps 2 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r2_bench_lsu_${N}gpu.json 2> gpurun_out/r2_bench_lsu_${N}gpu.err
if [ "$N" != 1 ]; then
  run --workload lsu --grid-shard --steps 2 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r2_bench_lsu_gridshard_${N}gpu.json 2>> gpurun_out/r2_bench_lsu_${N}gpu.err
fi
for f in gpurun_out/r2_bench_ssu_${N}gpu.json gpurun_out/r2_bench_ssu_mixed_${N}gpu.json gpurun_out/r2_bench_lsu_${N}gpu.json gpurun_out/r2_bench_lsu_gridshard_${N}gpu.json; do
  [ -s $f ] && python -c "
import json,sys; d=[json.loads(l) for l in open('$f') if l[0]=='{'][-1]; print('$f'.split('/')[-1], 'value %.3g ms %.2f e2e %.3g ms %.2f frac %.3f share %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['gram_share_of_step']))"
done
grep -h "phases" gpurun_out/r2_bench_ssu_${N}gpu.err | tail -$((2*N))
tail -3 gpurun_out/r2_bench_lsu_${N}gpu.err
if [ "$N" != 1 ]; then
  RSCAPE_B200_PEER_REDUCE=0 run --workload lsu --grid-shard --steps 2 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r2_bench_lsu_gridshard_${N}gpu_nccl.json 2>> gpurun_out/r2_bench_lsu_${N}gpu.err
  python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/r2_bench_lsu_gridshard_${N}gpu_nccl.json') if l[0]=='{'][-1]; print('lsu grid-shard with ncclAllReduce: value %.3g ms %.2f e2e %.3g ms %.2f share %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['gram_share_of_step']), d['config'].get('collectives'))"
fi
