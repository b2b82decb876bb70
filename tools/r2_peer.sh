#!/bin/bash
# the one-shot peer-memory all-reduce against NCCL on a sharded pair grid: tools/r2_peer.sh N   (gpurun --gpus N)
cd "$(dirname "$0")/.."
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
if [ "$N" = 2 ]; then timeout 300 python -m pytest tests/test_gpu_peer_reduce.py tests/test_gpu_sharded.py -q -x 2>&1 | tail -5; fi
for wl in lsu ssu; do
  for peer in 1 0; do
    RSCAPE_B200_PEER_REDUCE=$peer run --workload $wl --grid-shard --steps 2 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r2_bench_${wl}_gridshard_${N}gpu_peer${peer}.json 2> gpurun_out/r2_bench_${wl}_gridshard_${N}gpu_peer${peer}.err
    python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${wl}_gridshard_${N}gpu_peer${peer}.json')); print('$wl grid-shard N=$N peer=$peer: value %.3g ms %.2f e2e %.3g ms %.2f gram_ms %.3f share %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'], d['roofline']['gram_share_of_step']), d['config'].get('collectives'), d['config']['histogram'])" || tail -5 gpurun_out/r2_bench_${wl}_gridshard_${N}gpu_peer${peer}.err
  done
done
