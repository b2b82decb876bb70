#!/bin/bash
# replicates per launch in the sharded pipeline: tools/r2_slots.sh N   (gpurun --gpus N)
cd "$(dirname "$0")/.."
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@"; }
for wl in ssu lsu; do
  for slots in 2 4 8; do
    [ $wl = lsu ] && [ $slots = 8 ] && continue
    RSCAPE_B200_TRACE=1 run --workload $wl --grid-shard --slots $slots --steps 2 --warmup 2 --no-alt --no-cpu-baseline > gpurun_out/slots.json 2> gpurun_out/slots.err
    python -c "
import json; d=json.load(open('gpurun_out/slots.json')); print('$wl grid-shard N=$N slots=$slots: value %.3g ms %.2f e2e ms %.2f gram_ms %.3f share %.2f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'], d['roofline']['gram_share_of_step']))" || tail -5 gpurun_out/slots.err
    grep "rsb\] gram" gpurun_out/slots.err | tail -1
  done
done
