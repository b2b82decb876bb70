#!/bin/bash
# One GPU call: smoke(), the whole GPU suite and a short bench run.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1b_smoke.log
tail -n 2 gpurun_out/r1b_smoke.log
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r1b_gpu_all.log 2>&1
echo "all exit $?" >> gpurun_out/r1b_gpu_all.log
tail -n 4 gpurun_out/r1b_gpu_all.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r1b_bench_final.json 2> gpurun_out/r1b_bench_final.err
head -c 400 gpurun_out/r1b_bench_final.json; tail -n 3 gpurun_out/r1b_bench_final.err
