cd /root/repo
for lms in 25 100 200; do sed -i "s/\"-lms\", \"[0-9]*\"/\"-lms\", \"$lms\"/" bench.py; BENCH_PHASES=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-alt --nulls 13 2>&1 >/dev/null | grep "phases ms: nulls" | tail -2; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt | python -c "
import json,sys; d=json.load(sys.stdin); print('lms $lms', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'], d['clocks'])"; done
BENCH_NO_SAMPLER=1 BENCH_PHASES=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-alt --nulls 13 2>&1 >/dev/null | grep "phases ms: nulls" | tail -2
BENCH_NO_SAMPLER=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt | python -c "
import json,sys; d=json.load(sys.stdin); print('no sampler', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['gram_ms'])"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
BENCH_PHASES=1 python bench.py --workload sweep --stat all --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "phases" | tail -1
