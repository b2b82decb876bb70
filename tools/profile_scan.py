"""Profiling driver: a few null-histogram scans of one synthetic alignment of a bench workload, so that ncu sees every
kernel of the scan path a handful of times.  Usage (GPU box):
    ncu --set full --clock-control none -k regex:'stat_kernel|marg|correct_hist|pack|cov' -c 12 \
        -o gpurun_out/aux python tools/profile_scan.py ssu"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "ssu"
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 4
stat = getattr(pkg, sys.argv[3]) if len(sys.argv) > 3 else pkg.GT
w = bench.WORKLOADS[name]
N, L = w["N"], w["L"]
rng = np.random.default_rng(1)
msa = rng.integers(0, 5, (nrep, N, L)).astype(np.uint8)
ctx = pkg.Context(0)
ctx.configure(N, L, 2, 4)
ctx.set_weights(rng.gamma(2.0, 0.5, N))
ctx.hist_reset()
width = ctx.null_width(msa[0], stat)[0]
ctx.null_hist(msa, width, stat)
print("bins", int(ctx.hist_read(4000)[1]))
ctx.close()
