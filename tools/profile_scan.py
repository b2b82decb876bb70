"""Profiling driver: a few null-histogram scans of one synthetic alignment of a bench workload, so that ncu sees every
kernel of the scan path a handful of times: strict (nulls at the input's slices), mixed (nulls at 2 slices) and the
several-statistics-per-contraction loop.  Usage (GPU box):
    ncu --set full --clock-control none --import-source on -k regex:'gram_i8|gt_finish|pack_planes|correct_hist|multi_stat|marg_sum|stat_kernel' \
        -c 40 -o gpurun_out/r2_scan python tools/profile_scan.py ssu"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "ssu"
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
N, L = w["N"], w["L"]
rng = np.random.default_rng(1)
msa = torch.from_numpy(rng.integers(0, 5, (nrep, N, L)).astype(np.uint8)).cuda()
wgt = rng.gamma(2.0, 0.5, N)
for snull in (0, 2):
    ctx = pkg.Context(0)
    ctx.set_null_slices(snull)
    ctx.configure(N, L, 2, 4)
    ctx.set_weights(wgt)
    ctx.hist_reset()
    ctx.null_hist(msa, 0.05)
    print("null slices", snull, "bins", int(ctx.hist_read(4000)[1]))
    if snull == 0:
        ctx.scan(msa[0], pkg.GT, pkg.C16, pkg.APC, want_cov=False)                     # the input alignment's path: count epilogue + stat_kernel
        combos = [(s, a) for s in (pkg.GT, pkg.MI, pkg.MIr, pkg.MIg, pkg.CHI, pkg.OMES) for a in (pkg.APC, pkg.ASC)]
        ctx.hist_reset_multi()
        ctx.null_hist_multi(msa[:2], combos, [0.05] * len(combos))
    ctx.close()
