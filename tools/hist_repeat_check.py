"""Diagnostic: is the cumulative null histogram of the pipelined null loop reproducible run to run (a race between the
contraction and the statistics chain would show as differing bins), and does it equal the CPU oracle's on the same nulls?
    python tools/hist_repeat_check.py config1 | ssu"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
po = ge.load_oracle()
which = sys.argv[1] if len(sys.argv) > 1 else "config1"
if which == "config1":
    import _config1 as c1
    sub, wgt, keep, mask, tree, gold = c1.load(po)
    R, slots, seeds = 20, 4, (1001, 1002, 1003, 1004)
else:
    sub, wgt, _, tree = pkg.synth.synthetic_family(10000, 1800, seed=42)
    R, slots, seeds = 24, 2, (7,)
N, L = sub.shape
NB = 6000
for snull in (0, 2):
    ctx = pkg.Context(0)
    ctx.set_null_slices(snull)
    ctx.configure(N, L, slots, 4)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    for seed in seeds:
        ctx.null_fitch_shuffle(sub, seed, R)
        runs = []
        for rep in range(6):
            ctx.hist_reset()
            mm = ctx.null_hist_pool(0, R, 0.05)
            bins, n, _ = ctx.hist_read(NB)
            runs.append((bins.copy(), mm.copy()))
        nd = [int(np.abs(runs[k][0].astype(np.int64) - runs[0][0].astype(np.int64)).sum()) for k in range(1, 6)]
        md = [float(np.abs(runs[k][1] - runs[0][1]).max()) for k in range(1, 6)]
        msg = f"{which} null slices {snull} seed {seed}: sum |bins(run k) - bins(run 0)| = {nd}, max |minmax diff| = {md}"
        if which == "config1" and snull == 0:
            oracle = po.Oracle()
            nulls = ctx.pool_get(R, 0)
            cum = None
            for m in nulls:
                r = oracle.scan(m, wgt, po.GT, po.C16, po.APC)
                h = oracle.hist_from_cov(r["cov"], r["maxcov"], -10.0, 0.05, 1e-6)
                cum = oracle.accumulate(cum, h)
                oracle.free(h)
            view = oracle.view(cum)
            ref = np.zeros(NB, np.int64)
            ref[:min(NB, view.nb)] = view.obs[:NB]
            msg += f"; vs oracle on the same nulls: sum |diff| = {int(np.abs(runs[0][0].astype(np.int64) - ref).sum())}"
        print(msg, flush=True)
    ctx.close()
