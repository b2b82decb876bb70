"""One small-shape pass through every kernel family of the library, for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck).  Usage (GPU box):  compute-sanitizer --tool memcheck python tools/sanitize_driver.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
N, L, R = 96, 44, 3
msa, wgt, partner, tree = pkg.synth.synthetic_family(N, L, seed=5)
rng = np.random.default_rng(5)
nulls = np.stack([msa[:, rng.permutation(L)] for _ in range(R)])

for S, Snull in ((4, 0), (4, 2), (2, 0), (1, 0)):
    ctx = pkg.Context(0)
    ctx.set_null_slices(Snull)
    ctx.configure(N, L, 2, S)
    ctx.set_weights(wgt if S > 1 else None)
    for st in (pkg.GT, pkg.MI, pkg.MIr, pkg.MIg, pkg.CHI, pkg.OMES, pkg.RAF, pkg.RAFS, pkg.CCF):
        ctx.scan(msa, st, pkg.C16, pkg.APC, want_probs=(st == pkg.GT))
    ctx.scan(msa, pkg.GT, pkg.C2, pkg.ASC)
    ctx.scan(msa, pkg.GT, pkg.CWC, pkg.NOCORR)
    ctx.hist_reset()
    w, _, _ = ctx.null_width(nulls[0])
    ctx.null_hist(nulls, 0.01, pkg.RAFS)
    ctx.null_hist(nulls, 0.01, pkg.MI, pkg.C16, pkg.ASC)           # count epilogue + stat_kernel
    ctx.last_nseff()
    ctx.null_hist(nulls, w)                                        # record epilogue (pair-per-thread form for S = 1, 2, 4)
    ctx.hist_read(4000)
    ctx.last_nseff()
    combos = [(pkg.GT, pkg.APC), (pkg.MI, pkg.ASC), (pkg.CHI, pkg.APC), (pkg.OMES, pkg.NOCORR), (pkg.MIr, pkg.APC), (pkg.MIg, pkg.APC)]
    ctx.hist_reset_multi()
    ctx.null_hist_multi(nulls, combos, [0.05, 0.001, 0.05, 0.01, 0.001, 0.001])
    ctx.hist_read_multi(2, 1000)
    ctx.null_hist_multi(nulls, [(pkg.RAFS, pkg.APC), (pkg.RAF, pkg.ASC)], [0.01, 0.01])
    # generators + pool
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    ctx.null_fitch_shuffle(msa, 7, R)
    ctx.null_hist_pool(0, R, w)
    Q = np.array([[-1.00, 0.30, 0.50, 0.20], [0.25, -0.90, 0.15, 0.50], [0.60, 0.10, -0.95, 0.25], [0.20, 0.45, 0.30, -0.95]])
    ctx.null_simulate(Q, np.where(msa[0] < 4, msa[0], 0).astype(np.uint8), 9, R, gapmask=msa)
    ctx.null_hist_pool(0, R, w)
    # E-values / hit list, structure histograms
    res = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)
    bins, _, _ = ctx.hist_read(4000)
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    ctx.scan_hist(0.05, -10.0, 4000, mask)
    ctx.scan_hits(-10.0, 0.05, bins, float(res["maxcov"]), L * (L - 1) // 2, int(mask.sum()), mask, thresh=10.0)
    ctx.close()

# substitution counts over the tree; preprocessing
po = ge.load_oracle()
ora = po.Oracle()
r = ora.rng(3)
_, _, _, otree = po.synthetic_family(N, L, seed=5)              # the oracle's own tree type (same family, same seed)
_, allm, _ = ora.null_fitch_shuffle(r, otree, msa, want_all=True)
ora.rng_free(r)
ctx = pkg.Context(0)
ctx.configure(2 * (N - 1), L, 1, 1)
ctx.tree_substitutions(tree.left, tree.right, allm[:N], allm[N:])
ctx.close()
ctx = pkg.Context(0)
ctx.configure(N, L, 1, 0)
use = ctx.msa_gap_columns(msa, wgt)
ctx.msa_column_subset(msa, use)
ctx.msa_pb_weights(msa)
ctx.msa_pair_identity(msa)
ctx.close()
print("sanitize driver done")
