"""Diagnostic for tests/test_gpu_config1.py: the device flow and the CPU flow on the SAME device-generated nulls, per seed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import _config1 as c1  # noqa: E402

pkg = ge.load_package()
po = ge.load_oracle()
oracle = po.Oracle()
sub, wgt, keep, mask, tree, gold = c1.load(po)
N, L = sub.shape
P = L * (L - 1) // 2
real = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
for seed in (1, 2, 3, 4):
    ctx = pkg.Context(0)
    ctx.configure(N, L, 4, 4)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(c1.NSHUFFLE)
    ctx.null_fitch_shuffle(sub, 1000 + seed, c1.NSHUFFLE)
    ctx.hist_reset()
    w, lo, hi = ctx.null_width_pool(0)
    mm = ctx.null_hist_pool(0, c1.NSHUFFLE, w)
    res = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC)
    xmax = float(mm[:, 1].max())
    nb = c1.null_bins_needed(w, xmax, res["maxcov"])
    bins, n, _ = ctx.hist_read(nb)
    ha, hb, ht = ctx.scan_hist(w, c1.BMIN, nb, mask)
    Nb, Nt = int(hb.sum()), int(ht.sum())
    fit = po.nullfit_host(po.NullFit(c1.BMIN, w, bins, xmax=xmax), c1.PMASS, c1.FRACFIT, False)
    hits = ctx.scan_hits(fit.bmin, fit.w, fit.obs, fit.xmax, Nt, Nb, mask, fit.survfit, fit.phi, thresh=2000.0)
    E = {(int(keep[i]) + 1, int(keep[j]) + 1): e for i, j, e in zip(hits["i"], hits["j"], hits["eval"])}
    # CPU flow on the same nulls
    nulls = ctx.pool_get(c1.NSHUFFLE, 0)
    cum, xm = None, -np.inf
    for m in nulls:
        r = oracle.scan(m, wgt, po.GT, po.C16, po.APC)
        h = oracle.hist_from_cov(r["cov"], r["maxcov"], -10.0, 0.05, 1e-6)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
        xm = max(xm, r["maxcov"])
    view = oracle.view(cum)
    nb2 = c1.null_bins_needed(0.05, xm, real["maxcov"])
    obs = np.zeros(nb2, np.uint64)
    obs[:view.nb] = view.obs[:nb2]
    fit2 = po.nullfit_host(po.NullFit(-10.0, 0.05, obs, xmax=xm), c1.PMASS, c1.FRACFIT, False)
    ev = oracle.hitlist(real["cov"], fit2, mask, int(mask.sum()), P - int(mask.sum()), -1, 2000.0)
    E2 = {(int(keep[i]) + 1, int(keep[j]) + 1): e for i, j, e in zip(ev["i"], ev["j"], ev["eval"])}
    print(f"seed {seed}: device  w {w} xmax {xmax:.6f} nb {nb} lam {fit.lam:.5f} tau {fit.tau:.5f} Nb {Nb} Nt {Nt} E(104,130) {E.get((104, 130))} E(97,107) {E.get((97, 107))} "
          f"bins equal {np.array_equal(bins[:min(nb, nb2)], obs[:min(nb, nb2)])} ncalled {sum(1 for e in E.values() if e < 0.05)}")
    print(f"         cpu     xmax {xm:.6f} nb {nb2} lam {fit2.lam:.5f} tau {fit2.tau:.5f} E(104,130) {E2.get((104, 130))} E(97,107) {E2.get((97, 107))} ncalled {sum(1 for e in E2.values() if e < 0.05)}", flush=True)
    ctx.close()
