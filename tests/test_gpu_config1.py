"""BASELINE config 1 end to end on the DEVICE: `R-scape -s tutorial/updated_Arisong.sto` with R-scape's defaults.

Stockholm alignment (fixture) -> gap-column filter -> GSC weights -> FastTree tree rooted at the midpoint (fixture made by the
reference's own code, tests/test_config1_host.py) -> 20 tree-shuffled nulls on the device (generator A) -> width pass + cumulative
null histogram on the device -> scan of the input alignment -> two-set histograms with the SS_cons structure mask -> gamma tail fit
(host, cov_NullFit_b200) -> E-values and the significant-pair list on the device.

What must hold for every seed of the null generator:
  * the device flow and the CPU oracle's flow on the SAME (device-generated) nulls call the identical set of pairs with bit-equal
    E-values -- north_star's "identical significant pairs when the same null alignments are supplied";
  * the 11 significant pairs of documentation/tutorial.tex:187-212 are all called, with the transcript's scores to 5 decimals;
  * a call beyond those 11 can only be the transcript's borderline pair (104,130): with 20 shuffles its E-value scatters around
    0.18 (0.086 - 0.38 over 16 seeds of the REFERENCE's own generator, 0.03 - 0.29 over the device generator's seeds;
    tools/c1_generator_check.py, tools/c1_debug.py), i.e. a 20-shuffle null sample puts it below E = 0.05 now and then whatever
    the generator.  Seed 3 of the device generator is such a sample (fitted tail lambda 0.32 against 0.10 - 0.22)."""
import numpy as np
import pytest

import _config1 as c1

pytestmark = pytest.mark.gpu

BORDERLINE = {(104, 130)}          # see the module docstring


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("null_slices", [0, 2])
def test_tutorial_significant_pairs_on_the_device(ctx, pkg, po, seed, null_slices):
    sub, wgt, keep, mask, tree, gold = c1.load(po)
    N, L = sub.shape
    P = L * (L - 1) // 2
    ctx.set_null_slices(null_slices)
    ctx.configure(N, L, 4, 4)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(c1.NSHUFFLE)
    ctx.null_fitch_shuffle(sub, 1000 + seed, c1.NSHUFFLE)
    ctx.hist_reset()
    w, lo, hi = ctx.null_width_pool(0)                                  # calculate_width_histo on the first null
    assert w == c1.null_width(lo, hi) == 0.05
    mm = ctx.null_hist_pool(0, c1.NSHUFFLE, w)                          # run_rscape(RANSS) + null_add2cumranklist
    res = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC)                       # run_rscape(GIVSS)
    xmax = float(mm[:, 1].max())
    nb = c1.null_bins_needed(w, xmax, res["maxcov"])
    bins, n, _ = ctx.hist_read(nb)
    assert n == c1.NSHUFFLE * P == int(bins.sum())
    ha, hb, ht = ctx.scan_hist(w, c1.BMIN, nb, mask)                   # the input alignment's own histograms: Nb / Nt of the two-set test
    Nb, Nt = int(hb.sum()), int(ht.sum())
    assert (Nb, Nt) == (gold["nbpairs"], P - gold["nbpairs"]) and int(ha.sum()) == P
    fit = po.nullfit_host(po.NullFit(c1.BMIN, w, bins, xmax=xmax), c1.PMASS, c1.FRACFIT, False)
    hits = ctx.scan_hits(fit.bmin, fit.w, fit.obs, fit.xmax, Nt, Nb, mask, fit.survfit, fit.phi, thresh=c1.ETHRESH)
    want = {(p["i"], p["j"]): p for p in gold["pairs"]}
    called = c1.called_pairs(hits, keep)
    assert set(want) <= called and called - set(want) <= BORDERLINE, sorted(called ^ set(want))
    for i, j, sc, ev in zip(hits["i"], hits["j"], hits["sc"], hits["eval"]):
        assert ev < c1.ETHRESH
        p = want.get((int(keep[i]) + 1, int(keep[j]) + 1))
        if p is not None:
            assert round(float(sc), 5) == p["score"]
    # the CPU oracle's flow on the same nulls (oracle scans + histogram, same tail fit, the oracle's hit list): identical calls
    if null_slices == 0:
        oracle = po.Oracle()
        cum = None
        for m in ctx.pool_get(c1.NSHUFFLE, 0):
            r = oracle.scan(m, wgt, po.GT, po.C16, po.APC)
            h = oracle.hist_from_cov(r["cov"], r["maxcov"], c1.BMIN, w, c1.TOL)
            cum = oracle.accumulate(cum, h)
            oracle.free(h)
        view = oracle.view(cum)
        oracle.free(cum)
        obs = np.zeros(nb, np.uint64)
        obs[:min(nb, view.nb)] = view.obs[:nb]
        assert np.array_equal(obs, bins)
        real = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
        ref = oracle.hitlist(real["cov"], fit, mask, Nb, Nt, -1, c1.ETHRESH)
        assert [(int(a), int(b)) for a, b in zip(ref["i"], ref["j"])] == [(int(a), int(b)) for a, b in zip(hits["i"], hits["j"])]
        assert np.allclose(ref["eval"], hits["eval"], rtol=1e-6, atol=0.0)    # (the scores themselves agree to 1e-9)
    # E-values read off the empirical part of the null agree with the transcript within the noise of 20 shuffles (a factor of a few);
    # those extrapolated into the fitted tail depend on the RNG stream by orders of magnitude and are not compared
    emp = [(ev, want[(int(keep[i]) + 1, int(keep[j]) + 1)]["evalue"]) for i, j, ev in zip(hits["i"], hits["j"], hits["eval"])
           if (int(keep[i]) + 1, int(keep[j]) + 1) in want and want[(int(keep[i]) + 1, int(keep[j]) + 1)]["evalue"] > 1e-3]
    assert emp and all(0.1 < a / b < 10.0 for a, b in emp), emp
