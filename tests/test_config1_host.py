"""Host steps around the hot path that BASELINE config 1 needs end to end (R-scape's defaults on a Stockholm alignment): FastTree's
Newick -> rooted tree, and the tail fit of the cumulative null histogram -> survival table for the E-values.

Both are Easel routines in the reference (un-vendored, unpinned); the shim restates them (r-scape_b200/host/easel_shim_fit.c) and
the REFERENCE's own callers run on top: Tree_ReorderTaxaAccordingMSA + Tree_RootAtMidPoint (src/msatree.c:524-912) and
cov_histogram_pmass + cov_NullFitGamma / cov_NullFitExponential (src/covariation.c:459-487, 1915-1973), compiled unchanged into
oracle/_ref.  The product's mirror of the fit block (cov_NullFit_b200) must agree with the reference's bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _gamma_tail_hist(po, lam=0.15, tau=1.1, n=2_000_000, seed=5):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.normal(-2.0, 3.0, n), 20.0 + rng.gamma(tau, 1.0 / lam, n // 500)])
    x = np.maximum(x, -10 + 0.05)
    b = np.ceil((x + 10) / 0.05 - 1).astype(np.int64)
    obs = np.bincount(b, minlength=int(b.max()) + 8).astype(np.uint64)
    return po.NullFit(-10.0, 0.05, obs, xmax=float(x.max()))


def test_fasttree_fixture_is_what_the_reference_roots(po, reflib):
    """tests/golden/arisong_fasttree.npz = shim Newick reader + the reference's own re-ordering and midpoint rooting of the committed
    FastTree output; the tree is a valid rooted binary tree over the alignment's rows with the root at the midpoint."""
    z = np.load(os.path.join(GOLD, "arisong_fasttree.npz"))
    names = [str(x) for x in z["names"]]
    t = reflib.tree_from_newick(os.path.join(GOLD, "arisong_fasttree.nwk"), names, rootatmid=True)
    for k in ("left", "right", "parent", "ld", "rd"):
        assert np.array_equal(getattr(t, k), z[k]), k
    N = len(names)
    taxa = sorted(-c for c in np.concatenate([t.left, t.right]) if c <= 0)
    assert taxa == list(range(N))
    assert all(t.parent[c] == v for v in range(N - 1) for c in (t.left[v], t.right[v]) if c > 0)
    assert all(c > v for v in range(N - 1) for c in (t.left[v], t.right[v]) if c > 0)        # preorder numbering

    def depth(c):
        return 0.0 if c <= 0 else max(t.ld[c] + depth(t.left[c]), t.rd[c] + depth(t.right[c]))
    a, b = t.ld[0] + depth(t.left[0]), t.rd[0] + depth(t.right[0])
    assert abs(a - b) <= 1e-6 * max(a, b)                                                     # Tree_FindMidPoint works in float
    # unrooted, the tree keeps FastTree's total branch length
    nwk = open(os.path.join(GOLD, "arisong_fasttree.nwk")).read()
    import re
    total = sum(max(float(x), 0.0) for x in re.findall(r":(-?[0-9.eE+-]+)", nwk))
    assert abs(float(t.ld.sum() + t.rd.sum()) - total) <= 1e-5 * total


@pytest.mark.parametrize("doexpfit", [False, True])
def test_host_fit_block_equals_the_references(po, reflib, doexpfit):
    null = _gamma_tail_hist(po)
    for pmass, fracfit in ((0.0005, 1.0), (0.002, 1.0), (0.05, 0.3)):
        a = po.nullfit_host(null, pmass, fracfit, doexpfit)
        b = reflib.nullfit(null, pmass, fracfit, doexpfit)
        assert (a.cmin, a.phi, a.newmass, a.mu, a.lam, a.tau) == (b.cmin, b.phi, b.newmass, b.mu, b.lam, b.tau)
        assert np.array_equal(a.survfit, b.survfit)
        # the censored tail holds at least the requested mass, and one bin fewer would hold less
        cum = np.cumsum(null.obs[::-1].astype(np.float64))[::-1]
        assert cum[a.cmin] / null.Nc == pytest.approx(a.newmass)
        if fracfit >= 1.0:
            assert a.newmass >= pmass and cum[a.cmin + 1] / null.Nc < pmass
        assert a.survfit[a.cmin - 1] == 0.0 and a.survfit[a.cmin] > 0.0
        assert np.all(np.diff(a.survfit[a.cmin:]) <= 0.0)


def test_gamma_fit_recovers_a_gamma_tail(po):
    lam, tau = 0.15, 1.1
    null = _gamma_tail_hist(po, lam, tau)
    fit = po.nullfit_host(null, pmass=0.0015, fracfit=1.0, doexpfit=False)      # tail = mostly the planted gamma (mass 0.002 from 20 on)
    assert fit.mu == fit.phi and 19.0 < fit.phi < 24.0
    # a gamma conditioned on x > phi is not the same gamma, so compare survival ratios instead of parameters: the fitted tail
    # must decay like the planted one far out, where lambda dominates
    b1, b2 = int((40 + 10) / 0.05), int((60 + 10) / 0.05)
    decay = np.log(fit.survfit[b1] / fit.survfit[b2]) / 20.0
    assert abs(decay - lam) < 0.25 * lam, (decay, fit.lam, fit.tau)
    ex = po.nullfit_host(null, pmass=0.0015, fracfit=1.0, doexpfit=True)
    assert 0.5 * lam < ex.lam < 2.0 * lam


def test_incomplete_gamma_against_scipy(po):
    sp = pytest.importorskip("scipy.special")
    lib = C.CDLL(os.path.join(os.path.dirname(HERE), "r-scape_b200", "librscape_b200_host.so"))
    lib.esl_stats_IncompleteGamma.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(3)
    for a, x in zip(np.exp(rng.uniform(-3, 4, 300)), np.exp(rng.uniform(-6, 5, 300))):
        p, q = C.c_double(), C.c_double()
        assert lib.esl_stats_IncompleteGamma(a, x, C.byref(p), C.byref(q)) == 0
        assert p.value == pytest.approx(float(sp.gammainc(a, x)), rel=1e-10, abs=1e-300)
        assert q.value == pytest.approx(float(sp.gammaincc(a, x)), rel=1e-10, abs=1e-300)
