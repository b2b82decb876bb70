"""Several statistics from one contraction per null (rsb_null_hist_multi, BASELINE config 5: the statistic sweep).

cov_Calculate (src/covariation.c:100-258) computes corr_Probs once and dispatches to ONE corr_Calculate*; the sweep repeats the
whole scan per (statistic, correction).  The device evaluates every requested statistic from the same count planes; each
combination's histogram must be the one its own single-statistic null loop leaves, and the oracle's."""
import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu

STATS = ["CHI", "OMES", "GT", "MI", "MIr", "MIg"]


def _width(lo, hi, w_old=0.05, bmin=-10.0, hpts=400, tol=1e-6):
    """calculate_width_histo, src/R-scape.c:1355-1360"""
    w = min(w_old, (hi - max(bmin, lo)) / hpts)
    return 0.0 if w < tol else w


@pytest.mark.parametrize("cls", ["C16", "C2"])
def test_multi_equals_single_statistic_runs_and_oracle(ctx, pkg, po, oracle, cls):
    from test_gpu_nulls import oracle_null_loop
    N, L, R = 180, 60, 5
    wgt = po.synthetic_msa(N, L, seed=5)[1]
    nulls = np.stack([po.synthetic_msa(N, L, seed=900 + r)[0] for r in range(R)])
    combos = [(s, a) for s in STATS for a in ("APC", "ASC")] + [("GT", "NOCORR")]
    pc = [(getattr(pkg, s), getattr(pkg, a)) for s, a in combos]
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    # width pass: the score range of the first null for every combination, from one contraction
    mm0 = ctx.null_hist_multi(nulls[:1], pc, [0.0] * len(pc), getattr(pkg, cls))
    ref = [oracle_null_loop(po, oracle, nulls, wgt, getattr(po, s), getattr(po, cls), getattr(po, a)) + (oracle_null_loop.scores,) for s, a in combos]
    for k in range(len(pc)):
        assert abs(_width(mm0[k, 0, 0], mm0[k, 0, 1]) - ref[k][0]) <= 1e-9 * max(ref[k][0], 1e-3), combos[k]
    w = [r[0] for r in ref]                              # the oracle's widths, so that bins can be compared one to one
    ctx.hist_reset_multi()
    mm = ctx.null_hist_multi(nulls, pc, w, getattr(pkg, cls))
    P = L * (L - 1) // 2
    for k, (s, a) in enumerate(combos):
        # (1) the oracle's loop for this combination alone
        w_ref, view, mm_ref, scores = ref[k]
        bins, n, imax = ctx.hist_read_multi(k, view.nb + 8)
        assert n == R * P == int(bins.sum()), (s, a)
        assert_bins_identical(bins, view.obs, scores, -10.0, w_ref, scale=max(1.0, float(np.max(np.abs(mm_ref)))))
        assert np.max(np.abs(mm[k] - mm_ref)) <= 1e-9 * max(1.0, np.max(np.abs(mm_ref))), (s, a)
        # (2) the device's own single-statistic loop with the same width: identical integer bins
        ctx.hist_reset()
        mm1 = ctx.null_hist(nulls, w[k], getattr(pkg, s), getattr(pkg, cls), getattr(pkg, a))
        bins1, n1, _ = ctx.hist_read(view.nb + 8)
        assert n1 == n
        # same code per statistic, but compiled into another kernel (GT x C16 of the single run even goes through the record epilogue):
        # values may differ in the last bits, so bins are compared with the bin-edge rule and the score range to 1e-12
        assert_bins_identical(bins, bins1, scores, -10.0, w[k], rel=1e-12, scale=max(1.0, float(np.max(np.abs(mm_ref)))))
        assert np.max(np.abs(mm[k] - mm1)) <= 1e-12 * max(1.0, np.max(np.abs(mm1))), (s, a)


def test_multi_on_pool_entries_and_accumulation(ctx, pkg, po):
    N, L, R = 150, 48, 6
    msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=3)
    ctx.configure(N, L, 4, 0)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    ctx.null_fitch_shuffle(msa, 99, R)
    pc = [(pkg.GT, pkg.APC), (pkg.MI, pkg.APC), (pkg.OMES, pkg.ASC)]
    w = [0.05, 0.001, 0.01]
    ctx.hist_reset_multi()
    mm = ctx.null_hist_multi(R, pc, w, pkg.C16, first_rep=0)
    nulls = ctx.pool_get(R, 0)
    mm2 = ctx.null_hist_multi(nulls, pc, w, pkg.C16)              # host copies of the same alignments, on top: twice the counts
    assert np.array_equal(mm, mm2)
    P = L * (L - 1) // 2
    for k in range(3):
        bins, n, _ = ctx.hist_read_multi(k, 40000)
        assert n == 2 * R * P == int(bins.sum())
        assert not (bins % 2).any()
    ctx.hist_reset_multi()
    assert ctx.hist_read_multi(1, 1000)[1] == 0


def test_multi_unit_weight_statistics_share_their_own_contraction(ctx, pkg, po, oracle):
    """RAF / RAFS (unweighted, src/correlators.c:877-982) x corrections from one unit-weight contraction per null: identical to the
    single-statistic loops, which are bit-exact against the reference's O(N^2) loop (test_gpu_scan_parity::test_raf_bit_exact)."""
    N, L, R = 90, 40, 4
    wgt = po.synthetic_msa(N, L, seed=2)[1]
    nulls = np.stack([po.synthetic_msa(N, L, seed=40 + r)[0] for r in range(R)])
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    combos = [(pkg.RAFS, pkg.APC), (pkg.RAFS, pkg.ASC), (pkg.RAF, pkg.APC), (pkg.RAFS, pkg.NOCORR)]
    w = [0.01, 0.01, 0.01, 0.01]
    ctx.hist_reset_multi()
    mm = ctx.null_hist_multi(nulls, combos, w)
    for k, (st, ac) in enumerate(combos):
        ctx.hist_reset()
        mm1 = ctx.null_hist(nulls, w[k], st, pkg.C16, ac)
        bins1, n1, _ = ctx.hist_read(6000)
        bins, n, _ = ctx.hist_read_multi(k, 6000)
        assert n == n1 == R * L * (L - 1) // 2 == int(bins.sum())
        assert np.array_equal(bins, bins1) and np.array_equal(mm[k], mm1), (st, ac)


def test_multi_rejects_what_cannot_share_a_contraction(ctx, pkg, po):
    N, L = 60, 30
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(None)
    nulls = po.synthetic_msa(N, L, seed=1)[0][None]
    with pytest.raises(pkg.RscapeB200Error):
        ctx.null_hist_multi(nulls, [(pkg.RAFS, pkg.APC), (pkg.MI, pkg.APC)], [0.05, 0.05])   # unit-weight and weighted counts
    with pytest.raises(pkg.RscapeB200Error):
        ctx.null_hist_multi(nulls, [(pkg.CCF, pkg.APC)], [0.05])
    with pytest.raises(pkg.RscapeB200Error):
        ctx.null_hist_multi(nulls, [(pkg.MI, pkg.APC)], [0.05], pkg.CWC)       # CWC is defined for the G test only
