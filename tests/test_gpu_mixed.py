"""Mixed precision of the weighted counts (rsb_set_null_slices): the null alignments are contracted with fewer base-256 digits
of the sequence weights than the input alignment -- north_star's "split path with a stated bound".

What is exact: the nulls' counts are integer arithmetic on their own fixed-point weights wq' (bit for bit the oracle's counts on
wq'), and the null histogram equals the oracle's histogram computed with the weights wq' 2^-q' (doubles that hold them exactly).
What is bounded: against the input alignment's 4-slice weights the scores of a null move by
    |d score| <= MIXED_SCORE_BOUND * max(1, |score|)      (2 slices, ~21-bit weights; 2e-4, i.e. 4e-3 of a bin)
so that a fraction <= MIXED_BIN_FRACTION of a null's scores change histogram bin (bin width 0.05), and the set of significant
pairs of the input alignment is identical (its own scan never changes: same slices, same scores, bit for bit).
Reference: the nulls only feed the cumulative histogram ha (src/R-scape.c:1650-1697, src/covariation.c:415-435) from which
cov2evalue (:2370-2400) reads survival counts."""
import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu

MIXED_SCORE_BOUND = 2e-4      # measured 6.1e-5 (max |d score| 3.4e-4) on a whole SSU-shaped null; the tests print it with -s
MIXED_BIN_FRACTION = 1e-3     # measured 7.8e-5 of an SSU null's scores change bin; 1.1e-4 .. 1.4e-4 of the scores of 20 small nulls


def _mixed(pkg, N, L, slots, S, Snull, wgt):
    c = pkg.Context(0)
    c.set_null_slices(Snull)
    c.configure(N, L, slots, S)
    c.set_weights(wgt)
    return c


def test_null_counts_exact_on_their_own_weights(pkg, po, oracle):
    N, L = 300, 70
    msa, wgt, _ = po.synthetic_msa(N, L, seed=3)
    null = po.synthetic_msa(N, L, seed=4)[0]
    c = _mixed(pkg, N, L, 2, 4, 2, wgt)
    try:
        wq4, q4, S4 = c.quantisation()
        wq2, q2, S2, err2, bits2 = c.null_quantisation()
        assert (S4, S2) == (4, 2) and q2 < q4 and 19.0 < bits2 < 30.0
        assert np.max(np.abs(wq2 * 2.0 ** -q2 - wgt)) == pytest.approx(err2, rel=1e-12)
        # a null scored with a statistic that leaves the counts resident (MI: count epilogue)
        c.hist_reset()
        c.null_hist(null[None], 0.05, pkg.MI, want_minmax=False)
        assert np.array_equal(c.counts(), np.triu(oracle.counts_fixed(null, wq2).transpose(2, 0, 1), 1))
        # the input alignment keeps all four slices
        c.scan(msa, pkg.GT, pkg.C16, pkg.APC)
        assert np.array_equal(c.counts(), np.triu(oracle.counts_fixed(msa, wq4).transpose(2, 0, 1), 1))
    finally:
        c.close()


@pytest.mark.parametrize("stat", ["GT", "MI"])
def test_null_histogram_is_the_oracles_on_the_null_weights(pkg, po, oracle, stat):
    from test_gpu_nulls import oracle_null_loop
    N, L, R = 250, 70, 4
    wgt = po.synthetic_msa(N, L, seed=1)[1]
    nulls = np.stack([po.synthetic_msa(N, L, seed=100 + r)[0] for r in range(R)])
    c = _mixed(pkg, N, L, 2, 4, 2, wgt)
    try:
        wq2, q2, _, _, _ = c.null_quantisation()
        w2 = wq2 * 2.0 ** -q2                                         # exact doubles
        st = getattr(po, stat)
        w_ref, view, mm_ref = oracle_null_loop(po, oracle, nulls, w2, st, po.C16, po.APC)
        w, _, _ = c.null_width(nulls[0], getattr(pkg, stat))
        assert abs(w - w_ref) <= 1e-12
        c.hist_reset()
        mm = c.null_hist(nulls, w_ref, getattr(pkg, stat))
        bins, n, _ = c.hist_read(view.nb + 8)
        assert n == view.n
        assert_bins_identical(bins, view.obs, oracle_null_loop.scores, -10.0, w_ref)
        assert np.max(np.abs(mm - mm_ref)) <= 1e-9 * max(1.0, np.max(np.abs(mm_ref)))
    finally:
        c.close()


def _score_shift(pkg, msa_null, wgt, Snull):
    """Scores of one null alignment with Snull-slice and with 4-slice weights (two plain contexts: a context whose input
    alignment has Snull slices quantises exactly as a mixed context's nulls do)."""
    N, L = msa_null.shape
    out = []
    for S in (Snull, 4):
        c = pkg.Context(0)
        try:
            c.configure(N, L, 1, S)
            c.set_weights(wgt)
            out.append((c.scan(msa_null, pkg.GT, pkg.C16, pkg.APC)["cov"], c.quantisation()))
        finally:
            c.close()
    return out


def test_stated_bound_on_a_null_at_the_ssu_shape(pkg):
    """max |d score| of a whole SSU-shaped null between the 2-slice and the 4-slice weights, and the scores that change bin."""
    N, L = 10000, 1800
    msa, wgt, _, _ = pkg.synth.synthetic_family(N, L, seed=42)
    rng = np.random.default_rng(7)
    null = msa[:, rng.permutation(L)]                                 # a column permutation: the simplest null (msamanip_ShuffleColumns)
    (cov2, quant2), (cov4, _) = _score_shift(pkg, null, wgt, 2)
    c = _mixed(pkg, N, L, 1, 4, 2, wgt)
    try:
        assert np.array_equal(c.null_quantisation()[0], quant2[0])    # the mixed context's nulls carry exactly these weights
    finally:
        c.close()
    iu = np.triu_indices(L, 1)
    a, b = cov2[iu], cov4[iu]
    rel = np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))
    binof = lambda x: np.ceil((np.maximum(x, -10.0 + 0.05) + 10.0) / 0.05 - 1.0)
    moved = np.count_nonzero(binof(a) != binof(b)) / a.size
    print(f"\n[mixed] SSU null: max |d score| / max(1,|score|) = {rel:.3g}, max |d score| = {np.max(np.abs(a - b)):.3g}, "
          f"scores changing bin: {moved:.3g} of {a.size}")
    assert rel <= MIXED_SCORE_BOUND
    assert moved <= MIXED_BIN_FRACTION


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_identical_significant_pairs(pkg, po, oracle, seed):
    """Same input alignment, same null alignments (device generator A, same seed): the significant pairs called with the nulls at
    2 slices are those called with the nulls at 4 slices, and the E-values agree to 1e-3 relative."""
    N, L, R = 1200, 160, 20
    msa, wgt, partner, tree = pkg.synth.synthetic_family(N, L, seed=seed)
    P = L * (L - 1) // 2
    runs = []
    for Snull in (0, 2):
        c = _mixed(pkg, N, L, 2, 4, Snull, wgt)
        try:
            c.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
            c.pool_reserve(R)
            c.null_fitch_shuffle(msa, 777 + seed, R)
            c.hist_reset()
            w, _, _ = c.null_width_pool(0)
            mm = c.null_hist_pool(0, R, w)
            res = c.scan(msa, pkg.GT, pkg.C16, pkg.APC)
            xmax = max(float(mm[:, 1].max()), -10.0 + w)
            nb = int(np.ceil((max(xmax, res["maxcov"]) + 10.0) / w)) + 6
            bins, n, _ = c.hist_read(nb)
            assert n == R * P
            fits = [po.NullFit(-10.0, w, bins, xmax=xmax), po.NullFit(-10.0, w, bins, xmax=xmax).exp_tail(0.05)]
            hits = [c.scan_hits(f.bmin, f.w, f.obs, f.xmax, P, 0, None, f.survfit, f.phi, thresh=0.05) for f in fits]
            runs.append(dict(w=w, bins=bins, cov=res["cov"], hits=hits, mm=mm))
        finally:
            c.close()
    strict, mixed = runs
    assert strict["w"] == mixed["w"]
    assert np.array_equal(strict["cov"], mixed["cov"])                 # the input alignment's scan is untouched
    nb = min(len(strict["bins"]), len(mixed["bins"]))
    dbins = int(np.abs(strict["bins"][:nb].astype(np.int64) - mixed["bins"][:nb].astype(np.int64)).sum())
    dmm = float(np.max(np.abs(strict["mm"] - mixed["mm"]) / np.maximum(1.0, np.abs(strict["mm"]))))
    print(f"\n[mixed] seed {seed}: sum |d bins| = {dbins} of {R * P} scores, max rel shift of the replicates' score range {dmm:.3g}, "
          f"hits {[h['nhit'] for h in strict['hits']]}")
    assert dbins <= 2 * MIXED_BIN_FRACTION * R * P
    assert dmm <= MIXED_SCORE_BOUND
    for hs, hm in zip(strict["hits"], mixed["hits"]):
        assert hs["nhit"] == hm["nhit"] and hs["nhit"] > 0
        assert np.array_equal(hs["i"], hm["i"]) and np.array_equal(hs["j"], hm["j"]) and np.array_equal(hs["sc"], hm["sc"])
        assert np.allclose(hs["eval"], hm["eval"], rtol=1e-3, atol=0.0)


def test_null_slices_cannot_exceed_the_input_alignments(pkg):
    c = pkg.Context(0)
    try:
        c.set_null_slices(5)
        with pytest.raises(pkg.RscapeB200Error):
            c.configure(100, 40, 1, 4)
        c.set_null_slices(0)
        c.configure(100, 40, 1, 4)
    finally:
        c.close()
