"""E-values and the significant-pair list on the device (rsb_scan_hits) vs the oracle's restatement of the per-pair loop of
cov_CreateHitList (src/covariation.c:828-910) and cov2evalue (:2370-2400), itself pinned against the reference's own
static functions (tests/test_evalue_oracle.py).

The oracle is run on the score matrix the device holds, so the comparison isolates this stage and is exact: the same set of
significant pairs in the same order, bit-identical p-values, E-values and mi->Eval."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scan_and_null(ctx, pkg, po, N, L, seed, nnull=4):
    msa, wgt, partner = po.synthetic_msa(N, L, seed=seed)
    nulls = np.stack([po.synthetic_msa(N, L, seed=seed + 50 + r)[0] for r in range(nnull)])
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    ctx.hist_reset()
    w, _, _ = ctx.null_width(nulls[0])
    mm = ctx.null_hist(nulls, w)
    bmin = -10.0
    xmax = max(float(mm[:, 1].max()), bmin + w)
    res = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)                       # the input alignment last: its scores stay on the device
    nb = int(np.ceil((max(xmax, res["maxcov"]) - bmin) / w)) + 6
    bins, n, imax = ctx.hist_read(nb)
    assert n == nnull * L * (L - 1) // 2 == int(bins.sum())
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    return res, mask, po.NullFit(bmin, w, bins, xmax=xmax)


def _same_hits(a, b):
    assert a["nhit"] == len(b["i"]), (a["nhit"], len(b["i"]))
    for k in ("i", "j", "sc", "eval", "pval"):
        assert np.array_equal(a[k], b[k]), k


def _device_hits(ctx, null, mask, Nb, Nt, **kw):
    return ctx.scan_hits(null.bmin, null.w, null.obs, null.xmax, Nt, Nb, mask, null.survfit, null.phi, **kw)


@pytest.mark.parametrize("N,L,seed", [(200, 64, 5), (150, 131, 6), (40, 33, 7)])
@pytest.mark.parametrize("fit", [False, True])
def test_hits_equal_oracle(ctx, pkg, po, oracle, N, L, seed, fit):
    res, mask, null = _scan_and_null(ctx, pkg, po, N, L, seed)
    if fit:
        null = null.exp_tail(0.05)
    Nb = int(mask.sum())
    Nt = L * (L - 1) // 2 - Nb
    for m, nb_, nt_ in ((mask, Nb, Nt), (None, 0, Nt + Nb)):
        for thresh in (0.05, 10.0, 2000.0):                            # 2000 > MAX_EVAL: every pair is reported
            got = _device_hits(ctx, null, m, nb_, nt_, thresh=thresh)
            want = oracle.hitlist(res["cov"], null, m, nb_, nt_, -1, thresh)
            _same_hits(got, want)
            iu = np.triu_indices(L, 1)
            assert np.array_equal(got["Eval"][iu], want["Eval"][iu]) and np.array_equal(got["Eval"], got["Eval"].T)
            assert np.isposinf(np.diag(got["Eval"])).all()
            # significance is what the hit list says it is
            sig = (got["Eval"][iu] < thresh) | (thresh > 1000)
            assert int(sig.sum()) == got["nhit"]
    assert _device_hits(ctx, null, mask, Nb, Nt, thresh=2000.0)["nhit"] == L * (L - 1) // 2


@pytest.mark.parametrize("expBP", [1, 3, 40, 100000])
def test_hits_expbp_rule(ctx, pkg, po, oracle, expBP):
    """--structured: pairs outside the structure are multiplied by expBP until expBP hits are listed (src/covariation.c:852)."""
    res, mask, null = _scan_and_null(ctx, pkg, po, 160, 72, 9)
    null = null.exp_tail(0.05)
    Nt = 72 * 71 // 2
    for m in (None, mask):
        for thresh in (0.5, 50.0):
            got = _device_hits(ctx, null, m, 0 if m is None else int(m.sum()), Nt, expBP=expBP, thresh=thresh)
            want = oracle.hitlist(res["cov"], null, m, 0 if m is None else int(m.sum()), Nt, expBP, thresh)
            _same_hits(got, want)
            iu = np.triu_indices(72, 1)
            assert np.array_equal(got["Eval"][iu], want["Eval"][iu])


def test_hits_capacity_and_errors(ctx, pkg, po, oracle):
    res, mask, null = _scan_and_null(ctx, pkg, po, 120, 48, 11)
    P = 48 * 47 // 2
    full = _device_hits(ctx, null, None, 0, P, thresh=2000.0)
    part = _device_hits(ctx, null, None, 0, P, thresh=2000.0, cap=10, want_eval=False)
    assert part["nhit"] == P and len(part["i"]) == 10 and part["Eval"] is None
    assert set(zip(part["i"], part["j"])) <= set(zip(full["i"], full["j"]))
    none = _device_hits(ctx, null, None, 0, P, thresh=2000.0, cap=0, want_eval=False)
    assert none["nhit"] == P and len(none["i"]) == 0
    with pytest.raises(pkg.RscapeB200Error):
        ctx.scan_hits(null.bmin, null.w, np.zeros(8, np.uint64), 0.0, P)              # empty null histogram
    # a tail whose censoring point lies below the histogram makes the reference read survfit[-k]: reported, not computed
    bad = po.NullFit(null.bmin + 20.0, null.w, null.obs, null.xmax + 20.0, phi=-np.inf, cmin=0, survfit=np.zeros(2 * null.nb))
    with pytest.raises(pkg.RscapeB200Error, match="cannot find evalue"):
        _device_hits(ctx, bad, None, 0, P)
    # the context still works afterwards
    _same_hits(_device_hits(ctx, null, None, 0, P, thresh=1.0), oracle.hitlist(res["cov"], null, None, 0, P, -1, 1.0))


@pytest.mark.parametrize("world", [2, 3])
def test_hits_on_a_sharded_pair_grid(pkg, po, oracle, world):
    """Every rank lists the significant pairs of the rows it owns; concatenated and sorted they are the unsharded list
    (north_star item 4: only histograms and significant-pair lists cross GPUs)."""
    N, L = 220, 150
    msa, wgt, partner = po.synthetic_msa(N, L, seed=21)
    whole = pkg.Context(0)
    res, mask, null = _scan_and_null(whole, pkg, po, N, L, 21)
    null = null.exp_tail(0.05)
    Nb, P = int(mask.sum()), L * (L - 1) // 2
    want = _device_hits(whole, null, mask, Nb, P - Nb, thresh=5.0)
    ranks = []
    for k in range(world):
        c = pkg.Context(0)
        c.configure(N, L, 1, 0)
        c.set_shard(k, world)
        c.set_weights(wgt)
        ranks.append(c)
    msum = sum(c.sharded_counts(msa) for c in ranks)
    parts = [c.sharded_statistic(msum, pkg.GT, pkg.C16) for c in ranks]
    cs = np.sum(parts, axis=0)
    cs[L + 1] = min(p[L + 1] for p in parts)
    cs[L + 2] = max(p[L + 2] for p in parts)
    lists, evals = [], []
    for c in ranks:
        cov, _, _ = c.sharded_correct(cs, pkg.APC, want_cov=True)
        h = _device_hits(c, null, mask, Nb, P - Nb, thresh=5.0)
        # exact against the oracle on this rank's own scores
        own = np.zeros(L, bool)
        own[[i for i in range(L) if (i // 32) % world == c_rank(c, ranks)]] = True
        ref = oracle.hitlist(cov, null, mask, Nb, P - Nb, -1, 5.0)
        keep = own[ref["i"]]
        assert np.array_equal(h["i"], ref["i"][keep]) and np.array_equal(h["j"], ref["j"][keep]) and np.array_equal(h["eval"], ref["eval"][keep])
        lists.append(h); evals.append(np.triu(h["Eval"], 1))
        with pytest.raises(pkg.RscapeB200Error):
            _device_hits(c, null, mask, Nb, P - Nb, expBP=3)
    merged = pkg.parallel.merge_hit_lists(lists)
    assert np.array_equal(merged["i"], want["i"]) and np.array_equal(merged["j"], want["j"])
    scale = np.maximum(1e-300, np.abs(want["eval"]))
    assert np.max(np.abs(merged["eval"] - want["eval"]) / scale) < 1e-6        # sharded scores differ from unsharded ones by ~1e-11
    assert ((sum(evals) != 0).sum()) == P
    for c in ranks + [whole]:
        c.close()


def c_rank(c, ranks):
    return ranks.index(c)


def test_hits_full_size_properties(pkg, po):
    """SSU-sized L: the list is exactly the set of pairs whose returned E-value is below the threshold, in row-major order,
    and p-values are non-increasing in the score."""
    N, L = 512, 1800
    rng = np.random.default_rng(3)
    msa = rng.integers(0, 4, (N, L), dtype=np.uint8)
    msa[:, 100] = msa[:, 200]                                           # a few perfectly covarying column pairs
    msa[:, 300] = 3 - msa[:, 1500]
    null_msa = rng.integers(0, 4, (2, N, L), dtype=np.uint8)
    c = pkg.Context(0)
    c.configure(N, L, 2, 0)
    c.set_weights(None)
    w, _, _ = c.null_width(null_msa[0])
    mm = c.null_hist(null_msa, w)
    res = c.scan(msa, pkg.GT, pkg.C16, pkg.APC)
    nb = int(np.ceil((max(mm[:, 1].max(), res["maxcov"]) + 10.0) / w)) + 6
    bins, n, _ = c.hist_read(nb)
    P = L * (L - 1) // 2
    # without a fitted tail the smallest p-value is 1 / (number of null scores) = 1 / (2 P), i.e. E = 0.5: list the pairs that reach it
    got = c.scan_hits(-10.0, w, bins, float(mm[:, 1].max()), P, thresh=0.6)
    iu = np.triu_indices(L, 1)
    sig = got["Eval"][iu] < 0.6
    assert got["nhit"] == int(sig.sum()) >= 2
    assert np.array_equal(got["i"], iu[0][sig]) and np.array_equal(got["j"], iu[1][sig])
    assert (100, 200) in set(zip(got["i"].tolist(), got["j"].tolist())) and (300, 1500) in set(zip(got["i"].tolist(), got["j"].tolist()))
    order = np.argsort(res["cov"][iu], kind="stable")
    assert np.all(np.diff(got["Eval"][iu][order]) <= 0)
    c.close()


def test_cov_createhitlist_b200_matches_oracle(po, oracle):
    """The per-pair loop of cov_CreateHitList (src/covariation.c:828-910) through the host mirror: mi->Eval and the list of
    significant pairs from mi->COV and data->ranklist_null, against the oracle's loop run on the same mi->COV."""
    from test_gpu_host_api import _lib
    lib = _lib(po)
    glue = C.CDLL(os.path.join(ROOT, "oracle", "libglue_b200.so"))
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    glue.glue_hitlist.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_double, vp, C.c_int64,
                                  vp, vp, vp, vp, vp, vp, vp]
    N, L = 200, 70
    msa, wgt, partner = po.synthetic_msa(N, L, seed=41)
    P = L * (L - 1) // 2
    ap = np.ascontiguousarray(po.ALLOWPAIR_WC_GU)
    m = lib.glue_msa_create(N, L, msa.ctypes.data_as(C.POINTER(C.c_uint8)), wgt.ctypes.data_as(dp))
    mi = lib.corr_Create(L, N, 0, 8, 50, lib.glue_abc_rna(), po.C16)
    apm = lib.glue_allowpair_from(ap.ctypes.data_as(dp))
    data = lib.glue_data_create(mi, apm, 4, 1e-6)                     # GTp
    glue.glue_cov_calculate.argtypes = [vp, vp]
    assert glue.glue_cov_calculate(data, m) == 0, lib.glue_data_errbuf(data)
    cov = np.empty((L, L))
    lib.glue_mi_export(mi, None, None, None, None, None, cov.ctypes.data_as(dp), None, None)

    x = np.maximum(np.random.default_rng(4).normal(0, 6, 60000), -10 + 0.05)
    b = np.ceil((x + 10) / 0.05 - 1).astype(np.int64)
    null = po.NullFit(-10.0, 0.05, np.bincount(b, minlength=int(b.max()) + 6).astype(np.uint64), xmax=float(x.max())).exp_tail(0.05)
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    Nb = int(mask.sum())

    def run(thresh, expBP, use_mask, use_null=True, cap=P):
        hi, hj = np.zeros(cap + 1, np.int64), np.zeros(cap + 1, np.int64)
        sc, ev, pv = np.zeros(cap + 1), np.zeros(cap + 1), np.zeros(cap + 1)
        E = np.zeros((L, L))
        nhit = C.c_int64()
        geom = np.array([null.bmin, null.w, null.xmax, null.phi])
        ig = np.array([null.nb, null.imin, null.imax], np.int32)
        st = glue.glue_hitlist(data, mi, geom.ctypes.data, ig.ctypes.data, null.obs.ctypes.data if use_null else None,
                               null.survfit.ctypes.data, Nb if use_mask else 0, P - Nb if use_mask else P, expBP, thresh,
                               mask.ctypes.data if use_mask else None, cap, hi.ctypes.data, hj.ctypes.data, sc.ctypes.data,
                               ev.ctypes.data, pv.ctypes.data, E.ctypes.data, C.byref(nhit))
        assert st == 0, lib.glue_data_errbuf(data)
        k = min(nhit.value, cap)
        return dict(i=hi[:k], j=hj[:k], sc=sc[:k], eval=ev[:k], pval=pv[:k], nhit=nhit.value, Eval=E)

    iu = np.triu_indices(L, 1)
    for thresh, expBP, use_mask in ((0.05, -1, True), (5.0, -1, False), (5.0, 6, True), (2000.0, -1, True)):
        got = run(thresh, expBP, use_mask)
        want = oracle.hitlist(cov, null, mask if use_mask else None, Nb if use_mask else 0, P - Nb if use_mask else P, expBP, thresh)
        assert got["nhit"] == len(want["i"])
        for k in ("i", "j", "sc", "eval", "pval"):
            assert np.array_equal(got[k], want[k]), (thresh, expBP, k)
        assert np.array_equal(got["Eval"][iu], want["Eval"][iu]) and np.array_equal(got["Eval"], got["Eval"].T)
        assert np.isposinf(np.diag(got["Eval"])).all()
    # without a null rank list every pair is listed with pval = eval = 0 (the naive method, :844-847)
    naive = run(0.05, -1, False, use_null=False)
    assert naive["nhit"] == P and not naive["eval"].any() and np.array_equal(naive["sc"], cov[iu])
    lib.glue_data_destroy(data)
    lib.esl_dmatrix_Destroy(apm)
    lib.corr_Destroy(mi)
    lib.glue_msa_destroy(m)


def test_significant_pairs_identical_end_to_end(ctx, pkg, po, oracle):
    """north_star: "the set of significant pairs called is identical when the same null alignments are supplied".
    Whole chain on both sides from the same input alignment and the same nulls: width pass, null scans, cumulative histogram,
    scan of the input alignment, E-values, hit list.  The device's histogram bins equal the oracle's, so both sides see the same
    null distribution; the scores agree to 1e-9, so they fall in the same bins and the calls are the same."""
    from test_gpu_nulls import oracle_null_loop
    N, L, R = 180, 80, 6
    msa, wgt, partner = po.synthetic_msa(N, L, seed=123)
    nulls = np.stack([po.synthetic_msa(N, L, seed=900 + r)[0] for r in range(R)])
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    Nb, P = int(mask.sum()), L * (L - 1) // 2
    # reference side (oracle)
    w_ref, view, mm_ref = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)
    ref_scan = oracle.scan(msa, wgt, po.GT, po.C16, po.APC)
    ref_null = po.NullFit(view.bmin, view.w, view.obs, xmax=view.xmax).exp_tail(0.05)
    ref_hits = oracle.hitlist(ref_scan["cov"], ref_null, mask, Nb, P - Nb, -1, 0.5)
    # device side
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    w, _, _ = ctx.null_width(nulls[0])
    mm = ctx.null_hist(nulls, w)
    dev_scan = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)
    bins, n, imax = ctx.hist_read(view.nb)
    assert np.array_equal(bins, view.obs) and n == view.n
    xmax = max(float(mm[:, 1].max()), -10.0 + w)
    dev_null = po.NullFit(-10.0, w, bins, xmax=xmax).exp_tail(0.05)           # the same host-side "fit" on the device's histogram
    assert np.array_equal(dev_null.survfit, ref_null.survfit) or abs(w - w_ref) > 0
    dev_hits = ctx.scan_hits(-10.0, w, bins, xmax, P - Nb, Nb, mask, dev_null.survfit, dev_null.phi, thresh=0.5)
    assert len(ref_hits["i"]) >= 3
    assert np.array_equal(dev_hits["i"], ref_hits["i"]) and np.array_equal(dev_hits["j"], ref_hits["j"])
    assert np.allclose(dev_hits["eval"], ref_hits["eval"], rtol=1e-6, atol=0)
    assert np.max(np.abs(dev_hits["sc"] - ref_hits["sc"]) / np.maximum(1.0, np.abs(ref_hits["sc"]))) < 1e-8   # 1e-9 of the raw scale: test_gpu_scan_parity
