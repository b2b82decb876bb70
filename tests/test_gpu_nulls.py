"""Null loop on the device vs the oracle: the cumulative score histogram must have identical integer bins.

Reference: null_rscape's loop body (src/R-scape.c:1650-1697): calculate_width_histo on the first null, then
run_rscape(RANSS) + null_add2cumranklist for every null (histogram fill src/covariation.c:415-432).
"""
import ctypes as C

import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu


def oracle_null_loop(po, oracle, nulls, wgt, stat, cls, ac, w_old=0.05, bmin=-10.0, hpts=400, tol=1e-6):
    first = oracle.scan(nulls[0], wgt, stat, cls, ac)
    w = oracle.null_width(w_old, first["mincov"], first["maxcov"], bmin, hpts, tol)
    cum = None
    mm = []
    scores = []
    iu = np.triu_indices(nulls[0].shape[1], 1)
    for msa in nulls:
        res = oracle.scan(msa, wgt, stat, cls, ac)
        h = oracle.hist_from_cov(res["cov"], res["maxcov"], bmin, w, tol)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
        mm.append((res["mincov"], res["maxcov"]))
        scores.append(res["cov"][iu])
    view = oracle.view(cum)
    oracle.free(cum)
    oracle_null_loop.scores = np.concatenate(scores)          # the oracle's scores of the last call (edge-ambiguity check)
    return w, view, np.array(mm)


def _nulls(po, R, N, L, seed):
    return np.stack([po.synthetic_msa(N, L, seed=seed + r)[0] for r in range(R)])


@pytest.mark.parametrize("R,slots", [(1, 1), (5, 2), (7, 4), (6, 6)])
def test_null_histogram_matches_oracle(ctx, pkg, po, oracle, R, slots):
    N, L = 250, 70
    nulls = _nulls(po, R, N, L, 100)
    wgt = po.synthetic_msa(N, L, seed=1)[1]
    w_ref, view, mm_ref = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)
    ctx.configure(N, L, slots, 0)
    ctx.set_weights(wgt)
    w, mn, mx = ctx.null_width(nulls[0])
    assert abs(w - w_ref) <= 1e-12 * max(1.0, w_ref)
    mm = ctx.null_hist(nulls, w_ref)
    bins, n, imax = ctx.hist_read(view.nb + 8)
    assert n == R * L * (L - 1) // 2 == view.n
    assert np.array_equal(bins[:view.nb], view.obs), np.argwhere(bins[:view.nb] != view.obs)[:5]
    assert not bins[view.nb:].any()
    assert imax == view.imax
    assert np.max(np.abs(mm - mm_ref)) <= 1e-9 * max(1.0, np.max(np.abs(mm_ref)))


def test_null_histogram_device_resident_input(ctx, pkg, po, oracle):
    import torch
    N, L, R = 200, 64, 5
    nulls = _nulls(po, R, N, L, 300)
    wgt = po.synthetic_msa(N, L, seed=2)[1]
    w_ref, view, _ = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    dev = torch.from_numpy(nulls).cuda()
    ctx.null_hist(dev, w_ref, want_minmax=False)
    bins, n, imax = ctx.hist_read(view.nb)
    assert np.array_equal(bins, view.obs)
    # a second batch accumulates on top (cumulative histogram), reset clears
    ctx.null_hist(dev, w_ref, want_minmax=False)
    bins2, n2, _ = ctx.hist_read(view.nb)
    assert np.array_equal(bins2, 2 * view.obs) and n2 == 2 * n
    ctx.hist_reset()
    assert not ctx.hist_read(view.nb)[0].any()


@pytest.mark.parametrize("stat,cls,ac", [("MI", "C2", "ASC"), ("RAFS", "C2", "APC"), ("CHI", "C16", "NOCORR")])
def test_null_histogram_other_statistics(ctx, pkg, po, oracle, stat, cls, ac):
    N, L, R = 120, 40, 3
    nulls = _nulls(po, R, N, L, 500)
    wgt = po.synthetic_msa(N, L, seed=3)[1]
    a = (getattr(po, stat), getattr(po, cls), getattr(po, ac))
    w_ref, view, _ = oracle_null_loop(po, oracle, nulls, wgt, *a)
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    ctx.null_hist(nulls, w_ref, getattr(pkg, stat), getattr(pkg, cls), getattr(pkg, ac), want_minmax=False)
    bins, n, imax = ctx.hist_read(view.nb)
    # identical integer bins; a count may sit in the neighbouring bin only if the oracle's score lies within the score
    # tolerance (1e-9 relative) of that bin edge
    assert_bins_identical(bins, view.obs, oracle_null_loop.scores, -10.0, w_ref)


def test_last_null_nseff_quirk_q3(ctx, pkg, po, oracle):
    N, L, R = 150, 50, 3
    nulls = _nulls(po, R, N, L, 700)
    wgt = po.synthetic_msa(N, L, seed=4)[1]
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    w, _, _ = ctx.null_width(nulls[0])
    ctx.null_hist(nulls, w, want_minmax=False)
    ne, ng = ctx.last_nseff()
    ref = oracle.scan(nulls[-1], wgt, po.GT, po.C16, po.APC, want_probs=True)
    assert np.max(np.abs(ne - ref["nseff"])) <= 1e-9 * N
    assert np.max(np.abs(np.triu(ng, 1) - np.triu(ref["ngap"], 1))) <= 1e-9 * N


def test_scan_histograms_ha_hb_ht(ctx, pkg, po, oracle):
    """The three histograms of the input alignment's scan (src/covariation.c:420-457) with a structure pair mask."""
    N, L = 200, 60
    msa, wgt, partner = po.synthetic_msa(N, L, seed=77)
    ctx.configure(N, L, 1, 0)
    ctx.set_weights(wgt)
    res = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)
    ref = oracle.scan(msa, wgt, po.GT, po.C16, po.APC)
    w, bmin = 0.05, -10.0
    h = oracle.hist_from_cov(ref["cov"], ref["maxcov"], bmin, w)
    v = oracle.view(h)
    oracle.free(h)
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    ha, hb, ht = ctx.scan_hist(w, bmin, v.nb, mask)
    assert np.array_equal(ha, v.obs)
    assert np.array_equal(hb + ht, ha) and int(hb.sum()) == int(mask.sum())
    iu = np.triu_indices(L, 1)
    x = np.maximum(ref["cov"][iu], bmin + w)
    b = np.ceil((x - bmin) / w - 1).astype(int)
    want_hb = np.bincount(b[mask[iu] > 0], minlength=v.nb)
    assert np.array_equal(hb, want_hb.astype(np.uint64))


@pytest.mark.parametrize("N,L,R,S", [(300, 150, 5, 0), (257, 97, 3, 5), (64, 40, 4, 1)])
def test_record_epilogue_equals_count_path(pkg, po, oracle, monkeypatch, N, L, R, S):
    """Nulls scored with GT x C16 take the record epilogue of the tcgen05 kernel (64-byte records + gt_finish_kernel) instead of
    int64 counts + stat_kernel.  Both paths against the oracle and against each other: identical integer bins, min/max within
    1e-12, nseff/ngap of the last null identical (quirk Q3)."""
    nulls = _nulls(po, R, N, L, 900)
    wgt = po.synthetic_msa(N, L, seed=5)[1] if S != 1 else np.ones(N)
    w_ref, view, mm_ref = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("RSCAPE_B200_FUSED_GT", fused)
        c = pkg.Context(0)
        c.configure(N, L, 2, S)
        c.set_weights(wgt)
        w, mn, mx = c.null_width(nulls[0])
        mm = c.null_hist(nulls, w_ref)
        bins, n, imax = c.hist_read(view.nb + 8)
        ne, ng = c.last_nseff()
        out[fused] = (w, mn, mx, mm, bins, n, ne, ng)
        c.close()
        assert abs(w - w_ref) <= 1e-12 * max(1.0, w_ref)
        assert n == view.n
        assert_bins_identical(bins, view.obs, oracle_null_loop.scores, -10.0, w_ref)
        assert np.max(np.abs(mm - mm_ref)) <= 1e-9 * max(1.0, np.max(np.abs(mm_ref)))
    a, b = out["1"], out["0"]
    assert np.array_equal(a[4], b[4])                                                    # the two device paths: same bins
    assert np.max(np.abs(a[3] - b[3])) <= 1e-11 * max(1.0, np.max(np.abs(b[3])))
    assert np.array_equal(a[6], b[6]) and np.array_equal(a[7], b[7])


def test_pdb_distance_rule_keeps_pairs_out_of_the_histograms(ctx, pkg, po, oracle):
    """src/covariation.c:421-427: pairs with both columns in the PDB sequence and closer than `mind` are skipped by the histogram
    fill (null and input alignment alike); scores and min/max are not affected."""
    N, L, R = 150, 64, 3
    nulls = _nulls(po, R, N, L, 40)
    wgt = po.synthetic_msa(N, L, seed=6)[1]
    rng = np.random.default_rng(0)
    m2p = np.where(rng.random(L) < 0.8, np.cumsum(rng.integers(1, 3, L)), -1).astype(np.int32)
    mind = 5
    iu = np.triu_indices(L, 1)
    keep = ~((m2p[iu[0]] >= 0) & (m2p[iu[1]] >= 0) & (m2p[iu[1]] - m2p[iu[0]] < mind))
    assert 0 < keep.sum() < keep.size
    w_ref, view, mm_ref = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)
    sc = oracle_null_loop.scores.reshape(R, -1)
    x = np.maximum(sc[:, keep].ravel(), -10.0 + w_ref)
    want = np.bincount(np.ceil((x + 10.0) / w_ref - 1).astype(np.int64), minlength=view.nb)
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    ctx.set_pair_exclusion(m2p, mind)
    mm = ctx.null_hist(nulls, w_ref)
    bins, n, imax = ctx.hist_read(view.nb)
    assert n == R * int(keep.sum())
    assert_bins_identical(bins, want, sc[:, keep].ravel(), -10.0, w_ref)
    assert np.max(np.abs(mm - mm_ref)) <= 1e-9 * max(1.0, np.max(np.abs(mm_ref)))       # min/max over ALL pairs
    # the input alignment's three histograms follow the same rule
    res = ctx.scan(nulls[0], pkg.GT, pkg.C16, pkg.APC)
    ha, _, _ = ctx.scan_hist(w_ref, -10.0, view.nb)
    x0 = np.maximum(res["cov"][iu][keep], -10.0 + w_ref)
    assert np.array_equal(ha, np.bincount(np.ceil((x0 + 10.0) / w_ref - 1).astype(np.int64), minlength=view.nb).astype(np.uint64))
    ctx.set_pair_exclusion(None)
    ha2, _, _ = ctx.scan_hist(w_ref, -10.0, view.nb)
    assert int(ha2.sum()) == L * (L - 1) // 2
