"""BASELINE.json shapes at full size.  The CPU oracle needs minutes to hours there, so parity is checked through
(a) the direct (non tensor core) verification kernel: fixed-point counts must be bit-identical to the tcgen05 kernel's,
(b) the oracle on a slice of columns (pairs inside the slice do not depend on the other columns for counts / pp),
(c) size-independent properties: symmetry, -inf diagonal, the APC correction recomputed in numpy from the device's raw
    scores, histogram mass = number of pairs, determinism run to run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = {"trna": (1000, 76), "rnasep": (5000, 400), "ssu": (10000, 1800), "lsu": (20000, 3500)}


@pytest.mark.parametrize("name", ["trna", "rnasep", "ssu", "lsu"])
def test_full_size_counts_and_properties(ctx, pkg, po, oracle, name):
    N, L = SHAPES[name]
    msa, wgt, _ = pkg.synth.synthetic_msa(N, L, seed=42)
    ctx.configure(N, L, 2, 5)
    ctx.set_weights(wgt)
    res = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC, want_cov=True)
    cov = res["cov"]
    got = ctx.counts()
    # (a) tensor-core counts == direct kernel counts, every cell of every pair
    direct = ctx.counts_direct(msa)
    assert np.array_equal(got, direct)
    del direct
    # (b) oracle on a column slice: counts bit-exact, pp within 1e-12
    sl = slice(L // 3, L // 3 + min(24, L))
    sub = np.ascontiguousarray(msa[:, sl])
    wq, q, S = ctx.quantisation()
    ref_cnt = np.triu(oracle.counts_fixed(sub, wq).transpose(2, 0, 1), 1)
    assert np.array_equal(got[:, sl, sl], ref_cnt)
    # (c) properties
    off = ~np.eye(L, dtype=bool)
    assert np.array_equal(cov, cov.T) or np.allclose(cov[off], cov.T[off], rtol=0, atol=0)
    assert np.isneginf(np.diag(cov)).all() and np.isfinite(cov[off]).all()
    assert res["mincov"] == cov[off].min() and res["maxcov"] == cov[off].max()
    # the correction applied to the device's own raw scores, recomputed in numpy (corr_CalculateCOVCorrected :1093-1118)
    raw = ctx.scan(msa, pkg.GT, pkg.C16, pkg.NOCORR, want_cov=True)["cov"]
    rawz = np.where(off, raw, 0.0)
    avg = rawz.sum() / (L * (L - 1.0))
    covx = rawz.sum(1) / (L - 1.0)
    want = rawz - np.outer(covx, covx) / avg
    scale = np.abs(rawz).max()
    assert np.max(np.abs(cov[off] - want[off])) <= 1e-9 * scale
    del raw, rawz, want
    # determinism: a second scan gives the same bits
    res2 = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC, want_cov=True)
    assert np.array_equal(res2["cov"], cov)
    # null histogram mass
    w, _, _ = ctx.null_width(msa)
    ctx.hist_reset()
    ctx.null_hist(msa[None], w, want_minmax=False)
    bins, n, imax = ctx.hist_read(1 << 20)
    assert n == L * (L - 1) // 2 == int(bins.sum())


@pytest.mark.parametrize("name", ["trna", "rnasep"])
def test_whole_config_against_the_oracle(ctx, pkg, po, oracle, name):
    """BASELINE configs 1 and 2 are small enough for the CPU oracle to scan WHOLE (0.03 s and 3 s): every score of the full
    alignment, pm, nseff, min/max and the null histogram's integer bins against the oracle (double weights), not only a slice."""
    from _helpers import assert_bins_identical
    N, L = SHAPES[name]
    msa, wgt, _ = pkg.synth.synthetic_msa(N, L, seed=42)
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    got = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC, want_probs=True)
    ref = oracle.scan(msa, wgt, po.GT, po.C16, po.APC, want_probs=True)
    raw = oracle.scan(msa, wgt, po.GT, po.C16, po.NOCORR)
    scale = max(abs(raw["maxcov"]), abs(raw["mincov"]), 1.0)
    iu = np.triu_indices(L, 1)
    assert np.max(np.abs(got["cov"][iu] - ref["cov"][iu])) <= 1e-9 * scale
    assert np.max(np.abs(got["cov"].T[iu] - ref["cov"][iu])) <= 1e-9 * scale
    assert abs(got["mincov"] - ref["mincov"]) <= 1e-9 * scale and abs(got["maxcov"] - ref["maxcov"]) <= 1e-9 * scale
    for k in ("pm", "ps", "nseff", "pp"):
        assert np.max(np.abs(got[k] - ref[k])) <= 1e-9 * max(1.0, np.max(np.abs(ref[k]))), k
    # the same alignment as a "null": width, then the histogram's integer bins
    w_ref = oracle.null_width(0.05, ref["mincov"], ref["maxcov"], -10.0, 400, 1e-6)
    w, mn, mx = ctx.null_width(msa)
    assert abs(w - w_ref) <= 1e-12 * max(1.0, w_ref)
    h = oracle.hist_from_cov(ref["cov"], ref["maxcov"], -10.0, w_ref, 1e-6)
    v = oracle.view(h)
    oracle.free(h)
    ctx.hist_reset()
    ctx.null_hist(msa[None], w_ref, want_minmax=False)
    bins, n, imax = ctx.hist_read(v.nb + 8)
    assert n == L * (L - 1) // 2 == v.n
    assert_bins_identical(bins, v.obs, ref["cov"][iu], -10.0, w_ref, scale=scale)


@pytest.mark.parametrize("nslices", [4, 5])
def test_ssu_slice_scores_match_oracle(ctx, pkg, po, oracle, nslices):
    """Scores of a full-size scan cannot be compared pair by pair with an oracle run on a slice (pm and APC depend on all
    columns), so the comparison runs the device on the same slice: N = 10000 sequences, 160 columns.  The oracle gets the
    DOUBLE weights, so the fixed-point weight error (4 digit slices + 8-bit multiplier is the default) is inside the bound."""
    N, L = 10000, 160
    msa, wgt, _ = pkg.synth.synthetic_msa(N, 1800, seed=42)
    sub = np.ascontiguousarray(msa[:, 400:400 + L])
    ctx.configure(N, L, 1, nslices)
    ctx.set_weights(wgt)
    got = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC)
    ref = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
    raw = oracle.scan(sub, wgt, po.GT, po.C16, po.NOCORR)
    scale = max(abs(raw["maxcov"]), abs(raw["mincov"]), 1.0)
    off = ~np.eye(L, dtype=bool)
    assert np.max(np.abs(got["cov"][off] - ref["cov"][off])) <= 1e-9 * scale


@pytest.mark.parametrize("name", ["rnasep", "ssu"])
def test_full_size_statistic_sweep_is_consistent(ctx, pkg, name):
    """BASELINE config 5 (statistic sweep) at full size, through relations between the device's own outputs:
    G = 2 nseff MI and MIg = MI - ngap/nseff pair by pair (correlators.c:383-387, :560-570, :700-712), and the APC / ASC
    corrections of every statistic recomputed in numpy from its raw scores (:1093-1118)."""
    N, L = SHAPES[name]
    msa, wgt, _ = pkg.synth.synthetic_msa(N, L, seed=7)
    ctx.configure(N, L, 1, 0)
    ctx.set_weights(wgt)
    iu = np.triu_indices(L, 1)
    first = ctx.scan(msa, pkg.MI, pkg.C16, pkg.NOCORR, want_probs=(name == "rnasep"))
    mi = first["cov"][iu]
    gt = ctx.scan(msa, pkg.GT, pkg.C16, pkg.NOCORR)["cov"][iu]
    mig = ctx.scan(msa, pkg.MIg, pkg.C16, pkg.NOCORR)["cov"][iu]
    ne, ng = ctx.last_nseff()
    ne, ng = ne[iu], ng[iu]
    ok = ne > 0
    assert ok.mean() > 0.99
    scale = np.abs(gt).max()
    assert np.max(np.abs(gt[ok] - 2.0 * ne[ok] * mi[ok])) <= 1e-9 * scale
    assert np.max(np.abs(mig[ok] - (mi[ok] - ng[ok] / ne[ok]))) <= 1e-9 * max(1.0, np.abs(mig).max())
    if name == "rnasep":
        assert np.allclose(first["nseff"][iu], ne, rtol=1e-12, atol=0)
    off = ~np.eye(L, dtype=bool)
    for stat in (pkg.GT, pkg.MI, pkg.MIr, pkg.CHI, pkg.OMES, pkg.RAFS):
        raw = ctx.scan(msa, stat, pkg.C16, pkg.NOCORR)["cov"]
        rawz = np.where(off, raw, 0.0)
        avg = rawz.sum() / (L * (L - 1.0))
        covx = rawz.sum(1) / (L - 1.0)
        scale = max(np.abs(rawz).max(), 1e-300)
        apc = ctx.scan(msa, stat, pkg.C16, pkg.APC)["cov"]
        asc = ctx.scan(msa, stat, pkg.C16, pkg.ASC)["cov"]
        want_apc = rawz - np.outer(covx, covx) / avg
        want_asc = rawz - (covx[:, None] + covx[None, :] - avg)
        assert np.max(np.abs(apc[off] - want_apc[off])) <= 1e-9 * scale, stat
        assert np.max(np.abs(asc[off] - want_asc[off])) <= 1e-9 * scale, stat

