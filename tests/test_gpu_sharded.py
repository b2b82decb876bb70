"""One scan with the L x L pair grid sharded over ranks by 32-column row blocks (BASELINE config 4).
Here the ranks are contexts on one GPU and the "all-reduces" are numpy sums on the host, so the test needs one GPU;
the same three-phase protocol runs over NCCL in the multi-GPU bench."""
import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("stat,cls,ac", [("GT", "C16", "APC"), ("MI", "C2", "ASC")])
def test_sharded_scan_equals_unsharded(pkg, po, oracle, world, stat, cls, ac):
    N, L = 300, 150
    msa, wgt, _ = po.synthetic_msa(N, L, seed=61)
    S, Cc, A = getattr(pkg, stat), getattr(pkg, cls), getattr(pkg, ac)
    whole = pkg.Context(0)
    whole.configure(N, L, 1, 0)
    whole.set_weights(wgt)
    ref = whole.scan(msa, S, Cc, A)
    w = 0.05
    whole.hist_reset()
    whole.null_hist(msa[None], w, S, Cc, A, want_minmax=False)
    ref_bins, ref_n, _ = whole.hist_read(1 << 16)

    ranks = []
    for k in range(world):
        c = pkg.Context(0)
        c.configure(N, L, 1, 0)
        c.set_shard(k, world)
        c.set_weights(wgt)
        ranks.append(c)
    msum = sum(c.sharded_counts(msa) for c in ranks)                                   # all-reduce SUM
    parts = [c.sharded_statistic(msum, S, Cc) for c in ranks]
    cs = np.sum(parts, axis=0)                                                        # SUM on [0..L], MIN / MAX on the range
    cs[L + 1] = min(p[L + 1] for p in parts)
    cs[L + 2] = max(p[L + 2] for p in parts)
    covs, los, his, bins, n = [], [], [], 0, 0
    for c in ranks:
        cov, lo, hi = c.sharded_correct(cs, A, want_cov=True, hist_w=w)
        covs.append(cov); los.append(lo); his.append(hi)
        b, nn, _ = c.hist_read(1 << 16)
        bins = bins + b.astype(np.int64); n += nn
    up = np.triu(np.sum(covs, axis=0), 1)                                             # each pair is owned by exactly one rank
    owned = sum((np.triu(cv, 1) != 0).astype(int) for cv in covs)
    assert owned.max() <= 1
    iu = np.triu_indices(L, 1)
    scale = max(1.0, np.abs(ref["cov"][iu]).max())
    assert np.max(np.abs(up[iu] - ref["cov"][iu])) <= 1e-11 * scale
    assert abs(min(los) - ref["mincov"]) <= 1e-11 * scale and abs(max(his) - ref["maxcov"]) <= 1e-11 * scale
    assert n == ref_n == L * (L - 1) // 2
    assert_bins_identical(bins, ref_bins, ref["cov"][iu], -10.0, w, rel=1e-11, scale=scale)
    # ... and against the ORACLE (not only the unsharded device scan): scores within 1e-9, identical integer bins
    ora = oracle.scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac))
    raw = oracle.scan(msa, wgt, getattr(po, stat), getattr(po, cls), po.NOCORR)
    oscale = max(1.0, abs(raw["maxcov"]), abs(raw["mincov"]))
    assert np.max(np.abs(up[iu] - ora["cov"][iu]) / np.maximum(np.maximum(1.0, np.abs(ora["cov"][iu])), oscale)) <= 1e-9
    h = oracle.hist_from_cov(ora["cov"], ora["maxcov"], -10.0, w)
    v = oracle.view(h)
    oracle.free(h)
    assert_bins_identical(bins[:max(v.nb, int(np.nonzero(bins)[0][-1]) + 1)], v.obs, ora["cov"][iu], -10.0, w, scale=oscale)
    for c in ranks + [whole]:
        c.close()
