"""Every statistic x class x correction through the C-ABI vs the CPU oracle.

Tolerance (north_star: 1e-9 relative on the fp64 path): |gpu - oracle| <= 1e-9 * max(1, |oracle|, scale) where
scale is the largest |raw score| of the scan -- the corrected scores are differences of raw scores, so their absolute
error is set by the raw magnitude.  RAF / RAFS are integer arithmetic up to the last division: bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _close(got, ref, scale=1.0):
    iu = np.triu_indices(ref.shape[0], 1)
    g = np.concatenate([got[iu], got.T[iu]])
    r = np.concatenate([ref[iu], ref.T[iu]])
    err = np.abs(g - r) / np.maximum(np.maximum(1.0, np.abs(r)), scale)
    return float(err.max()) if err.size else 0.0


STATS = ["GT", "CHI", "MI", "MIr", "MIg", "OMES", "CCF"]


@pytest.mark.parametrize("stat", STATS)
@pytest.mark.parametrize("cls", ["C16", "C2", "CWC"])
@pytest.mark.parametrize("ac", ["APC", "ASC", "NOCORR"])
def test_scan_parity(ctx, pkg, po, oracle, stat, cls, ac):
    if cls == "CWC" and stat != "GT":
        pytest.skip("CWC exists for GT only (src/correlators.c:72)")
    if stat == "CCF" and cls != "C16":
        pytest.skip("CCF has a single class")
    N, L = 400, 83
    msa, wgt, _ = po.synthetic_msa(N, L, seed=17)
    ctx.configure(N, L, 1, 0)
    ctx.set_weights(wgt)
    got = ctx.scan(msa, getattr(pkg, stat), getattr(pkg, cls), getattr(pkg, ac), want_probs=True)
    ref = oracle.scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac), want_probs=True)
    raw = oracle.scan(msa, wgt, getattr(po, stat), getattr(po, cls), po.NOCORR)
    scale = max(1.0, abs(raw["maxcov"]), abs(raw["mincov"]))
    assert _close(got["cov"], ref["cov"], scale) <= TOL
    assert abs(got["mincov"] - ref["mincov"]) <= TOL * scale and abs(got["maxcov"] - ref["maxcov"]) <= TOL * scale
    assert np.all(np.isneginf(np.diag(got["cov"])))
    for k in ("pp", "pm", "ps", "nseff"):
        assert np.max(np.abs(got[k] - ref[k])) <= 1e-9 * max(1.0, np.max(np.abs(ref[k]))), k
    assert np.max(np.abs(np.triu(got["ngap"], 1) - np.triu(ref["ngap"], 1))) <= 1e-9 * N
    assert np.all(np.tril(got["ngap"]) == 0)              # quirk Q4: ngap is not mirrored


@pytest.mark.parametrize("stat", ["RAF", "RAFS"])
@pytest.mark.parametrize("ac", ["APC", "NOCORR"])
def test_raf_bit_exact(ctx, pkg, po, oracle, stat, ac):
    N, L = 150, 41
    msa, wgt, _ = po.synthetic_msa(N, L, seed=23)
    ctx.configure(N, L, 1, 0)
    ctx.set_weights(wgt)
    got = ctx.scan(msa, getattr(pkg, stat), pkg.C2, getattr(pkg, ac))
    direct, _, _ = oracle.raf_direct(msa, smooth=(stat == "RAFS"))        # the reference's O(N^2) loop
    if ac == "NOCORR":
        iu = np.triu_indices(L, 1)
        assert np.array_equal(got["cov"][iu], direct[iu])
    ref = oracle.scan(msa, wgt, getattr(po, stat), po.C2, getattr(po, ac))
    assert _close(got["cov"], ref["cov"]) <= TOL


def test_validation_failure_is_reported(ctx, pkg, po):
    """corr_ValidateProbs / corr_Marginals fail when a marginal does not sum to 1 within tol (src/correlators.c:1363, 1500-1545:
    esl_vec_DValidate -> eslFAIL "pm validation failed").  A tolerance no sum can meet must make the device scan fail the same
    way (the flag set by marg_norm_kernel), and the context must stay usable afterwards."""
    N, L = 120, 37
    msa, wgt, _ = po.synthetic_msa(N, L, seed=4)
    ctx.configure(N, L, 1, 0)
    ctx.set_weights(wgt)
    with pytest.raises(pkg.RscapeB200Error, match="validation failed"):
        ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC, tol=-1.0)
    got = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)          # the flag is cleared: the next scan succeeds
    assert np.isfinite(got["maxcov"])
