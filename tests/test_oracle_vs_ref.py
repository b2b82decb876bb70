"""Pin the oracle (oracle/oracle.c) against the reference's own code.

* live: oracle/_ref/librscape_ref.so = the reference's src/correlators.c compiled unchanged against the Easel shim;
  every statistic x class x correction must agree bit for bit (skipped when that build is absent);
* committed: tests/golden/ref_scans.npz holds outputs of that same build (made by tests/golden/make_golden.py), so the
  pin also holds where /root/reference never existed.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
STATS = ["GT", "CHI", "MI", "MIr", "MIg", "OMES", "CCF", "RAF", "RAFS"]
COMBOS = [(s, c, a) for s in STATS for c in ("C16", "C2", "CWC") for a in ("APC", "ASC", "NOCORR") if not (c == "CWC" and s != "GT")]


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("stat,cls,ac", COMBOS)
def test_oracle_matches_committed_reference_outputs(po, oracle, stat, cls, ac):
    z = np.load(os.path.join(HERE, "golden", "ref_scans.npz"))
    res = oracle.scan(z["msa"], z["wgt"], getattr(po, stat), getattr(po, cls), getattr(po, ac), want_probs=True)
    assert _same(res["cov"], z[f"{stat}_{cls}_{ac}_cov"])
    assert _same(np.array([res["mincov"], res["maxcov"]]), z[f"{stat}_{cls}_{ac}_mm"])
    if stat == "GT" and cls == "C16" and ac == "APC":
        for k in ("pp", "pm", "ps", "nseff", "ngap"):
            assert _same(res[k], z["probs_" + k]), k


@pytest.mark.parametrize("N,L,seed", [(60, 70, 1), (7, 12, 2), (300, 30, 3), (2, 5, 4), (1, 6, 5)])
def test_oracle_matches_live_reference(po, oracle, reflib, N, L, seed):
    msa, wgt, _ = po.synthetic_msa(N, L, seed=seed)
    for stat, cls, ac in COMBOS:
        a = reflib.scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac))
        b = oracle.scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac), want_probs=True)
        assert _same(a["cov"], b["cov"]), (stat, cls, ac)
        assert (a["mincov"], a["maxcov"]) == (b["mincov"], b["maxcov"]) or (np.isinf(a["mincov"]) and np.isinf(b["mincov"]))
        for k in ("pp", "pm", "ps", "nseff", "ngap"):
            assert _same(a[k], b[k]), (stat, cls, ac, k)


def test_cselect_rule(po, oracle, reflib):
    """CSELECT = C2 iff nseq <= nseqthresh or alen <= alenthresh (src/correlators.c:336)."""
    for N, L, want in ((6, 70, po.C2), (60, 40, po.C2), (60, 70, po.C16)):
        msa, wgt, _ = po.synthetic_msa(N, L, seed=N)
        a = reflib.scan(msa, wgt, po.GT, po.CSELECT, po.NOCORR)
        b = oracle.scan(msa, wgt, po.GT, want, po.NOCORR)
        assert _same(a["cov"], b["cov"]) and a["covclass"] == want


def test_raf_count_identity_equals_reference_loop(po, oracle):
    """RAF through the unweighted 4x4 table == the reference's O(N^2) sequence-pair loop, bit for bit."""
    msa, wgt, _ = po.synthetic_msa(80, 33, seed=9)
    for smooth, stat in ((False, po.RAF), (True, po.RAFS)):
        direct, mn, mx = oracle.raf_direct(msa, smooth=smooth)
        fast = oracle.scan(msa, wgt, stat, po.C2, po.NOCORR)
        assert _same(direct, fast["cov"]) and (mn, mx) == (fast["mincov"], fast["maxcov"])


def test_edge_cases(po, oracle):
    # a pair with no sequence where both residues are canonical: nseff = 0, pp uniform from the prior, not counted in pm
    msa = np.array([[0, 4, 1], [4, 2, 1], [0, 4, 3], [4, 1, 2]], dtype=np.uint8)
    res = oracle.scan(msa, np.ones(4), po.GT, po.C16, po.NOCORR, want_probs=True)
    assert res["nseff"][0, 1] == 0 and np.allclose(res["pp"][0, 1], 1 / 16)
    assert res["ngap"][0, 1] == 4 and res["ngap"][1, 0] == 0          # quirk Q4
    assert np.isneginf(np.diag(res["cov"])).all()
