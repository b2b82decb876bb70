"""BASELINE configs[0] -- R-scape's defaults on the tutorial alignment -- end to end on the CPU: the REFERENCE's own null generator
(Tree_FitchAlgorithmAncenstral + msamanip_ShuffleTreeSubstitutions, oracle/_ref) on the FastTree fixture, the oracle's scans and
histogram, the reference's own tail fit, the oracle's E-values: exactly the 11 significant pairs of documentation/tutorial.tex:187-212
for every seed (the transcript's E-values depend on Easel's RNG stream and are not compared: SURVEY 0.6)."""
import numpy as np
import pytest

import _config1 as c1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tutorial_significant_pairs_cpu(po, oracle, reflib, seed):
    sub, wgt, keep, mask, tree, gold = c1.load(po)
    assert sub.shape == (gold["nseq"], gold["alen"]) and int(mask.sum()) == gold["nbpairs"]
    N, L = sub.shape
    P = L * (L - 1) // 2
    nulls = np.stack([sh for sh, _, _ in reflib.fitch_shuffle(seed, tree, sub, nrep=c1.NSHUFFLE)])
    assert nulls.shape == (c1.NSHUFFLE, N, L)
    first = oracle.scan(nulls[0], wgt, po.GT, po.C16, po.APC)
    w = c1.null_width(first["mincov"], first["maxcov"])
    real = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
    cum, xmax = None, -np.inf
    for msa in nulls:
        r = oracle.scan(msa, wgt, po.GT, po.C16, po.APC)
        h = oracle.hist_from_cov(r["cov"], r["maxcov"], c1.BMIN, w, c1.TOL)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
        xmax = max(xmax, r["maxcov"])
    view = oracle.view(cum)
    oracle.free(cum)
    nb = c1.null_bins_needed(w, xmax, real["maxcov"])
    obs = np.zeros(nb, np.uint64)
    obs[:view.nb] = view.obs[:nb]
    assert int(obs.sum()) == c1.NSHUFFLE * P
    fit = reflib.nullfit(po.NullFit(c1.BMIN, w, obs, xmax=xmax), c1.PMASS, c1.FRACFIT, False)
    Nb = int(mask.sum())
    hits = oracle.hitlist(real["cov"], fit, mask, Nb, P - Nb, -1, c1.ETHRESH)
    want = {(p["i"], p["j"]) for p in gold["pairs"]}
    assert c1.called_pairs(hits, keep) == want
    # "[-9.95,121.66] [0 | 11 20 11 | 55.00 100.00 70.97]": all 11 calls are base pairs of the given structure
    assert all(mask[i, j] for i, j in zip(hits["i"], hits["j"]))
