"""The reference's only known-answer vector through the DEVICE path: `R-scape -s tutorial/updated_Arisong.sto`
(documentation/tutorial.tex:187-212).  The committed fixture (tests/golden/arisong_tutorial.npz, made by
tests/golden/make_golden.py) is preprocessed as R-scape does (gap-column filter src/msamanip.c:486-500, degenerate -> N,
GSC weights for nseq <= 1000, src/R-scape.c:1555), scanned by the CUDA path through the C-ABI, and the 11 significant
pairs must carry the transcript's GTp scores to the printed 5 decimals."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    z = np.load(os.path.join(HERE, "golden", "arisong_tutorial.npz"))
    with open(os.path.join(HERE, "golden", "arisong_tutorial.json")) as fh:
        gold = json.load(fh)
    return z["ax"], gold


@pytest.mark.parametrize("nslices", [0, 4, 5])
def test_tutorial_transcript_scores_on_the_device(ctx, pkg, po, oracle, nslices):
    ax, gold = _load()
    sub, keep = po.remove_gap_columns(ax)
    sub = po.degen_to_N(sub)
    assert sub.shape == (gold["nseq"], gold["alen"])
    wgt = po.weights_gsc(sub)
    N, L = sub.shape
    ctx.configure(N, L, 2, nslices)
    ctx.set_weights(wgt)
    res = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC, want_probs=True)
    col = {int(c) + 1: k for k, c in enumerate(keep)}
    for p in gold["pairs"]:
        got = res["cov"][col[p["i"]], col[p["j"]]]
        assert round(got, 5) == p["score"], (p, got)
    assert round(res["maxcov"], 2) == 121.66                      # "[cov_min,cov_max] = [-9.95,121.66]"
    # and pair by pair against the oracle on the same alignment and (double) weights
    ref = oracle.scan(sub, wgt, po.GT, po.C16, po.APC, want_probs=True)
    raw = oracle.scan(sub, wgt, po.GT, po.C16, po.NOCORR)
    scale = max(1.0, abs(raw["maxcov"]), abs(raw["mincov"]))
    iu = np.triu_indices(L, 1)
    assert np.max(np.abs(res["cov"][iu] - ref["cov"][iu])) <= 1e-9 * scale
    for k in ("pp", "pm", "ps", "nseff"):
        assert np.max(np.abs(res[k] - ref[k])) <= 1e-9 * max(1.0, np.max(np.abs(ref[k]))), k
    # the 11 listed pairs are the 11 best-scoring pairs of the structure on the device as well
    eleventh = min(p["score"] for p in gold["pairs"])
    listed = {(col[p["i"]], col[p["j"]]) for p in gold["pairs"]}
    higher = {(i, j) for i, j in zip(*iu) if res["cov"][i, j] >= eleventh - 1e-9}
    assert listed <= higher


def test_tutorial_unit_weights_on_the_device(ctx, pkg, po):
    """Unit weights (one 8-bit slice, the 'unweighted int8' special case): 96.46 instead of 121.66 for the top pair."""
    ax, gold = _load()
    sub, keep = po.remove_gap_columns(ax)
    sub = po.degen_to_N(sub)
    ctx.configure(sub.shape[0], sub.shape[1], 1, 0)
    ctx.set_weights(None)
    assert ctx.quantisation()[2] == 1
    res = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC)
    col = {int(c) + 1: k for k, c in enumerate(keep)}
    assert abs(res["cov"][col[98], col[106]] - 96.46) < 0.01
