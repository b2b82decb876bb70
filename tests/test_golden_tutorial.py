"""The reference's only known-answer vector: `R-scape -s tutorial/updated_Arisong.sto`
(documentation/tutorial.tex:187-212).  The oracle -- with the preprocessing that defines the analysed matrix
(gap-column filter src/msamanip.c:486-500, GSC weights) -- must reproduce the banner and all 11 significant
pairs' GTp scores to the printed 5 decimals.  Fixtures: tests/golden/ (made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    z = np.load(os.path.join(HERE, "golden", "arisong_tutorial.npz"))
    with open(os.path.join(HERE, "golden", "arisong_tutorial.json")) as fh:
        gold = json.load(fh)
    return z["ax"], gold


def test_tutorial_transcript_scores(po, oracle):
    ax, gold = _load()
    assert ax.shape == (gold["nseq"], gold["alen_orig"])
    sub, keep = po.remove_gap_columns(ax)                       # --gapthresh 0.75 (src/R-scape.c:289)
    sub = po.degen_to_N(sub)
    assert sub.shape == (gold["nseq"], gold["alen"])            # "nseq 95 (95) alen 66 (150)"
    wgt = po.weights_gsc(sub)                                   # nseq <= 1000 -> GSC (src/R-scape.c:1555)
    assert abs(wgt.sum() - gold["nseq"]) < 1e-9
    res = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
    col = {int(c) + 1: k for k, c in enumerate(keep)}           # 1-based input coordinates -> analysed column
    for p in gold["pairs"]:
        got = res["cov"][col[p["i"]], col[p["j"]]]
        assert round(got, 5) == p["score"], (p, got)
    # "[cov_min,cov_max] = [-9.95,121.66]": the maximum, and the histogram clamp bmin + w for the minimum
    assert round(res["maxcov"], 2) == 121.66
    assert res["mincov"] < -10 + 0.05 and round(-10 + 0.05, 2) == -9.95
    # the 11 listed pairs are the 11 best-scoring annotated pairs: every other score is lower than the 11th
    eleventh = min(p["score"] for p in gold["pairs"])
    listed = {(col[p["i"]], col[p["j"]]) for p in gold["pairs"]}
    iu = np.triu_indices(sub.shape[1], 1)
    higher = {(i, j) for i, j in zip(*iu) if res["cov"][i, j] >= eleventh - 1e-9}
    assert listed <= higher


def test_weights_matter(po, oracle):
    """With unit weights the same pairs score differently (96.46 instead of 121.66): the GSC restatement is exercised."""
    ax, gold = _load()
    sub, keep = po.remove_gap_columns(ax)
    sub = po.degen_to_N(sub)
    res = oracle.scan(sub, np.ones(sub.shape[0]), po.GT, po.C16, po.APC)
    col = {int(c) + 1: k for k, c in enumerate(keep)}
    assert abs(res["cov"][col[98], col[106]] - 96.46) < 0.01
