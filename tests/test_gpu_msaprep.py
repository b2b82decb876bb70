"""Preprocessing that defines the scanned alignment and its weights on the device, against the CPU restatement
in oracle/pyoracle.py (which reproduces the tutorial transcript: GSC weights -> scores, banner 'alen 66 (150) avgid 65.82') and
against that golden vector itself.

  gap-column filter  msamanip_RemoveGapColumns  src/msamanip.c:486-500       PB / GSC weights  src/R-scape.c:1545-1562
  average identity   msamanip_XStats            src/msamanip.c:1967
Easel itself is not vendored by the reference: PB weights and the sampled average identity have no reference-side vector
(parity unpinned; the restatement follows SURVEY 9.7)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _tutorial():
    z = np.load(os.path.join(HERE, "golden", "arisong_tutorial.npz"))
    with open(os.path.join(HERE, "golden", "arisong_tutorial.json")) as fh:
        return z["ax"], json.load(fh)


@pytest.fixture(scope="module")
def glue(pkg):
    path = os.path.join(ROOT, "oracle", "libglue_b200.so")
    if not os.path.exists(path):
        pytest.skip("oracle/libglue_b200.so not built")
    lib = C.CDLL(path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.glue_msaprep.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint8), dp, C.c_int, C.c_int, C.c_double, C.c_int, dp, ip, dp]
    return lib


def host_prep(glue, ax, wgt_in=None, what=0, maxsq_gsc=1000, gapthresh=0.75, max_comparisons=10000, want=("wgt", "useme", "avgid")):
    N, L = ax.shape
    ax = np.ascontiguousarray(ax, dtype=np.uint8)
    w_in = None if wgt_in is None else np.ascontiguousarray(wgt_in, dtype=np.float64)
    w, use, aid = np.empty(N), np.empty(L, np.int32), C.c_double()
    dp = C.POINTER(C.c_double)
    st = glue.glue_msaprep(N, L, ax.ctypes.data_as(C.POINTER(C.c_uint8)), None if w_in is None else w_in.ctypes.data_as(dp), what, maxsq_gsc,
                           gapthresh, max_comparisons, w.ctypes.data_as(dp) if "wgt" in want else None,
                           use.ctypes.data_as(C.POINTER(C.c_int)) if "useme" in want else None, C.byref(aid) if "avgid" in want else None)
    assert st == 0
    return w, use, aid.value


def test_gap_column_filter_matches_the_restatement(ctx, po):
    rng = np.random.default_rng(3)
    N, L = 300, 217
    ax = rng.choice(np.array([0, 1, 2, 3, 4, 15, 16, 17, 7], np.uint8), size=(N, L), p=[.2, .2, .2, .15, .2, .02, .01, .01, .01])
    gapfrac = rng.beta(0.6, 1.2, L)
    ax = np.where(rng.random((N, L)) < gapfrac[None, :], 4, ax).astype(np.uint8)
    ax[:, 5] = 4                                                            # an all-gap column (r = 0)
    ax[:, 6] = 17                                                           # all missing: r = tot = 0
    for gapthresh in (0.75, 0.5, 1.0):
        _, keep = po.remove_gap_columns(ax, None, gapthresh)
        use = ctx.msa_gap_columns(ax, None, gapthresh)
        assert np.array_equal(np.nonzero(use)[0], keep), gapthresh
    w = rng.gamma(2.0, 0.5, N)
    _, keep = po.remove_gap_columns(ax, w, 0.6)
    use = ctx.msa_gap_columns(ax, w, 0.6)
    assert np.array_equal(np.nonzero(use)[0], keep)
    sub = ctx.msa_column_subset(ax, use)
    assert np.array_equal(sub, ax[:, keep])


def test_gap_column_filter_device_resident_input(ctx, po):
    import torch
    msa, _, _ = po.synthetic_msa(500, 160, seed=5)
    use = ctx.msa_gap_columns(torch.from_numpy(msa).cuda(), None, 0.1)
    _, keep = po.remove_gap_columns(msa, None, 0.1)
    assert np.array_equal(np.nonzero(use)[0], keep) and 0 < len(keep) < 160


@pytest.mark.parametrize("N,L", [(1500, 120), (5000, 400), (37, 1000)])
def test_pb_weights_match_the_restatement(ctx, pkg, po, N, L):
    msa, _, _ = pkg.synth.synthetic_msa(N, L, seed=9)
    got = ctx.msa_pb_weights(msa)
    ref = po.weights_pb(msa)
    assert abs(got.sum() - N) <= 1e-9 * N
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, ref.max())


def test_pair_identity_and_distance_matrix(ctx, po):
    msa, _, _ = po.synthetic_msa(90, 133, seed=2)
    rng = np.random.default_rng(1)
    pairs = rng.integers(0, 90, size=(500, 2))
    got = ctx.msa_pair_identity(msa, pairs)
    ref = np.array([po.pair_identity(msa, a, b) for a, b in pairs])
    assert np.array_equal(got, ref)                                          # integer counts, one division: bit-identical
    D = ctx.msa_pair_identity(msa)
    assert np.array_equal(D, D.T) and not np.diag(D).any()
    i, j = np.triu_indices(90, 1)
    assert np.array_equal(D[i, j], 1.0 - np.array([po.pair_identity(msa, a, b) for a, b in zip(i, j)]))


def test_tutorial_banner_and_scores_from_device_preprocessing(ctx, pkg, po, glue):
    """documentation/tutorial.tex:187-212 end to end on the device: gap filter -> 'alen 66 (150)', average identity ->
    'avgid 65.82', GSC weights (nseq 95 <= 1000) -> the 11 GTp scores to 5 decimals."""
    ax, gold = _tutorial()
    use = ctx.msa_gap_columns(ax, None, 0.75)
    keep = np.nonzero(use)[0]
    assert len(keep) == gold["alen"] == 66
    sub = po.degen_to_N(ctx.msa_column_subset(ax, use))
    assert np.array_equal(sub, po.degen_to_N(po.remove_gap_columns(ax)[0]))
    wgt, _, avgid = host_prep(glue, sub, want=("wgt", "avgid"))              # msaweight_b200 (GSC) + esl_dst_XAverageId_b200
    assert round(100.0 * avgid, 2) == gold["avgid"] == 65.82
    assert avgid == po.average_id(sub)                                       # same pairs, same summation order: identical
    ref_w = po.weights_gsc(sub)
    assert abs(wgt.sum() - 95) < 1e-9 and np.max(np.abs(wgt - ref_w)) <= 1e-11
    ctx.configure(95, 66, 1, 0)
    ctx.set_weights(wgt)
    res = ctx.scan(sub, pkg.GT, pkg.C16, pkg.APC)
    col = {int(c) + 1: k for k, c in enumerate(keep)}
    for p in gold["pairs"]:
        assert round(res["cov"][col[p["i"]], col[p["j"]]], 5) == p["score"], p


def test_host_mirror_weights_gap_columns_and_sampled_identity(glue, pkg, po):
    """msaweight_b200 picks PB above maxsq_gsc; msamanip_GapColumns_b200 uses the alignment's weights; esl_dst_XAverageId_b200
    samples max_comparisons pairs from MT19937(42) when there are more pairs than that."""
    N, L = 1200, 90
    msa, w_in, _ = pkg.synth.synthetic_msa(N, L, seed=4)
    wgt, use, avgid = host_prep(glue, msa, wgt_in=w_in, what=0, maxsq_gsc=1000, gapthresh=0.3, max_comparisons=2000)
    assert np.max(np.abs(wgt - po.weights_pb(msa))) <= 1e-12 * wgt.max()
    _, keep = po.remove_gap_columns(msa, w_in, 0.3)
    assert np.array_equal(np.nonzero(use)[0], keep)
    # the sampled pair list of esl_dst_XAverageId, reproduced with the oracle's MT19937 (the same shim generator)
    ora = po.Oracle()
    rng = ora.rng(42)
    pairs = []
    while len(pairs) < 2000:
        i, j = int(ora.random(rng) * N), int(ora.random(rng) * N)
        while j == i:
            i, j = int(ora.random(rng) * N), int(ora.random(rng) * N)
        pairs.append((i, j))
    ora.rng_free(rng)
    assert avgid == po.average_id(msa, 2000, pairs)
    # GSC on a small alignment through the same entry
    small = msa[:60]
    w60, _, _ = host_prep(glue, small, what=0, want=("wgt",))
    assert np.max(np.abs(w60 - po.weights_gsc(small))) <= 1e-11
