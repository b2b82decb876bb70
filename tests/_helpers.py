"""Shared assertions of the parity tests."""
import numpy as np


def edge_ambiguous(scores, bmin, w, rel=1e-9, scale=1.0):
    """How many of the (oracle's) scores lie within the parity tolerance of a histogram bin edge: only those may
    legitimately fall into the neighbouring bin on the device ('assert on (bin index, call)')."""
    x = np.maximum(np.asarray(scores, dtype=np.float64), bmin + w)
    t = (x - bmin) / w
    d = np.abs(t - np.round(t)) * w                                   # distance to the nearest edge, in score units
    tol = rel * np.maximum(np.maximum(1.0, np.abs(x)), scale)
    return int(np.count_nonzero(d <= tol))


def assert_bins_identical(bins, ref_obs, scores, bmin, w, rel=1e-9, scale=1.0):
    """Integer bins identical to the oracle's; a difference is tolerated only as far as it is explained by scores
    sitting on a bin edge (each such score can move one count from one bin to its neighbour: 2 per score)."""
    bins = np.asarray(bins).astype(np.int64)
    ref = np.asarray(ref_obs).astype(np.int64)
    n = max(len(bins), len(ref))
    b, r = np.zeros(n, np.int64), np.zeros(n, np.int64)
    b[:len(bins)] = bins
    r[:len(ref)] = ref
    assert b.sum() == r.sum(), (b.sum(), r.sum())
    diff = int(np.abs(b - r).sum())
    if diff:
        amb = edge_ambiguous(scores, bmin, w, rel, scale)
        assert diff <= 2 * amb, f"{diff} counts differ from the oracle's bins, {amb} scores lie on a bin edge: {np.argwhere(b != r)[:6].ravel()}"
    return diff
