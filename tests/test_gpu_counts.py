"""Pair counts on the tensor cores vs the CPU oracle: bit-exact (integer fixed-point arithmetic).

Reference: mutual_naive_ppij accumulate loop, src/correlators.c:1724-1755.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, msa, wgt, nslices):
    N, L = msa.shape
    ctx.configure(N, L, 1, nslices)
    ctx.set_weights(wgt)
    ctx.scan(msa, want_cov=False)
    wq, q, S = ctx.quantisation()
    if nslices:
        assert S == nslices
    ref = np.triu(oracle.counts_fixed(msa, wq).transpose(2, 0, 1), 1)
    got = ctx.counts()
    if not np.array_equal(got, ref):
        bad = np.argwhere(got != ref)
        raise AssertionError(f"{len(bad)} of {ref.size} counts differ; first {bad[:5].tolist()} got {got[tuple(bad[0])]} want {ref[tuple(bad[0])]}")
    # the direct (non tensor core) verification kernel must agree as well
    assert np.array_equal(ctx.counts_direct(msa), ref)
    # quantisation: wq = u V with u < 256, V < 256^S; the library reports its own worst error, which must be what we see
    # and no worse than ~8 S bits below the largest weight
    if wgt is not None:
        err = float(np.max(np.abs(wgt - wq.astype(np.longdouble) * np.longdouble(2.0) ** (-q))))
        abs_err, bits = ctx.quantisation_error()
        assert err <= abs_err * (1 + 1e-9) + 1e-300
        if S > 1 or q != 0:
            assert bits >= 8 * S - 1, (bits, S)


@pytest.mark.parametrize("nslices", [1, 2, 3, 4, 5, 6])
def test_counts_all_slice_counts(ctx, po, oracle, nslices):
    msa, wgt, _ = po.synthetic_msa(333, 77, seed=nslices)
    if nslices == 1:
        wgt = np.ones(333)
    _check(ctx, oracle, msa, wgt, nslices)


@pytest.mark.parametrize("N,L", [(1, 2), (2, 2), (5, 3), (129, 33), (128, 32), (257, 64), (1000, 76), (640, 130)])
def test_counts_shapes(ctx, po, oracle, N, L):
    msa, wgt, _ = po.synthetic_msa(N, L, seed=N + L)
    _check(ctx, oracle, msa, wgt, 0)


def test_counts_accumulator_headroom(ctx, po, oracle):
    """Many sequences with near-maximal weights: the int32 accumulators hold sum_s u_s d_s <= N 255^2."""
    rng = np.random.default_rng(11)
    N, L = 20000, 36
    msa = rng.integers(0, 2, (N, L)).astype(np.uint8)        # two residues only: large counts per cell
    _check(ctx, oracle, msa, rng.uniform(0.97, 1.0, N), 4)


def test_counts_unit_weights_auto_single_slice(ctx, po, oracle):
    msa, _, _ = po.synthetic_msa(500, 60, seed=3)
    _check(ctx, oracle, msa, np.ones(500), 0)
    assert ctx.quantisation()[2] == 1 and ctx.quantisation()[1] == 0


def test_counts_all_gaps_and_unknowns(ctx, po, oracle):
    rng = np.random.default_rng(5)
    msa = rng.integers(0, 4, (200, 40)).astype(np.uint8)
    msa[:, 3] = 4            # an all-gap column
    msa[:, 7] = 15           # an all-N column
    msa[17, :] = 4           # an all-gap sequence
    wgt = rng.gamma(2.0, 0.5, 200)
    _check(ctx, oracle, msa, wgt, 0)


def test_counts_extreme_weights(ctx, po, oracle):
    rng = np.random.default_rng(9)
    msa = rng.integers(0, 5, (300, 50)).astype(np.uint8)
    wgt = np.concatenate([rng.uniform(1e-6, 1e-3, 100), rng.uniform(0.5, 2, 199), [250.0]])
    _check(ctx, oracle, msa, wgt, 6)
    wgt[0] = 0.0
    _check(ctx, oracle, msa, wgt, 5)


@pytest.mark.parametrize("nslices,N,L", [(1, 300, 70), (4, 333, 77), (5, 257, 130), (3, 1000, 33)])
def test_counts_cta_pair_kernel(ctx, po, oracle, monkeypatch, nslices, N, L):
    """The cta_group::2 variant of the contraction (two SMs per 256-row tile; opt-in) gives the same bits."""
    monkeypatch.setenv("RSCAPE_B200_GRAM_PAIR", "1")
    msa, wgt, _ = po.synthetic_msa(N, L, seed=N)
    _check(ctx, oracle, msa, np.ones(N) if nslices == 1 else wgt, nslices)


def test_counts_many_sequences_caps_the_multiplier(ctx, po, oracle):
    """N > 33 025: the 8-bit multiplier is capped so that sum_s u_s d_s stays inside the int32 accumulators."""
    rng = np.random.default_rng(21)
    N, L = 40000, 12
    msa = rng.integers(0, 3, (N, L)).astype(np.uint8)
    wgt = rng.uniform(0.5, 1.0, N)
    _check(ctx, oracle, msa, wgt, 4)
    wq, q, S = ctx.quantisation()
    assert int(wq.max()) < 210 * 256 ** 4          # u <= floor((2^31 - 1) / (255 Kpad)) = 210
