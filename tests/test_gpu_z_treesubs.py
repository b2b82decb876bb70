"""Tree_Substitutions on the device (rsb_tree_substitutions: one row per branch + the unweighted tcgen05 pair contraction)
vs the oracle's restatement of src/msatree.c:1455-1540, which equals the reference's own function on the shared RNG stream
(tests/test_nullgen_oracle.py).  Integer counts: exact."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(po, oracle, N, L, seed):
    msa = po.synthetic_msa(N, L, seed=seed)[0]
    tree = po.random_tree(N, np.random.default_rng(seed))
    rng = oracle.rng(seed)
    _, allm, _ = oracle.null_fitch_shuffle(rng, tree, msa, want_all=True)
    oracle.rng_free(rng)
    return tree, allm


@pytest.mark.parametrize("N,L,seed", [(12, 30, 1), (100, 131, 2), (300, 64, 3), (2, 9, 4), (65, 33, 5)])
@pytest.mark.parametrize("includegaps", [False, True])
def test_tree_substitutions_equal_oracle(ctx, pkg, po, oracle, N, L, seed, includegaps):
    tree, allm = _case(po, oracle, N, L, seed)
    want = oracle.tree_substitutions(tree, allm, includegaps)
    ctx.configure(2 * (N - 1), L, 1, 1)                                 # one row per branch
    got = ctx.tree_substitutions(tree.left, tree.right, allm[:N], allm[N:], includegaps)
    for a, b, name in zip(got, want, ("nsubs", "ndouble", "njoin")):
        assert np.array_equal(a, b), (name, np.argwhere(a != b)[:5])
    only = ctx.tree_substitutions(tree.left, tree.right, allm[:N], allm[N:], includegaps, want_pairs=False)
    assert np.array_equal(only[0], want[0]) and only[1] is None


def test_tree_substitutions_argument_checks(ctx, pkg, po, oracle):
    tree, allm = _case(po, oracle, 20, 16, 7)
    ctx.configure(20, 16, 1, 1)                                         # wrong: sequences instead of branches
    with pytest.raises(pkg.RscapeB200Error, match="one row per branch"):
        ctx.tree_substitutions(tree.left, tree.right, allm[:20], allm[20:])
    ctx.configure(38, 16, 1, 1)
    bad = tree.left.copy()
    bad[3] = 99
    with pytest.raises(pkg.RscapeB200Error, match="outside the tree"):
        ctx.tree_substitutions(bad, tree.right, allm[:20], allm[20:])
    got = ctx.tree_substitutions(tree.left, tree.right, allm[:20], allm[20:])
    assert np.array_equal(got[1], oracle.tree_substitutions(tree, allm)[1])


def test_tree_substitutions_host_mirror(po, oracle):
    """Tree_Substitutions_b200 (host mirror over ESL_MSA / ESL_TREE) allocates and fills the reference's three arrays."""
    glue = C.CDLL(os.path.join(ROOT, "oracle", "libglue_b200.so"))
    N, L = 90, 70
    tree, allm = _case(po, oracle, N, L, 8)
    vp = C.c_void_p
    glue.glue_tree_substitutions.argtypes = [C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_char_p]
    for includegaps in (0, 1):
        ns, nd, nj = np.zeros(L, np.int32), np.zeros((L, L), np.int32), np.zeros((L, L), np.int32)
        err = C.create_string_buffer(256)
        left, right = np.ascontiguousarray(tree.left, np.int32), np.ascontiguousarray(tree.right, np.int32)
        st = glue.glue_tree_substitutions(N, left.ctypes.data, right.ctypes.data, L, np.ascontiguousarray(allm).ctypes.data, includegaps, 1,
                                          ns.ctypes.data, nd.ctypes.data, nj.ctypes.data, err)
        assert st == 0, err.value
        want = oracle.tree_substitutions(tree, allm, bool(includegaps))
        assert np.array_equal(ns, want[0]) and np.array_equal(nd, want[1]) and np.array_equal(nj, want[2])


def test_tree_substitutions_full_size_properties(pkg):
    """RNase-P-sized tree (5000 taxa, L = 400; 9998 branch rows): relations that hold for any reconstruction."""
    import __graft_entry__ as ge
    synth = ge.load_package().synth
    N, L = 5000, 400
    rng = np.random.default_rng(12)
    tree = synth.random_tree(N, rng)
    # any rows serve as "ancestral sequences" for the counting identities: evolve residues down the tree with rare changes and gaps
    internal = np.zeros((N - 1, L), np.uint8)
    leaves = np.zeros((N, L), np.uint8)
    internal[0] = rng.integers(0, 4, L)
    for v in range(N - 1):
        for kid in (tree.left[v], tree.right[v]):
            row = internal[v].copy()
            m = rng.random(L) < 0.02
            row[m] = rng.integers(0, 5, int(m.sum()))
            if kid > 0:
                internal[kid] = row
            else:
                leaves[-kid] = row
    c = pkg.Context(0)
    c.configure(2 * (N - 1), L, 1, 1)
    iu = np.triu_indices(L, 1)
    ns_g, nd_g, nj_g = c.tree_substitutions(tree.left, tree.right, leaves, internal, includegaps=True)
    ns, nd, nj = c.tree_substitutions(tree.left, tree.right, leaves, internal, includegaps=False)
    c.close()
    # single-column counts straight from the rows
    kids = np.concatenate([tree.left, tree.right])
    par = np.concatenate([np.arange(N - 1), np.arange(N - 1)])
    child_rows = np.where(kids[:, None] > 0, internal[np.maximum(kids, 0)], leaves[np.maximum(-kids, 0)])
    par_rows = internal[par]
    changed_g = child_rows != par_rows
    valid = (child_rows < 4) & (par_rows < 4)
    assert np.array_equal(ns_g, changed_g.sum(0)) and np.array_equal(ns, (changed_g & valid).sum(0))
    # with gaps included every branch counts: |A or B| = |A| + |B| - |A and B|
    assert np.array_equal(nj_g[iu], ns_g[iu[0]] + ns_g[iu[1]] - nd_g[iu])
    assert (nd[iu] <= np.minimum(ns[iu[0]], ns[iu[1]])).all() and (nd[iu] <= nd_g[iu]).all() and (nj[iu] <= nj_g[iu]).all()
    # a few pairs in full
    ch = changed_g & valid
    for i, j in ((0, 1), (17, 399), (200, 201), (5, 250)):
        both = valid[:, i] & valid[:, j]
        assert nd[i, j] == int((ch[:, i] & ch[:, j]).sum())
        assert nj[i, j] == int((both & (ch[:, i] | ch[:, j])).sum())
    assert not np.tril(nd).any() and not np.tril(nj_g).any()
