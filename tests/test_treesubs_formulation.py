"""The device computes Tree_Substitutions' pair tables (src/msatree.c:1480-1530) as an unweighted pair-count table over one row
per branch (r-scape_b200/csrc/treesubs.cu).  This CPU test checks that formulation itself with the oracle's pair counter:
ndouble = C[0,0], njoin = C[0,0] + C[0,1] + C[1,0] over rows coded 0 = substituted, 1 = not substituted, 4 = branch not counted."""
import numpy as np
import pytest


@pytest.mark.parametrize("N,L,seed", [(60, 40, 1), (300, 25, 2)])
def test_branch_row_count_table_gives_ndouble_and_njoin(po, oracle, N, L, seed):
    msa = po.synthetic_msa(N, L, seed=seed)[0]
    tree = po.random_tree(N, np.random.default_rng(seed))
    rng = oracle.rng(seed)
    _, allm, _ = oracle.null_fitch_shuffle(rng, tree, msa, want_all=True)
    oracle.rng_free(rng)
    leaves, internal = allm[:N], allm[N:]
    kids = np.concatenate([np.stack([tree.left, tree.right], 1).ravel()])            # branch e = 2 v + side
    par = np.repeat(np.arange(N - 1), 2)
    child = np.where(kids[:, None] > 0, internal[np.maximum(kids, 0)], leaves[np.maximum(-kids, 0)])
    parent = internal[par]
    for includegaps in (False, True):
        valid = np.ones_like(child, bool) if includegaps else (child < 4) & (parent < 4)
        codes = np.where(valid, np.where(child != parent, 0, 1), 4).astype(np.uint8)
        cnt = oracle.counts_fixed(codes, np.ones(len(codes), np.int64))
        ns, nd, nj = oracle.tree_substitutions(tree, allm, includegaps)
        assert np.array_equal((codes == 0).sum(0), ns)
        assert np.array_equal(np.triu(cnt[:, :, 0], 1), nd)
        assert np.array_equal(np.triu(cnt[:, :, 0] + cnt[:, :, 1] + cnt[:, :, 4], 1), nj)
