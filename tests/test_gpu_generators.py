"""Device null generators vs the oracle's restatement of the reference generators (which is itself pinned residue for
residue against the reference code, tests/test_nullgen_oracle.py).  The device uses Philox streams, the reference one
Mersenne-Twister stream, so agreement is distributional (north_star: two-sample KS on null score histograms and on
substitution counts) plus exact checks on inputs where the generators are deterministic.

A: Fitch + shuffle (src/msatree.c:173-227,1700-1931; src/msamanip.c:1164-1233,1449-1780)
B: cov_GenerateAlignment, noss + noindels (src/cov_simulate.c:289-324,585-631,724-773; src/ratematrix.c:185-233)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q_TEST = np.array([[-1.00, 0.30, 0.50, 0.20], [0.25, -0.90, 0.15, 0.50], [0.60, 0.10, -0.95, 0.25], [0.20, 0.45, 0.30, -0.95]])


def ks_stat(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    ca = np.searchsorted(a, allv, side="right") / a.size
    cb = np.searchsorted(b, allv, side="right") / b.size
    return float(np.max(np.abs(ca - cb)))


@pytest.fixture(params=["rank", "position"])
def replay_form(request, monkeypatch):
    """Generator A's replay step has two kernels (nullgen.cu): the rank-table form (default) and the position form
    (RSCAPE_B200_REPLAY=position: copy the shuffled parent row, re-place each substitution on a uniformly drawn free column of its
    source class).  Same distribution, same exact invariants: the tests below run on both."""
    monkeypatch.setenv("RSCAPE_B200_REPLAY", request.param)
    return request.param


def _setup(ctx, po, N, L, seed, R):
    msa, wgt, _ = po.synthetic_msa(N, L, seed=seed)
    tree = po.random_tree(N, np.random.default_rng(seed))
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    return msa, wgt, tree


def _branch_subs(tree, allrows, N):
    """number of differing positions per branch given rows [leaves | internal nodes]"""
    out = []
    for v in range(N - 1):
        for ch in (tree.left[v], tree.right[v]):
            row = allrows[N + ch] if ch > 0 else allrows[-ch]
            out.append(int((allrows[N + v] != row).sum()))
    return np.array(out)


# ------------------------------------------------------------------------------------------------ generator A
def test_fitch_shuffle_identical_sequences_is_a_column_permutation(ctx, po, replay_form):
    N, L = 12, 50
    row = np.random.default_rng(1).integers(0, 5, L).astype(np.uint8)
    msa = np.tile(row, (N, 1))
    tree = po.random_tree(N, np.random.default_rng(2))
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(None)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(3)
    ctx.null_fitch_shuffle(msa, seed=5, nrep=3)
    out = ctx.pool_get(3)
    for r in range(3):
        assert (out[r] == out[r][0]).all()                                  # no substitutions on any branch
        assert np.array_equal(np.sort(out[r][0]), np.sort(row))             # a permutation of the columns
    assert not np.array_equal(out[0][0], out[1][0])                         # replicates use different permutations


def test_fitch_shuffle_is_keyed_by_replicate_id(ctx, po):
    msa, wgt, tree = _setup(ctx, po, 40, 64, 3, 6)
    ctx.null_fitch_shuffle(msa, seed=9, nrep=4, first_rep=0, first_id=10)
    a = ctx.pool_get(4)
    ctx.null_fitch_shuffle(msa, seed=9, nrep=2, first_rep=4, first_id=12)   # ids 12,13 again, other pool entries, other batch
    b = ctx.pool_get(2, first_rep=4)
    assert np.array_equal(a[2:], b)
    assert not np.array_equal(a[0], a[1])
    assert a.max() <= 4                                                      # only residues and gaps, never N (msamanip.c:1634-1645)


def test_fitch_shuffle_distribution_matches_oracle(ctx, pkg, po, oracle, replay_form):
    N, L, R = 48, 90, 24
    msa, wgt, tree = _setup(ctx, po, N, L, 7, R)
    ctx.null_fitch_shuffle(msa, seed=11, nrep=R)
    gpu = ctx.pool_get(R)
    rng = oracle.rng(11)
    cpu = np.stack([oracle.null_fitch_shuffle(rng, tree, msa) for _ in range(R)])
    oracle.rng_free(rng)
    # residue composition of every null equals (statistically) that of the oracle's nulls
    comp_g = np.stack([(gpu == a).mean(axis=(1, 2)) for a in range(5)])       # [5][R]
    comp_c = np.stack([(cpu == a).mean(axis=(1, 2)) for a in range(5)])
    assert np.max(np.abs(comp_g.mean(1) - comp_c.mean(1))) < 0.01
    # per-sequence composition is (approximately) kept, the author's intent at src/R-scape.c:1664-1667
    seqcomp_g = np.stack([(gpu == a).mean(axis=2).mean(axis=0) for a in range(5)])   # [5][N]
    seqcomp_c = np.stack([(cpu == a).mean(axis=2).mean(axis=0) for a in range(5)])
    assert np.max(np.abs(seqcomp_g - seqcomp_c)) < 0.03
    # pairwise differences between leaves (substitution counts along the tree) have the same distribution
    def pairdiff(x):
        i, j = np.triu_indices(N, 1)
        return np.concatenate([(x[r][i] != x[r][j]).sum(1) for r in range(R)])
    assert ks_stat(pairdiff(gpu), pairdiff(cpu)) < 0.05
    # null GTp score distributions (what the E-values are made of)
    def scores(x):
        iu = np.triu_indices(L, 1)
        return np.concatenate([oracle.scan(x[r], wgt, po.GT, po.C16, po.APC)["cov"][iu] for r in range(R)])
    sg, sc = scores(gpu), scores(cpu)
    assert ks_stat(sg, sc) < 0.03
    for qt in (0.5, 0.9, 0.99, 0.999):
        a, b = np.quantile(sg, qt), np.quantile(sc, qt)
        assert abs(a - b) <= 0.08 * max(1.0, abs(b)) + 0.5, (qt, a, b)


def _branch_tables(tree, leaves, internal, N):
    """5x5 table of (parent residue, child residue) over the columns of every branch, residues <= 4 only
    (the substitution counts of shuffle_tree_substitutions, src/msamanip.c:1634-1646): int [2(N-1)][25]"""
    rows = np.concatenate([leaves, internal])                                # [2N-1][L]: leaves 0..N-1, node v at N + v
    par = np.repeat(np.arange(N - 1) + N, 2)
    kid = np.stack([tree.left, tree.right], 1).ravel()
    kid = np.where(kid > 0, kid + N, -kid)
    out = np.zeros((2 * (N - 1), 25), np.int64)
    step = max(1, (1 << 24) // max(1, rows.shape[1]))
    for b0 in range(0, len(par), step):
        pa, kd = rows[par[b0:b0 + step]], rows[kid[b0:b0 + step]]
        ok = (pa <= 4) & (kd <= 4)
        code = np.where(ok, pa.astype(np.int64) * 5 + kd, 25)
        for c in range(25):
            out[b0:b0 + step, c] = (code == c).sum(1)
    return out


def _check_generator_a_invariants(ctx, tree, msa, R):
    """What shuffle_tree_substitutions keeps exactly, whatever the random stream: on every branch of every replicate the
    5x5 substitution table of the shuffled rows equals that of the Fitch rows (msamanip.c:1634-1646, 1718-1757); the
    shuffled root row is a permutation of the Fitch root row (msamanip_ShuffleColumns, :1164-1233); leaves hold residues
    and gaps only."""
    N, L = msa.shape
    out = ctx.pool_get(R)
    anc, shanc = ctx.pool_get_internal(0, R), ctx.pool_get_internal(1, R)
    assert out.max() <= 4 and anc.max() <= 4 and shanc.max() <= 4
    off = ~np.eye(5, dtype=bool).ravel()
    for r in range(R):
        assert np.array_equal(np.bincount(anc[r][0], minlength=5), np.bincount(shanc[r][0], minlength=5))
        want = _branch_tables(tree, msa, anc[r], N)
        got = _branch_tables(tree, out[r], shanc[r], N)
        # substitutions (off-diagonal cells) are reproduced exactly; N in a leaf of the input is not a substitution (:1639)
        assert np.array_equal(got[:, off], want[:, off]), np.argwhere(got[:, off] != want[:, off])[:5]
    return out, anc, shanc


@pytest.mark.parametrize("N,L,kernel", [(48, 92, "<4,seg>"), (40, 120, "<4,seg>"), (24, 4096, "<4,ballot>"), (48, 90, "<1,ballot>")])
def test_fitch_shuffle_exact_invariants_every_kernel_variant(ctx, po, N, L, kernel, replay_form):
    """The replay kernel variant depends on L (nullgen.cu: L % 4 == 0 && L <= 4092 -> segment form, L % 4 == 0 -> word/ballot,
    else byte/ballot); every BASELINE shape takes the first.  Exact invariants on each."""
    R = 6
    msa, wgt, tree = _setup(ctx, po, N, L, 21, R)
    ctx.null_fitch_shuffle(msa, seed=13, nrep=R)
    out, anc, shanc = _check_generator_a_invariants(ctx, tree, msa, R)
    assert not np.array_equal(out[0], out[1])


@pytest.mark.parametrize("N,L", [(48, 92), (40, 120)])
def test_fitch_shuffle_distribution_matches_oracle_word_kernels(ctx, pkg, po, oracle, N, L, replay_form):
    """The KS battery of test_fitch_shuffle_distribution_matches_oracle at L % 4 == 0, i.e. on the kernels every BASELINE
    shape runs: fitch_up/down_level_kernel<4> and replay_level_row_kernel<4, segment form>."""
    R = 24
    msa, wgt, tree = _setup(ctx, po, N, L, 7, R)
    ctx.null_fitch_shuffle(msa, seed=11, nrep=R)
    gpu = ctx.pool_get(R)
    rng = oracle.rng(11)
    cpu = np.stack([oracle.null_fitch_shuffle(rng, tree, msa) for _ in range(R)])
    oracle.rng_free(rng)
    comp_g = np.stack([(gpu == a).mean(axis=(1, 2)) for a in range(5)])
    comp_c = np.stack([(cpu == a).mean(axis=(1, 2)) for a in range(5)])
    assert np.max(np.abs(comp_g.mean(1) - comp_c.mean(1))) < 0.01
    seqcomp_g = np.stack([(gpu == a).mean(axis=2).mean(axis=0) for a in range(5)])
    seqcomp_c = np.stack([(cpu == a).mean(axis=2).mean(axis=0) for a in range(5)])
    assert np.max(np.abs(seqcomp_g - seqcomp_c)) < 0.03
    i, j = np.triu_indices(N, 1)
    pairdiff = lambda x: np.concatenate([(x[r][i] != x[r][j]).sum(1) for r in range(R)])
    assert ks_stat(pairdiff(gpu), pairdiff(cpu)) < 0.05
    iu = np.triu_indices(L, 1)
    scores = lambda x: np.concatenate([oracle.scan(x[r], wgt, po.GT, po.C16, po.APC)["cov"][iu] for r in range(R)])
    sg, sc = scores(gpu), scores(cpu)
    assert ks_stat(sg, sc) < 0.03
    for qt in (0.5, 0.9, 0.99, 0.999):
        a, b = np.quantile(sg, qt), np.quantile(sc, qt)
        assert abs(a - b) <= 0.08 * max(1.0, abs(b)) + 0.5, (qt, a, b)


def test_fitch_shuffle_distribution_long_alignment_ballot_kernel(ctx, pkg, po, oracle, replay_form):
    """L > 4092 with L % 4 == 0: replay_level_row_kernel<4, ballot form> against the oracle (composition, leaf distances)."""
    N, L, R = 24, 4096, 8
    msa, wgt, tree = _setup(ctx, po, N, L, 5, R)
    ctx.null_fitch_shuffle(msa, seed=3, nrep=R)
    gpu = ctx.pool_get(R)
    rng = oracle.rng(3)
    cpu = np.stack([oracle.null_fitch_shuffle(rng, tree, msa) for _ in range(R)])
    oracle.rng_free(rng)
    comp_g = np.stack([(gpu == a).mean(axis=(1, 2)) for a in range(5)])
    comp_c = np.stack([(cpu == a).mean(axis=(1, 2)) for a in range(5)])
    assert np.max(np.abs(comp_g.mean(1) - comp_c.mean(1))) < 0.01
    i, j = np.triu_indices(N, 1)
    pairdiff = lambda x: np.concatenate([(x[r][i] != x[r][j]).sum(1) for r in range(R)])
    assert ks_stat(pairdiff(gpu) / L, pairdiff(cpu) / L) < 0.08
    # leaf-pair distances, pair by pair (means over replicates): the tree's substitution counts are reproduced
    dg = np.stack([(gpu[r][i] != gpu[r][j]).mean(1) for r in range(R)]).mean(0)
    dc = np.stack([(cpu[r][i] != cpu[r][j]).mean(1) for r in range(R)]).mean(0)
    assert np.max(np.abs(dg - dc)) < 0.02


def test_fitch_shuffle_byte_and_word_kernels_agree(ctx, pkg, po, oracle, replay_form):
    """<1> (L % 4 != 0) against <4> on the same input: an appended all-gap column changes the kernel variant but carries no
    substitution and never makes two leaves differ, so leaf-pair differences must have the same distribution."""
    N, L, R = 48, 92, 32
    msa, wgt, tree = _setup(ctx, po, N, L, 7, R)
    msa = np.where(msa > 4, 0, msa).astype(np.uint8)
    ctx.null_fitch_shuffle(msa, seed=17, nrep=R)
    a = ctx.pool_get(R)
    msa1 = np.concatenate([msa, np.full((N, 1), 4, np.uint8)], 1)
    ctx.configure(N, L + 1, 2, 0)
    ctx.set_weights(None)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    ctx.null_fitch_shuffle(msa1, seed=18, nrep=R)
    b = ctx.pool_get(R)
    _check_generator_a_invariants(ctx, tree, msa1, 4)
    i, j = np.triu_indices(N, 1)
    pd = lambda x: np.concatenate([(x[r][i] != x[r][j]).sum(1) for r in range(R)])
    assert ks_stat(pd(a), pd(b)) < 0.04
    assert abs(pd(a).mean() - pd(b).mean()) < 0.02 * pd(a).mean() + 0.2


def test_fitch_shuffle_exact_invariants_at_the_ssu_shape(ctx, pkg, replay_form):
    """The BASELINE config 3 shape itself (N = 10000, L = 1800): every one of the 19 998 branches of two replicates keeps its
    5x5 substitution table; the root row is a permutation; only residues and gaps come out."""
    N, L, R = 10000, 1800, 2
    msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=42)
    ctx.configure(N, L, 2, 4)
    ctx.set_weights(wgt)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(R)
    ctx.null_fitch_shuffle(msa, seed=20261017, nrep=R)
    out, anc, shanc = _check_generator_a_invariants(ctx, tree, msa, R)
    # the Fitch rows are a most-parsimonious reconstruction: a child differs from its parent only where the parent's
    # residue is outside the child's set, so the number of substitutions per column is at most the number of leaves - 1
    assert not np.array_equal(out[0], out[1])
    # column composition is NOT kept (columns are permuted), the alignment's total composition nearly is
    comp_in = np.bincount(np.minimum(msa, 5).ravel(), minlength=6)[:5] / msa.size
    comp_out = np.bincount(out[0].ravel(), minlength=5) / out[0].size
    assert np.max(np.abs(comp_in - comp_out)) < 0.02


# ------------------------------------------------------------------------------------------------ generator B
def test_simulate_identity_and_stationary_limits(ctx, po):
    N, L = 30, 200
    tree = po.random_tree(N, np.random.default_rng(4))
    ctx.configure(N, L, 2, 0)
    ctx.set_weights(None)
    root = np.random.default_rng(5).integers(0, 4, L).astype(np.uint8)
    # zero rate: P = I on every branch, every leaf equals the root
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(2)
    ctx.null_simulate(np.zeros((4, 4)), root, seed=1, nrep=2)
    out = ctx.pool_get(2)
    assert (out == root[None, None, :]).all()
    # very long branches: leaves are draws from the stationary distribution of Q, whatever the root
    long_tree = po.Tree(tree.left, tree.right, tree.parent, np.full(N - 1, 200.0), np.full(N - 1, 200.0))
    ctx.set_tree(long_tree.left, long_tree.right, long_tree.parent, long_tree.ld, long_tree.rd)
    ctx.null_simulate(Q_TEST, root, seed=2, nrep=2)
    out = ctx.pool_get(2)
    w, v = np.linalg.eig(Q_TEST.T)
    pi = np.real(v[:, np.argmin(np.abs(w))]); pi /= pi.sum()
    freq = np.array([(out == a).mean() for a in range(4)])
    assert np.max(np.abs(freq - pi)) < 0.02


def test_simulate_gap_mask_and_replicate_ids(ctx, po):
    msa, wgt, tree = _setup(ctx, po, 35, 70, 8, 5)
    root = np.random.default_rng(1).integers(0, 4, 70).astype(np.uint8)
    ctx.null_simulate(Q_TEST, root, seed=3, nrep=3, gapmask=msa, first_id=100)
    a = ctx.pool_get(3)
    noncanon = msa >= 4
    assert (a[:, noncanon] == msa[noncanon][None, :]).all() and (a[:, ~noncanon] < 4).all()
    ctx.null_simulate(Q_TEST, root, seed=3, nrep=1, gapmask=msa, first_rep=4, first_id=101)
    assert np.array_equal(ctx.pool_get(1, first_rep=4)[0], a[1])


def test_simulate_distribution_matches_oracle(ctx, pkg, po, oracle):
    N, L, R = 40, 120, 24
    msa, wgt, tree = _setup(ctx, po, N, L, 12, R)
    root = np.random.default_rng(2).integers(0, 4, L).astype(np.uint8)
    ctx.null_simulate(Q_TEST, root, seed=21, nrep=R)
    gpu = ctx.pool_get(R)
    rng = oracle.rng(21)
    cpu = np.stack([oracle.null_simulate(rng, tree, Q_TEST, root) for _ in range(R)])
    oracle.rng_free(rng)
    # per-column substitution counts relative to the root (north_star's second KS)
    sub_g = (gpu != root[None, None, :]).sum(axis=1).ravel()
    sub_c = (cpu != root[None, None, :]).sum(axis=1).ravel()
    assert ks_stat(sub_g, sub_c) < 0.04
    # per-leaf distance to the root follows the branch lengths: leaf by leaf means agree
    dg = (gpu != root[None, None, :]).mean(axis=(0, 2))
    dc = (cpu != root[None, None, :]).mean(axis=(0, 2))
    assert np.max(np.abs(dg - dc)) < 0.04
    iu = np.triu_indices(L, 1)
    sg = np.concatenate([oracle.scan(gpu[r], wgt, po.GT, po.C16, po.APC)["cov"][iu] for r in range(R)])
    sc = np.concatenate([oracle.scan(cpu[r], wgt, po.GT, po.C16, po.APC)["cov"][iu] for r in range(R)])
    assert ks_stat(sg, sc) < 0.03


def test_pool_scan_equals_host_scan(ctx, pkg, po, oracle):
    """Nulls scanned in place from the device pool give the same histogram as the same alignments passed from the host."""
    N, L, R = 60, 50, 5
    msa, wgt, tree = _setup(ctx, po, N, L, 15, R)
    ctx.null_fitch_shuffle(msa, seed=4, nrep=R)
    nulls = ctx.pool_get(R)
    w, mn, mx = ctx.null_width_pool(0)
    w2, mn2, mx2 = ctx.null_width(nulls[0])
    assert (w, mn, mx) == (w2, mn2, mx2)
    ctx.null_hist_pool(0, R, w)
    a = ctx.hist_read(4096)
    ctx.hist_reset()
    ctx.null_hist(nulls, w)
    b = ctx.hist_read(4096)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] == R * L * (L - 1) // 2


# ------------------------------------------------------------------------------------------------ plumbing of generator A
def test_fitch_shuffle_is_deterministic_and_chunk_independent(ctx, po, replay_form):
    """Replicates are keyed by their global id: generating them in one call, again, or one by one into other pool
    entries gives the same alignments (the generation stream works in chunks of growing size)."""
    msa, wgt, tree = _setup(ctx, po, 60, 44, seed=9, R=24)
    ctx.null_fitch_shuffle(msa, seed=77, nrep=24)
    a = ctx.pool_get(24)
    ctx.null_fitch_shuffle(msa, seed=77, nrep=24)
    assert np.array_equal(ctx.pool_get(24), a)
    for r in (0, 5, 23):
        ctx.null_fitch_shuffle(msa, seed=77, nrep=1, first_rep=2, first_id=r)
        assert np.array_equal(ctx.pool_get(1, 2)[0], a[r])
    # an explicit id list (replicate 0 + a block, as a rank of a multi-GPU run asks for)
    ids = [0, 17, 18, 19, 20, 5]
    ctx.null_fitch_shuffle(msa, seed=77, nrep=len(ids), first_rep=3, ids=ids)
    got = ctx.pool_get(len(ids), 3)
    for k, r in enumerate(ids):
        assert np.array_equal(got[k], a[r])


def test_fitch_shuffle_shared_up_pass_equals_per_replicate(ctx, po, monkeypatch):
    """Without unknown residues the Fitch sets are computed once for all replicates; the result must be what the general
    per-replicate path gives."""
    msa, wgt, tree = _setup(ctx, po, 80, 52, seed=4, R=6)
    msa = np.where(msa > 4, 0, msa).astype(np.uint8)                          # no N
    ctx.null_fitch_shuffle(msa, seed=3, nrep=6)
    shared = ctx.pool_get(6)
    monkeypatch.setenv("RSCAPE_B200_FITCH_PER_REPLICATE", "1")
    ctx.null_fitch_shuffle(msa, seed=3, nrep=6)
    assert np.array_equal(ctx.pool_get(6), shared)


def test_scans_wait_for_the_generation_stream(ctx, po):
    """null_hist_pool issued right after the (asynchronous) generator call sees the finished alignments: its histogram
    equals the one from the same alignments uploaded from the host."""
    msa, wgt, tree = _setup(ctx, po, 70, 48, seed=12, R=20)
    ctx.null_fitch_shuffle(msa, seed=11, nrep=20)
    ctx.hist_reset()
    w, _, _ = ctx.null_width_pool(0)
    mm = ctx.null_hist_pool(0, 20, w)
    bins, n, imax = ctx.hist_read(4096)
    nulls = ctx.pool_get(20)
    ctx.hist_reset()
    w2, _, _ = ctx.null_width(nulls[0])
    mm2 = ctx.null_hist(nulls, w2)
    bins2, n2, imax2 = ctx.hist_read(4096)
    assert w == w2 and n == n2 and imax == imax2
    assert np.array_equal(bins, bins2) and np.array_equal(mm, mm2)


def test_fitch_shuffle_rejects_codes_the_reference_cannot_set_up(ctx, pkg, po):
    """Only esl_abc_XIsUnknown (N) gets the uniform Fitch set; any other non-canonical code makes the reference fail with
    "S not set up properly" (src/msatree.c:1731-1740).  The device refuses such an alignment instead of resolving the code at random."""
    N, L = 24, 40
    msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=9)
    ctx.configure(N, L, 2, 0)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.pool_reserve(2)
    ok = msa.copy()
    ok[3, 5] = 15                                                     # N: fine
    ctx.null_fitch_shuffle(ok, 1, 2)
    bad = msa.copy()
    bad[3, 5] = 5                                                     # R (degenerate): msamanip_ConvertDegen2N must have run first
    with pytest.raises(pkg.RscapeB200Error, match="Fitch"):
        ctx.null_fitch_shuffle(bad, 1, 2)
