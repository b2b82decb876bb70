"""The pipelined null loop on a pair grid sharded over TWO GPUs, the per-scan vectors (marginal sums, APC row sums, score range;
SURVEY 8e-2 / K7) summed by the library's one-shot all-reduce kernel over NVLink peer memory (csrc/peer_reduce.cu), and again
by ncclAllReduce: the summed histogram must be the unsharded device loop's and the oracle's, the two collective paths must give
identical bins and ranges.  One process, one context and one host thread per device (rsb_comm_init_all).  Skipped on a box
with a single GPU; run with `gpurun --gpus 2`."""
import threading

import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu


def _ndev(pkg):
    return int(pkg.lib().rsb_device_count())


def _sharded_loop(pkg, world, N, L, slots, wgt, nulls, w, msa):
    ctxs = []
    for k in range(world):
        c = pkg.Context(k)
        c.configure(N, L, slots, 0)
        ctxs.append(c)
    pkg.comm_init_all(ctxs)
    out, err = [None] * world, [None] * world

    def work(k):
        try:
            c = ctxs[k]
            c.set_shard(k, world)
            c.set_weights(wgt)
            c.hist_reset()
            mm = c.null_hist(nulls, w)                             # every rank scans its row blocks of EVERY null; ranges all-reduced
            res = c.sharded_scan(msa, want_cov=True)               # the input alignment on the sharded grid, matrix assembled on every rank
            bins, n, _ = c.hist_read(1 << 15)
            out[k] = (mm, bins.astype(np.int64), n, res, c.comm_info())
        except Exception as e:                                      # noqa: BLE001
            err[k] = e

    th = [threading.Thread(target=work, args=(k,)) for k in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "sharded null loop hung"
    for c in ctxs:
        c.comm_destroy()
        c.close()
    assert err == [None] * world, err
    return out


@pytest.mark.parametrize("slots", [1, 2])
def test_sharded_null_loop_peer_kernel_equals_nccl_and_unsharded(pkg, po, oracle, monkeypatch, slots):
    if _ndev(pkg) < 2:
        pytest.skip("needs two GPUs")
    from test_gpu_nulls import oracle_null_loop
    N, L, R, w = 260, 136, 5, 0.05
    msa, wgt, _ = po.synthetic_msa(N, L, seed=71)
    nulls = np.stack([po.synthetic_msa(N, L, seed=300 + r)[0] for r in range(R)])
    w_ref, view, mm_o = oracle_null_loop(po, oracle, nulls, wgt, po.GT, po.C16, po.APC)[:3]
    assert w_ref == w
    oscores, oscale = oracle_null_loop.scores, max(1.0, float(np.max(np.abs(mm_o))))
    whole = pkg.Context(0)
    whole.configure(N, L, slots, 0)
    whole.set_weights(wgt)
    whole.hist_reset()
    mm_ref = whole.null_hist(nulls, w)
    bins_ref, n_ref, _ = whole.hist_read(1 << 15)
    real_ref = whole.scan(msa)
    whole.close()

    monkeypatch.setenv("RSCAPE_B200_PEER_REDUCE", "1")
    peer = _sharded_loop(pkg, 2, N, L, slots, wgt, nulls, w, msa)
    monkeypatch.setenv("RSCAPE_B200_PEER_REDUCE", "0")
    nccl = _sharded_loop(pkg, 2, N, L, slots, wgt, nulls, w, msa)
    assert all(o[4]["peer_path"] and o[4]["reductions"] >= 3 * R for o in peer)
    assert not any(o[4]["peer_path"] for o in nccl)

    iu = np.triu_indices(L, 1)
    scale = max(1.0, float(np.abs(mm_ref).max()))
    for run in (peer, nccl):
        bins = run[0][1] + run[1][1]
        assert run[0][2] + run[1][2] == n_ref == R * L * (L - 1) // 2
        # ranges are all-reduced: every rank reports the global (min, max) of each null
        assert np.array_equal(run[0][0], run[1][0])
        assert np.max(np.abs(run[0][0] - mm_ref)) <= 1e-11 * scale
        assert_bins_identical(bins, bins_ref.astype(np.int64), oscores, -10.0, w, scale=oscale)
        for k in range(2):
            assert np.max(np.abs(run[k][3]["cov"][iu] - real_ref["cov"][iu])) <= 1e-11 * scale
    # the two collective paths sum in the same rank order: identical results
    assert np.array_equal(peer[0][1] + peer[1][1], nccl[0][1] + nccl[1][1])
    assert np.array_equal(peer[0][0], nccl[0][0])
    # ... and the oracle's loop on the same nulls
    bins = peer[0][1] + peer[1][1]
    assert not bins[view.nb:].any()
    assert_bins_identical(bins[:view.nb], view.obs, oscores, -10.0, w, scale=oscale)
    assert np.max(np.abs(peer[0][0] - mm_o)) <= 1e-9 * oscale


def test_pool_broadcast_makes_every_rank_hold_all_nulls(pkg, po):
    """rsb_pool_broadcast: each of two ranks generates its own block of nulls (generator A, keyed by the global replicate id), the
    blocks are exchanged over NVLink, and both pools equal the pool of one rank that generated everything."""
    if _ndev(pkg) < 2:
        pytest.skip("needs two GPUs")
    N, L, R = 120, 64, 6
    msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=9)
    whole = pkg.Context(0)
    whole.configure(N, L, 2, 0)
    whole.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    whole.pool_reserve(R)
    whole.null_fitch_shuffle(msa, 77, R)
    ref = whole.pool_get(R, 0)
    whole.close()
    ctxs = []
    for k in range(2):
        c = pkg.Context(k)
        c.configure(N, L, 2, 0)
        c.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
        c.pool_reserve(R)
        ctxs.append(c)
    pkg.comm_init_all(ctxs)
    blocks = [(0, 4), (4, 2)]
    got, err = [None, None], [None, None]

    def work(k):
        try:
            c = ctxs[k]
            c.null_fitch_shuffle(msa, 77, blocks[k][1], first_rep=blocks[k][0], first_id=blocks[k][0])
            for root, (f, n) in enumerate(blocks):
                c.pool_broadcast(f, n, root)
            got[k] = c.pool_get(R, 0)
        except Exception as e:                                      # noqa: BLE001
            err[k] = e

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "broadcast hung"
    for c in ctxs:
        c.comm_destroy()
        c.close()
    assert err == [None, None], err
    assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref)
