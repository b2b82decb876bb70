"""The N > 1 path on CPU: world_size 2 over gloo.  Null replicates are dealt round-robin to the ranks, every rank
repeats the width pass on replicate 0, scans its share and the integer histograms are summed with one all-reduce --
the same orchestration bench.py runs over NCCL, with the oracle standing in for the device scan."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, R, N, L, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    po = ge.load_oracle()
    ora = po.Oracle()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nulls = [po.synthetic_msa(N, L, seed=50 + r)[0] for r in range(R)]
    wgt = po.synthetic_msa(N, L, seed=1)[1]
    first = ora.scan(nulls[0], wgt, po.GT, po.C16, po.APC)                 # every rank: width pass on replicate 0
    w = ora.null_width(0.05, first["mincov"], first["maxcov"])
    nb = 4096
    bins = np.zeros(nb, np.uint64)
    lo, hi = np.inf, -np.inf
    for r in pkg.parallel.null_shard(R, world, rank):
        res = ora.scan(nulls[r], wgt, po.GT, po.C16, po.APC)
        h = ora.hist_from_cov(res["cov"], res["maxcov"], -10.0, w)
        v = ora.view(h)
        ora.free(h)
        bins[:v.nb] += v.obs
        lo, hi = min(lo, res["mincov"]), max(hi, res["maxcov"])
    total = pkg.parallel.reduce_histogram(bins)
    lo, hi = pkg.parallel.reduce_range(lo, hi)
    if rank == 0:
        np.savez(out, bins=total, w=w, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_null_sharding_two_ranks_gloo(tmp_path, po, oracle, pkg):
    import torch.multiprocessing as mp
    R, N, L = 5, 60, 30
    out = str(tmp_path / "dist.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, R, N, L, out), nprocs=2, join=True)
    z = np.load(out)
    # single-process answer
    nulls = [po.synthetic_msa(N, L, seed=50 + r)[0] for r in range(R)]
    wgt = po.synthetic_msa(N, L, seed=1)[1]
    first = oracle.scan(nulls[0], wgt, po.GT, po.C16, po.APC)
    w = oracle.null_width(0.05, first["mincov"], first["maxcov"])
    cum, lo, hi = None, np.inf, -np.inf
    for m in nulls:
        res = oracle.scan(m, wgt, po.GT, po.C16, po.APC)
        h = oracle.hist_from_cov(res["cov"], res["maxcov"], -10.0, w)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
        lo, hi = min(lo, res["mincov"]), max(hi, res["maxcov"])
    v = oracle.view(cum)
    oracle.free(cum)
    assert z["w"] == w and z["lo"] == lo and z["hi"] == hi
    assert np.array_equal(z["bins"][:v.nb], v.obs) and not z["bins"][v.nb:].any()
    assert int(z["bins"].sum()) == R * L * (L - 1) // 2


def _hits_worker(rank, world, port, L, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    po = ge.load_oracle()
    ora = po.Oracle()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cov, null, mask = _hits_case(po, ora, L)
    # this rank's rows of the pair grid (32-column row blocks dealt cyclically, as rsb_set_shard): hits of the other rows are dropped
    full = ora.hitlist(cov, null, mask, int(mask.sum()), L * (L - 1) // 2 - int(mask.sum()), -1, 5.0)
    mine = (full["i"] // 32) % world == rank
    part = {k: full[k][mine] for k in ("i", "j", "sc", "eval", "pval")}
    part["nhit"] = int(mine.sum())
    merged = pkg.parallel.gather_hit_lists(part)
    np.savez(out % rank, **{k: merged[k] for k in ("i", "j", "sc", "eval", "pval")}, nhit=merged["nhit"])
    dist.barrier()
    dist.destroy_process_group()


def _hits_case(po, ora, L):
    msa, wgt, partner = po.synthetic_msa(150, L, seed=33)
    cov = ora.scan(msa, wgt, po.GT, po.C16, po.APC)["cov"]
    x = np.maximum(np.random.default_rng(5).normal(0, 5, 40000), -10 + 0.05)
    b = np.ceil((x + 10) / 0.05 - 1).astype(np.int64)
    null = po.NullFit(-10.0, 0.05, np.bincount(b, minlength=int(b.max()) + 6).astype(np.uint64), xmax=float(x.max())).exp_tail(0.05)
    mask = np.zeros((L, L), np.uint8)
    for i, j in enumerate(partner):
        if j > i:
            mask[i, j] = 1
    return cov, null, mask


def test_hit_lists_gathered_over_two_ranks_gloo(tmp_path, po, oracle, pkg):
    """north_star item 4: the significant-pair lists of a sharded scan are the only per-pair data that crosses ranks."""
    import torch.multiprocessing as mp
    L = 100
    out = str(tmp_path / "hits%d.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_hits_worker, args=(2, port, L, out), nprocs=2, join=True)
    cov, null, mask = _hits_case(po, oracle, L)
    want = oracle.hitlist(cov, null, mask, int(mask.sum()), L * (L - 1) // 2 - int(mask.sum()), -1, 5.0)
    assert len(want["i"]) > 3
    for rank in range(2):
        z = np.load(out % rank)
        assert int(z["nhit"]) == len(want["i"])
        for k in ("i", "j", "sc", "eval", "pval"):
            assert np.array_equal(z[k], want[k]), (rank, k)


def test_merge_hit_lists_orders_row_major(pkg):
    a = dict(i=np.array([5, 0]), j=np.array([9, 3]), sc=np.array([1.0, 2.0]), eval=np.array([.1, .2]), pval=np.array([.01, .02]), nhit=2)
    b = dict(i=np.array([0, 5]), j=np.array([2, 7]), sc=np.array([3.0, 4.0]), eval=np.array([.3, .4]), pval=np.array([.03, .04]), nhit=2)
    m = pkg.parallel.merge_hit_lists([a, b])
    assert list(zip(m["i"], m["j"])) == [(0, 2), (0, 3), (5, 7), (5, 9)] and list(m["sc"]) == [3.0, 2.0, 4.0, 1.0] and m["nhit"] == 4
    e = pkg.parallel.merge_hit_lists([])
    assert len(e["i"]) == 0 and e["nhit"] == 0
    assert pkg.parallel.gather_hit_lists(a)["nhit"] == 2          # no process group: the rank's own list, ordered


def test_shard_covers_every_replicate_once(pkg):
    for R in (1, 7, 100):
        for world in (1, 2, 4, 8):
            got = sorted(sum((pkg.parallel.null_shard(R, world, k) for k in range(world)), []))
            assert got == list(range(R))
            sizes = [len(pkg.parallel.null_shard(R, world, k)) for k in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_null_shard_with_extra_work_on_the_last_rank(pkg):
    """The rank that also scans the input alignment takes fewer nulls (bench.py: last_rank_extra=2): blocks stay contiguous, cover
    every replicate once, and the largest block -- which sets the job's time with resident nulls -- is the even split's."""
    for R in (100, 20, 13, 1):
        for world in (1, 2, 4, 8):
            blocks = [pkg.parallel.null_shard(R, world, k, last_rank_extra=2) for k in range(world)]
            assert sum(blocks, []) == list(range(R))
            even = [len(pkg.parallel.null_shard(R, world, k)) for k in range(world)]
            assert len(blocks[-1]) <= even[-1]
            if world > 1:
                scans = max([len(b) for b in blocks[:-1]] + [len(blocks[-1]) + 1])
                assert scans <= max(even[:-1] + [even[-1] + 1])
