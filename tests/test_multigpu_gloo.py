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


def test_shard_covers_every_replicate_once(pkg):
    for R in (1, 7, 100):
        for world in (1, 2, 4, 8):
            got = sorted(sum((pkg.parallel.null_shard(R, world, k) for k in range(world)), []))
            assert got == list(range(R))
            sizes = [len(pkg.parallel.null_shard(R, world, k)) for k in range(world)]
            assert max(sizes) - min(sizes) <= 1
