"""parallel.py -- sharding of the null loop across the GPUs of one box (one process per GPU, torch.distributed).

The path shards along the null replicates (src/R-scape.c:1650-1697: every iteration only adds into the cumulative
histogram).  There is no data-path collective: each rank scans its own replicates; the only exchange is the sum of
the small integer histograms (and min/max of the score range) at the end -- one all-reduce over NCCL on the GPU box,
gloo in the CPU tests.  The histogram width w is defined by replicate 0 (src/R-scape.c:1681-1684); every rank repeats
that width pass on the same replicate 0, so no broadcast is needed to agree on w.
"""
import numpy as np


def null_shard(nnull, world, rank, last_rank_extra=0):
    """Replicate ids scanned by `rank`: one contiguous block per rank (so that a generator call covers it), sizes
    floor or ceil of nnull/world.

    last_rank_extra: work the LAST rank carries besides its nulls, in units of one null -- the rank that also scans the
    input alignment (run_rscape(GIVSS), src/R-scape.c:2548) gets that many nulls fewer, the others share them.  The
    blocks stay contiguous and cover 0..nnull-1 exactly once."""
    if last_rank_extra <= 0 or world == 1:
        base, extra = divmod(nnull, world)
        first = rank * base + min(rank, extra)
        return list(range(first, first + base + (1 if rank < extra else 0)))
    share = (nnull + last_rank_extra) / world                       # units of work per rank
    n_last = int(min(nnull, max(0, round(share - last_rank_extra))))
    if rank == world - 1:
        return list(range(nnull - n_last, nnull))
    return [r for r in null_shard(nnull - n_last, world - 1, rank)]


def reduce_histogram(bins, device=None, group=None):
    """Sum uint64 histogram bins over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bins
    t = torch.from_numpy(bins.astype(np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(np.uint64)


def reduce_histogram_on_device(ctx, nb, group=None):
    """Sum the first nb bins of the context's DEVICE histogram over all ranks (device copy out, NCCL all-reduce, device
    copy back), so that a following hist_read returns the cumulative histogram of the whole job on every rank.
    No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    t = torch.empty(int(nb), dtype=torch.int64, device=f"cuda:{ctx.device}")
    ctx.hist_exchange(t, to_library=False)                      # synchronous: the bins are in t when it returns
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream(t.device).synchronize()           # the library copies on its own stream
    ctx.hist_exchange(t, to_library=True)


def reduce_range(lo, hi, device=None, group=None):
    """Global (min, max) of the per-rank score ranges."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return lo, hi
    t = torch.tensor([-lo, hi], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return -float(t[0]), float(t[1])


def reduce_sum(vec, device=None, group=None):
    """Sum a float64 vector over all ranks (marginal sums of a sharded pair grid)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def reduce_cov_sums(cov_sums, L, device=None, group=None):
    """cov_sums of rsb_sharded_statistic: [0..L] are sums, [L+1] a minimum, [L+2] a maximum."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cov_sums
    out = np.array(cov_sums, dtype=np.float64)
    out[:L + 1] = reduce_sum(out[:L + 1], device, group)
    lo, hi = reduce_range(out[L + 1], out[L + 2], device, group)
    out[L + 1], out[L + 2] = lo, hi
    return out


HIT_FIELDS = ("i", "j", "sc", "eval", "pval")


def merge_hit_lists(lists):
    """Concatenate the ranks' significant-pair lists (dicts of rsb_scan_hits: i, j, sc, eval, pval) and restore the reference's
    row-major order of cov_CreateHitList (src/covariation.c:828-829).  On a sharded pair grid every pair is listed by the one
    rank that owns its row, so the merged list is the list of the whole scan."""
    out = {k: np.concatenate([np.asarray(h[k]) for h in lists]) if lists else np.empty(0) for k in HIT_FIELDS}
    order = np.lexsort((out["j"], out["i"]))
    out = {k: v[order] for k, v in out.items()}
    out["nhit"] = int(sum(int(h.get("nhit", len(h["i"]))) for h in lists))
    return out


def gather_hit_lists(hits, device=None, group=None):
    """All-gather the significant-pair lists of a sharded scan: one all-gather of the list lengths, one of the lists padded to the
    longest (5 x 8 bytes per hit; NCCL on the GPU box, gloo in the CPU tests).  Returns the merged list on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_hit_lists([hits])
    world = dist.get_world_size(group)
    n = len(hits["i"])
    counts = torch.tensor([n, int(hits.get("nhit", n))], dtype=torch.int64)
    if device is not None:
        counts = counts.to(device)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    lens = [int(c[0]) for c in all_counts]
    width = max(max(lens), 1)
    # i and j travel as float64 bit patterns next to the three double fields (exact for any int64)
    buf = np.zeros((len(HIT_FIELDS), width), dtype=np.float64)
    buf[0, :n] = np.asarray(hits["i"], dtype=np.int64).view(np.float64)
    buf[1, :n] = np.asarray(hits["j"], dtype=np.int64).view(np.float64)
    for f, k in enumerate(HIT_FIELDS[2:], start=2):
        buf[f, :n] = hits[k]
    t = torch.from_numpy(buf).view(torch.int64)                     # gathered as integers: no NaN canonicalisation on the way
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    lists = []
    for p, m, c in zip(parts, lens, all_counts):
        a = p.cpu().numpy()
        lists.append(dict(i=a[0, :m].copy(), j=a[1, :m].copy(), sc=a[2, :m].view(np.float64).copy(), eval=a[3, :m].view(np.float64).copy(),
                          pval=a[4, :m].view(np.float64).copy(), nhit=int(c[1])))
    return merge_hit_lists(lists)
