#!/usr/bin/env python
"""bench.py -- R-scape covariation scan throughput on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2], the SSU-rRNA-shaped alignment L=1800, N=10000 with 100
null replicates, statistic GTp (G-test + APC).  One *step* is the whole job of null_rscape + run_rscape
(src/R-scape.c:1616-1724, 2548-2724) for that alignment: the width pass on the first null, 100 null scans into
the cumulative histogram and the scan of the input alignment = 102 scans = 102 * L(L-1)/2 * N pair-cells...
reported with the metric's own definition of a pair-cell count per scan, L^2*N/2.

  value   whole-job pair-cells/s with every input already resident in HBM (CUDA events, max over ranks)
  e2e     the same job through the C-ABI with HOST buffers: the input alignment (pinned), tree and weights in, null
          alignments generated on the device (Fitch + shuffle), cumulative histogram and the input alignment's score
          matrix out, all copies inside the timed region
  N > 1   the 100 nulls are split into one contiguous block per rank (one process per GPU, torchrun); the width pass is fused
          with the scan of replicate 0 (every rank histograms with the default w = 0.05, the owner of replicate 0 checks it:
          32 bytes over NCCL), the last rank also scans the input alignment; the per-rank histograms are summed with one
          all-reduce of the bins the scores reach, on the library's own NCCL communicator.  --grid-shard: the L x L pair grid
          of EVERY scan is dealt to the ranks instead (BASELINE config 4) and the per-scan vectors (marginal sums, APC row
          sums, score range) are all-reduced on the device inside the pipeline.  Total work is fixed: "scaling": "strong".

--impl reference times the reference's own CPU implementation of the path (oracle/_ref: src/correlators.c compiled
unchanged; the oracle port if that build is absent) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

WORKLOADS = {
    "trna": dict(L=76, N=1000, nulls=20),
    "rnasep": dict(L=400, N=5000, nulls=20),
    "ssu": dict(L=1800, N=10000, nulls=100),
    "lsu": dict(L=3500, N=20000, nulls=20),
    "sweep": dict(L=1800, N=10000, nulls=500),      # BASELINE config 5: run once per --stat / --actype
}
METRIC = "pair-cells/s (L^2*N/2) GTp+APC incl. nulls"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm), sm_mhz_min=min(sm) if sm else None, power_w_median=float(np.median(pw)) if pw else None,
                    power_w_max=max(pw) if pw else None)


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    L, N = wl["L"], wl["N"]
    po = ge.load_oracle()
    msa, wgt, _, _ = po.synthetic_family(N, L, seed=42)
    cores = os.cpu_count() or 1
    nthreads = max(1, min(cores, 64))
    ncols = args.ref_cols                       # columns per thread task (bounded sample of the L x L pair grid)
    use_ref = po.RefLib.available()
    kind = "reference" if use_ref else "port"
    lib = po.RefLib() if use_ref else None
    ora = None if use_ref else po.Oracle()
    rng = np.random.default_rng(1)

    def task(c0):
        sub = np.ascontiguousarray(msa[:, c0:c0 + ncols])
        if use_ref:                               # the reference's own corr_Probs + corr_CalculateGT + COVCorrected
            lib.scan(sub, wgt, po.GT, po.C16, po.APC)
        else:
            ora.scan(sub, wgt, po.GT, po.C16, po.APC)

    def one_step():
        starts = [int(rng.integers(0, L - ncols)) for _ in range(nthreads)]
        th = [threading.Thread(target=task, args=(c0,)) for c0 in starts]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    cells_per_step = nthreads * (ncols * ncols * N / 2.0)
    total = sum(times)
    value = cells_per_step * args.steps / total
    sample = f"{nthreads} threads x one scan (corr_Probs+GT+APC) of a {ncols}-column slice of the L={L} N={N} alignment per step"
    line = dict(impl="reference", metric=METRIC, value=value, unit="pair-cells/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"{args.workload}: L={L} N={N} nulls={wl['nulls']} GTp+APC (bounded sample)", sample=sample),
                cpu_baseline=dict(value=value, unit="pair-cells/s", cores=nthreads, kind=kind, sample=sample),
                e2e=dict(value=value, unit="pair-cells/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def cpu_baseline_sample(args, msa, wgt):
    """The reference's CPU path on a bounded sample (about 10-30 s of CPU work), rank 0 at N=1 only."""
    po = ge.load_oracle()
    N, L = msa.shape
    cores = os.cpu_count() or 1
    nthreads = max(1, min(cores, 64))
    ncols = args.ref_cols
    use_ref = po.RefLib.available()
    lib = po.RefLib() if use_ref else None
    ora = None if use_ref else po.Oracle()

    def task(c0):
        sub = np.ascontiguousarray(msa[:, c0:c0 + ncols])
        (lib or ora).scan(sub, wgt, po.GT, po.C16, po.APC)

    reps, t_total = 0, 0.0
    rng = np.random.default_rng(2)
    while t_total < args.cpu_seconds and reps < 50:
        th = [threading.Thread(target=task, args=(int(rng.integers(0, L - ncols)),)) for _ in range(nthreads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        t_total += time.perf_counter() - t0
        reps += 1
    cells = reps * nthreads * (ncols * ncols * N / 2.0)
    return dict(value=cells / t_total, unit="pair-cells/s", cores=nthreads, kind="reference" if use_ref else "port",
                sample=f"{reps} x {nthreads} threads, each one scan (corr_Probs+GT+APC) of a {ncols}-column slice of the L={L} N={N} alignment "
                       f"({t_total:.1f} s)")


# ---------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ssu", choices=sorted(WORKLOADS))
    ap.add_argument("--slices", type=int, default=4,
                    help="8-bit digit slices S of the fixed-point weights wq = u V (8-bit multiplier u, V < 256^S): ~8(S+1)-bit weights")
    ap.add_argument("--stat", default="GT", choices=["GT", "MI", "MIr", "MIg", "CHI", "OMES", "RAFS"],
                    help="covariation statistic of the scans (BASELINE config 5 sweeps them; the headline metric is GT)")
    ap.add_argument("--actype", default="APC", choices=["APC", "ASC"], help="background correction")
    ap.add_argument("--slots", type=int, default=0, help="replicate slots (alignments in flight); 0 = choose from the shape")
    ap.add_argument("--ref-cols", type=int, default=160)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid-shard", action="store_true",
                    help="shard the L x L pair grid of every scan over the GPUs (BASELINE config 4) instead of the null replicates")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = ge.load_package()
    synth = pkg.synth
    if world == 1:
        args.grid_shard = False                                       # one GPU owns the whole pair grid

    wl = WORKLOADS[args.workload]
    L, N, R = wl["L"], wl["N"], wl["nulls"]
    peaks = load_peaks()

    # ---- synthetic inputs (same on every rank: seeded) ------------------------------------------------
    msa, wgt, _, tree = synth.synthetic_family(N, L, seed=42)        # alignment evolved on the tree the null generator is given
    stream = torch.cuda.current_stream()
    slots = args.slots if args.slots > 0 else pkg.replicate_slots(N, L, R, args.slices)
    ctx = pkg.Context(local, stream.cuda_stream)
    ctx.configure(N, L, slots, args.slices)
    ctx.set_weights(wgt)
    q_abs, q_bits = ctx.quantisation_error()
    if world > 1:
        # the library's own NCCL communicator (histogram sum; with --grid-shard also the per-scan vectors, inside the pipeline):
        # rank 0 makes the id, torch.distributed only carries its 128 bytes
        box = [pkg.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], world, rank)
    if args.grid_shard:
        ctx.set_shard(rank, world)                                    # row blocks of the pair grid of EVERY scan dealt to the ranks
        ctx.set_weights(wgt)
    my_ids = list(range(R)) if args.grid_shard else pkg.parallel.null_shard(R, world, rank)   # replicate ids held by this rank
    n_mine = len(my_ids)
    own0 = (n_mine > 0 and my_ids[0] == 0)                            # replicate 0 defines the histogram width (R-scape.c:1681-1684)
    real_rank = world - 1                                             # the input alignment is scanned by the rank with the fewest nulls
    ctx.pool_reserve(max(n_mine, 1))
    SEED = 20261017
    W0, BMIN, HPTS, TOL = 0.05, -10.0, 400, 1e-6                      # cfg->w, BMIN, HPTS, tol (src/R-scape.c:426, covariation.h:22)
    host_msa = torch.from_numpy(msa).pin_memory()
    dev_msa = torch.from_numpy(msa).cuda()
    cov_out = torch.empty((L, L), dtype=torch.float64).pin_memory().numpy()          # pinned: the score matrix of the input alignment lands here
    bins_pinned = torch.empty(1 << 22, dtype=torch.int64).pin_memory().numpy().view(np.uint64)   # the histogram lands here

    def generate():
        """R-scape's default null model on the device: Fitch + tree-substitution shuffle (null_rscape, R-scape.c:1653-1661).
        Replicates are keyed by their global id: a rank generates exactly its own block (every rank all of them with --grid-shard)."""
        ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
        if n_mine:
            ctx.null_fitch_shuffle(host_msa.numpy(), SEED, n_mine, first_rep=0, first_id=my_ids[0])

    STAT, ACT = getattr(pkg, args.stat), getattr(pkg, args.actype)
    PHASES = bool(os.environ.get("BENCH_PHASES"))                    # host-side phase times of every job on stderr
    if args.grid_shard and args.stat in ("RAFS",):
        raise SystemExit("--grid-shard does not offer RAFS")

    def width_of(lo, hi):
        """calculate_width_histo, src/R-scape.c:1355-1360, from the score range of replicate 0"""
        w = min(W0, (hi - max(BMIN, lo)) / HPTS)
        return 0.0 if w < TOL else w

    def job(real):
        """null_rscape + run_rscape for this rank's share of the work, nulls already in the device pool.

        The width pass is FUSED with the scan of replicate 0 (quirk Q2: the reference scans the first null twice, the scan consumes
        no randomness): every rank histograms its nulls at once with the default width w = 0.05, the owner of replicate 0 derives
        the width calculate_width_histo would return from that replicate's score range, and only if it differs from 0.05 (score
        range of the first null below 20) is the loop repeated with it -- the histogram is then exactly the reference's."""
        tp = [time.perf_counter()] if PHASES else None
        ctx.hist_reset()
        w = W0
        for attempt in range(2):
            lo, hi, w_true = np.inf, -np.inf, np.inf
            if n_mine:
                mm = ctx.null_hist_pool(0, n_mine, w, STAT, pkg.C16, ACT)                # run_rscape(RANSS) + null_add2cumranklist
                lo, hi = float(mm[:, 0].min()), float(mm[:, 1].max())
                if own0:
                    if not mm[0, 1] > BMIN:
                        raise SystemExit("bmin should be larger than maxCOV (R-scape.c:1355)")
                    w_true = width_of(mm[0, 0], mm[0, 1])
            if world > 1 and not args.grid_shard:
                lo, hi, w_true = ctx.comm_range(lo, hi, w_true)                          # 32 bytes over NCCL
            if w_true == w or attempt == 1:
                break
            w = w_true                                                                   # rare: redo with the width of replicate 0
            ctx.hist_reset()
        if PHASES: tp.append(time.perf_counter())
        out = None
        if args.grid_shard:
            out = ctx.sharded_scan(real, STAT, pkg.C16, ACT, want_cov=isinstance(real, np.ndarray), cov_out=cov_out)
        elif rank == real_rank:
            out = ctx.scan(real, STAT, pkg.C16, ACT, want_cov=isinstance(real, np.ndarray), cov_out=cov_out)   # run_rscape(GIVSS)
        if PHASES: tp.append(time.perf_counter())
        # sum over ranks and read only the bins the null scores reach: bin of the largest score + cov_GrowRankList's 5 w margin
        nb = int(min(1 << 22, max(64, np.ceil((hi - BMIN) / w) + 8))) if (w > 0 and np.isfinite(hi)) else 64
        if world > 1:
            ctx.hist_allreduce(nb)                                                       # null_add2cumranklist across ranks, on the device
        bins, n, imax = ctx.hist_read(nb, out=bins_pinned)
        if PHASES:
            tp.append(time.perf_counter())
            print("[bench] rank %d phases ms: nulls %.2f input %.2f hist %.2f" % ((rank,) + tuple((b - a) * 1e3 for a, b in zip(tp, tp[1:]))),
                  file=sys.stderr, flush=True)
        return w, bins, out

    t_gen0 = time.perf_counter()
    generate()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen0

    scans_total = R + 2
    cells_per_scan = L * L * N / 2.0
    cells_total = scans_total * cells_per_scan

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: inputs resident in HBM -------------------------------------------------------------------
    sampler = ClockSampler(local)
    ctx.counters(reset=True)
    ctx.profile_gram(True)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda: job(dev_msa), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    cnt = ctx.counters(reset=True)
    ctx.profile_gram(False)
    value = cells_total * args.steps / (ms_dev * 1e-3)
    # sanity of what was timed: the cumulative histogram holds every pair of every null exactly once.  The job reads (and sums
    # over ranks) a window of bins; scores beyond it stay in the tail of each rank's device histogram and are counted here.
    _, bins_chk, _ = job(dev_msa)
    expected = R * (L * (L - 1) // 2)
    mass = int(bins_chk.sum())
    hist_ok = (mass == expected)
    if not hist_ok:
        print(f"[bench] rank {rank}: cumulative null histogram holds {mass} scores, expected {expected}", file=sys.stderr, flush=True)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ------------------------------
    # in : the input alignment (pinned host memory, uploaded twice: generators + scan), the tree, the weights
    # out: cumulative null histogram and the input alignment's corrected score matrix
    def job_e2e():
        ctx.set_weights(wgt)
        generate()
        w, bins, out = job(host_msa.numpy())

    ms_e2e = timed(job_e2e, args.steps, 1)
    e2e_value = cells_total * args.steps / (ms_e2e * 1e-3)
    # whole-job bytes per step: every rank uploads the alignment (generator), its tree and weights, and reads its histogram;
    # the rank scanning the input alignment uploads it once more and reads the score matrix
    h2d = (world + 1) * N * L + world * ((N - 1) * (3 * 4 + 2 * 8) + N * 8)
    d2h = world * len(bins_chk) * 8 + L * L * 8

    # ---- roofline of the dominant kernel (tcgen05 gram): algorithmic ops / measured launch time -------------
    # launches during the value run: per step, gram launches = width(1) + ceil(nulls/slots) + real(1 on rank 0)
    pairs = L * (L - 1) / 2.0
    gram_ms_avg = cnt["gram_ms"] / max(1, cnt["gram_launches"])
    scans_this_rank = (n_mine + (1 if (rank == real_rank or args.grid_shard) else 0)) * (args.steps + args.warmup)
    if args.grid_shard:
        scans_this_rank /= world                                     # every rank contracts 1/world of each scan's tiles
    ops_alg_per_launch = 32.0 * pairs * N * scans_this_rank / max(1, cnt["gram_launches"])
    achieved = ops_alg_per_launch / (gram_ms_avg * 1e-3) / 1e12 if gram_ms_avg > 0 else 0.0
    peak_i8 = 2.0 * peaks["bf16"]
    # DRAM bytes of one gram launch from the committed `ncu --set full` capture (profiles/r1_ncu_full_gram_*): read + write
    traffic = {("ssu", 4): 1.079e9 + 0.229e9}.get((args.workload, args.slices)) if world == 1 and not args.grid_shard else None
    roofline = dict(bound="tensor", achieved=achieved, peak=peak_i8, unit="TOP/s", frac=achieved / peak_i8,
                    traffic=traffic,
                    note=f"algorithmic int8 ops (32 per pair-cell) of one gram launch / mean launch time {gram_ms_avg:.3f} ms; "
                         f"the kernel issues {args.slices}x that in tcgen05 kind::i8 MMAs (one pass per 8-bit digit slice of the weights): "
                         f"implementation rate {achieved * args.slices:.1f} TOP/s = {achieved * args.slices / peak_i8:.3f} of peak; "
                         f"peak = 2 x {peaks['src']} bf16 {peaks['bf16']} TFLOP/s (int8 runs at twice the bf16 rate)",
                    gram_share_of_step=cnt["gram_ms"] / (ms_dev * (args.steps + args.warmup) / args.steps) if ms_dev > 0 else None)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(args, msa, wgt)
        line = dict(metric=METRIC, value=value, unit="pair-cells/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="u8 x u8 -> s32 tensor-core counts (fixed-point weights), f64 statistics", data="synthetic",
                    config=dict(workload=f"{args.workload}: L={L} N={N} nulls={R} {args.stat}+{args.actype}, scans per step = {scans_total} "
                                         f"(width pass + {R} nulls + input alignment, the reference's count; the width pass is fused with the "
                                         f"scan of replicate 0, so {R + 1} contractions are executed)",
                                weight_slices=args.slices,
                                weights=f"fixed point wq = u V, u < 256, V < 256^{args.slices}: largest |wq 2^-q - w| = {q_abs:.3g} "
                                        f"({q_bits:.1f} bits below the largest weight); counts are exact integer arithmetic on wq",
                                replicate_slots=slots, histogram=dict(bins_read_per_step=int(len(bins_chk)), mass_ok=bool(hist_ok)), parallelism=(f"L x L pair grid of every scan sharded over {world} GPU(s) by 32-column row blocks" if args.grid_shard
                                             else f"nulls in contiguous blocks over {world} GPU(s)"),
                                l2="inputs larger than L2 (null alignments %.1f GB, operand planes %.1f GB per replicate)" %
                                   (R * N * L / 1e9, (4 + 4 * args.slices) * L * N / 1e9),
                                null_model="Fitch + tree-substitution shuffle generated on the device: resident before the timed region "
                                           f"for `value`, generated inside it for `e2e` (first generation incl. allocations {t_gen * 1e3:.0f} ms)"),
                    e2e=dict(value=e2e_value, unit="pair-cells/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(cnt["launches"] * args.steps / (args.steps + args.warmup)),
                    clocks=clocks, roofline=roofline, cpu_baseline=cpu)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
