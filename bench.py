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
    # the reference is single-threaded (src/Makefile has no OpenMP): its own rate is one such task alone on one core
    t0 = time.perf_counter()
    n1 = 0
    while time.perf_counter() - t0 < max(2.0, args.cpu_seconds / 4) and n1 < 20:
        task(int(rng.integers(0, L - ncols)))
        n1 += 1
    t1 = time.perf_counter() - t0
    return dict(value=cells / t_total, unit="pair-cells/s", cores=nthreads, kind="reference" if use_ref else "port",
                one_thread=dict(value=n1 * (ncols * ncols * N / 2.0) / t1, cores=1, sample=f"{n1} such scans on one thread ({t1:.1f} s)"),
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
    ap.add_argument("--null-slices", type=int, default=0,
                    help="mixed precision: digit slices of the weights the NULL alignments are contracted with (0 = --slices, one set of "
                         "weights everywhere); the input alignment always uses --slices")
    ap.add_argument("--stat", default="GT", choices=["GT", "MI", "MIr", "MIg", "CHI", "OMES", "RAFS", "all"],
                    help="covariation statistic of the scans (BASELINE config 5 sweeps them; the headline metric is GT)")
    ap.add_argument("--actype", default="APC", choices=["APC", "ASC"], help="background correction")
    ap.add_argument("--slots", type=int, default=0, help="replicate slots (alignments in flight); 0 = choose from the shape")
    ap.add_argument("--nulls", type=int, default=0, help="diagnostics: override the workload's number of null replicates")
    ap.add_argument("--ref-cols", type=int, default=160)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the second measurement in the other precision mode (mixed / strict)")
    ap.add_argument("--grid-shard", action="store_true",
                    help="shard the L x L pair grid of every scan over the GPUs (BASELINE config 4) instead of the null replicates")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE line, the JSON line: libraries that print there (NCCL's version banner under NCCL_DEBUG, torchrun's
    # children) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

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
    L, N, R = wl["L"], wl["N"], (args.nulls if args.nulls > 0 else wl["nulls"])
    peaks = load_peaks()

    # ---- synthetic inputs (same on every rank: seeded) ------------------------------------------------
    msa, wgt, _, tree = synth.synthetic_family(N, L, seed=42)        # alignment evolved on the tree the null generator is given
    stream = torch.cuda.current_stream()
    SEED = 20261017
    W0, BMIN, HPTS, TOL = 0.05, -10.0, 400, 1e-6                      # cfg->w, BMIN, HPTS, tol (src/R-scape.c:426, covariation.h:22)
    host_msa = torch.from_numpy(msa).pin_memory()
    dev_msa = torch.from_numpy(msa).cuda()
    cov_out = torch.empty((L, L), dtype=torch.float64).pin_memory().numpy()          # pinned: the score matrix of the input alignment lands here
    bins_pinned = torch.empty(1 << 22, dtype=torch.int64).pin_memory().numpy().view(np.uint64)   # the histogram lands here
    STAT, ACT = (getattr(pkg, args.stat) if args.stat != "all" else None), getattr(pkg, args.actype)
    PHASES = bool(os.environ.get("BENCH_PHASES"))                    # host-side phase times of every job on stderr
    if args.grid_shard and args.stat in ("RAFS",):
        raise SystemExit("--grid-shard does not offer RAFS")
    scans_total = R + 2
    cells_per_scan = L * L * N / 2.0
    cells_total = scans_total * cells_per_scan
    pairs = L * (L - 1) / 2.0
    peak_i8 = 2.0 * peaks["bf16"]

    def width_of(lo, hi):
        """calculate_width_histo, src/R-scape.c:1355-1360, from the score range of replicate 0"""
        w = min(W0, (hi - max(BMIN, lo)) / HPTS)
        return 0.0 if w < TOL else w

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

    def run_sweep():
        """BASELINE config 5: the statistic sweep {GT, MI, MIr, MIg, CHI, OMES} x {APC, ASC} + RAFS x {APC, ASC} = 14 combinations over
        the same nulls.  The reference repeats the whole scan per combination (cov_Calculate dispatches to one corr_Calculate*,
        src/covariation.c:100-258): 14 x (nulls + 2) scans.  Here every null is contracted ONCE for the twelve weighted combinations
        (rsb_null_hist_multi) and once more, unweighted and with a single digit slice, for the two RAFS ones."""
        if world > 1:
            raise SystemExit("--stat all runs on one GPU")
        weighted = [(st, ac) for st in ("GT", "MI", "MIr", "MIg", "CHI", "OMES") for ac in ("APC", "ASC")]
        pc = [(getattr(pkg, st), getattr(pkg, ac)) for st, ac in weighted]
        rafs = [(pkg.RAFS, pkg.APC), (pkg.RAFS, pkg.ASC)]
        slots = args.slots if args.slots > 0 else pkg.replicate_slots(N, L, R, args.slices)
        ctx = pkg.Context(local, stream.cuda_stream)
        ctx.set_null_slices(args.null_slices if args.null_slices != args.slices else 0)
        ctx.configure(N, L, slots, args.slices)
        ctx.set_weights(wgt)
        ctx.pool_reserve(R)
        ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
        ctx.null_fitch_shuffle(host_msa.numpy(), SEED, R)
        state = {}

        def job():
            tp = [time.perf_counter()]
            ctx.hist_reset_multi()
            mm0 = ctx.null_hist_multi(1, pc, [0.0] * len(pc), pkg.C16, first_rep=0)          # calculate_width_histo for every combination
            w = [width_of(mm0[k, 0, 0], mm0[k, 0, 1]) for k in range(len(pc))]
            mm = ctx.null_hist_multi(R, pc, w, pkg.C16, first_rep=0)
            hists = []
            for k in range(len(pc)):
                nb = int(min(1 << 22, max(64, np.ceil((float(mm[k, :, 1].max()) - BMIN) / w[k]) + 8))) if w[k] > 0 else 64
                hists.append(ctx.hist_read_multi(k, nb)[0])
            tp.append(time.perf_counter())
            # RAFS x {APC, ASC}: unit weights, one single-slice contraction per null shared by the two corrections
            mr0 = ctx.null_hist_multi(1, rafs, [0.0] * len(rafs), pkg.C16, first_rep=0)
            wr = [width_of(mr0[k, 0, 0], mr0[k, 0, 1]) for k in range(len(rafs))]
            ctx.hist_reset_multi()                                                           # (the weighted histograms have been read)
            mr = ctx.null_hist_multi(R, rafs, wr, pkg.C16, first_rep=0)
            for k in range(len(rafs)):
                nb = int(min(1 << 22, max(64, np.ceil((float(mr[k, :, 1].max()) - BMIN) / wr[k]) + 8))) if wr[k] > 0 else 64
                hists.append(ctx.hist_read_multi(k, nb)[0])
            tp.append(time.perf_counter())
            for st, ac in pc + rafs:                                                         # run_rscape(GIVSS) per combination
                ctx.scan(dev_msa, st, pkg.C16, ac, want_cov=False)
            state["hists"] = hists
            if PHASES:
                tp.append(time.perf_counter())
                print("[bench] sweep phases ms: 12 weighted combinations %.1f, RAFS x 2 %.1f, %d scans of the input alignment %.1f" %
                      ((tp[1] - tp[0]) * 1e3, (tp[2] - tp[1]) * 1e3, len(pc) + len(rafs), (tp[3] - tp[2]) * 1e3), file=sys.stderr, flush=True)

        sampler = ClockSampler(local)
        ctx.counters(reset=True)
        ctx.profile_gram(True)
        sampler.start()
        ms = timed(job, args.steps, args.warmup)
        clocks = sampler.stop()
        cnt = ctx.counters(reset=True)
        ncombo = len(pc) + len(rafs)
        expected = R * (L * (L - 1) // 2)
        ok = all(int(h.sum()) == expected for h in state["hists"])
        cells = ncombo * scans_total * cells_per_scan
        g0 = cnt["geometry"][2] if cnt["geometry"][2]["gram_launches"] else cnt["geometry"][0]
        line = dict(metric=METRIC.replace("GTp+APC", "statistic sweep"), value=cells * args.steps / (ms * 1e-3), unit="pair-cells/s", n_gpus=1, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="u8 x u8 -> s32 tensor-core counts (fixed-point weights), f64 statistics", data="synthetic",
                    config=dict(workload=f"sweep (BASELINE config 5): L={L} N={N} nulls={R}, {ncombo} (statistic, correction) combinations = {ncombo} x {scans_total} "
                                         f"reference scans per step; executed: {R + 1} weighted contractions shared by 12 combinations + "
                                         f"{R + 1} single-slice unweighted ones shared by RAFS+APC / RAFS+ASC (rsb_null_hist_multi) + {ncombo} scans of the input alignment",
                                combinations=[f"{a}+{b}" for a, b in weighted] + ["RAFS+APC", "RAFS+ASC"], weight_slices=args.slices,
                                histogram=dict(mass_ok=bool(ok))),
                    gpu_launches=int(cnt["launches"] * args.steps / (args.steps + args.warmup)), clocks=clocks,
                    contraction=dict(weighted_ms=g0["gram_ms"] / max(1, g0["gram_launches"]), weighted_launches=g0["gram_launches"], slices=g0["slices"],
                                     unit_weight_ms=cnt["geometry"][1]["gram_ms"] / max(1, cnt["geometry"][1]["gram_launches"]),
                                     unit_weight_launches=cnt["geometry"][1]["gram_launches"],
                                     share_of_step=cnt["gram_ms"] / (ms * (args.steps + args.warmup) / args.steps)))
        emit(line)
        ctx.close()
        return 0

    def measure(null_slices, steps, warmup, sample_clocks):
        """The whole benchmark for one precision mode: null_slices = 0 scores the nulls with the input alignment's weights
        (--slices digit slices everywhere), > 0 with that many slices (rsb_set_null_slices)."""
        slots = args.slots if args.slots > 0 else pkg.replicate_slots(N, L, R, args.slices)
        ctx = pkg.Context(local, stream.cuda_stream)
        ctx.set_null_slices(null_slices if null_slices != args.slices else 0)
        ctx.configure(N, L, slots, args.slices)
        ctx.set_weights(wgt)
        q_abs, q_bits = ctx.quantisation_error()
        _, _, s_null, qn_abs, qn_bits = ctx.null_quantisation()
        if world > 1:
            # the library's own NCCL communicator (histogram sum; with --grid-shard also the per-scan vectors, inside the pipeline):
            # rank 0 makes the id, torch.distributed only carries its 128 bytes
            box = [pkg.comm_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0], world, rank)
        if args.grid_shard:
            ctx.set_shard(rank, world)                                    # row blocks of the pair grid of EVERY scan dealt to the ranks
            ctx.set_weights(wgt)
        # replicate ids held by this rank; the rank that also scans the input alignment (uploaded and read back in `e2e`) takes two nulls
        # fewer -- the largest block, which sets `value`, stays what an even split gives (13 of 100 on 8 GPUs, 51 on 2)
        my_ids = list(range(R)) if args.grid_shard else pkg.parallel.null_shard(R, world, rank, last_rank_extra=2)
        n_mine = len(my_ids)
        own0 = (n_mine > 0 and my_ids[0] == 0)                            # replicate 0 defines the histogram width (R-scape.c:1681-1684)
        real_rank = world - 1                                             # the input alignment is scanned by the rank with the fewest nulls
        ctx.pool_reserve(max(n_mine, 1))

        def generate():
            """R-scape's default null model on the device: Fitch + tree-substitution shuffle (null_rscape, R-scape.c:1653-1661).
            Replicates are keyed by their global id: a rank generates exactly its own block (every rank all of them with --grid-shard)."""
            ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
            if args.grid_shard and world > 1 and not os.environ.get("BENCH_GEN_ALL"):
                # every rank scans every null: generate 1/world of them each and exchange the blocks over NVLink (rsb_pool_broadcast)
                blocks = [pkg.parallel.null_shard(R, world, r) for r in range(world)]
                if blocks[rank]:
                    ctx.null_fitch_shuffle(host_msa.numpy(), SEED, len(blocks[rank]), first_rep=blocks[rank][0], first_id=blocks[rank][0])
                for r, b in enumerate(blocks):
                    if b:
                        ctx.pool_broadcast(b[0], len(b), r)
            elif n_mine:
                ctx.null_fitch_shuffle(host_msa.numpy(), SEED, n_mine, first_rep=0, first_id=my_ids[0])

        def job(real):
            """null_rscape + run_rscape for this rank's share of the work, nulls already in the device pool.

            The width pass is FUSED with the scan of replicate 0 (quirk Q2: the reference scans the first null twice, the scan consumes
            no randomness): every rank histograms its nulls at once with the default width w = 0.05, the owner of replicate 0 derives
            the width calculate_width_histo would return from that replicate's score range, and only if it differs from 0.05 (score
            range of the first null below 20) is the loop repeated with it -- the histogram is then exactly the reference's."""
            tp = [time.perf_counter()] if PHASES else None
            ctx.hist_reset()
            out = None
            if not args.grid_shard and rank == real_rank:
                # run_rscape(GIVSS): independent of the nulls, and done BEFORE the ranks meet in the width agreement below, so that the
                # others do not wait for it
                out = ctx.scan(real, STAT, pkg.C16, ACT, want_cov=isinstance(real, np.ndarray), cov_out=cov_out)
            if PHASES: tp.append(time.perf_counter())
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
            if args.grid_shard:
                out = ctx.sharded_scan(real, STAT, pkg.C16, ACT, want_cov=isinstance(real, np.ndarray), cov_out=cov_out)
            if PHASES: tp.append(time.perf_counter())
            # sum over ranks and read only the bins the null scores reach: bin of the largest score + cov_GrowRankList's 5 w margin
            nb = int(min(1 << 22, max(64, np.ceil((hi - BMIN) / w) + 8))) if (w > 0 and np.isfinite(hi)) else 64
            if world > 1:
                ctx.hist_allreduce(nb)                                                       # null_add2cumranklist across ranks, on the device
            bins, n, imax = ctx.hist_read(nb, out=bins_pinned)
            if PHASES:
                tp.append(time.perf_counter())
                print("[bench] rank %d phases ms: input %.2f nulls %.2f sharded input %.2f hist %.2f" % ((rank,) + tuple((b - a) * 1e3 for a, b in zip(tp, tp[1:]))),
                      file=sys.stderr, flush=True)
            return w, bins, out

        t_gen0 = time.perf_counter()
        generate()
        torch.cuda.synchronize()
        t_gen = time.perf_counter() - t_gen0

        # ---- value: inputs resident in HBM -------------------------------------------------------------------
        sampler = ClockSampler(local) if (sample_clocks and rank == 0 and not os.environ.get("BENCH_NO_SAMPLER")) else None
        ctx.counters(reset=True)
        ctx.profile_gram(True)
        if sampler:
            sampler.start()
        ms_dev = timed(lambda: job(dev_msa), steps, warmup)
        clocks = sampler.stop() if sampler else None
        cnt = ctx.counters(reset=True)
        ctx.profile_gram(False)
        value = cells_total * steps / (ms_dev * 1e-3)
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
            t0 = time.perf_counter()
            ctx.set_weights(wgt)
            t1 = time.perf_counter()
            generate()
            t2 = time.perf_counter()
            job(host_msa.numpy())
            if PHASES:
                print("[bench] rank %d e2e phases ms: set_weights %.2f generate (enqueue) %.2f job %.2f" %
                      (rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), file=sys.stderr, flush=True)

        ms_e2e = timed(job_e2e, steps, 1)
        e2e_value = cells_total * steps / (ms_e2e * 1e-3)
        # whole-job bytes per step: every rank uploads the alignment (generator), its tree and weights, and reads its histogram;
        # the rank scanning the input alignment uploads it once more and reads the score matrix
        h2d = (world + 1) * N * L + world * ((N - 1) * (3 * 4 + 2 * 8) + N * 8)
        d2h = world * len(bins_chk) * 8 + L * L * 8

        # ---- roofline of the dominant kernel (tcgen05 gram): algorithmic ops / measured launch time -------------
        # The dominant launches are the contractions of the NULL alignments: their own operand geometry when the nulls carry their
        # own weights (mixed precision), else the one they share with the input alignment's scan.
        gi = 2 if cnt["geometry"][2]["gram_launches"] > 0 else (1 if args.stat == "RAFS" else 0)
        geo = cnt["geometry"][gi]
        gram_ms_avg = geo["gram_ms"] / max(1, geo["gram_launches"])
        scans_of_geo = n_mine * (steps + warmup)
        if gi != 2 and (rank == real_rank or args.grid_shard):
            scans_of_geo += (steps + warmup)                              # the input alignment's scan shares the geometry
        if args.grid_shard:
            scans_of_geo /= world                                         # every rank contracts 1/world of each scan's tiles
        ops_alg_per_launch = 32.0 * pairs * N * scans_of_geo / max(1, geo["gram_launches"])
        achieved = ops_alg_per_launch / (gram_ms_avg * 1e-3) / 1e12 if gram_ms_avg > 0 else 0.0
        S_dom = max(1, geo["slices"])
        roofline = dict(bound="tensor", achieved=achieved, peak=peak_i8, unit="TOP/s", frac=achieved / peak_i8,
                        traffic=None,
                        note=f"algorithmic int8 ops (32 per pair-cell) of one gram launch / mean launch time {gram_ms_avg:.3f} ms (CUDA events around every "
                             f"launch, in situ); the kernel issues {S_dom}x that in tcgen05 kind::i8 MMAs (one pass per 8-bit digit slice of the weights): "
                             f"implementation rate {achieved * S_dom:.1f} TOP/s = {achieved * S_dom / peak_i8:.3f} of peak; "
                             f"peak = 2 x {peaks['src']} bf16 {peaks['bf16']} TFLOP/s (int8 runs at twice the bf16 rate); traffic: not measured in this "
                             f"run (dram__bytes of the ncu --set full capture: profiles/)",
                        gram_ms=gram_ms_avg, gram_slices=S_dom,
                        gram_share_of_step=cnt["gram_ms"] / (ms_dev * (steps + warmup) / steps) if ms_dev > 0 else None)
        comm = ctx.comm_info() if world > 1 else None
        ctx.close()
        return dict(comm=comm, value=value, ms_dev=ms_dev, e2e_value=e2e_value, ms_e2e=ms_e2e, h2d=h2d, d2h=d2h, roofline=roofline, clocks=clocks, cnt=cnt,
                    hist_ok=hist_ok, nbins=int(len(bins_chk)), slots=slots, t_gen=t_gen, q_abs=q_abs, q_bits=q_bits, s_null=s_null,
                    qn_abs=qn_abs, qn_bits=qn_bits, steps=steps, warmup=warmup)

    if args.stat == "all":
        return run_sweep()

    m = measure(args.null_slices, args.steps, args.warmup, True)
    # the other precision mode beside it (same inputs, same run): nulls at 2 digit slices when the headline is strict, and vice versa
    alt = None
    if not args.no_alt and args.stat != "RAFS" and args.slices > 2:
        alt = measure(2 if (args.null_slices in (0, args.slices)) else 0, args.steps, args.warmup, False)
    value, ms_dev, e2e_value, ms_e2e, roofline, clocks, cnt = m["value"], m["ms_dev"], m["e2e_value"], m["ms_e2e"], m["roofline"], m["clocks"], m["cnt"]
    slots, q_abs, q_bits, hist_ok, t_gen, h2d, d2h = m["slots"], m["q_abs"], m["q_bits"], m["hist_ok"], m["t_gen"], m["h2d"], m["d2h"]
    bins_len = m["nbins"]

    def mode_text(r):
        if r["s_null"] == args.slices:
            return (f"strict: {args.slices} digit slices of the weights for every alignment (largest |wq 2^-q - w| = {r['q_abs']:.3g}, "
                    f"{r['q_bits']:.1f} bits below the largest weight)")
        return (f"mixed: nulls contracted with {r['s_null']} digit slices (largest |wq 2^-q - w| = {r['qn_abs']:.3g}, {r['qn_bits']:.1f} bits), the input "
                f"alignment with {args.slices} ({r['q_bits']:.1f} bits); stated bound on a null's scores |d score| <= 2e-4 max(1,|score|), identical "
                f"significant pairs (tests/test_gpu_mixed.py)")

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
                                weight_slices=args.slices, null_weight_slices=m["s_null"], precision=mode_text(m),
                                weights=f"fixed point wq = u V, u < 256, V < 256^{args.slices}: largest |wq 2^-q - w| = {q_abs:.3g} "
                                        f"({q_bits:.1f} bits below the largest weight); counts are exact integer arithmetic on wq",
                                replicate_slots=slots, histogram=dict(bins_read_per_step=bins_len, mass_ok=bool(hist_ok)), parallelism=(f"L x L pair grid of every scan sharded over {world} GPU(s) by 32-column row blocks" if args.grid_shard
                                             else f"nulls in contiguous blocks over {world} GPU(s)"),
                                l2="inputs larger than L2 (null alignments %.1f GB, operand planes %.1f GB per replicate)" %
                                   (R * N * L / 1e9, (4 + 4 * args.slices) * L * N / 1e9),
                                null_model="Fitch + tree-substitution shuffle generated on the device: resident before the timed region "
                                           f"for `value`, generated inside it for `e2e` (first generation incl. allocations {t_gen * 1e3:.0f} ms)"),
                    e2e=dict(value=e2e_value, unit="pair-cells/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(cnt["launches"] * args.steps / (args.steps + args.warmup)),
                    clocks=clocks, roofline=roofline, cpu_baseline=cpu)
        if m["comm"] is not None:
            # how the small vectors crossed the GPUs: the library's one-shot kernel over NVLink peer memory (csrc/peer_reduce.cu) or NCCL
            line["config"]["collectives"] = dict(small_vectors="one-shot all-reduce kernel over NVLink peer memory" if m["comm"]["peer_path"] else "ncclAllReduce",
                                                 peer_reductions=m["comm"]["reductions"], histogram="ncclAllReduce (uint64 bins, once per job)")
        if alt is not None:
            # the same job in the other precision mode, measured in the same run (not the headline)
            line["other_precision_mode"] = dict(precision=mode_text(alt), null_weight_slices=alt["s_null"], value=alt["value"],
                                                ms_per_step=alt["ms_dev"] / alt["steps"],
                                                e2e=dict(value=alt["e2e_value"], ms_per_step=alt["ms_e2e"] / alt["steps"]),
                                                roofline={k: alt["roofline"][k] for k in ("achieved", "peak", "unit", "frac", "gram_ms", "gram_slices", "gram_share_of_step")},
                                                histogram_mass_ok=bool(alt["hist_ok"]))
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
