"""Times the two stages after the scan that run on the device: the E-value / significant-pair stage (rsb_scan_hits) at the SSU
alignment length and Tree_Substitutions (rsb_tree_substitutions) at the RNase P and SSU shapes.  Whole C-ABI calls (host
buffers in and out, each call returns synchronised), wall clock; one JSON line per stage.  Kernel-only times come from the ncu
launch list taken over this script (profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
quick = "--quick" in sys.argv
reps = 2 if quick else 5


def timed(fn):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()                                    # returns after its results have landed in host memory (stream synchronised)
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.median(ts))


# ---- E-values + hit list, L = 1800 (the stage does not depend on the number of sequences) -----------------------
N, L = 256, 1800
rng = np.random.default_rng(1)
msa = rng.integers(0, 4, (N, L), dtype=np.uint8)
msa[:, 10] = msa[:, 900]
nulls = rng.integers(0, 4, (2, N, L), dtype=np.uint8)
ctx = pkg.Context(0)
ctx.configure(N, L, 2, 0)
ctx.set_weights(None)
w, _, _ = ctx.null_width(nulls[0])
mm = ctx.null_hist(nulls, w)
res = ctx.scan(msa, pkg.GT, pkg.C16, pkg.APC)
nb = int(np.ceil((max(mm[:, 1].max(), res["maxcov"]) + 10.0) / w)) + 6
bins, n, _ = ctx.hist_read(nb)
P = L * (L - 1) // 2
mask = np.zeros((L, L), np.uint8)
ev_pinned = np.empty((L, L))
pinned = pkg.pin(mask) and pkg.pin(ev_pinned)                         # what corr_Create does for mi->Eval (rsb_host_register)
for variant, kw in (("E-values + mi->Eval + hit list", dict(want_eval=True)), ("hit list only", dict(want_eval=False)),
                    ("E-values + mi->Eval + hit list, page-locked mi->Eval and mask", dict(want_eval=True, eval_out=ev_pinned))):
    best, med = timed(lambda: ctx.scan_hits(-10.0, w, bins, float(mm[:, 1].max()), P, 0, mask, thresh=0.6, **kw))
    print(json.dumps(dict(stage="rsb_scan_hits", variant=variant, L=L, pairs=P, ms_best=best, ms_median=med,
                          algorithmic_bytes=P * (8 + 1 + (16 if kw["want_eval"] else 0)), pinned=bool(pinned))), flush=True)
pkg.unpin(mask); pkg.unpin(ev_pinned)
ctx.close()

# ---- Tree_Substitutions ---------------------------------------------------------------------------------------
for ntaxa, L in ((5000, 400),) if quick else ((5000, 400), (10000, 1800)):
    tree = pkg.synth.random_tree(ntaxa, np.random.default_rng(2))
    leaves = rng.integers(0, 5, (ntaxa, L), dtype=np.uint8)
    internal = rng.integers(0, 4, (ntaxa - 1, L), dtype=np.uint8)
    ctx = pkg.Context(0)
    ctx.configure(2 * (ntaxa - 1), L, 1, 1)
    tables = (np.empty((L, L), np.int32), np.empty((L, L), np.int32))
    pins = [leaves, internal, tables[0], tables[1]]
    pinned = all(pkg.pin(a) for a in pins)
    for variant, kw in (("nsubs + ndouble + njoin", dict(want_pairs=True, out=tables)), ("nsubs only", dict(want_pairs=False))):
        best, med = timed(lambda: ctx.tree_substitutions(tree.left, tree.right, leaves, internal, False, **kw))
        cells = L * L * 2 * (ntaxa - 1) / 2.0
        print(json.dumps(dict(stage="rsb_tree_substitutions", variant=variant, ntaxa=ntaxa, L=L, branch_rows=2 * (ntaxa - 1),
                              pair_cells=cells, ms_best=best, ms_median=med, pinned=bool(pinned))), flush=True)
    for a in pins:
        pkg.unpin(a)
    ctx.close()
