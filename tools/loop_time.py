"""Fixed and per-replicate cost of the null loop: rsb_null_hist_pool over 1..100 resident nulls of the bench family.
RSCAPE_B200_TRACE=1 adds the library's stage timers."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
N, L, R = 10000, 1800, 100
msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=42)
for snull in (0, 2):
    ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    ctx.set_null_slices(snull)
    ctx.configure(N, L, 2, 4)
    ctx.set_weights(wgt)
    ctx.pool_reserve(R)
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.null_fitch_shuffle(msa, 1, R)
    torch.cuda.synchronize()
    ctx.null_hist_pool(0, 4, 0.05)
    for n in (1, 2, 4, 8, 13, 26, 50, 100):
        ts = []
        for _ in range(3):
            ctx.hist_reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.null_hist_pool(0, n, 0.05)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"null slices {snull}: {n:3d} nulls {min(ts):8.3f} ms  ({min(ts) / n:.3f} per null)", flush=True)
    ctx.profile_gram(True)
    ctx.counters(reset=True)
    ctx.null_hist_pool(0, 100, 0.05)
    c = ctx.counters()
    print("   gram", c["gram_ms"] / c["gram_launches"], flush=True)
    ctx.close()
