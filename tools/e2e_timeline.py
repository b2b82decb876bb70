"""Host-side timeline of one end-to-end step (generate + width pass + null scans) with and without a synchronisation
between generation and scanning, to see how much of the generation hides under the scans."""
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
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.set_null_slices(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ctx.configure(N, L, 2, 4)
ctx.set_weights(wgt)
ctx.pool_reserve(R)
host = torch.from_numpy(msa).pin_memory().numpy()


def step(sync_between):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
    ctx.null_fitch_shuffle(host, 1234, R)
    t1 = time.perf_counter()
    if sync_between:
        ctx.pool_get(1, R - 1)          # waits for the last chunk
    t2 = time.perf_counter()
    ctx.hist_reset()
    w, _, _ = ctx.null_width_pool(0)
    t3 = time.perf_counter()
    ctx.null_hist_pool(0, R, w, want_minmax=False)
    t4 = time.perf_counter()
    ctx.hist_read(1 << 18)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    return [round((b - a) * 1e3, 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t0, t5))]


for sync_between in (1, 0, 1, 0, 1, 0):
    print("sync" if sync_between else "overlap", "enqueue-gen, wait, width, nulls, read, TOTAL =", step(sync_between), flush=True)
ctx.close()
