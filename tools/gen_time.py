"""Times the device null generator (Fitch + tree-substitution shuffle) alone: ms per batch of replicates."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
N, L, R = 10000, 1800, int(sys.argv[1]) if len(sys.argv) > 1 else 100
if os.environ.get("GEN_UNRELATED_TREE"):
    msa, wgt, _ = pkg.synth.synthetic_msa(N, L, seed=42)
    tree = pkg.synth.random_tree(N, np.random.default_rng(42))
else:
    msa, wgt, _, tree = pkg.synth.synthetic_family(N, L, seed=42)
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.configure(N, L, 2, 4)
ctx.set_weights(wgt)
ctx.pool_reserve(R)
ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
host = torch.from_numpy(msa).pin_memory().numpy()
for it in range(int(os.environ.get("GEN_ITERS", "4"))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.null_fitch_shuffle(host, 1234, R)
    torch.cuda.synchronize()
    print(f"generate {R} replicates: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
if os.environ.get("GEN_SIMULATE"):
    # generator B (cov_GenerateAlignment, ungapped, structure-free): rate matrix + root sequence
    Q = np.array([[-1.00, 0.30, 0.50, 0.20], [0.25, -0.90, 0.15, 0.50], [0.60, 0.10, -0.95, 0.25], [0.20, 0.45, 0.30, -0.95]])
    root = np.where(msa[0] < 4, msa[0], 0).astype(np.uint8)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.null_simulate(Q, root, 99, R, gapmask=msa)
        torch.cuda.synchronize()
        print(f"simulate {R} replicates: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
chk = ctx.pool_get(1, 0)
print("checksum", int(chk.astype(np.int64).sum()), "subst vs input", int((chk[0] != msa).sum()))
ctx.close()
