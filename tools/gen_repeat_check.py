"""Diagnostic: is generator A reproducible call to call (same seed -> byte-identical nulls, Fitch rows and shuffled rows)?"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
po = ge.load_oracle()
which = sys.argv[1] if len(sys.argv) > 1 else "config1"
if which == "config1":
    import _config1 as c1
    sub, wgt, keep, mask, tree, gold = c1.load(po)
    R, seeds = 20, (1001, 1002, 1003, 1004)
else:
    sub, wgt, _, tree = pkg.synth.synthetic_family(10000, 1800, seed=42)
    R, seeds = 16, (7,)
N, L = sub.shape
ctx = pkg.Context(0)
ctx.configure(N, L, 4, 4)
ctx.set_weights(wgt)
ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
ctx.pool_reserve(R)
for seed in seeds:
    sums = []
    for rep in range(5):
        ctx.null_fitch_shuffle(sub, seed, R)
        a = ctx.pool_get(R, 0)
        anc = ctx.pool_get_internal(0, R, 0)
        sh = ctx.pool_get_internal(1, R, 0)
        sums.append((zlib.crc32(a.tobytes()), zlib.crc32(anc.tobytes()), zlib.crc32(sh.tobytes())))
        if rep == 0:
            first = (a.copy(), anc.copy(), sh.copy())
        elif sums[-1] != sums[0]:
            d = [np.argwhere(x != y) for x, y in zip((a, anc, sh), first)]
            print("   differing cells (pool, anc, shanc):", [len(x) for x in d], "first:", [x[0].tolist() if len(x) else None for x in d])
    print(which, "seed", seed, "reproducible" if len(set(sums)) == 1 else "NOT REPRODUCIBLE", sums[:2], flush=True)
ctx.close()
