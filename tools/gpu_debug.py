"""Manual GPU bring-up script (not a pytest file): staged checks of the tcgen05 count kernel with verbose
mismatch reports.  Usage on the GPU box:  timeout 300 python tools/gpu_debug.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
po = ge.load_oracle()
ora = po.Oracle()


def say(*a):
    print(*a, flush=True)


def numpy_counts(msa, wq):
    N, L = msa.shape
    out = np.zeros((16, L, L), dtype=np.int64)
    for a in range(4):
        A = (msa == a).astype(np.int64) * wq[:, None]
        for b in range(4):
            B = (msa == b).astype(np.int64)
            out[a * 4 + b] = np.triu(A.T @ B, 1)
    return out


def stage(N, L, S, unit=False, seed=0):
    rng = np.random.default_rng(seed)
    msa = rng.integers(0, 5, (N, L)).astype(np.uint8)
    wgt = np.ones(N) if unit else rng.gamma(2.0, 0.5, N)
    ctx = pkg.Context(0)
    ctx.configure(N, L, 1, S)
    ctx.set_weights(wgt)
    wq, q, Sx = ctx.quantisation()
    t0 = time.time()
    ctx.scan(msa, want_cov=False)
    got = ctx.counts()
    dt = time.time() - t0
    ref = numpy_counts(msa, wq)
    direct = ctx.counts_direct(msa)
    ok = np.array_equal(got, ref)
    okd = np.array_equal(direct, ref)
    say(f"N={N} L={L} S={Sx} q={q} unit={unit}: tcgen05 {'OK' if ok else 'MISMATCH'}  direct {'OK' if okd else 'MISMATCH'}  ({dt*1e3:.1f} ms)")
    if not ok:
        bad = np.argwhere(got != ref)
        say(f"   {len(bad)} / {ref.size} differ; planes hit {sorted(set(bad[:,0].tolist()))[:16]}")
        say(f"   rows i hit (first 20) {sorted(set(bad[:,1].tolist()))[:20]}  cols j hit (first 20) {sorted(set(bad[:,2].tolist()))[:20]}")
        for p, i, j in bad[:8]:
            say(f"   plane {p} i {i} j {j}: got {got[p,i,j]} want {ref[p,i,j]} ratio {got[p,i,j]/max(1,ref[p,i,j]):.4f}")
        nz = np.count_nonzero(got)
        say(f"   nonzero got {nz} want {np.count_nonzero(ref)}; sum got {got.sum()} want {ref.sum()}")
    ctx.close()
    return ok


if __name__ == "__main__":
    import torch
    say(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    allok = True
    for args in [(128, 32, 1, True), (128, 64, 1, True), (256, 64, 1, True), (1000, 76, 1, True), (300, 40, 2, False),
                 (300, 40, 4, False), (700, 100, 5, False), (700, 100, 6, False), (2000, 300, 5, False)]:
        allok &= stage(*args)
    say("ALL OK" if allok else "SOME FAILED")
