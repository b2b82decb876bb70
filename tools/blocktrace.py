"""Experiment: per-block (kernel, SM, start, end) timeline of the pipelined null loop, to see which kernels really overlap.
Needs the library built with -DRSB_BLOCKTRACE:
    make -C r-scape_b200 -B NVFLAGS='-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DRSB_BLOCKTRACE'
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
N, L, nrep = 10000, 1800, 8
rng = np.random.default_rng(1)
msa = rng.integers(0, 5, (nrep, N, L)).astype(np.uint8)
ctx = pkg.Context(0)
ctx.configure(N, L, 2, 5)
ctx.set_weights(rng.gamma(2.0, 0.5, N))
ctx.pool_reserve(nrep)
ctx.pool_put(msa)
ctx.hist_reset()
width = ctx.null_width_pool(0)[0]
ctx.null_hist_pool(0, nrep, width)
ctx.hist_read(4000)
lib = pkg.lib()
lib.rsb_blocktrace(1, None)
ctx.null_hist_pool(0, nrep, width)
ctx.hist_read(4000)
lib.rsb_blocktrace(0, b"gpurun_out/blocktrace.bin")
raw = np.fromfile("gpurun_out/blocktrace.bin", dtype=np.uint64)
n = int(raw[0]); rec = raw[1:1 + 3 * n].reshape(n, 3)
kid = (rec[:, 0] >> np.uint64(32)).astype(int); sm = (rec[:, 0] & np.uint64(0xffffffff)).astype(int)
t0 = rec[:, 1].astype(np.int64); t1 = rec[:, 2].astype(np.int64)
base = t0.min()
names = {1: "gram", 2: "marg_partial", 3: "stat"}
# split each kernel's records into launches by clustering start times
for k in (1, 2, 3):
    m = kid == k
    if not m.any():
        continue
    order = np.argsort(t0[m]); a0 = t0[m][order] - base; a1 = t1[m][order] - base
    per = {1: 148, 2: None, 3: None}[k]
    cnt = m.sum()
    nl = nrep
    sz = cnt // nl
    print(f"{names[k]}: {cnt} blocks, {sz} per launch")
    # launches identified by block count order is unreliable for overlapped kernels; print percentile timeline instead
    for q in range(nl):
        s = slice(q * sz, (q + 1) * sz)
        d = (a1[s] - a0[s]) / 1e3
        print(f"   launch {q}: first start {a0[s].min()/1e3:9.1f} us  last end {a1[s].max()/1e3:9.1f} us   block dur mean {d.mean():7.1f} max {d.max():7.1f} us   SMs used {len(set(sm[m][order][s]))}")

# co-residency: how many blocks of a kernel are in flight on one SM at the same time (time-averaged over its busy span)
for k in (2, 3):
    m = kid == k
    if not m.any():
        continue
    conc = []
    for s_ in range(0, 148, 37):
        mm = m & (sm == s_)
        ev = sorted([(t, 1) for t in t0[mm]] + [(t, -1) for t in t1[mm]])
        cur = 0; last = None; area = 0; busy = 0; peak = 0
        for t, d in ev:
            if last is not None and cur > 0:
                area += cur * (t - last); busy += (t - last)
            cur += d; peak = max(peak, cur); last = t
        conc.append((s_, round(area / max(busy, 1), 2), peak))
    print(f"{names[k]}: (SM, mean blocks in flight while busy, peak) {conc}")
    live = (t1[m] - t0[m]) > 5000
    print(f"   live blocks: {live.sum()} of {m.sum()}, mean duration {(t1[m] - t0[m])[live].mean() / 1e3:.1f} us")
