"""Times the tcgen05 contraction in situ (CUDA events around every launch, library counters) inside the pipelined null
loop, for a list of slice counts.  Usage (GPU box):  python tools/gram_time.py [workload] [S,S,...] [nrep]
Environment: RSCAPE_B200_LIB=<variant .so> (tools/build_variant.sh), RSCAPE_B200_FUSED_GT=0 (count epilogue + stat_kernel)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "ssu"
slices = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4").split(",")]
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 12
w = bench.WORKLOADS[name]
N, L = w["N"], w["L"]
rng = np.random.default_rng(1)
msa = torch.from_numpy(rng.integers(0, 5, (nrep, N, L)).astype(np.uint8)).cuda()
wgt = rng.gamma(2.0, 0.5, N)
for S in slices:
    ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    ctx.configure(N, L, 2, S)
    ctx.set_weights(wgt)
    ctx.hist_reset()
    def run():
        try:
            ctx.null_hist(msa, 0.05)
        except pkg.RscapeB200Error as e:        # experimental variants leave garbage scores: only the timing matters
            print("   (", str(e)[:60], ")")
    run()                                       # warm-up
    ctx.counters(reset=True)
    ctx.profile_gram(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    c = ctx.counters(reset=True)
    print(f"{os.path.basename(pkg.LIB_PATH)} fused={os.environ.get('RSCAPE_B200_FUSED_GT', '1')} {name} S={S}: gram {c['gram_ms'] / max(1, c['gram_launches']):.3f} ms/launch "
          f"({c['gram_launches']} launches), loop {wall / nrep:.3f} ms/replicate", flush=True)
    # the input alignment's path (count epilogue + stat_kernel + correction, results staying on the device)
    ctx.scan(msa[0], want_cov=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.scan(msa[0], want_cov=False)
    torch.cuda.synchronize()
    print(f"   rsb_scan of a device-resident alignment, no host outputs: {(time.perf_counter() - t0) * 200:.3f} ms", flush=True)
    if os.environ.get("RSCAPE_B200_TRACE"):
        ctx.counters()
    ctx.close()
