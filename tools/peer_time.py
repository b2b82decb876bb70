"""Latency of the library's small all-reduce over the GPUs of one box, one process driving all devices (rsb_comm_init_all, one host
thread per device): the one-shot kernel over NVLink peer memory against ncclAllReduce, for the vector sizes of the scan
(score range, APC row sums [L+4], marginal sums [L][4] at the SSU and LSU shapes).  Usage: python tools/peer_time.py [ngpus]"""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
W = int(sys.argv[1]) if len(sys.argv) > 1 else int(pkg.lib().rsb_device_count())
COUNTS = [2, 1804, 3504, 7200, 14000]
for peer in ("1", "0"):
    os.environ["RSCAPE_B200_PEER_REDUCE"] = peer
    ctxs = []
    for k in range(W):
        c = pkg.Context(k)
        c.configure(2000, 3500, 2, 1)
        ctxs.append(c)
    pkg.comm_init_all(ctxs)
    res = [[None] * len(COUNTS) for _ in range(W)]

    def work(k):
        for j, n in enumerate(COUNTS):
            res[k][j] = ctxs[k].comm_selftest(n, 500)

    th = [threading.Thread(target=work, args=(k,)) for k in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    info = ctxs[0].comm_info()
    for j, n in enumerate(COUNTS):
        print(f"{W} GPUs, {'peer kernel' if info['peer_path'] else 'ncclAllReduce'}: {n:6d} doubles  {max(r[j][0] for r in res):7.2f} us per all-reduce"
              f"  (max error of the checked sum {max(r[j][1] for r in res):g})", flush=True)
    for c in ctxs:
        c.comm_destroy()
        c.close()
