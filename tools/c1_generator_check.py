"""Diagnostic: the tutorial alignment's null score distribution from the device generator A vs the reference's own generator
(oracle/_ref), both scored by the CPU oracle, over many seeds: tail counts, fitted tail, E-values of the two borderline pairs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import _config1 as c1  # noqa: E402

po = ge.load_oracle()
oracle = po.Oracle()
reflib = po.RefLib()
pkg = ge.load_package()
sub, wgt, keep, mask, tree, gold = c1.load(po)
N, L = sub.shape
P = L * (L - 1) // 2
real = oracle.scan(sub, wgt, po.GT, po.C16, po.APC)
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iu = np.triu_indices(L, 1)


def stats(nulls, tag, seed):
    cum, xmax, sc = None, -np.inf, []
    for msa in nulls:
        r = oracle.scan(msa, wgt, po.GT, po.C16, po.APC)
        h = oracle.hist_from_cov(r["cov"], r["maxcov"], -10.0, 0.05, 1e-6)
        cum = oracle.accumulate(cum, h)
        oracle.free(h)
        xmax = max(xmax, r["maxcov"])
        sc.append(r["cov"][iu])
    sc = np.concatenate(sc)
    view = oracle.view(cum)
    oracle.free(cum)
    nb = c1.null_bins_needed(0.05, xmax, real["maxcov"])
    obs = np.zeros(nb, np.uint64)
    obs[:view.nb] = view.obs[:nb]
    fit = reflib.nullfit(po.NullFit(-10.0, 0.05, obs, xmax=xmax), c1.PMASS, c1.FRACFIT, False)
    Nb = int(mask.sum())
    ev = oracle.hitlist(real["cov"], fit, mask, Nb, P - Nb, -1, 2000.0)
    E = {(int(keep[i]) + 1, int(keep[j]) + 1): e for i, j, e in zip(ev["i"], ev["j"], ev["eval"])}
    nsub = float(np.mean([(m != sub).sum() for m in nulls]))
    print(f"{tag} seed {seed:3d}: >20 {int((sc > 20).sum()):5d} >40 {int((sc > 40).sum()):4d} >50 {int((sc > 50).sum()):3d} >60 {int((sc > 60).sum()):3d} xmax {xmax:6.1f} "
          f"mean {sc.mean():8.4f} sd {sc.std():7.4f} lam {fit.lam:.3f} tau {fit.tau:.3f} E(104,130) {E[(104, 130)]:.3g} E(97,107) {E[(97, 107)]:.3g} "
          f"cells differing from the input {nsub:.0f}", flush=True)
    return [(sc > 20).sum(), (sc > 40).sum(), (sc > 50).sum(), (sc > 60).sum(), xmax, sc.mean(), sc.std(), np.log(E[(104, 130)]), np.log(E[(97, 107)]), nsub]


acc = {"ref": [], "dev": []}
ctx = pkg.Context(0)
ctx.configure(N, L, 4, 4)
ctx.set_weights(wgt)
ctx.set_tree(tree.left, tree.right, tree.parent, tree.ld, tree.rd)
ctx.pool_reserve(20)
for seed in range(1, nseeds + 1):
    acc["ref"].append(stats(np.stack([sh for sh, _, _ in reflib.fitch_shuffle(seed, tree, sub, nrep=20)]), "ref", seed))
    ctx.null_fitch_shuffle(sub, 1000 + seed, 20)
    acc["dev"].append(stats(ctx.pool_get(20, 0), "dev", seed))
for k, v in acc.items():
    a = np.array(v, dtype=float)
    print(k, "mean", np.round(a.mean(0), 4), "sd", np.round(a.std(0), 4))
