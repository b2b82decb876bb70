"""E-values and the significant-pair list: pin the oracle (orc_cov2evalue / orc_evalue2cov / orc_hitlist) against the
reference's own `static cov2evalue / evalue2cov` (src/covariation.c:2370-2435), reached through oracle/ref_glue_evalue.c,
which includes that source file where it lies.  The tail fit is Easel code outside the path: the same stand-in tail
(NullFit.exp_tail) is handed to both sides, so what is compared is the arithmetic from (histogram, survfit) to E-values."""
import numpy as np
import pytest


def _null(po, seed, n=200000, w=0.05, bmin=-10.0, shape=2.0, scale=2.5):
    rng = np.random.default_rng(seed)
    x = np.maximum(rng.gamma(shape, scale, n) - 8.0, bmin + w)        # null scores, clamped as covariation.c:431
    b = np.ceil((x - bmin) / w - 1.0).astype(np.int64)
    nb = int(b.max()) + 6                                             # bmax = max + 5 w
    obs = np.bincount(b, minlength=nb).astype(np.uint64)
    return po.NullFit(bmin, w, obs, xmax=float(x.max()))


def _scores(null, rng, n=400):
    edges = null.bmin + null.w * np.arange(0, 2 * null.nb + 3)       # scores exactly on bin bounds and just around them
    pick = rng.choice(edges, 60)
    near = np.concatenate([np.nextafter(pick, np.inf), np.nextafter(pick, -np.inf), pick])
    return np.concatenate([rng.uniform(null.bmin - 3, null.bmin + null.w * (2 * null.nb + 4), n), near,
                           [null.xmax, np.nextafter(null.xmax, -np.inf), null.phi if np.isfinite(null.phi) else 0.0]])


@pytest.mark.parametrize("seed,fit", [(1, False), (2, True), (3, True)])
def test_cov2evalue_matches_reference(po, oracle, reflib, seed, fit):
    null = _null(po, seed)
    if fit:
        null = null.exp_tail(0.02 if seed == 2 else 0.2)
    rng = np.random.default_rng(100 + seed)
    for Nc in (1, 1225):
        for x in _scores(null, rng):
            a = reflib.cov2evalue(x, null, Nc)
            b = oracle.cov2evalue(x, null, Nc)
            assert a == b, (x, Nc, a, b)


@pytest.mark.parametrize("seed,fit", [(4, False), (5, True)])
def test_evalue2cov_matches_reference(po, oracle, reflib, seed, fit):
    null = _null(po, seed)
    if fit:
        null = null.exp_tail(0.05)
    for Nc in (1, 300, 79800):
        for e in (1e-6, 1e-3, 0.05, 1.0, 10.0, 1e4):
            assert reflib.evalue2cov(e, null, Nc) == oracle.evalue2cov(e, null, Nc), (e, Nc)


def test_evalue_is_monotone_and_thresholds_agree(po, oracle):
    """Significance by E-value == significance by the score threshold evalue2cov returns (src/covariation.c:491-497)."""
    null = _null(po, 6).exp_tail(0.05)
    xs = np.sort(np.random.default_rng(6).uniform(-12, 60, 3000))
    ev = np.array([oracle.cov2evalue(x, null, 1000) for x in xs])
    assert np.all(np.diff(ev) <= 0)
    for e in (0.05, 1.0):
        sc = oracle.evalue2cov(e, null, 1000)
        sig = ev < e
        assert np.all(xs[sig] >= sc - null.w)              # the threshold is a bin bound: no significant score lies a bin below it


def _pairs_case(po, oracle, L=40, seed=7):
    msa, wgt, _ = po.synthetic_msa(120, L, seed=seed)
    return oracle.scan(msa, wgt, po.GT, po.C16, po.APC)["cov"]


def test_hitlist_loop(po, oracle):
    """orc_hitlist against a direct Python transcription of the loop at src/covariation.c:828-910."""
    cov = _pairs_case(po, oracle)
    L = cov.shape[0]
    rng = np.random.default_rng(8)
    x = np.maximum(rng.normal(0, 4, 50000), -10 + 0.05)
    b = np.ceil((x + 10) / 0.05 - 1).astype(np.int64)
    null = po.NullFit(-10.0, 0.05, np.bincount(b, minlength=int(b.max()) + 6).astype(np.uint64), xmax=float(x.max())).exp_tail(0.05)
    mask = np.zeros((L, L), np.uint8)
    for i in range(0, 12):
        mask[i, L - 1 - i] = 1
    Nb, Nt = int(mask.sum()), L * (L - 1) // 2 - int(mask.sum())
    for expBP, thresh in ((-1, 0.05), (-1, 5.0), (4, 5.0), (0, 2000.0)):
        got = oracle.hitlist(cov, null, mask, Nb, Nt, expBP, thresh)
        h, want = 0, []
        for i in range(L - 1):
            for j in range(i + 1, L):
                p = oracle.cov2evalue(cov[i, j], null, 1)
                e = p * Nb if mask[i, j] else (p * expBP if h < expBP else p * Nt)
                assert got["Eval"][i, j] == e and got["Eval"][j, i] == e
                if e < thresh or thresh > 1000:
                    want.append((i, j, cov[i, j], e, p))
                    h += 1
        assert len(want) == len(got["i"])
        assert [(int(a), int(b_)) for a, b_ in zip(got["i"], got["j"])] == [(w_[0], w_[1]) for w_ in want]
        assert np.array_equal(got["eval"], [w_[3] for w_ in want]) and np.array_equal(got["pval"], [w_[4] for w_ in want])
        assert np.isinf(np.diag(got["Eval"])).all()
    assert len(oracle.hitlist(cov, null, mask, Nb, Nt, 0, 2000.0)["i"]) == L * (L - 1) // 2      # -E > MAX_EVAL reports all pairs


def test_evalues_match_committed_reference_outputs(po, oracle):
    """tests/golden/ref_evalues.npz holds outputs of the reference's static cov2evalue / evalue2cov (made by
    tests/golden/make_golden.py from oracle/_ref), so the pin also holds where /root/reference never existed."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_evalues.npz"))
    plain = po.NullFit(-10.0, 0.05, z["obs"], xmax=float(z["xmax"]))
    fit = po.NullFit(-10.0, 0.05, z["obs"], float(z["xmax"]), float(z["phi"]), int(z["cmin"]), z["survfit"])
    for name, null in (("plain", plain), ("fit", fit)):
        for Nc in (1, 1225):
            got = np.array([oracle.cov2evalue(v, null, Nc) for v in z["scores"]])
            assert np.array_equal(got, z[f"{name}_cov2evalue_{Nc}"]), (name, Nc)
            got = np.array([oracle.evalue2cov(e, null, Nc) for e in z["thresholds"]])
            assert np.array_equal(got, z[f"{name}_evalue2cov_{Nc}"]), (name, Nc)
