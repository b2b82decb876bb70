"""The score -> p-value function the E-value kernel runs (r-scape_b200/csrc/rsb_evalue.cuh, __host__ __device__) is compiled
here with g++ and compared with the oracle's restatement of cov2evalue (src/covariation.c:2370-2400) -- which is itself
pinned against the reference's own static function (tests/test_evalue_oracle.py) -- on scores that hit every branch, bin
bounds included.  This checks the kernel's arithmetic on the CPU; the kernel itself is checked by tests/test_gpu_z_hits.py."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "rsb_evalue.cuh"
extern "C" int hdr_pvals(const double *x, int n, double bmin, double w, double xmax, double phi, double Nc, int nb, int imin, int imax,
                         const unsigned long long *csum, const double *survfit, double *out)
{
  rsb_nullview h;
  h.bmin = bmin; h.w = w; h.xmax = xmax; h.phi = phi; h.Nc = Nc; h.nb = nb; h.imin = imin; h.imax = imax; h.csum = csum; h.survfit = survfit;
  int nbad = 0;
  for (int k = 0; k < n; k++) { int bad = 0; out[k] = rsb_cov2pval(x[k], h, &bad); nbad += bad; }
  return nbad;
}
'''


@pytest.fixture(scope="module")
def hdr():
    td = tempfile.mkdtemp()
    src, so = os.path.join(td, "h.cpp"), os.path.join(td, "h.so")
    open(src, "w").write(SRC)
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so,
                    "-I" + os.path.join(ROOT, "r-scape_b200", "csrc")], check=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hdr_pvals.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                              C.POINTER(C.c_uint64), dp, dp]
    return lib


def _pvals(lib, x, null):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    csum = np.zeros(null.nb, np.uint64)
    csum[:null.imax + 1] = np.cumsum(null.obs[:null.imax + 1][::-1])[::-1]
    dp = C.POINTER(C.c_double)
    sf = None if null.survfit is None else null.survfit.ctypes.data_as(dp)
    nbad = lib.hdr_pvals(x.ctypes.data_as(dp), len(x), null.bmin, null.w, null.xmax, null.phi, float(null.Nc), null.nb, null.imin, null.imax,
                         csum.ctypes.data_as(C.POINTER(C.c_uint64)), sf, out.ctypes.data_as(dp))
    return out, nbad


@pytest.mark.parametrize("seed,pmass", [(11, None), (12, 0.02), (13, 0.3)])
def test_header_pvalues_equal_the_oracle(po, oracle, hdr, seed, pmass):
    from test_evalue_oracle import _null, _scores
    null = _null(po, seed, n=300000)
    if pmass is not None:
        null = null.exp_tail(pmass)
    x = _scores(null, np.random.default_rng(seed), n=3000)
    got, nbad = _pvals(hdr, x, null)
    want = np.array([oracle.cov2evalue(v, null, 1) for v in x])
    assert nbad == 0 and np.array_equal(got, want)


def test_header_small_histograms(po, oracle, hdr):
    """one- and two-bin histograms: imin == imax, icov >= imax - 1 everywhere"""
    for obs in ([0, 0, 7, 0, 0, 0, 0, 0], [0, 3, 4, 0, 0, 0, 0, 0], [5, 0, 0, 0, 0, 0, 0, 9]):
        null = po.NullFit(-10.0, 0.5, np.array(obs, np.uint64))
        x = np.linspace(-12, -4, 161)
        got, nbad = _pvals(hdr, x, null)
        want = np.array([oracle.cov2evalue(v, null, 1) for v in x])
        assert nbad == 0 and np.array_equal(got, want)


def test_header_random_histograms(po, oracle, hdr):
    """Random small histograms (1 to 40 bins, sparse bins, tails starting anywhere) and scores on and around every bin bound:
    the kernel's p-value function, the oracle and -- when built -- the reference's own cov2evalue agree bit for bit."""
    ref = po.RefLib() if po.RefLib.available() else None
    rng = np.random.default_rng(0)
    ncmp = 0
    for trial in range(120):
        nb = int(rng.integers(1, 40))
        w = float(rng.choice([0.05, 0.5, 1e-3, 2.0]))
        bmin = float(rng.choice([-10.0, 0.0, -3.3]))
        obs = np.zeros(nb, np.uint64)
        k = int(rng.integers(1, nb + 1))
        obs[rng.choice(nb, k, replace=False)] = rng.integers(1, 1000, k).astype(np.uint64)
        null = po.NullFit(bmin, w, obs)
        null = po.NullFit(bmin, w, obs, xmax=bmin + w * (null.imax + float(rng.uniform(0.01, 1.0))))
        if rng.random() < 0.6:
            cmin = int(rng.integers(null.imin, null.imax + 1))
            surv = np.zeros(2 * nb)
            surv[cmin:] = np.sort(rng.uniform(0, 0.3, 2 * nb - cmin))[::-1]
            null = po.NullFit(bmin, w, obs, null.xmax, bmin + w * cmin, cmin, surv)
        x = np.concatenate([rng.uniform(bmin - 2 * w, bmin + w * (2 * nb + 3), 40), bmin + w * np.arange(-1, 2 * nb + 3)])
        got, nbad = _pvals(hdr, x, null)
        assert nbad == 0
        want = np.array([oracle.cov2evalue(v, null, 1) for v in x])
        assert np.array_equal(got, want), trial
        if ref is not None:
            assert np.array_equal(want, np.array([ref.cov2evalue(v, null, 1) for v in x])), trial
        ncmp += len(x)
    assert ncmp > 5000
