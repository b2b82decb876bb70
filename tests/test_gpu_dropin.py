"""The drop-in proof on the device: the reference's UNMODIFIED src/covariation.c (object code built by oracle/Makefile from the
file where it lies, oracle/_ref/libdropin_b200.so) calls its corr_* functions in librscape_b200_host.so -- i.e. the real
cov_Calculate (src/covariation.c:64-306) and the real histogram fill of cov_SignificantPairs_Ranking (:415-457) run over the
reference's own struct data_s / struct mutual_s with the B200 kernels underneath.  Scores must equal the oracle's within 1e-9,
the rank list's integer bins must be the oracle's."""
import ctypes as C
import os

import numpy as np
import pytest

from _helpers import assert_bins_identical

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libdropin_b200.so")

COVTYPE = dict(CHI=0, CHIp=1, CHIa=2, GT=3, GTp=4, GTa=5, MI=6, MIp=7, MIa=8, MIr=9, MIrp=10, MIra=11, MIg=12, MIgp=13, MIga=14,
               OMES=15, OMESp=16, OMESa=17, RAF=18, RAFp=19, RAFa=20, RAFS=21, RAFSp=22, RAFSa=23, CCF=24, CCFp=25, CCFa=26)


@pytest.fixture(scope="module")
def dropin(pkg):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libdropin_b200.so not built (needs /root/reference at build time)")
    lib = C.CDLL(DROPIN)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.dropin_cov_calculate.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint8), dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         dp, dp, ip, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    lib.dropin_errbuf.restype = C.c_char_p
    return lib


def run(lib, msa, wgt, covtype, covclass, analyze, w=0.05, bmin=-10.0, nb_cap=1 << 16):
    N, L = msa.shape
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    wgt = np.ascontiguousarray(wgt, dtype=np.float64)
    cov, meta, imeta = np.empty((L, L)), np.zeros(7), np.zeros(5, np.int32)
    n, bins = C.c_uint64(0), np.zeros(nb_cap, np.uint64)
    dp = C.POINTER(C.c_double)
    st = lib.dropin_cov_calculate(N, L, msa.ctypes.data_as(C.POINTER(C.c_uint8)), wgt.ctypes.data_as(dp), covtype, covclass, int(analyze), w, bmin,
                                  cov.ctypes.data_as(dp), meta.ctypes.data_as(dp), imeta.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n),
                                  bins.ctypes.data_as(C.POINTER(C.c_uint64)), nb_cap)
    assert st == 0, lib.dropin_errbuf().decode()
    return dict(cov=cov, mincov=meta[0], maxcov=meta[1], type=int(imeta[0]), cls=int(imeta[1]), bmin=meta[2], bmax=meta[3], w=meta[4],
                xmin=meta[5], xmax=meta[6], nb=int(imeta[2]), imin=int(imeta[3]), imax=int(imeta[4]), n=n.value, bins=bins[:int(imeta[2])])


@pytest.mark.parametrize("name,cls", [("GTp", "C16"), ("GTp", "CSELECT"), ("MIa", "C2"), ("CHI", "C16"), ("OMESp", "C16"), ("MIrp", "C16"),
                                      ("MIgp", "C2"), ("RAFSp", "C2"), ("GTa", "CWC")])
def test_reference_cov_calculate_on_the_device(dropin, po, oracle, name, cls):
    N, L = 260, 77
    msa, wgt, _ = po.synthetic_msa(N, L, seed=31)
    stat = name.rstrip("pa") if name not in ("CHI",) else name
    ac = po.APC if name.endswith("p") else po.ASC if name.endswith("a") and name != "CHI" else po.NOCORR
    pcls = {"C16": po.C16, "C2": po.C2, "CWC": po.CWC, "CSELECT": po.C16}[cls]          # CSELECT: N > 8 and L > 50 -> C16 (correlators.c:336)
    got = run(dropin, msa, wgt, COVTYPE[name], {"C16": 0, "C2": 1, "CWC": 2, "CSELECT": 3}[cls], analyze=True)
    ref = oracle.scan(msa, wgt, getattr(po, stat), pcls, ac)
    raw = oracle.scan(msa, wgt, getattr(po, stat), pcls, po.NOCORR)
    scale = max(1.0, abs(raw["maxcov"]), abs(raw["mincov"]))
    iu = np.triu_indices(L, 1)
    if stat in ("RAF", "RAFS") and ac == po.NOCORR:
        assert np.array_equal(got["cov"][iu], ref["cov"][iu])
    assert np.max(np.abs(got["cov"][iu] - ref["cov"][iu])) <= 1e-9 * scale
    assert np.array_equal(got["cov"], got["cov"].T)
    assert abs(got["mincov"] - ref["mincov"]) <= 1e-9 * scale and abs(got["maxcov"] - ref["maxcov"]) <= 1e-9 * scale
    assert got["type"] == COVTYPE[name]                                                 # mi->type as corr_CalculateCOVCorrected renames it (:1078-1090)
    # the rank list the reference's own loop filled from the device's scores
    h = oracle.hist_from_cov(ref["cov"], ref["maxcov"], -10.0, 0.05)
    v = oracle.view(h)
    oracle.free(h)
    assert got["n"] == L * (L - 1) // 2 == v.n and got["w"] == 0.05 and got["bmin"] == -10.0
    assert_bins_identical(got["bins"], v.obs, ref["cov"][iu], -10.0, 0.05, scale=scale)


def test_reference_cov_calculate_without_ranking(dropin, po, oracle):
    """analyze = FALSE, as calculate_width_histo calls it (src/R-scape.c:1299)"""
    msa, wgt, _ = po.synthetic_msa(120, 64, seed=8)
    got = run(dropin, msa, wgt, COVTYPE["GTp"], 0, analyze=False)
    ref = oracle.scan(msa, wgt, po.GT, po.C16, po.APC)
    iu = np.triu_indices(64, 1)
    assert np.max(np.abs(got["cov"][iu] - ref["cov"][iu])) <= 1e-9 * max(1.0, abs(ref["maxcov"]) * 4)
    assert got["nb"] == 0
