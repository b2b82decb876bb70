"""The drop-in boundary, CPU side: the host layer compiles against the REFERENCE's own headers
(-DRSB_USE_RSCAPE_HEADERS, the build INTEGRATION.md tells a maintainer to use), and the reference's unmodified
src/covariation.c links against librscape_b200_host.so in place of correlators.o (oracle/_ref/libdropin_b200.so,
built by oracle/Makefile; exercised on the device by tests/test_gpu_dropin.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libdropin_b200.so")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present (build container only)")
@pytest.mark.parametrize("src", ["correlators_b200.c", "covariation_b200.c", "msatree_b200.c"])
def test_host_layer_compiles_against_the_reference_headers(tmp_path, src):
    cmd = ["gcc", "-O1", "-fPIC", "-Wall", "-Wno-unused-function", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
           "-DRSB_USE_RSCAPE_HEADERS", f"-I{REF}/src", f"-I{REF}/lib/R-view/src", f"-I{ROOT}/include", f"-I{ROOT}/include/easel_compat",
           "-c", os.path.join(ROOT, "r-scape_b200", "host", src), "-o", str(tmp_path / "o.o")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/libdropin_b200.so not built (needs /root/reference at build time)")
def test_reference_covariation_links_against_the_b200_host_library():
    """cov_Calculate is the reference's own object code; every corr_* symbol it calls is left undefined in the test library and
    resolved by librscape_b200_host.so -- none of them is stubbed, none comes from the reference's correlators.c."""
    out = subprocess.run(["nm", "-D", DROPIN], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    undefined = {l.split()[-1] for l in out if " U " in l}
    defined = {l.split()[-1] for l in out if " T " in l}
    assert {"cov_Calculate", "cov_SignificantPairs_Ranking", "dropin_cov_calculate"} <= defined
    for sym in ("corr_Create", "corr_Destroy", "corr_Probs", "corr_CalculateGT", "corr_CalculateMI", "corr_CalculateCHI", "corr_CalculateOMES",
                "corr_CalculateRAFS", "corr_CalculateCOVCorrected", "corr_COVTYPEString"):
        assert sym in undefined, sym
    assert not any(s.startswith("corr_") for s in defined)
    host = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "r-scape_b200", "librscape_b200_host.so")],
                          stdout=subprocess.PIPE, text=True, check=True).stdout
    provided = {l.split()[-1] for l in host.splitlines()}
    assert {s for s in undefined if s.startswith("corr_")} <= provided
    stubs = open(os.path.join(ROOT, "oracle", "_ref", "dropin_stubs.c")).read()
    assert "void corr_" not in stubs and "void esl_histogram_Add(" not in stubs and "void esl_histogram_CreateFull(" not in stubs


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/libdropin_b200.so not built (needs /root/reference at build time)")
def test_null_histogram_files_through_the_reference_io(tmp_path):
    """--savenull / --givennull (src/R-scape.c:2466, 2480): the reference's unmodified cov_WriteNullHistogram and cov_ReadNullHistogram
    (src/covariation.c:1845-1866, 1718-1843) on a cumulative null histogram in the form the B200 null loop leaves it (bmin, w, integer
    bins).  The file is `<lower edge of the bin, %f> <count>` per bin; read back, the list holds the same counts in the same order
    (the reader re-derives bmin and w from the file and may grow the list below by two bins: a constant index shift) and the same mass."""
    import ctypes as C

    import numpy as np
    lib = C.CDLL(DROPIN)
    u64p, dp, ip = C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.dropin_null_histogram_roundtrip.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_int, u64p, dp, ip, u64p, u64p, C.c_int]
    bmin, w, nb = -10.0, 0.05, 1500
    rng = np.random.default_rng(1)
    sc = rng.gamma(2.0, 4.0, 200000) - 9.0
    idx = np.ceil((np.maximum(sc, bmin + w) - bmin) / w - 1).astype(int)          # Easel's bin rule, as the device histogram applies it
    bins = np.zeros(nb, np.uint64)
    np.add.at(bins, idx[idx < nb], 1)
    path = str(tmp_path / "family.null")
    meta, imeta, n, out = np.zeros(5), np.zeros(3, np.int32), C.c_uint64(), np.zeros(4 * nb, np.uint64)
    rc = lib.dropin_null_histogram_roundtrip(path.encode(), bmin, w, nb, bins.ctypes.data_as(u64p), meta.ctypes.data_as(dp), imeta.ctypes.data_as(ip),
                                             C.byref(n), out.ctypes.data_as(u64p), len(out))
    assert rc == 0
    lines = open(path).read().splitlines()
    assert len(lines) == nb
    for i in (0, 1, 417, nb - 1):
        assert lines[i] == "%f %d" % (bmin + i * w, int(bins[i]))
    assert n.value == int(bins.sum())
    assert abs(meta[2] - w) <= 1e-6
    nz, nzb = np.nonzero(bins)[0], np.nonzero(out)[0]
    shift = int(nzb[0] - nz[0])
    assert 0 <= shift <= 2
    assert np.array_equal(out[shift:shift + nb], bins)
    assert abs((meta[0] + shift * meta[2]) - bmin) <= 1e-6                       # ... i.e. the same score intervals
