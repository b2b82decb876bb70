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
