"""The C-ABI libraries load without a GPU and export every function the headers declare.
No compute call is made here (the product path needs a B200 and has no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return sorted({n for n in names if n not in ("defined", "sizeof")})


def test_capi_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(pkg.LIB_PATH)
    names = declared_functions("rscape_b200.h")
    assert len(names) >= 25 and "rsb_scan" in names and "rsb_null_hist" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_library_exports_the_reference_api(pkg):
    lib = C.CDLL(pkg.HOST_LIB_PATH)
    names = declared_functions("rscape_b200_host.h")
    for must in ("corr_Create", "corr_Probs", "corr_CalculateGT", "corr_CalculateCOVCorrected", "corr_CalculateRAFS", "corr_Destroy",
                 "cov_CalculateCOV", "null_rscape_b200", "null_add2cumranklist"):
        assert must in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product refuses to create a context (and says why) instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.RscapeB200Error, match="no CUDA device|no CPU fallback"):
        pkg.Context(0)
    # the reference-API layer behaves the same way: corr_Create returns NULL
    host = C.CDLL(pkg.HOST_LIB_PATH)
    host.esl_alphabet_Create.restype = C.c_void_p
    host.corr_Create.restype = C.c_void_p
    host.corr_Create.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    abc = host.esl_alphabet_Create(1)
    assert host.corr_Create(20, 10, 0, 8, 50, abc, 0) is None


def test_product_does_not_reference_the_oracle():
    """Nothing under r-scape_b200/ may include, link or import oracle/ (the oracle is test infrastructure)."""
    pkgdir = os.path.join(ROOT, "r-scape_b200")
    for dirpath, _, files in os.walk(pkgdir):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle.h" not in txt and "../oracle" not in txt, os.path.join(dirpath, f)


def test_headers_compile_as_c_and_cxx(tmp_path):
    """include/*.h are usable from a C and from a C++ translation unit (extern "C" guards, no C++-only or C-only constructs)."""
    import subprocess
    inc = [f"-I{os.path.join(ROOT, 'include')}", f"-I{os.path.join(ROOT, 'include', 'easel_compat')}"]
    body = '#include "rscape_b200.h"\n#include "rscape_b200_host.h"\n' \
           'int probe(void) { rsb_nullfit nf; HITLIST hl; nf.nb = 0; hl.nhit = 0; return nf.nb + hl.nhit + (int) sizeof(struct mutual_s); }\n'
    for name, cc, std in (("t.c", "gcc", "-std=c99"), ("t.cpp", "g++", "-std=c++17")):
        src = tmp_path / name
        src.write_text(body)
        subprocess.run([cc, std, "-Wall", "-Werror", "-c", str(src), "-o", str(tmp_path / (name + ".o"))] + inc, check=True)


def test_pair_mask_from_ct(pkg):
    """cov_PairMaskFromCT: the base pairs of a ct array (Easel convention) as the pair mask the histogram / hit-list stages take."""
    import numpy as np
    host = C.CDLL(pkg.HOST_LIB_PATH)
    host.cov_PairMaskFromCT.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L = 12
    ct = np.zeros(L + 1, np.int32)
    for i, j in ((1, 12), (2, 11), (4, 8)):
        ct[i], ct[j] = j, i
    mask = np.full((L, L), 7, np.uint8)
    assert host.cov_PairMaskFromCT(ct.ctypes.data, L, mask.ctypes.data) == 0
    want = np.zeros((L, L), np.uint8)
    want[0, 11] = want[1, 10] = want[3, 7] = 1
    assert np.array_equal(mask, want)
    ct[5] = 9                                            # 9 does not point back
    assert host.cov_PairMaskFromCT(ct.ctypes.data, L, mask.ctypes.data) != 0
