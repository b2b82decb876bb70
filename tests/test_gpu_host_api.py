"""The reference's own API (corr_Create / corr_Probs / corr_Calculate* / corr_CalculateCOVCorrected over
struct mutual_s, src/correlators.h:445-482) served by librscape_b200_host.so, driven exactly like
cov_Calculate drives it (src/covariation.c:78-258) and compared with the oracle and, when built, with the
reference's own correlators.c (oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(po):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "libglue_b200.so"))
    host = C.CDLL(os.path.join(ROOT, "r-scape_b200", "librscape_b200_host.so"))
    # the glue library re-exports nothing: bind the API from the host library and the helpers from the glue
    class Both:
        def __getattr__(self, name):
            for l in (lib, host):
                try:
                    return getattr(l, name)
                except AttributeError:
                    continue
            raise AttributeError(name)
    return po._bind_corr_api(Both())


STATS = [("GT", "C16", "APC"), ("GT", "CSELECT", "APC"), ("GT", "CWC", "NOCORR"), ("MI", "C16", "ASC"), ("MIr", "C2", "APC"),
         ("MIg", "C16", "NOCORR"), ("CHI", "C2", "APC"), ("OMES", "C16", "ASC"), ("CCF", "C16", "APC"), ("RAF", "C2", "NOCORR"),
         ("RAFS", "C2", "APC")]


@pytest.mark.parametrize("stat,cls,ac", STATS)
def test_corr_api_matches_oracle_and_reference(po, oracle, stat, cls, ac):
    lib = _lib(po)
    N, L = 220, 61
    msa, wgt, _ = po.synthetic_msa(N, L, seed=31)
    got = po.corr_api_scan(lib, msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac))
    ecls = getattr(po, cls) if cls != "CSELECT" else po.C16          # N > 8 and L > 50
    ref = oracle.scan(msa, wgt, getattr(po, stat), ecls, getattr(po, ac), want_probs=True)
    raw = oracle.scan(msa, wgt, getattr(po, stat), ecls, po.NOCORR)
    scale = max(1.0, abs(raw["maxcov"]), abs(raw["mincov"]))
    off = ~np.eye(L, dtype=bool)
    assert np.max(np.abs(got["cov"][off] - ref["cov"][off])) <= 1e-9 * scale
    assert np.all(np.isneginf(np.diag(got["cov"])))
    assert abs(got["mincov"] - ref["mincov"]) <= 1e-9 * scale and abs(got["maxcov"] - ref["maxcov"]) <= 1e-9 * scale
    for k in ("pp", "pm", "ps", "nseff", "ngap"):
        assert np.max(np.abs(got[k] - ref[k])) <= 1e-9 * max(1.0, np.max(np.abs(ref[k]))), k
    # type / class labels as the reference leaves them in mutual_s
    if po.RefLib.available():
        r = po.RefLib().scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac))
        assert (got["type"], got["covclass"]) == (r["type"], r["covclass"])
        assert np.max(np.abs(got["cov"][off] - r["cov"][off])) <= 1e-9 * scale


def test_corr_api_error_convention(po):
    lib = _lib(po)
    N, L = 50, 20
    msa, wgt, _ = po.synthetic_msa(N, L, seed=5)
    with pytest.raises(RuntimeError, match="CWC not implemented"):
        po.corr_api_scan(lib, msa, wgt, po.CHI, po.CWC, po.NOCORR)
    # AKMAEV is refused with a message, not silently computed
    m = lib.glue_msa_create(N, L, msa.ctypes.data_as(C.POINTER(C.c_uint8)), wgt.ctypes.data_as(C.POINTER(C.c_double)))
    mi = lib.corr_Create(L, N, 0, 8, 50, lib.glue_abc_rna(), po.C16)
    err = C.create_string_buffer(256)
    assert lib.corr_Probs(None, m, None, None, mi, 2, 1e-6, 0, err) != 0 and b"AKMAEV" in err.value
    lib.corr_Destroy(mi)
    lib.glue_msa_destroy(m)


@pytest.mark.parametrize("covtype,stat", [(4, "GT"), (7, "MI")])            # GTp: w stays 0.05; MIp: the first null asks for a narrower bin -> the loop is redone
@pytest.mark.parametrize("gpus", [None, "3", "all"])
def test_null_rscape_b200_matches_oracle_cumulative_ranklist(po, oracle, monkeypatch, gpus, covtype, stat):
    """gpus: RSCAPE_B200_GPUS -- the C-level multi-GPU entry (one host thread + one context per device, histograms summed on the
    host).  On a one-GPU box the three contexts share the device (RSCAPE_B200_GPUS_OVERSUBSCRIBE): same protocol, same result."""
    from test_gpu_nulls import oracle_null_loop
    if gpus is not None:
        monkeypatch.setenv("RSCAPE_B200_GPUS", gpus)
        monkeypatch.setenv("RSCAPE_B200_GPUS_OVERSUBSCRIBE", "1")
    lib = _lib(po)
    glue = C.CDLL(os.path.join(ROOT, "oracle", "libglue_b200.so"))
    N, L, R = 180, 55, 6
    nulls = np.stack([po.synthetic_msa(N, L, seed=900 + r)[0] for r in range(R)])
    wgt = po.synthetic_msa(N, L, seed=9)[1]
    w_ref, view, _ = oracle_null_loop(po, oracle, nulls, wgt, getattr(po, stat), po.C16, po.APC)
    assert (w_ref == 0.05) == (stat == "GT")
    ap = np.ascontiguousarray(po.ALLOWPAIR_WC_GU)
    mi = lib.corr_Create(L, N, 0, 8, 50, lib.glue_abc_rna(), po.C16)
    apm = lib.glue_allowpair_from(ap.ctypes.data_as(C.POINTER(C.c_double)))
    data = lib.glue_data_create(mi, apm, covtype, 1e-6)
    meta, imeta, cnt = np.zeros(5), np.zeros(3, np.int32), np.zeros(3, np.uint64)
    bins = np.zeros(view.nb + 16, np.uint64)
    glue.glue_null_rscape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]
    st = glue.glue_null_rscape(data, R, N, L, nulls.ctypes.data, wgt.ctypes.data, 400, meta.ctypes.data, imeta.ctypes.data,
                               cnt.ctypes.data, bins.ctypes.data, len(bins))
    assert st == 0, lib.glue_data_errbuf(data)
    assert abs(meta[2] - w_ref) <= 1e-12 and meta[0] == view.bmin
    assert imeta[0] == view.nb and imeta[1] == view.imin and imeta[2] == view.imax
    assert abs(meta[1] - view.bmax) <= 1e-9 * abs(view.bmax)
    assert cnt[0] == view.n and cnt[1] == view.Nc and cnt[2] == view.No
    if stat == "GT":
        assert np.array_equal(bins[:view.nb], view.obs)
    else:
        from _helpers import assert_bins_identical
        assert_bins_identical(bins[:view.nb], view.obs, oracle_null_loop.scores, -10.0, w_ref)
    assert abs(meta[3] - view.xmin) <= 1e-9 * max(1, abs(view.xmin)) and abs(meta[4] - view.xmax) <= 1e-9 * max(1, abs(view.xmax))
    lib.glue_data_destroy(data)
    lib.esl_dmatrix_Destroy(apm)
    lib.corr_Destroy(mi)

