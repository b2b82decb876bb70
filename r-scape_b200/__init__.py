"""rscape_b200 -- Python plumbing over the C-ABI of include/rscape_b200.h (librscape_b200.so).

The product is the shared library (hand-written sm_100a kernels behind a plain C-ABI) and the C host layer
that mirrors the reference's covariation API (librscape_b200_host.so).  This module only loads them through
ctypes for the tests, the benchmark and multi-GPU orchestration (torch.distributed); it contains no
arithmetic and no fallback: if the CUDA library is missing or no B200 is present, it raises.

The package directory is named ``r-scape_b200`` (not importable by name); ``__graft_entry__.load_package()``
registers it as ``rscape_b200``.
"""
import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (seeded synthetic workloads, host-side utility)
from . import parallel  # noqa: F401  (null-replicate sharding over torch.distributed)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSCAPE_B200_LIB") or os.path.join(HERE, "librscape_b200.so")     # (override: experimental builds, tools/build_variant.sh)
HOST_LIB_PATH = os.path.join(HERE, "librscape_b200_host.so")

CHI, GT, MI, MIr, MIg, OMES, RAF, RAFS, CCF = 0, 3, 6, 9, 12, 15, 18, 21, 24
C16, C2, CWC, CSELECT = 0, 1, 2, 3
APC, ASC, NOCORR = 0, 1, 2

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

_lib = None


class NullFitStruct(C.Structure):
    """rsb_nullfit of include/rscape_b200.h"""
    _fields_ = [("bmin", C.c_double), ("w", C.c_double), ("nb", C.c_int), ("imin", C.c_int), ("imax", C.c_int), ("xmax", C.c_double),
                ("phi", C.c_double), ("Nc", C.c_uint64), ("obs", _u64p), ("survfit", _dp)]


class RscapeB200Error(RuntimeError):
    pass


def lib():
    """The C-ABI library.  Fails loudly when it has not been built: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RscapeB200Error(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C r-scape_b200)")
        L = C.CDLL(LIB_PATH)
        L.rsb_create.argtypes = [C.c_int, _vp, C.POINTER(_vp)]
        L.rsb_destroy.argtypes = [_vp]
        L.rsb_error.restype = C.c_char_p
        L.rsb_error.argtypes = [_vp]
        L.rsb_create_error.restype = C.c_char_p
        L.rsb_configure.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.rsb_set_weights.argtypes = [_vp, _dp]
        L.rsb_get_quantisation.argtypes = [_vp, _i64p, _ip, _ip]
        L.rsb_get_quantisation_error.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.rsb_set_null_slices.argtypes = [_vp, C.c_int]
        L.rsb_get_null_quantisation.argtypes = [_vp, _i64p, _ip, _ip, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.rsb_probs.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp]
        L.rsb_fetch_probs.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
        L.rsb_statistic.argtypes = [_vp, C.c_int, C.c_int, _dp, _vp, C.c_int64, C.c_int, _dp, _dp, _dp]
        L.rsb_correct.argtypes = [_vp, C.c_int, _dp, _dp, _dp]
        L.rsb_correct_host.argtypes = [_vp, C.c_int, _dp, _dp, _dp]
        L.rsb_scan.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double,
                               _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.rsb_scan_hist.argtypes = [_vp, _u8p, C.c_double, C.c_double, C.c_int, _u64p, _u64p, _u64p]
        L.rsb_scan_hits.argtypes = [_vp, C.POINTER(NullFitStruct), _u8p, C.c_uint64, C.c_uint64, C.c_int, C.c_double, _dp, C.c_int64,
                                    _i64p, _i64p, _dp, _dp, _dp, _i64p]
        L.rsb_load_scores.argtypes = [_vp, _dp]
        L.rsb_set_pair_exclusion.argtypes = [_vp, _ip, C.c_int]
        L.rsb_tree_substitutions.argtypes = [_vp, C.c_int, _ip, _ip, _u8p, C.c_int64, _u8p, C.c_int64, C.c_int, _ip, _ip, _ip]
        L.rsb_set_shard.argtypes = [_vp, C.c_int, C.c_int]
        L.rsb_sharded_counts.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_double, _dp]
        L.rsb_sharded_counts_pool.argtypes = [_vp, C.c_int, C.c_double, _dp]
        L.rsb_sharded_statistic.argtypes = [_vp, _dp, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.rsb_sharded_correct.argtypes = [_vp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.rsb_get_counts.argtypes = [_vp, _i64p]
        L.rsb_get_counts_direct.argtypes = [_vp, _vp, C.c_int64, _i64p]
        L.rsb_null_width.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double,
                                     C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
        L.rsb_null_hist.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                    C.c_double, C.c_double, C.c_double, _dp]
        L.rsb_null_hist_pool.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double,
                                         C.c_double, _dp]
        L.rsb_null_width_pool.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                          _dp, _dp, _dp]
        L.rsb_null_hist_multi.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, _ip, _ip, C.c_int, _dp, C.c_double, _dp,
                                          C.c_double, _dp]
        L.rsb_null_hist_multi_pool.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _dp, C.c_double, _dp, C.c_double, _dp]
        L.rsb_hist_read_multi.argtypes = [_vp, C.c_int, _u64p, C.c_int, _u64p, _ip]
        L.rsb_hist_reset_multi.argtypes = [_vp]
        L.rsb_pool_reserve.argtypes = [_vp, C.c_int]
        L.rsb_pool_get.argtypes = [_vp, C.c_int, C.c_int, _u8p]
        L.rsb_pool_put.argtypes = [_vp, C.c_int, C.c_int, _u8p]
        L.rsb_pool_get_internal.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _u8p]
        L.rsb_hist_reset.argtypes = [_vp]
        L.rsb_hist_exchange.argtypes = [_vp, C.c_void_p, C.c_int, C.c_int]
        L.rsb_hist_read.argtypes = [_vp, _u64p, C.c_int, _u64p, _ip]
        L.rsb_last_nseff.argtypes = [_vp, _dp, _dp]
        L.rsb_set_tree.argtypes = [_vp, _ip, _ip, _ip, _dp, _dp]
        L.rsb_null_simulate.argtypes = [_vp, _dp, _u8p, _u8p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.rsb_null_fitch_shuffle.argtypes = [_vp, _u8p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.rsb_null_fitch_shuffle_ids.argtypes = [_vp, _u8p, C.c_int64, C.c_uint64, _u64p, C.c_int, C.c_int]
        L.rsb_counters.argtypes = [_vp, _i64p, _dp, _i64p, C.c_int]
        L.rsb_msa_gap_columns.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int64, C.c_int, _dp, C.c_double, _u8p]
        L.rsb_msa_column_subset.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int64, C.c_int, _u8p, _vp, C.c_int, _ip]
        L.rsb_msa_pb_weights.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int64, C.c_int, _dp]
        L.rsb_msa_pair_identity.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int64, C.c_int, _ip, C.c_int64, _dp]
        L.rsb_comm_id.argtypes = [_u8p]
        L.rsb_comm_init.argtypes = [_vp, _u8p, C.c_int, C.c_int]
        L.rsb_comm_init_all.argtypes = [C.POINTER(_vp), C.c_int]
        L.rsb_comm_destroy.argtypes = [_vp]
        L.rsb_comm_info.argtypes = [_vp, _ip, _ip, _ip, _i64p]
        L.rsb_pool_broadcast.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
        L.rsb_comm_selftest.argtypes = [_vp, C.c_int, C.c_int, _dp, _dp]
        L.rsb_hist_allreduce.argtypes = [_vp, C.c_int]
        L.rsb_comm_range.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.rsb_sharded_scan.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, _dp, _dp, _dp]
        L.rsb_profile_gram.argtypes = [_vp, C.c_int]
        L.rsb_host_register.argtypes = [_vp, C.c_size_t]
        L.rsb_host_unregister.argtypes = [_vp]
        L.rsb_counters_geometry.argtypes = [_vp, C.c_int, C.POINTER(C.c_double), _i64p, _ip]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _ptr(x):
    """(pointer, on_device) of a numpy array or a torch CUDA tensor."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(_vp), 0
    return C.c_void_p(x.data_ptr()), (1 if x.is_cuda else 0)


def pin(array):
    """Page-lock a numpy buffer that is handed to the library repeatedly (rsb_host_register): copies then run at the PCIe rate.
    Returns True if it is now pinned; call unpin() before the array is freed."""
    assert isinstance(array, np.ndarray) and array.flags.c_contiguous
    return lib().rsb_host_register(array.ctypes.data_as(_vp), array.nbytes) == 0


def unpin(array):
    return lib().rsb_host_unregister(array.ctypes.data_as(_vp)) == 0


def comm_id():
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    if lib().rsb_comm_id(buf) != 0:
        raise RscapeB200Error(lib().rsb_create_error().decode())
    return bytes(buf)


def comm_init_all(contexts):
    """One process driving several devices: a communicator over the contexts (rank k = contexts[k])."""
    arr = (_vp * len(contexts))(*[c._h for c in contexts])
    if lib().rsb_comm_init_all(arr, len(contexts)) != 0:
        raise RscapeB200Error(lib().rsb_error(contexts[0]._h).decode())


class Context:
    """One device + one stream.  Mirrors the C-ABI one to one; numpy in, numpy out."""

    def __init__(self, device=0, stream=None):
        self._h = _vp()
        self.N = self.L = 0
        self.device = device
        L = lib()
        if L.rsb_create(device, _vp(stream) if stream else None, C.byref(self._h)) != 0:
            raise RscapeB200Error(L.rsb_create_error().decode())

    def close(self):
        if self._h:
            lib().rsb_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RscapeB200Error(lib().rsb_error(self._h).decode())

    # ---- configuration ------------------------------------------------------------------------
    def configure(self, nseq, alen, max_replicates=1, nslices=0):
        self._ck(lib().rsb_configure(self._h, nseq, alen, max_replicates, nslices))
        self.N, self.L, self.R = nseq, alen, max_replicates

    def set_weights(self, wgt=None):
        w = None if wgt is None else np.ascontiguousarray(wgt, dtype=np.float64)
        self._ck(lib().rsb_set_weights(self._h, _d(w)))

    def quantisation(self):
        wq = np.zeros(self.N, dtype=np.int64)
        q, S = C.c_int(), C.c_int()
        self._ck(lib().rsb_get_quantisation(self._h, wq.ctypes.data_as(_i64p), C.byref(q), C.byref(S)))
        return wq, q.value, S.value

    def quantisation_error(self):
        """(largest |wq 2^-q - w| in weight units, log2(max weight / that error))."""
        a, b = C.c_double(), C.c_double()
        self._ck(lib().rsb_get_quantisation_error(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_null_slices(self, nslices):
        """Mixed precision: the nulls are contracted with `nslices` digit slices of the weights (0 = off); call before set_weights."""
        self._ck(lib().rsb_set_null_slices(self._h, nslices))

    def null_quantisation(self):
        """(wq, q, S, largest |wq 2^-q - w|, effective bits) of the weights the null alignments are scored with."""
        wq = np.zeros(self.N, dtype=np.int64)
        q, S, a, b = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        self._ck(lib().rsb_get_null_quantisation(self._h, wq.ctypes.data_as(_i64p), C.byref(q), C.byref(S), C.byref(a), C.byref(b)))
        return wq, q.value, S.value, a.value, b.value

    # ---- one alignment ------------------------------------------------------------------------
    def _msa(self, msa):
        if isinstance(msa, np.ndarray):
            msa = np.ascontiguousarray(msa, dtype=np.uint8)
            assert msa.shape == (self.N, self.L), (msa.shape, self.N, self.L)
        return msa

    def scan(self, msa, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, want_cov=True, want_probs=False, cov_out=None):
        """cov_out: optional float64 [L][L] array to receive the scores (e.g. a view of pinned memory)."""
        msa = self._msa(msa)
        p, dev = _ptr(msa)
        L = self.L
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        if cov_out is not None:
            assert cov_out.shape == (L, L) and cov_out.dtype == np.float64 and cov_out.flags.c_contiguous
        out = dict(cov=(cov_out if cov_out is not None else np.empty((L, L))) if want_cov else None, pp=None, pm=None, ps=None, nseff=None, ngap=None)
        if want_probs:
            out.update(pp=np.empty((L, L, 16)), pm=np.empty((L, 4)), ps=np.empty((L, 5)), nseff=np.empty((L, L)), ngap=np.empty((L, L)))
        mn, mx = C.c_double(), C.c_double()
        self._ck(lib().rsb_scan(self._h, p, L, dev, stat, covclass, actype, _d(ap), tol, _d(out["cov"]), C.byref(mn), C.byref(mx),
                                _d(out["pp"]), _d(out["pm"]), _d(out["ps"]), _d(out["nseff"]), _d(out["ngap"])))
        out.update(mincov=mn.value, maxcov=mx.value)
        return out

    def scan_hist(self, w, bmin, nb, pairmask=None):
        """ha / hb / ht of the last scan (cov_SignificantPairs_Ranking, src/covariation.c:415-457)."""
        ha = np.zeros(nb, np.uint64)
        hb = np.zeros(nb, np.uint64) if pairmask is not None else None
        ht = np.zeros(nb, np.uint64) if pairmask is not None else None
        pm = None if pairmask is None else np.ascontiguousarray(pairmask, dtype=np.uint8)
        up = lambda a: None if a is None else a.ctypes.data_as(_u64p)
        self._ck(lib().rsb_scan_hist(self._h, None if pm is None else pm.ctypes.data_as(_u8p), w, bmin, nb, up(ha), up(hb), up(ht)))
        return ha, hb, ht

    def set_pair_exclusion(self, msa2pdb=None, mind=1):
        """Pairs with both columns in the PDB sequence and closer than mind stay out of the histograms (covariation.c:421-427)."""
        m = None if msa2pdb is None else np.ascontiguousarray(msa2pdb, dtype=np.int32)
        assert m is None or len(m) == self.L
        self._ck(lib().rsb_set_pair_exclusion(self._h, None if m is None else m.ctypes.data_as(_ip), int(mind)))

    def load_scores(self, cov):
        """Replace the device score matrix by a host matrix [L][L] (the stages after the scan read whatever mi->COV holds)."""
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        assert cov.shape == (self.L, self.L)
        self._ck(lib().rsb_load_scores(self._h, _d(cov)))

    def scan_hits(self, bmin, w, obs, xmax, Nt, Nb=0, pairmask=None, survfit=None, phi=np.inf, expBP=-1, thresh=0.05, want_eval=True, cap=None, eval_out=None):
        """E-values and significant pairs of the last scan (cov_CreateHitList's per-pair loop, src/covariation.c:828-910) against
        the cumulative null histogram obs[nb] (geometry bmin, w; largest null score xmax) and, optionally, its fitted tail
        survfit[2 nb] with censoring point phi.  Returns dict(i, j, sc, eval, pval, nhit, Eval)."""
        obs = np.ascontiguousarray(obs, dtype=np.uint64)
        nz = np.nonzero(obs)[0]
        if len(nz) == 0:
            raise RscapeB200Error("scan_hits: the null histogram is empty")
        sf = None if survfit is None else np.ascontiguousarray(survfit, dtype=np.float64)
        assert sf is None or len(sf) == 2 * len(obs)
        nf = NullFitStruct(float(bmin), float(w), len(obs), int(nz[0]), int(nz[-1]), float(xmax), float(phi), int(obs.sum()),
                           obs.ctypes.data_as(_u64p), _d(sf))
        pm = None if pairmask is None else np.ascontiguousarray(pairmask, dtype=np.uint8)
        assert pm is None or pm.shape == (self.L, self.L)
        P = self.L * (self.L - 1) // 2
        ev = (eval_out if eval_out is not None else np.empty((self.L, self.L))) if want_eval else None
        assert ev is None or (ev.shape == (self.L, self.L) and ev.dtype == np.float64 and ev.flags.c_contiguous)
        # cap None: lists are short unless every pair is reported; start small and repeat the call once if the list is longer
        retry = cap is None
        cap = (P if (thresh > 1000 or expBP > 0) else min(P, 1 << 16)) if cap is None else int(cap)
        while True:
            hi, hj = np.empty(max(cap, 1), np.int64), np.empty(max(cap, 1), np.int64)
            sc, he, hp = np.empty(max(cap, 1)), np.empty(max(cap, 1)), np.empty(max(cap, 1))
            n = C.c_int64()
            self._ck(lib().rsb_scan_hits(self._h, C.byref(nf), None if pm is None else pm.ctypes.data_as(_u8p), int(Nb), int(Nt), int(expBP),
                                         float(thresh), _d(ev), cap, hi.ctypes.data_as(_i64p), hj.ctypes.data_as(_i64p), _d(sc), _d(he), _d(hp),
                                         C.byref(n)))
            if not retry or n.value <= cap:
                break
            cap, retry = n.value, False
        k = min(n.value, cap)
        return dict(i=hi[:k].copy(), j=hj[:k].copy(), sc=sc[:k].copy(), eval=he[:k].copy(), pval=hp[:k].copy(), nhit=n.value, Eval=ev)

    def tree_substitutions(self, left, right, leaves, internal, includegaps=False, want_pairs=True, out=None):
        """Tree_Substitutions after its Fitch pass (src/msatree.c:1455-1540) -> (nsubs [L], ndouble [L][L], njoin [L][L]).
        The context must be configured with nseq = 2 (ntaxa - 1) rows (one per branch)."""
        leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
        internal = np.ascontiguousarray(internal, dtype=np.uint8)
        ntaxa, L = leaves.shape
        assert internal.shape == (ntaxa - 1, L) and L == self.L
        lf, rt = np.ascontiguousarray(left, dtype=np.int32), np.ascontiguousarray(right, dtype=np.int32)
        ns = np.empty(L, np.int32)
        nd = (out[0] if out is not None else np.empty((L, L), np.int32)) if want_pairs else None       # out: caller's (e.g. pinned) tables
        nj = (out[1] if out is not None else np.empty((L, L), np.int32)) if want_pairs else None
        ip = lambda a: None if a is None else a.ctypes.data_as(_ip)
        self._ck(lib().rsb_tree_substitutions(self._h, ntaxa, ip(lf), ip(rt), leaves.ctypes.data_as(_u8p), L, internal.ctypes.data_as(_u8p), L,
                                              1 if includegaps else 0, ip(ns), ip(nd), ip(nj)))
        return ns, nd, nj

    # ---- alignment preprocessing (any shape, independent of configure) ----------------------------
    @staticmethod
    def _any_msa(msa):
        if isinstance(msa, np.ndarray):
            msa = np.ascontiguousarray(msa, dtype=np.uint8)
        return msa, int(msa.shape[0]), int(msa.shape[1])

    def msa_gap_columns(self, msa, wgt=None, gapthresh=0.75):
        """useme uint8 [alen]: the column test of msamanip_RemoveGapColumns (src/msamanip.c:486-500)."""
        msa, N, L = self._any_msa(msa)
        p, dev = _ptr(msa)
        w = None if wgt is None else np.ascontiguousarray(wgt, dtype=np.float64)
        use = np.empty(L, np.uint8)
        self._ck(lib().rsb_msa_gap_columns(self._h, p, N, L, L, dev, _d(w), gapthresh, use.ctypes.data_as(_u8p)))
        return use

    def msa_column_subset(self, msa, useme):
        msa, N, L = self._any_msa(msa)
        p, dev = _ptr(msa)
        use = np.ascontiguousarray(useme, dtype=np.uint8)
        out = np.empty((N, int(np.count_nonzero(use))), np.uint8)
        n = C.c_int()
        self._ck(lib().rsb_msa_column_subset(self._h, p, N, L, L, dev, use.ctypes.data_as(_u8p), out.ctypes.data_as(_vp), 0, C.byref(n)))
        assert n.value == out.shape[1]
        return out

    def msa_pb_weights(self, msa):
        msa, N, L = self._any_msa(msa)
        p, dev = _ptr(msa)
        w = np.empty(N)
        self._ck(lib().rsb_msa_pb_weights(self._h, p, N, L, L, dev, _d(w)))
        return w

    def msa_pair_identity(self, msa, pairs=None):
        """pairs int [npairs][2] -> pid [npairs]; None -> the distance matrix 1 - pid [N][N]."""
        msa, N, L = self._any_msa(msa)
        p, dev = _ptr(msa)
        if pairs is None:
            out = np.empty((N, N))
            self._ck(lib().rsb_msa_pair_identity(self._h, p, N, L, L, dev, None, 0, _d(out)))
            return out
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty(len(pr))
        self._ck(lib().rsb_msa_pair_identity(self._h, p, N, L, L, dev, pr.ctypes.data_as(_ip), len(pr), _d(out)))
        return out

    # ---- communicator ----------------------------------------------------------------------------
    def comm_init(self, id128, nranks, rank):
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self._ck(lib().rsb_comm_init(self._h, buf, nranks, rank))

    def comm_destroy(self):
        self._ck(lib().rsb_comm_destroy(self._h))

    def pool_broadcast(self, first_rep, nrep, root):
        """Pool entries generated by rank `root` made resident on every rank of the communicator (collective)."""
        self._ck(lib().rsb_pool_broadcast(self._h, first_rep, nrep, root))

    def comm_selftest(self, count, iters=200):
        """(microseconds per all-reduce of `count` doubles, largest error of a checked sum) -- collective."""
        us, err = C.c_double(), C.c_double()
        self._ck(lib().rsb_comm_selftest(self._h, count, iters, C.byref(us), C.byref(err)))
        return us.value, err.value

    def comm_info(self):
        """dict(nranks, rank, peer_path, reductions): peer_path = the small per-scan all-reduces go through the one-shot kernel
        over NVLink peer memory (csrc/peer_reduce.cu) rather than NCCL; reductions = how many it has done."""
        n, r, p, k = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._ck(lib().rsb_comm_info(self._h, C.byref(n), C.byref(r), C.byref(p), C.byref(k)))
        return dict(nranks=n.value, rank=r.value, peer_path=bool(p.value), reductions=k.value)

    def hist_allreduce(self, nb):
        """Sum the first nb bins of the device histograms of all ranks, in place (null_add2cumranklist across ranks)."""
        self._ck(lib().rsb_hist_allreduce(self._h, int(nb)))

    def comm_range(self, lo, hi, aux_min=np.inf):
        """(min over ranks of lo, max over ranks of hi, min over ranks of aux_min)."""
        a, b, c = C.c_double(lo), C.c_double(hi), C.c_double(aux_min)
        self._ck(lib().rsb_comm_range(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sharded_scan(self, msa, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, want_cov=True, cov_out=None):
        """rsb_scan on a pair grid sharded over the communicator's ranks; the corrected matrix is assembled on every rank."""
        msa = self._msa(msa)
        p, dev = _ptr(msa)
        L = self.L
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        cov = (cov_out if cov_out is not None else np.empty((L, L))) if want_cov else None
        mn, mx = C.c_double(), C.c_double()
        self._ck(lib().rsb_sharded_scan(self._h, p, L, dev, stat, covclass, actype, _d(ap), tol, _d(cov), C.byref(mn), C.byref(mx)))
        return dict(cov=cov, mincov=mn.value, maxcov=mx.value)

    # ---- one scan with the pair grid sharded over ranks ------------------------------------------
    def set_shard(self, rank, world):
        self._ck(lib().rsb_set_shard(self._h, rank, world))

    def sharded_counts(self, msa, tol=1e-6):
        msa = self._msa(msa)
        p, dev = _ptr(msa)
        sums = np.empty((self.L, 4))
        self._ck(lib().rsb_sharded_counts(self._h, p, self.L, dev, tol, _d(sums)))
        return sums

    def sharded_counts_pool(self, rep, tol=1e-6):
        sums = np.empty((self.L, 4))
        self._ck(lib().rsb_sharded_counts_pool(self._h, rep, tol, _d(sums)))
        return sums

    def sharded_statistic(self, marg_sums, stat=GT, covclass=C16, allowpair=None, tol=1e-6):
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        ms = np.ascontiguousarray(marg_sums, dtype=np.float64)
        out = np.empty(self.L + 4)
        self._ck(lib().rsb_sharded_statistic(self._h, _d(ms), tol, stat, covclass, _d(ap), _d(out)))
        return out

    def sharded_correct(self, cov_sums, actype=APC, want_cov=True, hist_w=None, bmin=-10.0):
        cs = np.ascontiguousarray(cov_sums, dtype=np.float64)
        cov = np.empty((self.L, self.L)) if want_cov else None
        mm = np.empty(2)
        mode = (1 if want_cov else 0) | (2 if hist_w is not None else 0)
        self._ck(lib().rsb_sharded_correct(self._h, _d(cs), actype, mode, 0.0 if hist_w is None else hist_w, bmin, _d(cov), _d(mm)))
        return cov, mm[0], mm[1]

    def counts(self):
        out = np.empty((16, self.L, self.L), dtype=np.int64)
        self._ck(lib().rsb_get_counts(self._h, out.ctypes.data_as(_i64p)))
        return out

    def counts_direct(self, msa):
        msa = self._msa(msa)
        out = np.empty((16, self.L, self.L), dtype=np.int64)
        self._ck(lib().rsb_get_counts_direct(self._h, msa.ctypes.data_as(_vp), self.L, out.ctypes.data_as(_i64p)))
        return out

    # ---- nulls --------------------------------------------------------------------------------
    def null_width(self, null0, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, w_old=0.05, bmin=-10.0, hpts=400):
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        if null0 is None:
            p, dev = None, 0
        else:
            null0 = self._msa(null0)
            p, dev = _ptr(null0)
        w, mn, mx = C.c_double(), C.c_double(), C.c_double()
        self._ck(lib().rsb_null_width(self._h, p, self.L, dev, stat, covclass, actype, _d(ap), tol, w_old, bmin, hpts,
                                      C.byref(w), C.byref(mn), C.byref(mx)))
        return w.value, mn.value, mx.value

    def null_hist(self, nulls, w, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, bmin=-10.0, want_minmax=True):
        """nulls: uint8 [R][N][L], numpy (host) or torch CUDA tensor (device)."""
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        if isinstance(nulls, np.ndarray):
            nulls = np.ascontiguousarray(nulls, dtype=np.uint8)
        R = nulls.shape[0]
        assert tuple(nulls.shape[1:]) == (self.N, self.L)
        p, dev = _ptr(nulls)
        mm = np.empty((R, 2)) if want_minmax else None
        self._ck(lib().rsb_null_hist(self._h, p, R, self.L, self.N * self.L, dev, stat, covclass, actype, _d(ap), tol, w, bmin, _d(mm)))
        return mm

    def null_width_pool(self, rep=0, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, w_old=0.05, bmin=-10.0, hpts=400):
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        w, mn, mx = C.c_double(), C.c_double(), C.c_double()
        self._ck(lib().rsb_null_width_pool(self._h, rep, stat, covclass, actype, _d(ap), tol, w_old, bmin, hpts,
                                           C.byref(w), C.byref(mn), C.byref(mx)))
        return w.value, mn.value, mx.value

    def null_hist_pool(self, first_rep, nrep, w, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, bmin=-10.0, want_minmax=True):
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        mm = np.empty((nrep, 2)) if want_minmax else None
        self._ck(lib().rsb_null_hist_pool(self._h, first_rep, nrep, stat, covclass, actype, _d(ap), tol, w, bmin, _d(mm)))
        return mm

    def null_hist_multi(self, nulls, combos, w, covclass=C16, allowpair=None, tol=1e-6, bmin=-10.0, first_rep=None):
        """Several (statistic, correction) combinations from one contraction per null (rsb_null_hist_multi).
        combos: list of (stat, actype); w: one bin width per combination (<= 0: score range only).  nulls: uint8 [R][N][L] numpy /
        torch CUDA tensor, or an int R with first_rep for pool entries.  Returns minmax [ncombo][R][2]."""
        ap = None if allowpair is None else np.ascontiguousarray(allowpair, dtype=np.float64)
        st = np.ascontiguousarray([c[0] for c in combos], dtype=np.int32)
        ac = np.ascontiguousarray([c[1] for c in combos], dtype=np.int32)
        ww = np.ascontiguousarray(w, dtype=np.float64)
        assert len(ww) == len(combos)
        if first_rep is not None:
            R = int(nulls)
            mm = np.empty((len(combos), R, 2))
            self._ck(lib().rsb_null_hist_multi_pool(self._h, first_rep, R, len(combos), st.ctypes.data_as(_ip), ac.ctypes.data_as(_ip), covclass,
                                                    _d(ap), tol, _d(ww), bmin, _d(mm)))
            return mm
        if isinstance(nulls, np.ndarray):
            nulls = np.ascontiguousarray(nulls, dtype=np.uint8)
        R = nulls.shape[0]
        assert tuple(nulls.shape[1:]) == (self.N, self.L)
        p, dev = _ptr(nulls)
        mm = np.empty((len(combos), R, 2))
        self._ck(lib().rsb_null_hist_multi(self._h, p, R, self.L, self.N * self.L, dev, len(combos), st.ctypes.data_as(_ip), ac.ctypes.data_as(_ip),
                                           covclass, _d(ap), tol, _d(ww), bmin, _d(mm)))
        return mm

    def hist_read_multi(self, combo, nb):
        bins = np.empty(nb, dtype=np.uint64)
        n, imax = C.c_uint64(), C.c_int()
        self._ck(lib().rsb_hist_read_multi(self._h, combo, bins.ctypes.data_as(_u64p), nb, C.byref(n), C.byref(imax)))
        return bins, n.value, imax.value

    def hist_reset_multi(self):
        self._ck(lib().rsb_hist_reset_multi(self._h))

    def hist_reset(self):
        self._ck(lib().rsb_hist_reset(self._h))

    def hist_read(self, nb, out=None):
        """First nb bins of the cumulative histogram -> (bins, n, imax).  out: optional uint64 array of at least nb elements
        to receive them (e.g. a view of pinned memory)."""
        bins = np.empty(nb, dtype=np.uint64) if out is None else out[:nb]           # fully overwritten by the copy
        assert bins.dtype == np.uint64 and bins.flags.c_contiguous and len(bins) == nb
        n, imax = C.c_uint64(), C.c_int()
        self._ck(lib().rsb_hist_read(self._h, bins.ctypes.data_as(_u64p), nb, C.byref(n), C.byref(imax)))
        return bins, n.value, imax.value

    def hist_exchange(self, tensor, to_library):
        """Copy the first len(tensor) bins between the device histogram and a torch int64/uint64 CUDA tensor (device to device)."""
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.element_size() == 8
        self._ck(lib().rsb_hist_exchange(self._h, C.c_void_p(tensor.data_ptr()), tensor.numel(), 1 if to_library else 0))

    def last_nseff(self):
        ne, ng = np.empty((self.L, self.L)), np.empty((self.L, self.L))
        self._ck(lib().rsb_last_nseff(self._h, _d(ne), _d(ng)))
        return ne, ng

    # ---- generators ---------------------------------------------------------------------------
    def set_tree(self, left, right, parent, ld, rd):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (left, right, parent)]
        b = [np.ascontiguousarray(x, dtype=np.float64) for x in (ld, rd)]
        self._ck(lib().rsb_set_tree(self._h, a[0].ctypes.data_as(_ip), a[1].ctypes.data_as(_ip), a[2].ctypes.data_as(_ip), _d(b[0]), _d(b[1])))

    def pool_reserve(self, nrep):
        self._ck(lib().rsb_pool_reserve(self._h, nrep))

    def null_simulate(self, Q, root, seed, nrep, gapmask=None, first_rep=0, first_id=None):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        root = np.ascontiguousarray(root, dtype=np.uint8)
        gm = None if gapmask is None else np.ascontiguousarray(gapmask, dtype=np.uint8)
        self._ck(lib().rsb_null_simulate(self._h, _d(Q), root.ctypes.data_as(_u8p), None if gm is None else gm.ctypes.data_as(_u8p),
                                         self.L, seed, first_rep if first_id is None else first_id, first_rep, nrep))

    def null_fitch_shuffle(self, msa, seed, nrep, first_rep=0, first_id=None, ids=None):
        """ids: optional explicit global replicate ids (len nrep) instead of first_id, first_id + 1, ..."""
        msa = self._msa(msa)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            assert len(ids) == nrep
            self._ck(lib().rsb_null_fitch_shuffle_ids(self._h, msa.ctypes.data_as(_u8p), self.L, seed, ids.ctypes.data_as(_u64p), first_rep, nrep))
            return
        self._ck(lib().rsb_null_fitch_shuffle(self._h, msa.ctypes.data_as(_u8p), self.L, seed,
                                              first_rep if first_id is None else first_id, first_rep, nrep))

    def pool_get(self, nrep, first_rep=0):
        out = np.empty((nrep, self.N, self.L), dtype=np.uint8)
        self._ck(lib().rsb_pool_get(self._h, first_rep, nrep, out.ctypes.data_as(_u8p)))
        return out

    def pool_get_internal(self, which, nrep, first_rep=0):
        """Generator A's internal-node rows [nrep][N-1][L]: which = 0 Fitch reconstruction, 1 shuffled rows (parity tests)."""
        out = np.empty((nrep, self.N - 1, self.L), dtype=np.uint8)
        self._ck(lib().rsb_pool_get_internal(self._h, which, first_rep, nrep, out.ctypes.data_as(_u8p)))
        return out

    def pool_put(self, nulls, first_rep=0):
        nulls = np.ascontiguousarray(nulls, dtype=np.uint8)
        self._ck(lib().rsb_pool_put(self._h, first_rep, nulls.shape[0], nulls.ctypes.data_as(_u8p)))

    # ---- instrumentation ------------------------------------------------------------------------
    def profile_gram(self, enable=True):
        self._ck(lib().rsb_profile_gram(self._h, 1 if enable else 0))

    def counters(self, reset=False):
        a, b, c = C.c_int64(), C.c_double(), C.c_int64()
        self._ck(lib().rsb_counters(self._h, C.byref(a), C.byref(b), C.byref(c), 1 if reset else 0))
        out = dict(launches=a.value, gram_ms=b.value, gram_launches=c.value, geometry=[])
        for which in range(3):                                       # contractions per operand geometry (input weights, unit weights, nulls' weights)
            ms, n, S = C.c_double(), C.c_int64(), C.c_int()
            self._ck(lib().rsb_counters_geometry(self._h, which, C.byref(ms), C.byref(n), C.byref(S)))
            out["geometry"].append(dict(gram_ms=ms.value, gram_launches=n.value, slices=S.value))
        return out


def replicate_slots(nseq, alen, nnull, nslices=4, budget_bytes=24e9, sm_count=148):
    """How many null replicates to keep in flight: enough tiles for every SM, bounded by HBM use."""
    cj = {1: 64, 2: 32, 3: 20, 4: 16, 5: 12, 6: 10}[nslices]
    tiles = (alen / 32.0) * (alen / cj) * 0.5 + 1.0
    per_rep = (4 + 4 * nslices) * alen * nseq + 16 * 8 * alen * alen + 3 * 8 * alen * alen + nseq * alen
    r = int(np.ceil(4.0 * sm_count / tiles))
    r = max(r, 4)                       # four slot groups: the statistics chain of a chunk has three contractions to finish in
    return int(max(1, min(r, int(budget_bytes // per_rep), max(nnull, 2), 64)))
