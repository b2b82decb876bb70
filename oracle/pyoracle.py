"""pyoracle -- ctypes/numpy front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under r-scape_b200/ does.

* `Oracle`     wraps oracle/liboracle.so (oracle.c: flat-array restatement of src/correlators.c,
               the histogram fill of src/covariation.c:415-457, the null accumulation of
               src/R-scape.c:1565-1612 and both null generators).
* `RefLib`     wraps oracle/_ref/librscape_ref.so -- the reference's own src/correlators.c (+ msatree.c,
               msamanip.c, cov_simulate.c) compiled unchanged against the Easel shim -- when present.
* numpy helpers restate the host-side preprocessing that defines the analysed matrix: Stockholm
  reader, gap-column filter (src/msamanip.c:461-542), GSC / PB weights (Easel), and a
  seeded synthetic-MSA generator used by tests and bench.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_synth():
    import importlib.util
    path = os.path.join(os.path.dirname(HERE), "r-scape_b200", "synth.py")
    spec = importlib.util.spec_from_file_location("rscape_b200_synth", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_synth = _load_synth()          # workload generators live in the package (bench.py uses them without touching oracle/)
Tree = _synth.Tree
_Tree = _synth._TreeStruct
random_tree = _synth.random_tree
synthetic_msa = _synth.synthetic_msa
synthetic_family = _synth.synthetic_family

CHI, GT, MI, MIr, MIg, OMES, RAF, RAFS, CCF = 0, 3, 6, 9, 12, 15, 18, 21, 24
C16, C2, CWC, CSELECT = 0, 1, 2, 3
APC, ASC, NOCORR = 0, 1, 2
STAT_NAMES = {CHI: "CHI", GT: "GT", MI: "MI", MIr: "MIr", MIg: "MIg", OMES: "OMES", RAF: "RAF", RAFS: "RAFS", CCF: "CCF"}

ALLOWPAIR_WC_GU = np.zeros((4, 4))
for _a, _b in ((0, 3), (3, 0), (1, 2), (2, 1), (2, 3), (3, 2)):     # src/R-scape.c:883-887
    ALLOWPAIR_WC_GU[_a, _b] = 1.0

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _u8(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class _Hist(C.Structure):
    _fields_ = [("bmin", C.c_double), ("bmax", C.c_double), ("w", C.c_double), ("nb", C.c_int),
                ("imin", C.c_int), ("imax", C.c_int), ("xmin", C.c_double), ("xmax", C.c_double),
                ("n", C.c_uint64), ("Nc", C.c_uint64), ("No", C.c_uint64), ("obs", C.POINTER(C.c_uint64))]


class Hist:
    """Python view of an ORC_HIST (copy of the bins + geometry)."""

    def __init__(self, h):
        self.bmin, self.bmax, self.w, self.nb = h.bmin, h.bmax, h.w, h.nb
        self.imin, self.imax, self.xmin, self.xmax = h.imin, h.imax, h.xmin, h.xmax
        self.n, self.Nc, self.No = h.n, h.Nc, h.No
        self.obs = np.array([h.obs[b] for b in range(h.nb)], dtype=np.uint64)


class NullFit:
    """What cov2evalue reads of data->ranklist_null (src/covariation.c:2370-2400): the cumulative null histogram `ha`
    (geometry, bins, imin/imax, xmax, Nc), its censoring point phi / first fitted bin cmin and the fitted tail survfit[2 nb]
    (cov_histogram_SetSurvFitTail, :1677-1699).  The tail FIT itself is Easel code and stays outside the path: tests pass any
    non-increasing tail, e.g. exp_tail() below, to the oracle, the reference and the device alike."""

    def __init__(self, bmin, w, obs, xmax=None, phi=np.inf, cmin=None, survfit=None):
        self.obs = np.ascontiguousarray(obs, dtype=np.uint64)
        nb = len(self.obs)
        nz = np.nonzero(self.obs)[0]
        self.bmin, self.w, self.nb = float(bmin), float(w), nb
        self.imin = int(nz[0]) if len(nz) else nb
        self.imax = int(nz[-1]) if len(nz) else -1
        self.Nc = int(self.obs.sum())
        self.No = self.Nc
        self.xmax = float(bmin + w * (self.imax + 1)) if xmax is None else float(xmax)   # a score on the upper edge of bin imax
        self.phi = float(phi)
        self.cmin = int(nb if cmin is None else cmin)
        self.survfit = None if survfit is None else np.ascontiguousarray(survfit, dtype=np.float64)
        assert self.survfit is None or len(self.survfit) == 2 * nb
        self.chist = _Hist(self.bmin, self.bmin + self.w * nb, self.w, nb, self.imin, self.imax, self.bmin, self.xmax,
                           self.Nc, self.Nc, self.No, self.obs.ctypes.data_as(C.POINTER(C.c_uint64)))

    def exp_tail(self, pmass=0.01):
        """A stand-in for the reference's tail fit: censor the top `pmass` of the scores (what esl_histogram_SetTailByMass
        does to phi / cmin) and give that tail an exponential survival with the decay of its own mean excess."""
        cum = np.cumsum(self.obs[::-1].astype(np.float64))[::-1]                  # scores in bins >= b
        tail = np.nonzero(cum <= pmass * self.Nc)[0]
        cmin = int(tail[0]) if len(tail) else self.imax
        cmin = max(cmin, self.imin + 1)
        phi = self.bmin + self.w * cmin                                           # lower bound of bin cmin
        b = np.arange(cmin, self.imax + 1)
        n = self.obs[cmin:self.imax + 1].astype(np.float64)
        mass = n.sum() / self.Nc
        mean_excess = float((n * (self.bmin + self.w * (b + 0.5) - phi)).sum() / max(n.sum(), 1.0))
        lam = 1.0 / max(mean_excess, self.w)
        surv = np.zeros(2 * self.nb)
        ub = self.bmin + self.w * (np.arange(cmin, 2 * self.nb) + 1)
        surv[cmin:] = mass * np.exp(-lam * (ub - phi))                            # survfit[b] = pmass * surv(UBound(b)), :1692
        return NullFit(self.bmin, self.w, self.obs, self.xmax, phi, cmin, surv)


def _nullfit_call(fn, null, pmass, fracfit, doexpfit):
    geom = np.array([null.bmin, null.w, null.xmax])
    ig = np.array([null.nb, null.imin, null.imax], dtype=np.int32)
    surv, out = np.zeros(2 * null.nb), np.zeros(6)
    fn.argtypes = [_dp, _ip, C.c_uint64, C.POINTER(C.c_uint64), C.c_double, C.c_double, C.c_int, _dp, _dp]
    rc = fn(_d(geom), _i(ig), null.Nc, null.obs.ctypes.data_as(C.POINTER(C.c_uint64)), pmass, fracfit, 1 if doexpfit else 0, _d(surv), _d(out))
    if rc != 0:
        raise RuntimeError(f"tail fit failed with status {rc}")
    fit = NullFit(null.bmin, null.w, null.obs, null.xmax, phi=out[4], cmin=int(out[5]), survfit=surv if surv.any() else None)
    fit.newmass, fit.mu, fit.lam, fit.tau = out[0], out[1], out[2], out[3]
    return fit


def nullfit_host(null, pmass=0.0005, fracfit=1.0, doexpfit=False):
    """The "Histogram and Fit" block of cov_SignificantPairs_Ranking (src/covariation.c:459-487) as the host library runs it
    (cov_NullFit_b200 in r-scape_b200/host/covariation_b200.c, through oracle/libglue_b200.so): R-scape's defaults --pmass 0.0005,
    --fracfit 1.0, gamma fit (src/R-scape.c:428-429).  Returns a NullFit with phi / cmin / survfit set."""
    lib = C.CDLL(os.path.join(HERE, "libglue_b200.so"))
    return _nullfit_call(lib.glue_nullfit, null, pmass, fracfit, doexpfit)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(path)
        L.orc_scan.restype = C.c_int
        L.orc_scan.argtypes = [_u8p, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_double,
                               _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_pair_counts_fixed.argtypes = [_u8p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_raf.argtypes = [_u8p, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_rafs.argtypes = [_u8p, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_raf_from_counts.argtypes = [_u8p, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_time_pair_probs.restype = C.c_double
        L.orc_time_pair_probs.argtypes = [_u8p, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp]
        L.orc_hist_from_cov.restype = C.POINTER(_Hist)
        L.orc_hist_from_cov.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_hist_destroy.argtypes = [C.POINTER(_Hist)]
        L.orc_hist_accumulate.argtypes = [C.POINTER(C.POINTER(_Hist)), C.POINTER(_Hist)]
        L.orc_null_width.restype = C.c_double
        L.orc_null_width.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        i64p = C.POINTER(C.c_int64)
        L.orc_cov2evalue.restype = C.c_double
        L.orc_cov2evalue.argtypes = [C.c_double, C.c_int, C.POINTER(_Hist), C.c_double, _dp]
        L.orc_evalue2cov.restype = C.c_double
        L.orc_evalue2cov.argtypes = [C.c_double, C.c_int, C.POINTER(_Hist), C.c_int, _dp]
        L.orc_hitlist.restype = C.c_int64
        L.orc_hitlist.argtypes = [_dp, C.c_int, C.POINTER(_Hist), C.c_double, _dp, _u8p, C.c_uint64, C.c_uint64, C.c_int, C.c_double,
                                  _dp, C.c_int64, i64p, i64p, _dp, _dp, _dp]
        L.orc_rng_create.restype = C.c_void_p
        L.orc_rng_create.argtypes = [C.c_uint32]
        L.orc_rng_destroy.argtypes = [C.c_void_p]
        L.orc_rng_uniform.restype = C.c_double
        L.orc_rng_uniform.argtypes = [C.c_void_p]
        L.orc_ptime.argtypes = [_dp, C.c_double, _dp]
        L.orc_null_simulate.argtypes = [C.c_void_p, C.POINTER(_Tree), _dp, _u8p, C.c_int, _u8p, _u8p]
        L.orc_null_fitch_shuffle.argtypes = [C.c_void_p, C.POINTER(_Tree), _u8p, C.c_int, _u8p, _u8p, _ip]
        L.orc_tree_substitutions.argtypes = [C.POINTER(_Tree), _u8p, C.c_int, C.c_int, _ip, _ip, _ip]

    # ---- one scan -------------------------------------------------------------------------
    def scan(self, msa, wgt, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6, want_probs=False):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        wgt = np.ascontiguousarray(wgt, dtype=np.float64)
        N, L = msa.shape
        ap = np.ascontiguousarray(ALLOWPAIR_WC_GU if allowpair is None else allowpair, dtype=np.float64)
        cov = np.empty((L, L))
        mn, mx = C.c_double(), C.c_double()
        out = dict(pp=None, pm=None, ps=None, nseff=None, ngap=None)
        if want_probs:
            out = dict(pp=np.zeros((L, L, 16)), pm=np.zeros((L, 4)), ps=np.zeros((L, 5)),
                       nseff=np.zeros((L, L)), ngap=np.zeros((L, L)))
        st = self.lib.orc_scan(_u8(msa), N, L, _d(wgt), stat, covclass, actype, _d(ap), tol, _d(cov),
                               C.byref(mn), C.byref(mx), _d(out["pp"]), _d(out["pm"]), _d(out["ps"]),
                               _d(out["nseff"]), _d(out["ngap"]))
        if st != 0:
            raise RuntimeError(f"orc_scan failed with status {st}")
        out.update(cov=cov, mincov=mn.value, maxcov=mx.value)
        return out

    def counts_fixed(self, msa, wq):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        wq = np.ascontiguousarray(wq, dtype=np.int64)
        N, L = msa.shape
        out = np.zeros((L, L, 16), dtype=np.int64)
        self.lib.orc_pair_counts_fixed(_u8(msa), N, L, wq.ctypes.data_as(C.POINTER(C.c_int64)),
                                       out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def raf_direct(self, msa, allowpair=None, smooth=False):
        """RAF / RAFS exactly as the reference loops (O(P N^2)); small inputs only."""
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        N, L = msa.shape
        ap = np.ascontiguousarray(ALLOWPAIR_WC_GU if allowpair is None else allowpair, dtype=np.float64)
        cov = np.empty((L, L))
        mn, mx = C.c_double(), C.c_double()
        if smooth:
            st = self.lib.orc_rafs(_u8(msa), N, L, _d(ap), 0, _d(cov), C.byref(mn), C.byref(mx))
        else:
            st = self.lib.orc_raf(_u8(msa), N, L, _d(ap), _d(cov), C.byref(mn), C.byref(mx))
        assert st == 0
        return cov, mn.value, mx.value

    def time_pair_probs(self, msa, wgt, row_stride=1, nthreads=1):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        wgt = np.ascontiguousarray(wgt, dtype=np.float64)
        N, L = msa.shape
        chk = C.c_double()
        secs = self.lib.orc_time_pair_probs(_u8(msa), N, L, _d(wgt), row_stride, nthreads, C.byref(chk))
        rows = range(0, L - 1, row_stride)
        pairs = sum(L - 1 - i for i in rows)
        return secs, pairs

    # ---- histograms -----------------------------------------------------------------------
    def hist_from_cov(self, cov, maxcov, bmin=-10.0, w=0.05, tol=1e-6):
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        h = self.lib.orc_hist_from_cov(_d(cov), cov.shape[0], maxcov, bmin, w, tol)
        if not h:
            raise RuntimeError("orc_hist_from_cov failed")
        return h            # opaque pointer; use view()/accumulate()/free()

    def view(self, hptr):
        return Hist(hptr.contents)

    def accumulate(self, cum, one):
        """cum: POINTER(_Hist) or None; returns the new cumulative pointer."""
        holder = C.POINTER(_Hist)() if cum is None else cum
        ref = C.pointer(holder)
        st = self.lib.orc_hist_accumulate(ref, one)
        assert st == 0
        return ref.contents

    def free(self, hptr):
        self.lib.orc_hist_destroy(hptr)

    def null_width(self, w_old, mincov, maxcov, bmin=-10.0, hpts=400, tol=1e-6):
        return self.lib.orc_null_width(w_old, mincov, maxcov, bmin, hpts, tol)

    # ---- E-values and hit list ---------------------------------------------------------------
    def cov2evalue(self, cov, null, Nc=1):
        """null: a NullFit (cumulative null histogram + optional fitted tail)"""
        return self.lib.orc_cov2evalue(float(cov), Nc, C.byref(null.chist), null.phi, _d(null.survfit))

    def evalue2cov(self, eval_thresh, null, Nc=1):
        return self.lib.orc_evalue2cov(float(eval_thresh), Nc, C.byref(null.chist), null.cmin, _d(null.survfit))

    def hitlist(self, cov, null, pairmask=None, Nb=0, Nt=None, expBP=-1, thresh=0.05, want_eval=True):
        """The per-pair loop of cov_CreateHitList (src/covariation.c:828-910): dict(i, j, sc, eval, pval, Eval matrix)."""
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        L = cov.shape[0]
        P = L * (L - 1) // 2
        Nt = P if Nt is None else Nt
        mask = None if pairmask is None else np.ascontiguousarray(pairmask, dtype=np.uint8)
        ev = np.full((L, L), np.inf) if want_eval else None
        hi, hj = np.empty(P, np.int64), np.empty(P, np.int64)
        sc, he, hp = np.empty(P), np.empty(P), np.empty(P)
        i64p = C.POINTER(C.c_int64)
        n = self.lib.orc_hitlist(_d(cov), L, C.byref(null.chist), null.phi, _d(null.survfit), _u8(mask), int(Nb), int(Nt), int(expBP),
                                 float(thresh), _d(ev), P, hi.ctypes.data_as(i64p), hj.ctypes.data_as(i64p), _d(sc), _d(he), _d(hp))
        if n < 0:
            raise RuntimeError("cannot find evalue for a covariation score (src/covariation.c:2394)")
        return dict(i=hi[:n].copy(), j=hj[:n].copy(), sc=sc[:n].copy(), eval=he[:n].copy(), pval=hp[:n].copy(), Eval=ev)

    # ---- null generators --------------------------------------------------------------------
    def rng(self, seed):
        return self.lib.orc_rng_create(seed)

    def random(self, r):
        """next uniform deviate of the stream (esl_random)"""
        return self.lib.orc_rng_uniform(r)

    def rng_free(self, r):
        self.lib.orc_rng_destroy(r)

    def ptime(self, Q, t):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        P = np.empty((4, 4))
        st = self.lib.orc_ptime(_d(Q), t, _d(P))
        assert st == 0, st
        return P

    def null_simulate(self, rng, tree, Q, root):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        root = np.ascontiguousarray(root, dtype=np.uint8)
        L = root.shape[0]
        leaves = np.empty((tree.N, L), dtype=np.uint8)
        t = tree.cstruct()
        st = self.lib.orc_null_simulate(rng, C.byref(t), _d(Q), _u8(root), L, _u8(leaves), None)
        assert st == 0, st
        return leaves

    def null_fitch_shuffle(self, rng, tree, msa, want_all=False):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        N, L = msa.shape
        assert N == tree.N
        sh = np.empty((N, L), dtype=np.uint8)
        allm = np.empty((2 * N - 1, L), dtype=np.uint8) if want_all else None
        sc = C.c_int()
        t = tree.cstruct()
        st = self.lib.orc_null_fitch_shuffle(rng, C.byref(t), _u8(msa), L, _u8(sh), _u8(allm), C.byref(sc))
        assert st == 0, st
        return (sh, allm, sc.value) if want_all else sh


    def tree_substitutions(self, tree, allmsa, includegaps=False):
        """Tree_Substitutions after its Fitch pass (src/msatree.c:1455-1540) on allmsa [2N-1][L] -> (nsubs, ndouble, njoin)."""
        allmsa = np.ascontiguousarray(allmsa, dtype=np.uint8)
        assert allmsa.shape[0] == 2 * tree.N - 1
        L = allmsa.shape[1]
        ns, nd, nj = np.zeros(L, np.int32), np.zeros((L, L), np.int32), np.zeros((L, L), np.int32)
        t = tree.cstruct()
        st = self.lib.orc_tree_substitutions(C.byref(t), _u8(allmsa), L, 1 if includegaps else 0, _i(ns), _i(nd), _i(nj))
        assert st == 0
        return ns, nd, nj


class RefLib:
    """The reference's own code (oracle/_ref/librscape_ref.so), driven through oracle/mi_glue.c."""

    COVTYPE_OF = {}

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_ref", "librscape_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = _bind_corr_api(C.CDLL(path))
        g = self.lib
        g.glue_tree_create.restype = C.c_void_p
        g.glue_tree_create.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp]
        g.glue_tree_destroy.argtypes = [C.c_void_p]
        g.glue_ref_fitch_shuffle.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _u8p, _u8p, _u8p, _ip]
        g.glue_ref_simulate.argtypes = [C.c_void_p, C.c_void_p, _dp, _u8p, C.c_int, _u8p]
        g.glue_ref_ptime.argtypes = [_dp, C.c_double, _dp]
        g.esl_randomness_Create.restype = C.c_void_p
        g.esl_randomness_Create.argtypes = [C.c_uint32]
        g.esl_randomness_Destroy.argtypes = [C.c_void_p]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "librscape_ref.so"))

    def scan(self, msa, wgt, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6,
             nseqthresh=8, alenthresh=50):
        return corr_api_scan(self.lib, msa, wgt, stat, covclass, actype, allowpair, tol, nseqthresh, alenthresh)

    def _tree(self, tree):
        return self.lib.glue_tree_create(tree.N, _i(tree.left), _i(tree.right), _i(tree.parent), _d(tree.ld), _d(tree.rd))

    def fitch_shuffle(self, seed, tree, msa, nrep=1):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        N, L = msa.shape
        T = self._tree(tree)
        r = self.lib.esl_randomness_Create(seed)
        outs = []
        for _ in range(nrep):
            sh = np.empty((N, L), np.uint8)
            allm = np.empty((2 * N - 1, L), np.uint8)
            sc = C.c_int()
            st = self.lib.glue_ref_fitch_shuffle(r, T, L, _u8(msa), _u8(sh), _u8(allm), C.byref(sc))
            assert st == 0, st
            outs.append((sh, allm, sc.value))
        self.lib.esl_randomness_Destroy(r)
        self.lib.glue_tree_destroy(T)
        return outs

    def tree_substitutions(self, seed, tree, msa, includegaps=False):
        """The reference's Tree_Substitutions (its own Fitch pass on the shim's MT19937 stream seeded with `seed`)."""
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        N, L = msa.shape
        T = self._tree(tree)
        r = self.lib.esl_randomness_Create(seed)
        ns, nd, nj = np.zeros(L, np.int32), np.zeros((L, L), np.int32), np.zeros((L, L), np.int32)
        self.lib.glue_ref_tree_substitutions.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _u8p, C.c_int, _ip, _ip, _ip]
        st = self.lib.glue_ref_tree_substitutions(r, T, L, _u8(msa), 1 if includegaps else 0, _i(ns), _i(nd), _i(nj))
        self.lib.esl_randomness_Destroy(r)
        self.lib.glue_tree_destroy(T)
        assert st == 0, st
        return ns, nd, nj

    def simulate(self, seed, tree, Q, root, nrep=1):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        root = np.ascontiguousarray(root, dtype=np.uint8)
        L = root.shape[0]
        T = self._tree(tree)
        r = self.lib.esl_randomness_Create(seed)
        outs = []
        for _ in range(nrep):
            leaves = np.empty((tree.N, L), np.uint8)
            st = self.lib.glue_ref_simulate(r, T, _d(Q), _u8(root), L, _u8(leaves))
            assert st == 0, st
            outs.append(leaves)
        self.lib.esl_randomness_Destroy(r)
        self.lib.glue_tree_destroy(T)
        return outs

    def ptime(self, Q, t):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        P = np.empty((4, 4))
        st = self.lib.glue_ref_ptime(_d(Q), t, _d(P))
        assert st == 0
        return P

    # the reference's own static cov2evalue / evalue2cov (oracle/ref_glue_evalue.c includes src/covariation.c where it lies)
    def nullfit(self, null, pmass=0.0005, fracfit=1.0, doexpfit=False):
        """The reference's own cov_histogram_pmass + cov_NullFitGamma / cov_NullFitExponential (src/covariation.c:459-487, 1915-1973)."""
        return _nullfit_call(self.lib.glue_ref_nullfit, null, pmass, fracfit, doexpfit)

    def tree_from_newick(self, path, names, rootatmid=True):
        """FastTree's Newick -> the rooted tree of Tree_CalculateExtFromMSA (src/msatree.c:84-95): the shim's Newick reader, then the
        reference's own Tree_ReorderTaxaAccordingMSA + Tree_RootAtMidPoint."""
        n = len(names)
        arr = (C.c_char_p * n)(*[x.encode() for x in names])
        left, right, parent = (np.zeros(n - 1, np.int32) for _ in range(3))
        ld, rd = np.zeros(n - 1), np.zeros(n - 1)
        g = self.lib
        g.glue_ref_tree_from_newick.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, _ip, _ip, _ip, _dp, _dp]
        rc = g.glue_ref_tree_from_newick(path.encode(), n, arr, 1 if rootatmid else 0, _i(left), _i(right), _i(parent), _d(ld), _d(rd))
        if rc != 0:
            raise RuntimeError(f"tree_from_newick failed with status {rc}")
        return Tree(left, right, parent, ld, rd)

    def _evalue_args(self, null):
        geom = np.array([null.bmin, null.w, null.xmax, null.phi])
        ig = np.array([null.nb, null.imin, null.imax, null.cmin], dtype=np.int32)
        return geom, ig

    def cov2evalue(self, cov, null, Nc=1):
        g = self.lib
        g.glue_ref_cov2evalue.restype = C.c_double
        g.glue_ref_cov2evalue.argtypes = [C.c_double, C.c_int, _dp, _ip, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), _dp]
        geom, ig = self._evalue_args(null)
        return g.glue_ref_cov2evalue(float(cov), Nc, _d(geom), _i(ig), null.Nc, null.No, null.obs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     _d(null.survfit))

    def evalue2cov(self, eval_thresh, null, Nc=1):
        g = self.lib
        g.glue_ref_evalue2cov.restype = C.c_double
        g.glue_ref_evalue2cov.argtypes = [C.c_double, C.c_int, _dp, _ip, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), _dp]
        geom, ig = self._evalue_args(null)
        return g.glue_ref_evalue2cov(float(eval_thresh), Nc, _d(geom), _i(ig), null.Nc, null.No, null.obs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     _d(null.survfit))


# ---------------------------------------------------------------------------------------------
# driving an implementation of the reference API (corr_* over struct mutual_s) through mi_glue.c
# ---------------------------------------------------------------------------------------------
def _bind_corr_api(lib):
    vp = C.c_void_p
    lib.glue_abc_rna.restype = vp
    lib.glue_msa_create.restype = vp
    lib.glue_msa_create.argtypes = [C.c_int, C.c_int, _u8p, _dp]
    lib.glue_msa_destroy.argtypes = [vp]
    lib.glue_allowpair_from.restype = vp
    lib.glue_allowpair_from.argtypes = [_dp]
    lib.glue_data_create.restype = vp
    lib.glue_data_create.argtypes = [vp, vp, C.c_int, C.c_double]
    lib.glue_data_destroy.argtypes = [vp]
    lib.glue_data_errbuf.restype = C.c_char_p
    lib.glue_data_errbuf.argtypes = [vp]
    lib.glue_mi_export.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
    lib.esl_dmatrix_Destroy.argtypes = [vp]
    lib.corr_Create.restype = vp
    lib.corr_Create.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    lib.corr_Reuse.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.corr_ReuseCOV.argtypes = [vp, C.c_int, C.c_int]
    lib.corr_Destroy.argtypes = [vp]
    lib.corr_Probs.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_double, C.c_int, C.c_char_p]
    for name in ("CHI", "OMES", "GT", "MI", "MIr", "MIg", "CCF"):
        getattr(lib, "corr_Calculate" + name).argtypes = [C.c_int, vp]
    for name in ("RAF", "RAFS"):
        getattr(lib, "corr_Calculate" + name).argtypes = [C.c_int, vp, vp]
    lib.corr_CalculateCOVCorrected.argtypes = [C.c_int, vp, C.c_int]
    return lib


def corr_api_scan(lib, msa, wgt, stat=GT, covclass=C16, actype=APC, allowpair=None, tol=1e-6,
                  nseqthresh=8, alenthresh=50, want_pp=True):
    """What cov_Calculate does for one alignment (src/covariation.c:78-258), through the API:
    corr_Create -> corr_Reuse -> corr_Probs -> corr_Calculate<stat> -> corr_CalculateCOVCorrected."""
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    wgt = np.ascontiguousarray(wgt, dtype=np.float64)
    N, L = msa.shape
    ap = np.ascontiguousarray(ALLOWPAIR_WC_GU if allowpair is None else allowpair, dtype=np.float64)
    m = lib.glue_msa_create(N, L, _u8(msa), _d(wgt))
    mi = lib.corr_Create(L, N, 0, nseqthresh, alenthresh, lib.glue_abc_rna(), covclass)
    assert mi
    apm = lib.glue_allowpair_from(_d(ap))
    covtype = stat + {APC: 1, ASC: 2, NOCORR: 0}[actype]
    data = lib.glue_data_create(mi, apm, covtype, tol)
    errbuf = C.create_string_buffer(256)
    try:
        st = lib.corr_Reuse(mi, 0, covtype, covclass)
        assert st == 0
        if stat not in (RAF, RAFS):
            st = lib.corr_Probs(None, m, None, None, mi, 0, tol, 0, errbuf)
            if st != 0:
                raise RuntimeError(f"corr_Probs: {errbuf.value!r}")
        name = STAT_NAMES[stat]
        fn = getattr(lib, "corr_Calculate" + name)
        st = fn(covclass, data, m) if stat in (RAF, RAFS) else fn(covclass, data)
        if st != 0:
            raise RuntimeError(f"corr_Calculate{name}: {lib.glue_data_errbuf(data)!r}")
        if actype in (APC, ASC):
            st = lib.corr_CalculateCOVCorrected(actype, data, 0)
            if st != 0:
                raise RuntimeError(f"corr_CalculateCOVCorrected: {lib.glue_data_errbuf(data)!r}")
        out = dict(pp=np.zeros((L, L, 16)) if want_pp else None, pm=np.zeros((L, 4)), ps=np.zeros((L, 5)),
                   nseff=np.zeros((L, L)), ngap=np.zeros((L, L)), cov=np.zeros((L, L)))
        mm = np.zeros(2)
        tc = np.zeros(2, dtype=np.int32)
        lib.glue_mi_export(mi, _d(out["pp"]), _d(out["pm"]), _d(out["ps"]), _d(out["nseff"]), _d(out["ngap"]),
                           _d(out["cov"]), _d(mm), _i(tc))
        out.update(mincov=mm[0], maxcov=mm[1], type=int(tc[0]), covclass=int(tc[1]))
        return out
    finally:
        lib.glue_data_destroy(data)
        lib.esl_dmatrix_Destroy(apm)
        lib.corr_Destroy(mi)
        lib.glue_msa_destroy(m)


# ---------------------------------------------------------------------------------------------
# numpy restatement of the host-side preprocessing
# ---------------------------------------------------------------------------------------------
_RNA_CODE = {c: i for i, c in enumerate("ACGU")}
_RNA_CODE.update({"T": 3, "-": 4, ".": 4, "_": 4, "~": 17, "*": 16, "N": 15, "X": 15})
_RNA_CODE.update({c: 5 + i for i, c in enumerate("RYMKSWHBVD")})


def read_stockholm(path):
    """First alignment of a Stockholm file -> (names, digital residues uint8 [N][alen], ss_cons str|None)."""
    names, rows, ss = [], {}, []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("//"):
                break
            if not line.strip() or line.startswith("# STOCKHOLM"):
                continue
            if line.startswith("#=GC SS_cons"):
                ss.append(line.split()[2])
                continue
            if line.startswith("#"):
                continue
            name, seq = line.split()[:2]
            if name not in rows:
                rows[name] = []
                names.append(name)
            rows[name].append(seq)
    seqs = ["".join(rows[n]) for n in names]
    alen = len(seqs[0])
    ax = np.empty((len(seqs), alen), dtype=np.uint8)
    for s, seq in enumerate(seqs):
        assert len(seq) == alen
        ax[s] = [_RNA_CODE[c.upper()] for c in seq]
    return names, ax, ("".join(ss) if ss else None)


def remove_gap_columns(ax, wgt=None, gapthresh=0.75):
    """msamanip_RemoveGapColumns, src/msamanip.c:486-500: keep a column iff the weighted residue
    fraction r/(r+gap) is >= 1-gapthresh and > 0.  Returns (filtered ax, kept column indices)."""
    N, L = ax.shape
    w = np.ones(N) if wgt is None else wgt
    is_res = (ax < 4) | ((ax > 4) & (ax < 16))         # esl_abc_XIsResidue
    is_gap = (ax == 4)                                 # esl_abc_XIsGap; missing data (17) and '*' (16) count on neither side (:497)
    r = (w[:, None] * is_res).sum(0)
    tot = (w[:, None] * (is_res | is_gap)).sum(0)
    frac = np.where(tot > 0, r / np.maximum(tot, 1e-300), 0.0)
    keep = np.nonzero((frac >= 1.0 - gapthresh) & (frac > 0))[0]
    return np.ascontiguousarray(ax[:, keep]), keep


def degen_to_N(ax):
    """msamanip_ConvertDegen2N + ConvertMissingNonresidue2Gap, src/R-scape.c:1855-1856."""
    out = ax.copy()
    out[(ax > 4) & (ax < 15)] = 15
    out[ax >= 16] = 4
    return out


def weights_gsc(ax):
    """esl_msaweight_GSC as restated in SURVEY 9.7: 1-pid distances, UPGMA, Gerstein/Sonnhammer/Chothia."""
    N, L = ax.shape
    if N == 1:
        return np.ones(1)
    canon = ax < 4
    lens = canon.sum(1)
    D = np.zeros((N, N))
    for a in range(N):
        both = canon[a][None, :] & canon
        same = (both & (ax[a][None, :] == ax)).sum(1)
        denom = np.minimum(lens[a], lens)
        pid = np.where(denom > 0, same / np.maximum(denom, 1), 0.0)
        D[a] = 1.0 - pid
    np.fill_diagonal(D, 0.0)
    # UPGMA with Easel's bookkeeping (first strict minimum, swap-to-end)
    nn = N - 1
    left = np.zeros(nn, int)
    right = np.zeros(nn, int)
    ld = np.zeros(nn)
    rd = np.zeros(nn)
    height = np.zeros(nn)
    idx = [-s for s in range(N)]
    nin = [1] * N
    Dm = D.copy()
    for M in range(N, 1, -1):
        sub = Dm[:M, :M]
        iu = np.triu_indices(M, 1)
        k = int(np.argmin(sub[iu]))                 # row-major first minimum
        i, j = int(iu[0][k]), int(iu[1][k])
        minD = sub[i, j]
        v = M - 2
        left[v], right[v] = idx[i], idx[j]
        height[v] = minD / 2.0
        ld[v] = height[v] - (height[idx[i]] if idx[i] > 0 else 0.0)
        rd[v] = height[v] - (height[idx[j]] if idx[j] > 0 else 0.0)
        # swap j -> M-1, then i -> M-2
        for a, b in ((j, M - 1), (i, M - 2)):
            if a != b:
                Dm[[a, b], :] = Dm[[b, a], :]
                Dm[:, [a, b]] = Dm[:, [b, a]]
                idx[a], idx[b] = idx[b], idx[a]
                nin[a], nin[b] = nin[b], nin[a]
            if (a, b) == (j, M - 1) and i == M - 1:
                i = j                                # i was moved by the first swap
        i, j = M - 2, M - 1
        tot = nin[i] + nin[j]
        Dm[i, :M] = (nin[i] * Dm[i, :M] + nin[j] * Dm[j, :M]) / tot
        Dm[:M, i] = Dm[i, :M]
        Dm[i, i] = 0.0
        nin[i] = tot
        idx[i] = v
    # GSC weights on the tree
    x = np.zeros(nn)
    for v in range(nn - 1, -1, -1):
        x[v] = ld[v] + rd[v]
        if left[v] > 0:
            x[v] += x[left[v]]
        if right[v] > 0:
            x[v] += x[right[v]]
    csize = np.zeros(nn, int)
    for v in range(nn - 1, -1, -1):
        csize[v] = (csize[left[v]] if left[v] > 0 else 1) + (csize[right[v]] if right[v] > 0 else 1)
    wgt = np.zeros(N)
    x[0] = 0.0
    for v in range(nn):
        lw = ld[v] + (x[left[v]] if left[v] > 0 else 0.0)
        rw = rd[v] + (x[right[v]] if right[v] > 0 else 0.0)
        if lw + rw == 0.0:
            ls = csize[left[v]] if left[v] > 0 else 1
            rs = csize[right[v]] if right[v] > 0 else 1
            lx = x[v] * ls / (ls + rs)
            rx = x[v] * rs / (ls + rs)
        else:
            lx = x[v] * lw / (lw + rw)
            rx = x[v] * rw / (lw + rw)
        if left[v] > 0:
            x[left[v]] = lx + ld[v]
        else:
            wgt[-left[v]] = lx + ld[v]
        if right[v] > 0:
            x[right[v]] = rx + rd[v]
        else:
            wgt[-right[v]] = rx + rd[v]
    s = wgt.sum()
    return wgt * (N / s) if s > 0 else np.ones(N)


def pair_identity(ax, a, b):
    """esl_dst_XPairId: identical canonical positions / min(canonical lengths) (0 if that is 0)."""
    ca, cb = ax[a] < 4, ax[b] < 4
    same = int((ca & cb & (ax[a] == ax[b])).sum())
    ln = min(int(ca.sum()), int(cb.sum()))
    return same / ln if ln > 0 else 0.0


def average_id(ax, max_comparisons=10000, pairs=None):
    """esl_dst_XAverageId (src/msamanip.c:1967): exhaustive when N(N-1)/2 <= max_comparisons; else over `pairs` (the sampled list)."""
    N = ax.shape[0]
    if N <= 1:
        return 1.0
    if pairs is None:
        assert N * (N - 1) // 2 <= max_comparisons
        pairs = [(i, j) for i in range(N) for j in range(i + 1, N)]
    s = 0.0
    for i, j in pairs:
        s += pair_identity(ax, i, j)
    return s / len(pairs)


def weights_pb(ax):
    """esl_msaweight_PB (Henikoff position-based) as restated in SURVEY 9.7."""
    N, L = ax.shape
    canon = ax < 4
    w = np.zeros(N)
    for a in range(4):
        is_a = ax == a
        n_a = is_a.sum(0)                                           # [L]
        r = sum(((ax == b).sum(0) > 0).astype(int) for b in range(4))
        contrib = np.where(is_a, 1.0 / np.maximum(r * n_a, 1)[None, :], 0.0)
        w += contrib.sum(1)
    rlen = canon.sum(1)
    w = np.where(rlen > 0, w / np.maximum(rlen, 1), 0.0)
    s = w.sum()
    return w * (N / s) if s > 0 else np.ones(N)
