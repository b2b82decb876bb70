/* rscape_b200.h -- C-ABI of the B200-native covariation scan (librscape_b200.so).
 *
 * Plain C: pointers and sizes only, no CUDA or torch types in any signature (a CUDA stream crosses
 * as void*).  This is the layer the reference's covariation API would bind to; the functions a
 * maintainer actually calls keep the reference's own names and signatures and live one level up in
 * include/rscape_b200_host.h (corr_Create / corr_Probs / corr_Calculate* / corr_CalculateCOVCorrected
 * over struct mutual_s, src/correlators.h:445-482), implemented in r-scape_b200/host/ on top of this.
 *
 * Each entry point names the reference code it replaces.  All return 0 on success; on failure a
 * message is available from rsb_error() (the host layer copies it into the caller's errbuf and
 * returns eslFAIL, the reference's error convention, src/correlators.c:72-88).
 *
 * There is no CPU fallback: every call needs a CUDA device of compute capability 10.x.
 *
 * Matrix layouts on the host side follow the reference: residues uint8 [nseq][row_stride] with the
 * digital RNA codes A0 C1 G2 U3 gap4 N15 (ESL_MSA ax[s]+1); cov/nseff/ngap double [L][L];
 * pp double [L][L][16]; pm double [L][4]; ps double [L][5].
 */
#ifndef RSCAPE_B200_INCLUDED
#define RSCAPE_B200_INCLUDED
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsb_ctx rsb_ctx;

/* statistic = base COVTYPE (src/correlators.h:43-87), class = COVCLASS (:36-41), correction = ACTYPE (:89-92) */
enum { RSB_STAT_CHI = 0, RSB_STAT_GT = 3, RSB_STAT_MI = 6, RSB_STAT_MIr = 9, RSB_STAT_MIg = 12, RSB_STAT_OMES = 15,
       RSB_STAT_RAF = 18, RSB_STAT_RAFS = 21, RSB_STAT_CCF = 24 };
enum { RSB_CLASS_C16 = 0, RSB_CLASS_C2 = 1, RSB_CLASS_CWC = 2, RSB_CLASS_CSELECT = 3 };
enum { RSB_CORR_APC = 0, RSB_CORR_ASC = 1, RSB_CORR_NONE = 2 };

/* ---- lifetime ------------------------------------------------------------------------------- */
/* device: CUDA ordinal.  stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL
 * for a stream owned by the context. */
int         rsb_create(int device, void *stream, rsb_ctx **out);
/* number of CUDA devices visible to the process (0 if none) */
int         rsb_device_count(void);
void        rsb_destroy(rsb_ctx *ctx);
const char *rsb_error(const rsb_ctx *ctx);
/* last error of a failed rsb_create (no context exists yet) */
const char *rsb_create_error(void);

/* Shape of the alignments to be scanned and how many of them are in flight at once (replicate slots).
 * nslices: number of 8-bit digit slices S (1..6) of the fixed-point weights, 0 = choose from the weights (1 if they
 * are all small integers, else 4: ~40-bit weights, see rsb_set_weights).  Replaces corr_Create's allocations, src/correlators.c:1161-1219. */
int rsb_configure(rsb_ctx *ctx, int nseq, int alen, int max_replicates, int nslices);

/* Sequence weights (host, double[nseq]); NULL = all 1.  Quantised to fixed point wq = u V ~ w 2^q with an 8-bit
 * multiplier u (carried by the one-hot operand) and S base-256 digits of V (the weighted operand), so S slices give
 * ~8(S+1)-bit weights; every count is then exact integer arithmetic on wq.
 * The same weights serve the input alignment and every null (src/R-scape.c:1668). */
int rsb_set_weights(rsb_ctx *ctx, const double *wgt);
/* the quantisation actually used: wq[s] (may be NULL), q, S */
int rsb_get_quantisation(rsb_ctx *ctx, int64_t *wq, int *q, int *nslices);
/* Largest |wq_s 2^-q - w_s| over the sequences (weight units) and log2(max weight / that error): how faithfully the
 * fixed-point weights follow the double weights of msa->wgt (src/correlators.c:1716 reads them as double). */
int rsb_get_quantisation_error(rsb_ctx *ctx, double *max_abs_err, double *effective_bits);
/* Mixed precision (the "split path with a stated bound" of the weighted counts): contract the NULL alignments (rsb_null_width*,
 * rsb_null_hist*) with nslices <= S base-256 digits of the same weights, the input alignment keeps all S.  The nulls only feed
 * the score histogram whose tail gives the E-values (src/R-scape.c:1650-1697); with 2 slices their weights carry ~21 bits and
 * a null's scores move by at most 2e-4 max(1,|score|), under 1e-2 of the histogram's bin width (bound measured in tests/test_gpu_mixed.py).
 * 0 = off (default): one set of weights everywhere.  Takes effect at the next rsb_set_weights. */
int rsb_set_null_slices(rsb_ctx *ctx, int nslices);
/* the quantisation the nulls are scored with (= rsb_get_quantisation / _error when rsb_set_null_slices is off); any pointer may be NULL */
int rsb_get_null_quantisation(rsb_ctx *ctx, int64_t *wq, int *q, int *nslices, double *max_abs_err, double *effective_bits);

/* ---- one alignment: the corr_* sequence of cov_Calculate (src/covariation.c:78-258) ----------- */
/* corr_Probs (src/correlators.c:1424): counts -> pp, nseff, ngap, ps, pm.  Host outputs may be NULL.
 * on_device != 0: msa is a device pointer. */
int rsb_probs(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, double tol,
              double *pp, double *pm, double *ps, double *nseff, double *ngap);
/* host copies of the state left by the last rsb_probs, without recomputing it (corr_NaivePS / corr_Marginals
 * called on their own, src/correlators.c:1320-1375) */
int rsb_fetch_probs(rsb_ctx *ctx, double *pp, double *pm, double *ps, double *nseff, double *ngap);
/* corr_Calculate{CHI,OMES,GT,MI,MIr,MIg,CCF} (:50-1061) on the state left by rsb_probs, or
 * corr_Calculate{RAF,RAFS} (:877-982) on msa (which is then required; unweighted).
 * covclass must already be resolved (no CSELECT).  allowpair: double[16], > 0 = allowed. */
int rsb_statistic(rsb_ctx *ctx, int stat, int covclass, const double *allowpair,
                  const uint8_t *msa, int64_t row_stride, int on_device,
                  double *cov, double *mincov, double *maxcov);
/* corr_CalculateCOVCorrected (:1064-1157) on the state left by rsb_statistic */
int rsb_correct(rsb_ctx *ctx, int actype, double *cov, double *mincov, double *maxcov);
/* the same correction applied to a raw matrix held by the host (cov is read and overwritten): the reference
 * corrects whatever mi->COV contains, e.g. Potts scores written on the host (src/covariation.c:100-106) */
int rsb_correct_host(rsb_ctx *ctx, int actype, double *cov, double *mincov, double *maxcov);
/* the three above back to back with no host round trip in between */
int rsb_scan(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device,
             int stat, int covclass, int actype, const double *allowpair, double tol,
             double *cov, double *mincov, double *maxcov,
             double *pp, double *pm, double *ps, double *nseff, double *ngap);
/* The score histograms of that scan, cov_SignificantPairs_Ranking (src/covariation.c:415-457): ha[b] counts every pair
 * i<j in bin b = ceil((max(x, bmin+w) - bmin)/w - 1), b < nb; with pairmask (uint8 [L][L], nonzero = pair belongs to the
 * structure set chosen by data->samplesize: contacts / base pairs / WC pairs) hb gets the flagged pairs and ht the others. */
int rsb_scan_hist(rsb_ctx *ctx, const uint8_t *pairmask, double w, double bmin, int nb, uint64_t *ha, uint64_t *hb, uint64_t *ht);
/* The PDB-distance rule of the histogram fill, src/covariation.c:421-427: pairs i<j with msa2pdb[i] >= 0, msa2pdb[j] >= 0 and
 * msa2pdb[j] - msa2pdb[i] < mind are left out of ha / hb / ht and of the null histogram (data->msa2pdb, data->clist->mind;
 * R-scape's defaults msa2pdb[i] = i, mind = 1 exclude nothing).  msa2pdb: int[alen] on the host, NULL = no exclusion.
 * Stays in force until changed or until rsb_configure.  Scores, min/max, E-values and the hit list are not affected. */
int rsb_set_pair_exclusion(rsb_ctx *ctx, const int *msa2pdb, int mind);
/* Replace the device copy of the score matrix by the host's (double [L][L], upper triangle read): the reference's ranking and
 * hit-list code reads whatever mi->COV holds (src/covariation.c:431, :845), which host code may have written or shifted
 * (e.g. Potts scores, shiftnonneg, src/correlators.c:1130-1134) after the scan. */
int rsb_load_scores(rsb_ctx *ctx, const double *cov);
/* What cov2evalue (src/covariation.c:2370-2400) reads of data->ranklist_null: the cumulative null histogram ha and the
 * fitted tail.  The tail FIT (esl_gam_FitCompleteBinned / esl_exp_FitCompleteBinned, :1915-1973) stays host code of the
 * caller; this struct carries its result. */
typedef struct rsb_nullfit {
  double          bmin, w;    /* ha->bmin, ha->w: bin b covers (bmin + b w, bmin + (b+1) w] */
  int             nb;         /* ha->nb */
  int             imin, imax; /* lowest / highest non-empty bin */
  double          xmax;       /* ha->xmax: the largest null score */
  double          phi;        /* ha->phi: censoring point of the fit (read only when survfit != NULL) */
  uint64_t        Nc;         /* ha->Nc: number of null scores */
  const uint64_t *obs;        /* ha->obs, [nb] */
  const double   *survfit;    /* ranklist_null->survfit, [2 nb] (cov_histogram_SetSurvFitTail, :1677-1699), or NULL */
} rsb_nullfit;
/* E-values and the significant-pair list of the scan left by rsb_scan / rsb_correct (or, on a sharded pair grid, by
 * rsb_sharded_correct with bit 0 of mode set): the per-pair loop of cov_CreateHitList, src/covariation.c:828-910.
 * For every pair i<j: pval = cov2evalue(score, 1, null), E = pval * Nb if pairmask[i][j] (pair of the given structure,
 * by data->samplesize), else pval * Nt -- or pval * expBP while fewer than expBP hits are listed (expBP > 0, :852);
 * hit iff E < thresh, every pair if thresh > 1000 (MAX_EVAL).  eval (double [L][L], may be NULL) receives mi->Eval: both
 * triangles, +inf on the diagonal.  Hits come back in the reference's row-major order; *nhit is their total number.  When
 * *nhit <= cap the arrays hold all of them; when *nhit > cap their contents are UNSPECIFIED (some subset of the hits, not a prefix)
 * and the call must be repeated with cap >= *nhit, as cov_CreateHitList_b200 does (hit_sc / hit_eval / hit_pval may be NULL).
 * On a sharded pair grid a rank lists the pairs of the rows it owns (entries of other rows in eval are 0), and the caller
 * concatenates the ranks' lists: the only data besides the histograms that crosses GPUs. */
int rsb_scan_hits(rsb_ctx *ctx, const rsb_nullfit *null, const uint8_t *pairmask, uint64_t Nb, uint64_t Nt, int expBP, double thresh,
                  double *eval, int64_t cap, int64_t *hit_i, int64_t *hit_j, double *hit_sc, double *hit_eval, double *hit_pval,
                  int64_t *nhit);
/* fixed-point counts of the last rsb_probs/rsb_scan: int64 [16][L][L], upper triangle (parity tests) */
int rsb_get_counts(rsb_ctx *ctx, int64_t *counts);
/* the same counts recomputed by the direct verification kernel (no tensor cores); tests only */
int rsb_get_counts_direct(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int64_t *counts);

/* ---- the pair grid of ONE scan sharded over ranks (LSU-scale L) -----------------------------------------
 * Rank `rank` of `world` owns the 32-column row blocks ib with ib % world == rank and computes counts, statistic, correction
 * and histogram for the pairs (i,j), i<j, whose row i it owns.  The operand planes are replicated, so the contraction needs
 * no exchange; what all pairs feed -- the marginal sums, the APC row sums -- leaves each phase as a small host vector that
 * the caller sums over ranks (one all-reduce each) before the next phase:
 *     rsb_set_shard; rsb_set_weights;
 *     rsb_sharded_counts    -> marg_sums[L][4]   (all-reduce SUM)      corr_Probs      src/correlators.c:1424
 *     rsb_sharded_statistic -> cov_sums[L+4]     (SUM on [0..L], MIN on [L+1], MAX on [L+2])   corr_Calculate* :50-874
 *     rsb_sharded_correct   -> corrected scores of the owned rows (0 elsewhere: SUM assembles the upper triangle), local
 *                              min/max, histogram (rsb_hist_read + all-reduce SUM)   corr_CalculateCOVCorrected :1064
 * mode of rsb_sharded_correct: bit 0 = write the corrected scores (needed for cov), bit 1 = add them to the histogram. */
int rsb_set_shard(rsb_ctx *ctx, int rank, int world);
int rsb_sharded_counts(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, double tol, double *marg_sums);
/* the same on pool entry `rep` (a null generated on the device; every rank generates the same replicate ids) */
int rsb_sharded_counts_pool(rsb_ctx *ctx, int rep, double tol, double *marg_sums);
int rsb_sharded_statistic(rsb_ctx *ctx, const double *marg_sums, double tol, int stat, int covclass, const double *allowpair, double *cov_sums);
int rsb_sharded_correct(rsb_ctx *ctx, const double *cov_sums, int actype, int mode, double w, double bmin, double *cov, double *minmax);

/* ---- communicator over the GPUs of the job (NCCL, bound at run time) --------------------------------------
 * One process per GPU: rank 0 calls rsb_comm_id, ships the 128 bytes to the other ranks by any means (the bench uses
 * torch.distributed's store), every rank calls rsb_comm_init.  One process driving several GPUs: rsb_comm_init_all over its
 * contexts (one per device).  With a communicator
 *   - rsb_hist_allreduce sums the cumulative null histograms of all ranks on the device (null_add2cumranklist across ranks,
 *     src/R-scape.c:1565-1612): the only exchange when the NULL REPLICATES are dealt to the ranks;
 *   - when the PAIR GRID is sharded as well (rsb_set_shard with world = the communicator's size), the null loop
 *     (rsb_null_hist / _pool, rsb_null_width / _pool) and rsb_sharded_scan all-reduce the marginal sums [L][4], the APC row
 *     sums [L+1] and the score range on the device, inside the pipeline: no host round trip per scan. */
int rsb_comm_id(uint8_t *id128);
int rsb_comm_init(rsb_ctx *ctx, const uint8_t *id128, int nranks, int rank);
int rsb_comm_init_all(rsb_ctx **ctxs, int n);
int rsb_comm_destroy(rsb_ctx *ctx);
/* pool entries [first_rep, first_rep + nrep) generated by rank `root` made resident on every rank (ncclBroadcast over NVLink, on the
 * generation stream; scans of those entries wait for it).  With a sharded pair grid every rank scans every null: each rank generates
 * its share (rsb_null_fitch_shuffle with first_id = first_rep = the block's first replicate) and every rank calls this once per
 * block, in the same order. */
int rsb_pool_broadcast(rsb_ctx *ctx, int first_rep, int nrep, int root);
/* The per-scan vectors of a sharded pair grid (marginal sums [L][4], APC row sums [L+4], score range) are summed over the ranks
 * by a one-shot kernel over NVLink peer memory (csrc/peer_reduce.cu) when the ranks can map each other's memory
 * (cudaIpc between processes, peer access inside one), else by ncclAllReduce; RSCAPE_B200_PEER_REDUCE=0 forces NCCL.
 * *peer_path = 1 when the kernel is in use, *reductions = all-reduces it has done so far. */
int rsb_comm_info(rsb_ctx *ctx, int *nranks, int *rank, int *peer_path, int64_t *reductions);
/* self-test and latency of that all-reduce (collective): `count` doubles summed once and checked (*max_err, 0 expected), then
 * all-reduced `iters` times between two CUDA events (*us_per_allreduce, device time) */
int rsb_comm_selftest(rsb_ctx *ctx, int count, int iters, double *us_per_allreduce, double *max_err);
int rsb_hist_allreduce(rsb_ctx *ctx, int nb);
/* in place: lo = min over ranks, hi = max over ranks, aux_min (may be NULL) = min over ranks */
int rsb_comm_range(rsb_ctx *ctx, double *lo, double *hi, double *aux_min);
/* rsb_scan on a pair grid sharded over the communicator's ranks; cov [L][L] (may be NULL) is assembled on every rank */
int rsb_sharded_scan(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, int stat, int covclass, int actype,
                     const double *allowpair, double tol, double *cov, double *mincov, double *maxcov);

/* ---- null alignments: the loop body of null_rscape (src/R-scape.c:1650-1697) ------------------- */
/* calculate_width_histo (src/R-scape.c:1281-1371): scan one null, w = min(w_old, (max - max(bmin,min))/hpts). */
int rsb_null_width(rsb_ctx *ctx, const uint8_t *null0, int64_t row_stride, int on_device,
                   int stat, int covclass, int actype, const double *allowpair, double tol,
                   double w_old, double bmin, int hpts, double *w_out, double *mincov, double *maxcov);
/* run_rscape(RANSS) + null_add2cumranklist for nrep nulls [nrep][nseq][row_stride]: scores are added to the
 * context's cumulative histogram (bin b = ceil((max(x,bmin+w) - bmin)/w - 1)).  minmax: double[nrep][2] or NULL. */
int rsb_null_hist(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device,
                  int stat, int covclass, int actype, const double *allowpair, double tol,
                  double w, double bmin, double *minmax);
int rsb_hist_reset(rsb_ctx *ctx);
/* bins[0..nb_cap), number of scores added, highest non-empty bin (-1 if none) */
int rsb_hist_read(rsb_ctx *ctx, uint64_t *bins, int nb_cap, uint64_t *n_out, int *imax_out);
/* Device-to-device copy of the first nb bins between the cumulative histogram and a caller's device buffer (uint64[nb]),
 * to_library = 0 copies out, 1 copies in; returns when the copy is complete (the caller must have completed its own work on
 * device_buf before copying in).  Lets a multi-GPU driver sum the histograms of its
 * ranks with one NCCL all-reduce on the device (null_add2cumranklist across ranks, src/R-scape.c:1565-1612). */
int rsb_hist_exchange(rsb_ctx *ctx, void *device_buf, int nb, int to_library);
/* nseff / ngap of the last replicate scanned (quirk Q3: what .cov prints as nseff(%), src/power.c:94-95) */
int rsb_last_nseff(rsb_ctx *ctx, double *nseff, double *ngap);

/* ---- null generators on the device --------------------------------------------------------------- */
/* tree in Easel convention: nseq leaves, internal nodes 0..nseq-2 with parents before children,
 * child <= 0 is leaf -child.  Stays resident for the simulators. */
int rsb_set_tree(rsb_ctx *ctx, const int *left, const int *right, const int *parent, const double *ld, const double *rd);
/* The generators write into a device-resident pool of null alignments uint8 [nrep][nseq][alen] that the scan reads in
 * place, so a null alignment never has to exist on the host. */
int rsb_pool_reserve(rsb_ctx *ctx, int nrep);
/* cov_GenerateAlignment, ungapped noss path (src/cov_simulate.c:289-324,585-631,724-773): evolve every
 * (replicate, column) independently down the tree with P(t) = exp(tQ) (src/ratematrix.c:185-233),
 * Philox4x32-10 keyed by (seed, replicate id, column).  root: uint8[alen] residues 0..3.  gapmask: optional
 * uint8 [nseq][alen] alignment whose non-canonical cells are copied over the result.
 * Replicates with global ids [first_id, first_id+nrep) land in pool entries [first_rep, first_rep+nrep); the
 * residues depend on (seed, id) only, so ranks that generate the same id get the same alignment.  Both generators
 * return once their kernels are queued on the context's generation stream; calls that read the pool wait on the device
 * for the entries they use. */
int rsb_null_simulate(rsb_ctx *ctx, const double *Q, const uint8_t *root, const uint8_t *gapmask, int64_t gap_stride,
                      uint64_t seed, uint64_t first_id, int first_rep, int nrep);
/* default null of R-scape: Fitch ancestral reconstruction + one column permutation + per-branch
 * substitution re-placement (src/msatree.c:173-227,1700-1931; src/msamanip.c:1164-1233,1449-1780) */
int rsb_null_fitch_shuffle(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, uint64_t seed, uint64_t first_id, int first_rep, int nrep);
/* The same with an explicit list of global replicate ids (host, uint64[nrep]) instead of first_id, first_id+1, ...: a rank
 * of a multi-GPU run generates replicate 0 (for the width pass, src/R-scape.c:1281) and its own block in one call. */
int rsb_null_fitch_shuffle_ids(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, uint64_t seed, const uint64_t *ids,
                               int first_rep, int nrep);
/* calculate_width_histo / the null loop on pool entries */
int rsb_null_width_pool(rsb_ctx *ctx, int rep, int stat, int covclass, int actype, const double *allowpair, double tol,
                        double w_old, double bmin, int hpts, double *w_out, double *mincov, double *maxcov);
int rsb_null_hist_pool(rsb_ctx *ctx, int first_rep, int nrep, int stat, int covclass, int actype, const double *allowpair,
                       double tol, double w, double bmin, double *minmax);
/* Several (statistic, correction) combinations from ONE contraction per null -- the statistic sweep of BASELINE config 5.
 * cov_Calculate (src/covariation.c:100-258) computes the probabilities once (corr_Probs) and then dispatches on covtype; a sweep
 * over covtypes repeats everything.  Here each null is packed and contracted once, all requested statistics of
 * {CHI, OMES, GT, MI, MIr, MIg} are evaluated from the same count planes in one pass, and combination k = (stat[k], actype[k])
 * is corrected and added to its own cumulative histogram with its own bin width w[k] (w[k] <= 0: score range only, no histogram
 * -- the width pass).  covclass is shared.  minmax: double [ncombo][nrep][2] (may be NULL).  Every histogram equals the one
 * rsb_null_hist leaves for that combination alone.  RAF / RAFS use unit weights, i.e. another contraction: a call holds either
 * weighted statistics or {RAF, RAFS} x corrections (which then share one single-slice contraction per null), not both. */
int rsb_null_hist_multi(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device, int ncombo,
                        const int *stat, const int *actype, int covclass, const double *allowpair, double tol, const double *w, double bmin,
                        double *minmax);
int rsb_null_hist_multi_pool(rsb_ctx *ctx, int first_rep, int nrep, int ncombo, const int *stat, const int *actype, int covclass,
                             const double *allowpair, double tol, const double *w, double bmin, double *minmax);
/* histogram of combination `combo` (as rsb_hist_read); rsb_hist_reset_multi clears all of them */
int rsb_hist_read_multi(rsb_ctx *ctx, int combo, uint64_t *bins, int nb_cap, uint64_t *n_out, int *imax_out);
int rsb_hist_reset_multi(rsb_ctx *ctx);
/* copy pool entries to / from the host: uint8 [nrep][nseq][alen] (e.g. for --outnull, or host-made nulls scanned repeatedly) */
int rsb_pool_get(rsb_ctx *ctx, int first_rep, int nrep, uint8_t *out);
int rsb_pool_put(rsb_ctx *ctx, int first_rep, int nrep, const uint8_t *in);
/* parity tests: the internal-node rows of generator A for pool entries [first_rep, first_rep + nrep), uint8 [nrep][nseq-1][alen]:
 * which = 0 the Fitch reconstruction (Tree_FitchAlgorithmAncenstral's rows nseq.. of allmsa, src/msatree.c:173-227),
 * which = 1 the shuffled rows (msamanip_ShuffleTreeSubstitutions' allmsa, src/msamanip.c:1449-1531) */
int rsb_pool_get_internal(rsb_ctx *ctx, int which, int first_rep, int nrep, uint8_t *out);

/* ---- substitution counts over the tree (input of the power calculation) ------------------------------ */
/* Tree_Substitutions after its Fitch pass, src/msatree.c:1455-1540 (callers src/R-scape.c:2784-2840): the O(L^2 N) loops
 * over all column pairs and branches run as the unweighted pair contraction over one row per branch.
 * The context must be configured with nseq = 2 (ntaxa - 1) (one row per branch), alen = L and at least one replicate slot.
 * Tree in Easel convention (see rsb_set_tree; only left / right are read); leaves uint8 [ntaxa][leaf_stride] = the alignment,
 * internal uint8 [ntaxa-1][internal_stride] = the ancestral sequence of every internal node as reconstructed by
 * Tree_FitchAlgorithmAncenstral (rows ntaxa.. of its allmsa, src/msatree.c:173-227), which stays with the caller: its random
 * tie-breaking belongs to the caller's RNG stream.
 * nsubs int [L]; ndouble, njoin int [L][L] with the entries i<j filled and the rest 0, as the reference leaves them.
 * Any of the three may be NULL (the contraction is skipped when only nsubs is asked for). */
int rsb_tree_substitutions(rsb_ctx *ctx, int ntaxa, const int *left, const int *right, const uint8_t *leaves, int64_t leaf_stride,
                           const uint8_t *internal, int64_t internal_stride, int includegaps, int *nsubs, int *ndouble, int *njoin);

/* ---- alignment preprocessing: what defines the scanned matrix and its weights ----------------------------
 * Independent of rsb_configure (any shape); msa: uint8 [nseq][row_stride], host or device.
 * Easel is not vendored by the reference: these follow SURVEY 9.7's restatement of the Easel routines R-scape calls. */
/* The column test of msamanip_RemoveGapColumns (src/msamanip.c:486-500): useme[c] = 1 iff r > 0 and r / (r + gap) >= 1 - gapthresh,
 * r / gap = summed weights of the sequences holding a residue / a gap in column c (wgt NULL: all 1, counted as integers). */
int rsb_msa_gap_columns(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const double *wgt,
                        double gapthresh, uint8_t *useme);
/* the residues of the kept columns (struct_ColumnSubset): out uint8 [nseq][nkeep], host or device */
int rsb_msa_column_subset(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const uint8_t *useme,
                          uint8_t *out, int out_on_device, int *nkeep);
/* esl_msaweight_PB, Henikoff position-based weights normalised to sum nseq (R-scape's choice for nseq > 1000, src/R-scape.c:1555-1556) */
int rsb_msa_pb_weights(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, double *wgt);
/* esl_dst_XPairId: pid = identical canonical positions / min(canonical lengths).  pairs int [npairs][2] -> out[npairs] (the sampled or
 * exhaustive pair list of esl_dst_XAverageId, src/msamanip.c:1967); pairs NULL -> out double [nseq][nseq] = 1 - pid, the distance
 * matrix esl_msaweight_GSC builds its UPGMA tree from (nseq <= 1000). */
int rsb_msa_pair_identity(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const int *pairs, int64_t npairs,
                          double *out);

/* ---- pinned host buffers --------------------------------------------------------------------------- */
/* Page-lock / release a host buffer that is handed to the library repeatedly (mi->COV, mi->Eval, the pp slab, alignment rows):
 * copies then run at the PCIe rate instead of through the driver's pageable staging.  0 = registered, 1 = left pageable
 * (never an error of the path).  corr_Create registers the buffers of struct mutual_s it allocates; corr_Destroy releases them. */
int rsb_host_register(void *ptr, size_t bytes);
int rsb_host_unregister(void *ptr);

/* ---- instrumentation ------------------------------------------------------------------------------ */
/* kernels launched by this context so far; device milliseconds spent in the gram kernel and number of
 * gram launches since the last call with reset != 0 (CUDA events on the context's stream). */
int rsb_counters(rsb_ctx *ctx, int64_t *launches, double *gram_ms, int64_t *gram_launches, int reset);
/* enable (1) / disable (0) event timing around the gram kernel */
int rsb_profile_gram(rsb_ctx *ctx, int enable);
/* the same per operand geometry, as of the last rsb_counters call (before its reset): which = 0 the input alignment's weights,
 * 1 unit weights (RAF tables, substitution counts), 2 the nulls' own weights (rsb_set_null_slices); *nslices = its digit slices */
int rsb_counters_geometry(rsb_ctx *ctx, int which, double *gram_ms, int64_t *gram_launches, int *nslices);

#ifdef __cplusplus
}
#endif
#endif
