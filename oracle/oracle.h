/* oracle.h -- CPU restatement of R-scape's covariation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under r-scape_b200/ may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and only as the checker.
 *
 * Plain C over flat arrays (no Easel, no struct mutual_s): every function names the
 * reference file:line whose arithmetic it restates.  Parity of this restatement is pinned by
 *   (a) oracle/_ref  -- the reference's own src/correlators.c compiled unchanged against the
 *       Easel shim (tests/test_oracle_vs_ref.py), and
 *   (b) the tutorial transcript documentation/tutorial.tex:202-212 (tests/test_golden_tutorial.py).
 * Easel itself is not vendored by the reference, so the Easel-side routines (weights, histogram
 * tail fit) are "parity unpinned" beyond (b); see DESIGN.md.
 *
 * Layouts (all row-major, L = alignment length, K = 4):
 *   msa    uint8  [nseq][L]      digital residues A0 C1 G2 U3, gap 4, N 15 (anything >= 4 is "not canonical")
 *   wgt    double [nseq]
 *   pp     double [L][L][16]     pp[i][j][a*4+b]; both triangles filled (pp[j][i][b*4+a] mirror), diagonal 0
 *   nseff  double [L][L]         mirrored; ngap double [L][L] upper triangle only (quirk Q4)
 *   pm     double [L][4]         ps double [L][5]
 *   cov    double [L][L]         symmetric, diagonal -inf
 */
#ifndef RSB_ORACLE_INCLUDED
#define RSB_ORACLE_INCLUDED
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* statistic codes = base COVTYPE of src/correlators.h:43-87 */
enum { ORC_CHI = 0, ORC_GT = 3, ORC_MI = 6, ORC_MIr = 9, ORC_MIg = 12, ORC_OMES = 15, ORC_RAF = 18, ORC_RAFS = 21, ORC_CCF = 24 };
/* COVCLASS src/correlators.h:36-41 */
enum { ORC_C16 = 0, ORC_C2 = 1, ORC_CWC = 2, ORC_CSELECT = 3 };
/* ACTYPE src/correlators.h:89-92 (+ none) */
enum { ORC_APC = 0, ORC_ASC = 1, ORC_NOCORR = 2 };

int  orc_pair_probs(const uint8_t *msa, int nseq, int L, const double *wgt,
                    double *pp, double *nseff, double *ngap);
int  orc_pair_counts_fixed(const uint8_t *msa, int nseq, int L, const int64_t *wq,
                           int64_t *counts /* [L][L][16], upper triangle */);
int  orc_single_probs(const uint8_t *msa, int nseq, int L, const double *wgt, double *ps);
int  orc_marginals(const double *pp, const double *nseff, int L, double tol, double *pm);
int  orc_validate_probs(const double *pp, const double *pm, const double *ps, int L, double tol);
int  orc_resolve_class(int covclass, int nseq, int L, int nseqthresh, int alenthresh);
int  orc_statistic(int stat, int covclass, int L, const double *pp, const double *pm,
                   const double *nseff, const double *ngap, const double *allowpair /* [4][4] */,
                   double *cov, double *mincov, double *maxcov);
int  orc_raf(const uint8_t *msa, int nseq, int L, const double *allowpair,
             double *cov, double *mincov, double *maxcov);
int  orc_raf_from_counts(const uint8_t *msa, int nseq, int L, const double *allowpair,
                         double *cov, double *mincov, double *maxcov);
int  orc_rafs(const uint8_t *msa, int nseq, int L, const double *allowpair, int use_count_identity,
              double *cov, double *mincov, double *maxcov);
int  orc_correct(int actype, int L, double *cov, double *mincov, double *maxcov);

/* one whole scan = corr_Probs + corr_Calculate<stat> + corr_CalculateCOVCorrected as dispatched by
 * cov_Calculate (src/covariation.c:78-258).  Any output pointer may be NULL. */
int  orc_scan(const uint8_t *msa, int nseq, int L, const double *wgt,
              int stat, int covclass, int actype, const double *allowpair, double tol,
              double *cov, double *mincov, double *maxcov,
              double *pp, double *pm, double *ps, double *nseff, double *ngap);

/* loop-faithful timing variant of the pair counter (per-pair column gather + malloc, as
 * src/correlators.c:1696-1780); nthreads > 1 deals rows i to pthreads. */
double orc_time_pair_probs(const uint8_t *msa, int nseq, int L, const double *wgt, int row_stride, int nthreads,
                           double *checksum);

/* ---- score histogram (src/covariation.c:415-457 over the Easel "full" histogram) ---- */
typedef struct {
  double    bmin, bmax, w;
  int       nb;
  int       imin, imax;
  double    xmin, xmax;
  uint64_t  n, Nc, No;
  uint64_t *obs;
} ORC_HIST;

ORC_HIST *orc_hist_create(double bmin, double bmax, double w);
void      orc_hist_destroy(ORC_HIST *h);
int       orc_hist_score2bin(const ORC_HIST *h, double x);
int       orc_hist_add(ORC_HIST *h, double x);
/* fill from a covariation matrix: bmax = maxcov + 5w, value = max(cov, bmin + w), all i<j */
ORC_HIST *orc_hist_from_cov(const double *cov, int L, double maxcov, double bmin, double w, double tol);
/* null_add2cumranklist (src/R-scape.c:1565-1612) + cov_GrowRankList (src/covariation.c:683-736) */
int       orc_hist_accumulate(ORC_HIST **cum, const ORC_HIST *one);
/* histogram bin width from the first null: src/R-scape.c:1357-1360 */
double    orc_null_width(double w_old, double mincov, double maxcov, double bmin, int hpts, double tol);

/* ---- E-values and the significant-pair list (src/covariation.c:2370-2435, 828-910) ---- */
double    orc_cov2evalue(double cov, int Nc, const ORC_HIST *h, double phi, const double *survfit /* [2 nb] or NULL */);
double    orc_evalue2cov(double eval_thresh, int Nc, const ORC_HIST *h, int cmin, const double *survfit);
int64_t   orc_hitlist(const double *cov, int L, const ORC_HIST *null, double phi, const double *survfit, const uint8_t *pairmask,
                      uint64_t Nb, uint64_t Nt, int expBP, double thresh, double *eval, int64_t cap,
                      int64_t *hit_i, int64_t *hit_j, double *hit_sc, double *hit_eval, double *hit_pval);

/* ---- null alignment generators ---- */
/* tree in Easel convention: N leaves, nodes 0..N-2, root 0, child <= 0 means leaf -child */
typedef struct {
  int           N;
  const int    *left, *right, *parent;
  const double *ld, *rd;
} ORC_TREE;

typedef struct orc_rng_s ORC_RNG;        /* MT19937 with the shim's seeding (r-scape_b200/host/easel_shim.c) */
ORC_RNG *orc_rng_create(uint32_t seed);
void     orc_rng_destroy(ORC_RNG *r);
double   orc_rng_uniform(ORC_RNG *r);

/* generator B: cov_GenerateAlignment ungapped/noss path, src/cov_simulate.c:289-324,388-451,585-631,724-773 */
int  orc_ptime(const double *Q /* [4][4] */, double t, double *P /* [4][4] */);
int  orc_null_simulate(ORC_RNG *r, const ORC_TREE *T, const double *Q, const uint8_t *root, int L,
                       uint8_t *leaves /* [N][L] */, uint8_t *internal /* [N-1][L] or NULL */);
/* generator A: Fitch + column shuffle + per-branch substitution re-placement,
 * src/msatree.c:173-227,1700-1931 and src/msamanip.c:1164-1233,1449-1531,1597-1780 */
int  orc_null_fitch_shuffle(ORC_RNG *r, const ORC_TREE *T, const uint8_t *msa, int L,
                            uint8_t *shmsa /* [N][L] */, uint8_t *allmsa /* [2N-1][L] or NULL */, int *fitch_sc);

/* Tree_Substitutions after its Fitch call, src/msatree.c:1455-1540; all = [2N-1][L] as written by orc_null_fitch_shuffle */
int  orc_tree_substitutions(const ORC_TREE *T, const uint8_t *all, int L, int includegaps, int *nsubs, int *ndouble, int *njoin);

#ifdef __cplusplus
}
#endif
#endif
