/* oracle.c -- CPU restatement of R-scape's covariation hot path over flat arrays.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Each function cites the reference lines whose
 * arithmetic it follows; loop orders follow the reference wherever a floating-point sum's
 * order could matter, so that oracle/_ref (the reference's correlators.c compiled unchanged)
 * and this file agree to the last bit on the same inputs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <time.h>
#include <pthread.h>

#include "easel.h"      /* shim: Kahan sums, MT19937, expm -- the Easel-side semantics */
#include "oracle.h"

#define K4   4
#define K16 16
#define PP(pp, L, i, j)    ((pp) + (((size_t)(i) * (size_t)(L) + (size_t)(j)) * K16))
#define AT(m, L, i, j)     ((m)[(size_t)(i) * (size_t)(L) + (size_t)(j)])

/* esl_vec_DNorm semantics: Kahan sum; zero sum -> uniform */
static void
normalise(double *v, int n)
{
  double s = esl_vec_DSum(v, n);
  int    k;
  if (s != 0.0) for (k = 0; k < n; k++) v[k] /= s;
  else          for (k = 0; k < n; k++) v[k] = 1.0 / (double) n;
}

static int
is_probability_vector(const double *v, int n, double tol)
{
  double s = 0.0;
  int    k;
  for (k = 0; k < n; k++) {
    if (!isfinite(v[k]) || v[k] < 0.0 || v[k] > 1.0) return 0;
    s += v[k];
  }
  return fabs(s - 1.0) <= tol;
}

/* src/correlators.c:1664-1669 */
static int
allowed(int x, int y, const double *allowpair)
{
  return (x < K4 && y < K4 && allowpair[x * K4 + y] > 0.0);
}

/* ------------------------------------------------------------------------------------------
 * pair probabilities: mutual_naive_ppij, src/correlators.c:1696-1780 (GAPASCHAR 0 branch),
 * driven over all i<j by corr_NaivePP :1301-1317.
 */
int
orc_pair_probs(const uint8_t *msa, int nseq, int L, const double *wgt, double *pp, double *nseff, double *ngap)
{
  int i, j, s, k, a, b;

  memset(pp,    0, sizeof(double) * (size_t) L * (size_t) L * K16);   /* corr_Reuse :1237-1246 */
  memset(nseff, 0, sizeof(double) * (size_t) L * (size_t) L);
  memset(ngap,  0, sizeof(double) * (size_t) L * (size_t) L);

  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double *p  = PP(pp, L, i, j);
      double *pt = PP(pp, L, j, i);
      double  ne = 0.0, ng = 0.0;

      for (k = 0; k < K16; k++) p[k] = 1e-10;                         /* :1713 prior first, weights after */
      for (s = 0; s < nseq; s++) {
        int ri = msa[(size_t) s * L + i];
        int rj = msa[(size_t) s * L + j];
        if (ri < K4 && rj < K4) { ne += wgt[s]; p[ri * K4 + rj] += wgt[s]; }   /* :1747-1750 */
        else                      ng += wgt[s];                                   /* :1751-1753 */
      }
      normalise(p, K16);                                               /* :1758 */
      for (a = 0; a < K4; a++)
        for (b = 0; b < K4; b++) pt[b * K4 + a] = p[a * K4 + b];        /* :1761-1763 */
      AT(nseff, L, i, j) = AT(nseff, L, j, i) = ne;                    /* :1765; ngap is not mirrored */
      AT(ngap,  L, i, j) = ng;
    }
  return 0;
}

/* Integer version used to check the GPU's fixed-point counts bit for bit:
 * counts[i][j][a*4+b] = sum_s wq[s] [x_si=a][x_sj=b] for i<j, wq = round(w * 2^q) as int64. */
int
orc_pair_counts_fixed(const uint8_t *msa, int nseq, int L, const int64_t *wq, int64_t *counts)
{
  int i, j, s;
  memset(counts, 0, sizeof(int64_t) * (size_t) L * (size_t) L * K16);
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      int64_t *c = counts + ((size_t) i * L + j) * K16;
      for (s = 0; s < nseq; s++) {
        int ri = msa[(size_t) s * L + i];
        int rj = msa[(size_t) s * L + j];
        if (ri < K4 && rj < K4) c[ri * K4 + rj] += wq[s];
      }
    }
  return 0;
}

/* single-column probabilities: mutual_naive_psi, src/correlators.c:1783-1817 */
int
orc_single_probs(const uint8_t *msa, int nseq, int L, const double *wgt, double *ps)
{
  int i, s, k;
  for (i = 0; i < L; i++) {
    double *p = ps + (size_t) i * 5;
    for (k = 0; k < 5; k++) p[k] = 1e-5;                              /* :1792 */
    for (s = 0; s < nseq; s++) {
      int r = msa[(size_t) s * L + i];
      if (r < 5) p[r] += wgt[s];                                      /* :1801 gaps count, N does not */
    }
    normalise(p, 5);
  }
  return 0;
}

/* corr_Marginals, src/correlators.c:1338-1375.  Sum order: x outer, j middle, y inner. */
int
orc_marginals(const double *pp, const double *nseff, int L, double tol, double *pm)
{
  int i, j, x, y;
  for (i = 0; i < L; i++) {
    double *m = pm + (size_t) i * K4;
    for (x = 0; x < K4; x++) {
      double acc = 0.0;
      for (j = 0; j < L; j++)
        for (y = 0; y < K4; y++)
          if (AT(nseff, L, i, j) > 0) acc += PP(pp, L, i, j)[x * K4 + y];
      m[x] = acc;
    }
    normalise(m, K4);
    if (!is_probability_vector(m, K4, tol)) return 1;
  }
  return 0;
}

/* corr_ValidateProbs, src/correlators.c:1500-1545 */
int
orc_validate_probs(const double *pp, const double *pm, const double *ps, int L, double tol)
{
  int i, j;
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++)
      if (!is_probability_vector(PP(pp, L, i, j), K16, tol)) return 1;
  for (i = 0; i < L; i++) if (!is_probability_vector(pm + (size_t) i * K4, K4, tol)) return 2;
  if (ps) for (i = 0; i < L; i++) if (!is_probability_vector(ps + (size_t) i * 5, 5, tol)) return 3;
  return 0;
}

/* CSELECT rule shared by every corr_Calculate<X>: src/correlators.c:336 */
int
orc_resolve_class(int covclass, int nseq, int L, int nseqthresh, int alenthresh)
{
  if (covclass != ORC_CSELECT) return covclass;
  return (nseq <= nseqthresh || L <= alenthresh) ? ORC_C2 : ORC_C16;
}

/* ------------------------------------------------------------------------------------------
 * per-pair statistics.  One function per (statistic, class) cell of the reference:
 *   CHI  C16 :93-128   C2 :130-181        OMES C16 :227-260  C2 :262-314
 *   GT   C16 :361-395  CWC :398-434  C2 :437-490
 *   MI   C16 :536-566  C2 :568-614        MIr  C16 :660-694  C2 :696-745
 *   MIg  C16 :789-823  C2 :825-874
 */
typedef struct { double p_in, p_out, q_in, q_out; } pooled_t;

/* the two-class pooling every _C2 variant starts with (e.g. :463-474) */
static pooled_t
pool_two_classes(const double *p, const double *mi, const double *mj, const double *allowpair)
{
  pooled_t t = { 0.0, 0.0, 0.0, 0.0 };
  int x, y;
  for (x = 0; x < K4; x++)
    for (y = 0; y < K4; y++) {
      if (allowed(x, y, allowpair)) { t.p_in  += p[x * K4 + y]; t.q_in  += mi[x] * mj[y]; }
      else                          { t.p_out += p[x * K4 + y]; t.q_out += mi[x] * mj[y]; }
    }
  return t;
}

static double
stat_pair(int stat, int cls, const double *p, const double *mi, const double *mj,
          double ne, double ng, const double *allowpair)
{
  double v = 0.0, H = 0.0;
  int    x, y;

  if (cls == ORC_C2) {
    pooled_t t = pool_two_classes(p, mi, mj, allowpair);
    double exp_in = ne * t.q_in,  exp_out = ne * t.q_out;
    double obs_in = ne * t.p_in,  obs_out = ne * t.p_out;
    switch (stat) {
    case ORC_CHI:
      v += (exp_in  > 0.) ? (obs_in  - exp_in)  * (obs_in  - exp_in)  / exp_in  : 0.0;
      v += (exp_out > 0.) ? (obs_out - exp_out) * (obs_out - exp_out) / exp_out : 0.0;
      return v;
    case ORC_OMES:
      v += (exp_in  > 0.) ? (obs_in  - exp_in)  * (obs_in  - exp_in)  / ne : 0.0;
      v += (exp_out > 0.) ? (obs_out - exp_out) * (obs_out - exp_out) / ne : 0.0;
      return v;
    case ORC_GT:
      v += (exp_in  > 0. && obs_in  > 0.) ? obs_in  * log(obs_in  / exp_in)  : 0.0;
      v += (exp_out > 0. && obs_out > 0.) ? obs_out * log(obs_out / exp_out) : 0.0;
      return 2.0 * v;
    case ORC_MI:                                   /* guards p only: :606-607 */
      v += (t.p_in  > 0.) ? t.p_in  * (log(t.p_in)  - log(t.q_in))  : 0.0;
      v += (t.p_out > 0.) ? t.p_out * (log(t.p_out) - log(t.q_out)) : 0.0;
      return v;
    case ORC_MIg:                                  /* :865-869 */
      v += (t.p_in  > 0.) ? t.p_in  * (log(t.p_in)  - log(t.q_in))  : 0.0;
      v += (t.p_out > 0.) ? t.p_out * (log(t.p_out) - log(t.q_out)) : 0.0;
      v -= (ne > 0) ? ng / ne : 0.0;
      return v;
    case ORC_MIr:                                  /* guards p and q: :733-739 */
      H -= (t.p_in  > 0.) ? t.p_in  * log(t.p_in)  : 0.0;
      H -= (t.p_out > 0.) ? t.p_out * log(t.p_out) : 0.0;
      v += (t.p_in  > 0. && t.q_in  > 0.) ? t.p_in  * (log(t.p_in)  - log(t.q_in))  : 0.0;
      v += (t.p_out > 0. && t.q_out > 0.) ? t.p_out * (log(t.p_out) - log(t.q_out)) : 0.0;
      return (H > 1e-2) ? v / H : 0.0;
    }
    return NAN;
  }

  /* C16 and CWC (CWC exists for GT only and just skips the non-allowed cells, :420) */
  for (x = 0; x < K4; x++)
    for (y = 0; y < K4; y++) {
      double pxy = p[x * K4 + y];
      double ex  = ne * mi[x] * mj[y];
      double ob  = ne * pxy;
      if (cls == ORC_CWC && !allowed(x, y, allowpair)) continue;
      switch (stat) {
      case ORC_CHI:  v += (ex > 0.) ? (ob - ex) * (ob - ex) / ex : 0.0; break;
      case ORC_OMES: v += (ex > 0.) ? (ob - ex) * (ob - ex) / ne : 0.0; break;
      case ORC_GT:   v += (ex > 0. && ob > 0.) ? ob * log(ob / ex) : 0.0; break;
      case ORC_MIr:  H -= (pxy > 0.0) ? pxy * log(pxy) : 0.0;   /* fallthrough to the MI term */
      case ORC_MI:
      case ORC_MIg:  v += (pxy > 0.0 && mi[x] > 0.0 && mj[y] > 0.0) ? pxy * (log(pxy) - log(mi[x]) - log(mj[y])) : 0.0; break;
      }
    }
  if (stat == ORC_GT)  v *= 2.0;
  if (stat == ORC_MIg) v -= (ne > 0) ? ng / ne : 0.0;          /* :816 */
  if (stat == ORC_MIr) v  = (H > 1e-2) ? v / H : 0.0;          /* :687 */
  return v;
}

/* corr_CalculateCCF_C16, src/correlators.c:1011-1061 */
static int
ccf_all(int L, const double *pm, const double *nseff, double *cov, double *mn, double *mx)
{
  double meanp[K4];
  int    i, j, x, y;

  for (x = 0; x < K4; x++) {
    meanp[x] = 0.0;
    for (i = 0; i < L; i++)
      for (j = i + 1; j < L; j++) meanp[x] += AT(nseff, L, i, j) * pm[(size_t) i * K4 + x];
  }
  normalise(meanp, K4);
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double ne = AT(nseff, L, i, j), acc = 0.0, cc;
      for (x = 0; x < K4; x++)
        for (y = 0; y < K4; y++) {
          cc   = (ne * pm[(size_t) i * K4 + x] - meanp[x]) * (ne * pm[(size_t) j * K4 + y] - meanp[y]);
          acc += cc * cc;
        }
      acc = sqrt(acc);
      AT(cov, L, i, j) = AT(cov, L, j, i) = acc;
      if (acc < *mn) *mn = acc;
      if (acc > *mx) *mx = acc;
    }
  return 0;
}

/* corr_ReuseCOV state, src/correlators.c:1255-1268: COV = -inf everywhere, min = +inf, max = -inf */
static void
reset_cov(double *cov, int L, double *mn, double *mx)
{
  size_t k, tot = (size_t) L * (size_t) L;
  for (k = 0; k < tot; k++) cov[k] = -INFINITY;
  *mn = INFINITY;
  *mx = -INFINITY;
}

int
orc_statistic(int stat, int covclass, int L, const double *pp, const double *pm, const double *nseff,
              const double *ngap, const double *allowpair, double *cov, double *mincov, double *maxcov)
{
  int i, j;

  reset_cov(cov, L, mincov, maxcov);
  if (stat == ORC_CCF) return ccf_all(L, pm, nseff, cov, mincov, maxcov);
  if (covclass == ORC_CWC && stat != ORC_GT) return 1;          /* "CWC not implemented", e.g. :72 */
  if (covclass != ORC_C16 && covclass != ORC_C2 && covclass != ORC_CWC) return 1;

  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double v = stat_pair(stat, covclass, PP(pp, L, i, j), pm + (size_t) i * K4, pm + (size_t) j * K4,
                           AT(nseff, L, i, j), AT(ngap, L, i, j), allowpair);
      AT(cov, L, i, j) = AT(cov, L, j, i) = v;
      if (v < *mincov) *mincov = v;
      if (v > *maxcov) *maxcov = v;
    }
  return 0;
}

/* corr_CalculateRAF as written, O(P N^2): src/correlators.c:877-933.  Unweighted. */
int
orc_raf(const uint8_t *msa, int nseq, int L, const double *allowpair, double *cov, double *mincov, double *maxcov)
{
  const double psi = 1.0;
  int i, j, s1, s2;

  reset_cov(cov, L, mincov, maxcov);
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double cij = 0.0, qij = 0.0, v;
      for (s1 = 0; s1 < nseq; s1++) {
        int ai = msa[(size_t) s1 * L + i], aj = msa[(size_t) s1 * L + j];
        int ok1 = allowed(ai, aj, allowpair);
        if (!ok1) qij += 1.0;
        for (s2 = s1 + 1; s2 < nseq; s2++) {
          int bi = msa[(size_t) s2 * L + i], bj = msa[(size_t) s2 * L + j];
          if (ok1 && allowed(bi, bj, allowpair)) {
            if      (ai != bi && aj != bj) cij += 2.0;
            else if (ai != bi || aj != bj) cij += 1.0;
          }
        }
      }
      qij /= nseq;
      cij /= (nseq > 1) ? (double) nseq * ((double) nseq - 1.0) : 1.0;
      cij *= 2.0;
      v = cij - psi * qij;
      AT(cov, L, i, j) = AT(cov, L, j, i) = v;
      if (v < *mincov) *mincov = v;
      if (v > *maxcov) *maxcov = v;
    }
  return 0;
}

/* The same RAF through the unweighted 4x4 count table: with n_ab the number of
 * sequences showing (a,b) at (i,j),
 *    sum_{s1<s2} H = sum over unordered pairs of distinct allowed cells (ab),(cd) of n_ab n_cd ([a!=c]+[b!=d])
 *    qij           = (N - sum_allowed n_ab) / N.
 * All partial sums are integers below 2^53, so the result is bit-identical to orc_raf(). */
int
orc_raf_from_counts(const uint8_t *msa, int nseq, int L, const double *allowpair, double *cov, double *mincov, double *maxcov)
{
  const double psi = 1.0;
  int i, j, s, c1, c2;

  reset_cov(cov, L, mincov, maxcov);
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      int64_t n[K16], nallowed = 0, h = 0;
      double  cij, qij, v;
      memset(n, 0, sizeof(n));
      for (s = 0; s < nseq; s++) {
        int ri = msa[(size_t) s * L + i], rj = msa[(size_t) s * L + j];
        if (ri < K4 && rj < K4) n[ri * K4 + rj]++;
      }
      for (c1 = 0; c1 < K16; c1++) {
        if (!allowed(c1 / K4, c1 % K4, allowpair)) continue;
        nallowed += n[c1];
        for (c2 = c1 + 1; c2 < K16; c2++) {
          if (!allowed(c2 / K4, c2 % K4, allowpair)) continue;
          h += n[c1] * n[c2] * (((c1 / K4) != (c2 / K4)) + ((c1 % K4) != (c2 % K4)));
        }
      }
      qij  = (double) (nseq - nallowed);
      qij /= nseq;
      cij  = (double) h;
      cij /= (nseq > 1) ? (double) nseq * ((double) nseq - 1.0) : 1.0;
      cij *= 2.0;
      v = cij - psi * qij;
      AT(cov, L, i, j) = AT(cov, L, j, i) = v;
      if (v < *mincov) *mincov = v;
      if (v > *maxcov) *maxcov = v;
    }
  return 0;
}

/* corr_CalculateRAFS, src/correlators.c:936-982: 3-point anti-diagonal stencil over RAF */
int
orc_rafs(const uint8_t *msa, int nseq, int L, const double *allowpair, int use_count_identity,
         double *cov, double *mincov, double *maxcov)
{
  double *B = malloc(sizeof(double) * (size_t) L * (size_t) L);
  int     i, j, st;

  if (B == NULL) return 1;
  st = use_count_identity ? orc_raf_from_counts(msa, nseq, L, allowpair, B, mincov, maxcov)
                          : orc_raf            (msa, nseq, L, allowpair, B, mincov, maxcov);
  if (st != 0) { free(B); return st; }
  reset_cov(cov, L, mincov, maxcov);
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double v = 2.0 * AT(B, L, i, j);
      if (i > 0 && j < L - 1)              v += AT(B, L, i - 1, j + 1);
      if (j > 0 && i < L - 1 && i < j - 2) v += AT(B, L, i + 1, j - 1);
      v *= 0.25;
      AT(cov, L, i, j) = AT(cov, L, j, i) = v;
      if (v < *mincov) *mincov = v;
      if (v > *maxcov) *maxcov = v;
    }
  free(B);
  return 0;
}

/* corr_CalculateCOVCorrected, src/correlators.c:1064-1157 (shiftnonneg = FALSE) */
int
orc_correct(int actype, int L, double *cov, double *mincov, double *maxcov)
{
  double *raw, *rowmean, avg = 0.0;
  int     i, j, bad = 0;

  if (actype != ORC_APC && actype != ORC_ASC) return 1;
  raw     = malloc(sizeof(double) * (size_t) L * (size_t) L);
  rowmean = malloc(sizeof(double) * (size_t) L);
  if (!raw || !rowmean) { free(raw); free(rowmean); return 1; }
  memcpy(raw, cov, sizeof(double) * (size_t) L * (size_t) L);

  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) avg += AT(raw, L, i, j);               /* :1093-1096 */
  if (L > 1) avg /= (double) L * ((double) L - 1.);
  avg *= 2.;

  for (i = 0; i < L; i++) {                                            /* :1101-1108 */
    double acc = 0.0;
    for (j = 0; j < L; j++) if (j != i) acc += AT(raw, L, i, j);
    if (L > 1) acc /= (double) L - 1.;
    rowmean[i] = acc;
  }

  *mincov = INFINITY;
  *maxcov = -INFINITY;
  for (i = 0; i < L; i++) {
    AT(cov, L, i, i) = -INFINITY;                                       /* corr_ReuseCOV :1260 */
    for (j = 0; j < L; j++) {
      double v;
      if (i == j) continue;
      if (actype == ORC_APC) v = (avg != 0.0) ? AT(raw, L, i, j) - rowmean[i] * rowmean[j] / avg : 0.0;   /* :1118 */
      else                   v = AT(raw, L, i, j) - (rowmean[i] + rowmean[j] - avg);                      /* :1120 */
      if (isnan(v)) bad = 1;                                                                               /* :1124 */
      AT(cov, L, i, j) = v;
      if (v < *mincov) *mincov = v;
      if (v > *maxcov) *maxcov = v;
    }
  }
  free(raw);
  free(rowmean);
  return bad;
}

/* cov_Calculate's dispatch for one alignment, src/covariation.c:78-258 */
int
orc_scan(const uint8_t *msa, int nseq, int L, const double *wgt, int stat, int covclass, int actype,
         const double *allowpair, double tol, double *cov, double *mincov, double *maxcov,
         double *pp_out, double *pm_out, double *ps_out, double *nseff_out, double *ngap_out)
{
  size_t LL = (size_t) L * (size_t) L;
  double *pp = pp_out, *pm = pm_out, *ps = ps_out, *ne = nseff_out, *ng = ngap_out;
  double  mn = INFINITY, mx = -INFINITY;
  int     st = 0;

  if (!pp) pp = calloc(LL * K16, sizeof(double));
  if (!pm) pm = calloc((size_t) L * K4, sizeof(double));
  if (!ps) ps = calloc((size_t) L * 5, sizeof(double));
  if (!ne) ne = calloc(LL, sizeof(double));
  if (!ng) ng = calloc(LL, sizeof(double));
  if (!pp || !pm || !ps || !ne || !ng) { st = 100; goto DONE; }

  if (stat == ORC_RAF || stat == ORC_RAFS) {                     /* corr_Probs is skipped: :82-84 */
    if (pp_out) memset(pp, 0, LL * K16 * sizeof(double));
    if (nseff_out) memset(ne, 0, LL * sizeof(double));
    if (ngap_out)  memset(ng, 0, LL * sizeof(double));
    st = (stat == ORC_RAF) ? orc_raf_from_counts(msa, nseq, L, allowpair, cov, &mn, &mx)
                           : orc_rafs(msa, nseq, L, allowpair, 1, cov, &mn, &mx);
  } else {
    orc_pair_probs(msa, nseq, L, wgt, pp, ne, ng);               /* corr_Probs :1424-1456 */
    orc_single_probs(msa, nseq, L, wgt, ps);
    if ((st = orc_marginals(pp, ne, L, tol, pm)) != 0)        { st = 10 + st; goto DONE; }
    if ((st = orc_validate_probs(pp, pm, ps, L, tol)) != 0)   { st = 20 + st; goto DONE; }
    st = orc_statistic(stat, covclass, L, pp, pm, ne, ng, allowpair, cov, &mn, &mx);
  }
  if (st != 0) { st = 30 + st; goto DONE; }
  if (actype == ORC_APC || actype == ORC_ASC)
    if ((st = orc_correct(actype, L, cov, &mn, &mx)) != 0) st = 40 + st;

 DONE:
  if (mincov) *mincov = mn;
  if (maxcov) *maxcov = mx;
  if (!pp_out) free(pp);
  if (!pm_out) free(pm);
  if (!ps_out) free(ps);
  if (!nseff_out) free(ne);
  if (!ngap_out)  free(ng);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline timing: the pair counter exactly as the reference runs it -- per pair, two
 * malloc'd int columns gathered from the row-major alignment, branchy accumulate, normalise,
 * mirror (src/correlators.c:1716-1768) -- on every row_stride-th row i.  Returns seconds;
 * *checksum defeats dead-code elimination.  nthreads > 1 = rows dealt to pthreads ("generous" baseline;
 * the reference itself is single-threaded, src/Makefile:44-46).
 */
typedef struct {
  const uint8_t *msa; const double *wgt;
  int nseq, L, row_stride, tid, nthreads;
  double total;
} timing_job_t;

static void *
timing_worker(void *arg)
{
  timing_job_t *jb = (timing_job_t *) arg;
  const uint8_t *msa = jb->msa;
  const double  *wgt = jb->wgt;
  int    nseq = jb->nseq, L = jb->L, i, j, s, k, r = 0;
  double total = 0.0;

  for (i = 0; i < L - 1; i += jb->row_stride, r++) {
    if (r % jb->nthreads != jb->tid) continue;                /* rows dealt round-robin to threads */
    for (j = i + 1; j < L; j++) {
      double p[K16], ne = 0.0, ng = 0.0;
      int   *ci = malloc(sizeof(int) * (size_t) nseq);
      int   *cj = malloc(sizeof(int) * (size_t) nseq);
      for (k = 0; k < K16; k++) p[k] = 1e-10;
      for (s = 0; s < nseq; s++) { ci[s] = msa[(size_t) s * L + i]; cj[s] = msa[(size_t) s * L + j]; }
      for (s = 0; s < nseq; s++) {
        if (ci[s] < K4 && cj[s] < K4) { ne += wgt[s]; p[ci[s] * K4 + cj[s]] += wgt[s]; }
        else                            ng += wgt[s];
      }
      normalise(p, K16);
      total += p[5] + ne * 1e-9 + ng * 1e-12;
      free(ci);
      free(cj);
    }
  }
  jb->total = total;
  return NULL;
}

double
orc_time_pair_probs(const uint8_t *msa, int nseq, int L, const double *wgt, int row_stride, int nthreads, double *checksum)
{
  struct timespec t0, t1;
  timing_job_t   *jobs;
  pthread_t      *th;
  double          total = 0.0;
  int             t;

  if (row_stride < 1) row_stride = 1;
  if (nthreads   < 1) nthreads   = 1;
  jobs = calloc((size_t) nthreads, sizeof(timing_job_t));
  th   = calloc((size_t) nthreads, sizeof(pthread_t));
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (t = 0; t < nthreads; t++) {
    jobs[t].msa = msa; jobs[t].wgt = wgt; jobs[t].nseq = nseq; jobs[t].L = L;
    jobs[t].row_stride = row_stride; jobs[t].tid = t; jobs[t].nthreads = nthreads;
    if (nthreads > 1) pthread_create(&th[t], NULL, timing_worker, &jobs[t]);
    else              timing_worker(&jobs[t]);
  }
  for (t = 0; t < nthreads; t++) { if (nthreads > 1) pthread_join(th[t], NULL); total += jobs[t].total; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(jobs); free(th);
  if (checksum) *checksum = total;
  return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------
 * score histogram.  Geometry and bin rule of Easel's ESL_HISTOGRAM as used at
 * src/covariation.c:415-432,641-665: nb = (int)((bmax-bmin)/w), bin b covers
 * (bmin + b w, bmin + (b+1) w], b = ceil((x-bmin)/w - 1); growth above doubles the overshoot.
 */
ORC_HIST *
orc_hist_create(double bmin, double bmax, double w)
{
  ORC_HIST *h = calloc(1, sizeof(ORC_HIST));
  if (!h) return NULL;
  h->bmin = bmin; h->bmax = bmax; h->w = w;
  h->nb   = (int) ((bmax - bmin) / w);
  h->imin = h->nb;
  h->imax = -1;
  h->xmin = DBL_MAX;
  h->xmax = -DBL_MAX;
  h->obs  = calloc((size_t) (h->nb > 0 ? h->nb : 1), sizeof(uint64_t));
  if (!h->obs) { free(h); return NULL; }
  return h;
}

void orc_hist_destroy(ORC_HIST *h) { if (h) { free(h->obs); free(h); } }

int
orc_hist_score2bin(const ORC_HIST *h, double x)
{
  return (int) ceil(((x - h->bmin) / h->w) - 1.0);
}

int
orc_hist_add(ORC_HIST *h, double x)
{
  int b, k, grow;

  if (!isfinite(x)) return 1;
  b = orc_hist_score2bin(h, x);
  if (b < 0) {
    grow = -b * 2;
    h->obs = realloc(h->obs, sizeof(uint64_t) * (size_t) (h->nb + grow));
    memmove(h->obs + grow, h->obs, sizeof(uint64_t) * (size_t) h->nb);
    for (k = 0; k < grow; k++) h->obs[k] = 0;
    h->nb += grow; b += grow; h->bmin -= grow * h->w; h->imin += grow;
    if (h->imax > -1) h->imax += grow;
  } else if (b >= h->nb) {
    grow = (b - h->nb + 1) * 2;
    h->obs = realloc(h->obs, sizeof(uint64_t) * (size_t) (h->nb + grow));
    for (k = h->nb; k < h->nb + grow; k++) h->obs[k] = 0;
    if (h->imin == h->nb) h->imin += grow;
    h->bmax += grow * h->w;
    h->nb   += grow;
  }
  h->obs[b]++; h->n++; h->Nc++; h->No++;
  if (b > h->imax) h->imax = b;
  if (b < h->imin) h->imin = b;
  if (x > h->xmax) h->xmax = x;
  if (x < h->xmin) h->xmin = x;
  return 0;
}

/* the "ha" histogram of one scan: src/covariation.c:415-432 (msa2pdb filter is a no-op without a PDB) */
ORC_HIST *
orc_hist_from_cov(const double *cov, int L, double maxcov, double bmin, double w, double tol)
{
  ORC_HIST *h;
  double    bmax = maxcov + 5 * w;
  int       i, j;

  while (fabs(bmax - bmin) < tol) bmax += w;
  if ((h = orc_hist_create(bmin, bmax, w)) == NULL) return NULL;
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double x = AT(cov, L, i, j);
      if (x < bmin + w) x = bmin + w;                                   /* ESL_MAX(cov, bmin+w) :431 */
      if (orc_hist_add(h, x) != 0) { orc_hist_destroy(h); return NULL; }
    }
  return h;
}

/* cov_ranklist_Bin2Bin, src/covariation.c:2334-2362 */
static int
rebin(int b, const ORC_HIST *from, const ORC_HIST *to)
{
  double lo = from->w * b + from->bmin;
  return (int) round((lo - to->bmin) / to->w);
}

/* null_add2cumranklist, src/R-scape.c:1565-1612, with cov_GrowRankList, src/covariation.c:683-736 */
int
orc_hist_accumulate(ORC_HIST **cum, const ORC_HIST *one)
{
  ORC_HIST *c = *cum;
  int       b, nb2;

  if (one == NULL) return 0;
  if (c == NULL) {
    c = orc_hist_create(one->bmin, one->bmax, one->w);
    if (!c) return 1;
    c->n = one->n; c->xmin = one->xmin; c->xmax = one->xmax; c->imin = one->imin; c->imax = one->imax;
  } else {
    double    new_bmin = c->bmin;
    ORC_HIST *g;
    if (one->bmin < c->bmin) new_bmin -= fabs(one->bmin) * 2. * c->w;
    g = orc_hist_create(new_bmin, (one->bmax > c->bmax) ? one->bmax : c->bmax, c->w);
    if (!g) return 1;
    g->n = c->n; g->xmin = c->xmin; g->xmax = c->xmax; g->imin = c->imin; g->imax = c->imax; g->Nc = c->Nc; g->No = c->No;
    for (b = c->imin; b <= c->imax; b++) {
      nb2 = rebin(b, c, g);
      if (nb2 < g->nb) g->obs[nb2] = c->obs[b];
    }
    orc_hist_destroy(c);
    c = g;
    c->n   += one->n;
    c->xmin = (one->xmin < c->xmin) ? one->xmin : c->xmin;
    c->xmax = (one->xmax > c->xmax) ? one->xmax : c->xmax;
    c->imin = (one->imin < c->imin) ? one->imin : c->imin;
    c->imax = (one->imax > c->imax) ? one->imax : c->imax;
  }
  for (b = one->imin; b <= one->imax; b++) {
    nb2 = rebin(b, one, c);
    if (nb2 < c->nb) { c->obs[nb2] += one->obs[b]; c->Nc += one->obs[b]; c->No += one->obs[b]; }
  }
  *cum = c;
  return 0;
}

/* calculate_width_histo, src/R-scape.c:1357-1360 */
double
orc_null_width(double w_old, double mincov, double maxcov, double bmin, int hpts, double tol)
{
  double lo    = (bmin > mincov) ? bmin : mincov;
  double w_new = (maxcov - lo) / (double) hpts;
  double w     = (w_old < w_new) ? w_old : w_new;
  if (w < tol) w = 0.0;
  return w;
}

/* ------------------------------------------------------------------------------------------
 * null alignment generators
 */
struct orc_rng_s { ESL_RANDOMNESS *r; };

ORC_RNG *orc_rng_create(uint32_t seed) { ORC_RNG *g = malloc(sizeof(ORC_RNG)); if (g) g->r = esl_randomness_Create(seed); return g; }
void     orc_rng_destroy(ORC_RNG *g)   { if (g) { esl_randomness_Destroy(g->r); free(g); } }
double   orc_rng_uniform(ORC_RNG *g)   { return esl_random(g->r); }

/* P(t) = exp(tQ), small negatives clipped, rows renormalised:
 * ratematrix_ConditionalsFromRate, src/ratematrix.c:185-233, reached with the float-rounded,
 * floored branch length of e1_model_Create, src/e1_model.c:64-75,96-98. */
int
orc_ptime(const double *Q, double t, double *P)
{
  ESL_DMATRIX *q = esl_dmatrix_Create(K4, K4), *p = esl_dmatrix_Create(K4, K4);
  float        rt = (float) t;
  double       time;
  int          i, j, st = 0;

  if (!q || !p) return 1;
  rt = (rt >= 0.0 && rt < 1e-5) ? 1e-5 : rt;
  if (rt < 0.0) { if (rt > -1e-5) rt = 1e-5; else { st = 1; goto DONE; } }
  time = (rt > 10000.) ? 10000.0 : (double) rt;
  for (i = 0; i < K4; i++) for (j = 0; j < K4; j++) q->mx[i][j] = Q[i * K4 + j];
  if (esl_dmx_Exp(q, time, p) != eslOK) { st = 1; goto DONE; }
  for (i = 0; i < K4; i++) {
    for (j = 0; j < K4; j++)
      if (p->mx[i][j] < 0.0) { if (fabs(p->mx[i][j]) < 0.001) p->mx[i][j] = 0.0; else { st = 2; goto DONE; } }
    normalise(p->mx[i], K4);
    for (j = 0; j < K4; j++) P[i * K4 + j] = p->mx[i][j];
  }
 DONE:
  esl_dmatrix_Destroy(q);
  esl_dmatrix_Destroy(p);
  return st;
}

/* cov_addres, src/cov_simulate.c:757-773: inverse CDF with one uniform */
static uint8_t
draw_residue(ORC_RNG *g, const double *row)
{
  double x = orc_rng_uniform(g), cdf = 0.0;
  int    k;
  for (k = 0; k < K4; k++) { cdf += row[k]; if (cdf > x) break; }
  if (k == K4) k = K4 - 1;
  return (uint8_t) k;
}

/* generator B, ungapped + noss: every column evolves independently down the tree.
 * cov_evolve_root_ungapped_tree :289-324 visits v = 0..N-2 and emits left then right child;
 * cov_emit_ungapped :585-631 walks positions 1..L, one uniform each (cov_substitute :724-742). */
int
orc_null_simulate(ORC_RNG *g, const ORC_TREE *T, const double *Q, const uint8_t *root, int L,
                  uint8_t *leaves, uint8_t *internal)
{
  int      N = T->N, v, c, side, st = 0;
  uint8_t *nodes = internal ? internal : malloc((size_t) (N - 1) * (size_t) L);
  double   P[K16];

  if (!nodes) return 1;
  memcpy(nodes, root, (size_t) L);                                    /* cov_add_root :204-238 */
  for (v = 0; v < N - 1 && st == 0; v++)
    for (side = 0; side < 2; side++) {
      int            child = side ? T->right[v] : T->left[v];
      double         t     = side ? T->rd[v]    : T->ld[v];
      const uint8_t *par   = nodes + (size_t) v * L;
      uint8_t       *dst   = (child > 0) ? nodes + (size_t) child * L : leaves + (size_t) (-child) * L;
      if (orc_ptime(Q, t, P) != 0) { st = 2; break; }
      for (c = 0; c < L; c++) {
        if (par[c] > K4) { st = 3; break; }                           /* :732 */
        dst[c] = draw_residue(g, P + (par[c] < K4 ? par[c] : 0) * K4);
      }
    }
  if (!internal) free(nodes);
  return st;
}

/* ---- generator A ---- */
#define FDIM 6     /* K+2: residues 0..3, gap 4, "set" flag 5 (src/msatree.c:1714-1723) */

static int
set_ok(const int *S)
{
  int k, n = 0;
  for (k = 0; k < FDIM - 1; k++) n += S[k];
  return (n > 0 && S[FDIM - 1]);
}

/* tree_fitch_choose with frq == NULL, src/msatree.c:1834-1849: rejection-sample a set member */
static uint8_t
choose_member(ORC_RNG *g, const int *S)
{
  int k = (int) (orc_rng_uniform(g) * (FDIM - 1));
  while (!S[k]) k = (int) (orc_rng_uniform(g) * (FDIM - 1));
  return (uint8_t) k;
}

/* tree_fitch_upwards, src/msatree.c:1869-1905 */
static void
fitch_up(const int *Sl, const int *Sr, int *S, int *sc)
{
  int k, any = 0;
  for (k = 0; k < FDIM - 1; k++) if (Sl[k] && Sr[k]) { S[k] = 1; any = 1; }
  if (!any) {
    (*sc)++;
    for (k = 0; k < FDIM - 2; k++) if (Sl[k] || Sr[k]) S[k] = 1;     /* residues only; a gap never joins a union */
  }
  S[FDIM - 1] = 1;
}

/* one column: tree_fitch_column, src/msatree.c:1700-1831.  all = [2N-1][L] with leaves first. */
static int
fitch_column(ORC_RNG *g, const ORC_TREE *T, uint8_t *all, int L, int c, int *S, int *stack, int *sc)
{
  int N = T->N, n, v, sp = 0, k;

  memset(S, 0, sizeof(int) * (size_t) (2 * N - 1) * FDIM);
  for (n = 0; n < N; n++) {                                           /* leaves :1726-1741 */
    int r = all[(size_t) n * L + c];
    if (r <= K4)      S[n * FDIM + r] = 1;
    else if (r == 15) S[n * FDIM + (int) (orc_rng_uniform(g) * (FDIM - 1))] = 1;   /* unknown -> random */
    S[n * FDIM + FDIM - 1] = 1;
    if (!set_ok(S + n * FDIM)) return 1;
  }
  stack[sp++] = 0;                                                    /* post-order by re-pushing :1758-1777 */
  while (sp > 0) {
    int il, ir;
    v  = stack[--sp];
    il = (T->left[v]  <= 0) ? -T->left[v]  : N + T->left[v];
    ir = (T->right[v] <= 0) ? -T->right[v] : N + T->right[v];
    if (!S[il * FDIM + FDIM - 1]) { stack[sp++] = T->left[v];  continue; }
    if (!S[ir * FDIM + FDIM - 1]) { stack[sp++] = T->right[v]; continue; }
    if (set_ok(S + (N + v) * FDIM)) return 2;
    fitch_up(S + il * FDIM, S + ir * FDIM, S + (N + v) * FDIM, sc);
    if (!set_ok(S + (N + v) * FDIM)) return 3;
    if (v > 0) stack[sp++] = T->parent[v];
  }
  all[(size_t) N * L + c] = choose_member(g, S + N * FDIM);           /* root :1779 */

  stack[sp++] = 0;                                                    /* pre-order :1787-1815 */
  while (sp > 0) {
    int side;
    v = stack[--sp];
    for (side = 0; side < 2; side++) {
      int child = side ? T->right[v] : T->left[v];
      if (child > 0) {
        int  ax = all[(size_t) (N + v) * L + c];
        int *Sc = S + (N + child) * FDIM;
        if (Sc[ax]) for (k = 0; k < FDIM - 1; k++) if (k != ax) Sc[k] = 0;        /* tree_fitch_downwards :1922-1931 */
        all[(size_t) (N + child) * L + c] = choose_member(g, Sc);
      }
    }
    if (T->left[v]  > 0) stack[sp++] = T->left[v];
    if (T->right[v] > 0) stack[sp++] = T->right[v];
  }
  return 0;
}

/* shuffle_tree_substitutions + shuffle_tree_substitute_all, src/msamanip.c:1597-1780:
 * count the branch's substitutions a->d (5x5, gap is class 4) on the ORIGINAL rows, copy the
 * shuffled parent row to the child, then for each source class a pick (Fisher-Yates over the
 * positions currently holding a in the shuffled parent) as many positions as there were
 * substitutions out of a and overwrite them class by class. */
static int
replay_branch(ORC_RNG *g, const uint8_t *orig_par, const uint8_t *orig_kid, const uint8_t *sh_par, uint8_t *sh_kid,
              int L, int *pos, int *perm)
{
  int nsub[25], a, d, c, k;

  memset(nsub, 0, sizeof(nsub));
  for (c = 0; c < L; c++) {
    int pa = orig_par[c], kd = orig_kid[c];
    if (pa != kd && pa <= K4 && kd <= K4) nsub[pa * 5 + kd]++;        /* :1634-1643 */
    sh_kid[c] = sh_par[c];                                            /* :1645 */
  }
  for (a = 0; a < 5; a++) {
    int total = 0, n = 0, idx;
    for (d = 0; d < 5; d++) total += nsub[a * 5 + d];
    if (total == 0) continue;
    for (c = 0; c < L; c++) if (sh_par[c] == a) pos[n++] = c;         /* :1718-1730 */
    if (n == 0) continue;
    for (k = 0; k < n; k++) perm[k] = k;
    esl_vec_IShuffle(g->r, perm, n);                                  /* :1734 */
    idx = n - 1;
    for (d = 0; d < 5; d++) {
      int s = nsub[a * 5 + d];
      while (s > 0 && idx >= 0) { sh_kid[pos[perm[idx]]] = (uint8_t) d; idx--; s--; total--; }
    }
    if (total > 0 && nsub[a * 5] + nsub[a * 5 + 1] + nsub[a * 5 + 2] + nsub[a * 5 + 3] + nsub[a * 5 + 4] <= n) return 1;
  }
  return 0;
}

int
orc_null_fitch_shuffle(ORC_RNG *g, const ORC_TREE *T, const uint8_t *msa, int L, uint8_t *shmsa, uint8_t *allmsa, int *fitch_sc)
{
  int      N = T->N, tot = 2 * N - 1, c, v, n, st = 0, sc = 0;
  uint8_t *all = allmsa ? allmsa : malloc((size_t) tot * (size_t) L);
  uint8_t *sh  = malloc((size_t) tot * (size_t) L);
  int     *S   = malloc(sizeof(int) * (size_t) tot * FDIM);
  int     *stk = malloc(sizeof(int) * (size_t) (2 * N + 4));
  int     *perm = malloc(sizeof(int) * (size_t) (L > 0 ? L : 1));
  int     *pos  = malloc(sizeof(int) * (size_t) (L > 0 ? L : 1));

  if (!all || !sh || !S || !stk || !perm || !pos) { st = 100; goto DONE; }
  memcpy(all, msa, (size_t) N * (size_t) L);
  memset(all + (size_t) N * L, K4, (size_t) (N - 1) * (size_t) L);
  for (c = 0; c < L && st == 0; c++) st = fitch_column(g, T, all, L, c, S, stk, &sc);   /* msatree.c:211-214 */
  if (st != 0) goto DONE;

  /* msamanip_ShuffleColumns, src/msamanip.c:1164-1233: one permutation for all 2N-1 rows */
  for (c = 0; c < L; c++) perm[c] = c;
  esl_vec_IShuffle(g->r, perm, L);
  for (n = 0; n < tot; n++)
    for (c = 0; c < L; c++) sh[(size_t) n * L + c] = all[(size_t) n * L + perm[c]];

  /* msamanip_ShuffleTreeSubstitutions, src/msamanip.c:1486-1504: parents before children */
  for (v = 0; v < N - 1 && st == 0; v++) {
    int ip = N + v;
    int il = (T->left[v]  <= 0) ? -T->left[v]  : N + T->left[v];
    int ir = (T->right[v] <= 0) ? -T->right[v] : N + T->right[v];
    st = replay_branch(g, all + (size_t) ip * L, all + (size_t) il * L, sh + (size_t) ip * L, sh + (size_t) il * L, L, pos, perm);
    if (st == 0)
      st = replay_branch(g, all + (size_t) ip * L, all + (size_t) ir * L, sh + (size_t) ip * L, sh + (size_t) ir * L, L, pos, perm);
  }
  if (st == 0) memcpy(shmsa, sh, (size_t) N * (size_t) L);             /* leaves only :1511 */
  if (fitch_sc) *fitch_sc = sc;

 DONE:
  if (!allmsa) free(all);
  free(sh); free(S); free(stk); free(perm); free(pos);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * E-values and the significant-pair list.
 */
/* cov2evalue, src/covariation.c:2370-2400: survival of the null scores at `cov`, times Nc.  The fitted tail
 * (survfit, 2*nb entries, src/covariation.c:1677-1699) serves scores at or above phi, the sampled distribution the rest.
 * The reference accumulates the bin counts in an int (:2374); the 64-bit sum here agrees up to 2^31 scores. */
double
orc_cov2evalue(double cov, int Nc, const ORC_HIST *h, double phi, const double *survfit)
{
  double   eval = INFINITY;
  int      icov = orc_hist_score2bin(h, cov);
  int64_t  c    = 0;
  int      i;

  if      (survfit && icov >= 2 * h->nb - 1) eval = survfit[2 * h->nb - 1] * (double) Nc;
  else if (survfit && cov >= phi)            eval = survfit[icov + 1]      * (double) Nc;
  else {
    if (cov >= h->xmax) return (double) Nc / (double) h->Nc;
    if (icov <= h->imax) {
      if (icov <  h->imin)     icov = h->imin;
      if (icov >= h->imax - 1) eval = (double) Nc / (double) h->Nc;
      else {
        for (i = h->imax; i >= icov; i--) c += (int64_t) h->obs[i];
        eval = (double) c * (double) Nc / (double) h->Nc;
      }
    }
    else return NAN;                                                    /* "cannot find evalue", exit(1) at :2394 */
  }
  return eval;
}

/* evalue2cov, src/covariation.c:2404-2435: the score at which the E-value reaches eval_thresh */
double
orc_evalue2cov(double eval_thresh, int Nc, const ORC_HIST *h, int cmin, const double *survfit)
{
  double  cov = -INFINITY, eval;
  int64_t c = 0;
  int     i, b = cmin - 1;

  if (h->No >= 10 && survfit) {
    for (b = 2 * h->nb - 1; b >= cmin; b--) {
      eval = survfit[b] * (double) Nc;
      if (eval >= eval_thresh) break;
    }
    cov = h->w * b + h->bmin;
  }
  if (b == cmin - 1) {
    for (i = h->imax; i >= h->imin; i--) {
      c += (int64_t) h->obs[i];
      eval = (double) c * (double) Nc / (double) h->Nc;
      if (eval >= eval_thresh) break;
    }
    cov = h->w * (i + 1) + h->bmin;
  }
  return cov;
}

/* The per-pair loop of cov_CreateHitList, src/covariation.c:828-910: p-value of every pair i<j from the null histogram,
 * E-value = p * (number of tests of the pair's set), hit iff E < thresh (or thresh > MAX_EVAL = 1000: report all).
 * pairmask (uint8 [L][L], may be NULL) flags the pairs of the structure set (isbp by data->samplesize); expBP > 0 applies
 * the `h < expBP` rule of :852.  eval (double [L][L], may be NULL) gets mi->Eval (both triangles; diagonal untouched).
 * Hits are appended in the reference's row-major order; returns their number (entries beyond cap are counted, not stored),
 * or -1 if a score has no E-value. */
int64_t
orc_hitlist(const double *cov, int L, const ORC_HIST *null, double phi, const double *survfit, const uint8_t *pairmask,
            uint64_t Nb, uint64_t Nt, int expBP, double thresh, double *eval, int64_t cap,
            int64_t *hit_i, int64_t *hit_j, double *hit_sc, double *hit_eval, double *hit_pval)
{
  int64_t h = 0;
  int     i, j;

  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      double sc   = AT(cov, L, i, j);
      double pval = orc_cov2evalue(sc, 1, null, phi, survfit);
      double ev;
      int    isbp = pairmask ? (pairmask[(size_t) i * L + j] != 0) : 0;

      if (isnan(pval)) return -1;
      if (isbp)             ev = pval * Nb;
      else if (h < expBP)   ev = pval * expBP;
      else                  ev = pval * Nt;
      if (eval) { AT(eval, L, i, j) = ev; AT(eval, L, j, i) = ev; }
      if (ev < thresh || thresh > 1000.) {
        if (h < cap) {
          if (hit_i)    hit_i[h]    = i;
          if (hit_j)    hit_j[h]    = j;
          if (hit_sc)   hit_sc[h]   = sc;
          if (hit_eval) hit_eval[h] = ev;
          if (hit_pval) hit_pval[h] = pval;
        }
        h++;
      }
    }
  return h;
}

/* ------------------------------------------------------------------------------------------
 * Tree_Substitutions, src/msatree.c:1455-1540, on the rows left by the Fitch pass (all = [2N-1][L]: leaves 0..N-1, then
 * internal node v at row N+v): per column the number of parent->child substitutions over the 2(N-1) branches, per pair
 * i<j the number of branches on which both columns change (ndouble) / at least one does (njoin).  Without includegaps a
 * branch only counts where parent and child residues are canonical in the column(s) concerned.
 * nsubs int [L]; ndouble, njoin int [L][L] (entries i<j, the rest 0).  Any output may be NULL. */
int
orc_tree_substitutions(const ORC_TREE *T, const uint8_t *all, int L, int includegaps, int *nsubs, int *ndouble, int *njoin)
{
  int N = T->N, v, side, i, j;

  if (nsubs)   memset(nsubs,   0, sizeof(int) * (size_t) L);
  if (ndouble) memset(ndouble, 0, sizeof(int) * (size_t) L * (size_t) L);
  if (njoin)   memset(njoin,   0, sizeof(int) * (size_t) L * (size_t) L);
  for (v = 0; v < N - 1; v++)
    for (side = 0; side < 2; side++) {
      int            kid = side ? T->right[v] : T->left[v];
      const uint8_t *ax  = all + (size_t) (N + v) * L;
      const uint8_t *axk = (kid > 0) ? all + (size_t) (N + kid) * L : all + (size_t) (-kid) * L;
      for (i = 0; i < L; i++) {
        int oki = includegaps || (ax[i] < 4 && axk[i] < 4);
        int chi = oki && axk[i] != ax[i];
        if (nsubs && chi) nsubs[i]++;
        if (!ndouble && !njoin) continue;
        if (!oki) continue;
        for (j = i + 1; j < L; j++) {
          int okj = includegaps || (ax[j] < 4 && axk[j] < 4);
          int chj = okj && axk[j] != ax[j];
          if (!okj) continue;
          if (ndouble && chi && chj)   ndouble[(size_t) i * L + j]++;
          if (njoin   && (chi || chj)) njoin[(size_t) i * L + j]++;
        }
      }
    }
  return 0;
}
