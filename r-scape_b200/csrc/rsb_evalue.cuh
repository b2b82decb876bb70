// rsb_evalue.cuh -- score -> p-value against the cumulative null histogram: cov2evalue(cov, 1, h, survfit),
// src/covariation.c:2370-2400, as one branch-only function without the reference's loop over the bins.
//
// The reference sums obs[imax..icov] for every pair (:2390); here the suffix sums csum[b] = sum_{i=b..imax} obs[i] are
// built once on the host (exact integers), so a pair costs one lookup.  Every floating-point operation is the
// reference's own, in its order: with Nc = 1 the products `x * (double) Nc` are exact and the quotients are
// `(double) c / (double) h->Nc` and `1 / (double) h->Nc`, so p-values are bit-identical to the reference's.
// (The reference accumulates c in an int, :2374; the 64-bit sums here agree with it below 2^31 null scores.)
//
// The function is __host__ __device__ so that tests/test_evalue_header.py can compile this very header with g++ and
// compare it with the oracle on the CPU; the product only ever calls it from evalue_hits_kernel (hits.cu).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define RSB_HD __host__ __device__ __forceinline__
#else
#define RSB_HD static inline
#endif

struct rsb_nullview {
  double bmin, w;                       // bin b covers (bmin + b w, bmin + (b+1) w]
  double xmax;                          // largest null score (h->xmax)
  double phi;                           // censoring point of the fitted tail (h->phi)
  double Nc;                            // (double) h->Nc: number of null scores
  int    nb, imin, imax;
  const unsigned long long *csum;       // [nb] suffix sums of the bins
  const double *survfit;                // [2 nb] fitted survival at the upper bound of every bin (:1677-1699), or NULL
};

// p-value of one score; *bad is set where the reference prints "cannot find evalue" and exits (:2394)
RSB_HD double rsb_cov2pval(double cov, const rsb_nullview &h, int *bad)
{
  const double bd = ceil(((cov - h.bmin) / h.w) - 1.);                  // esl_histogram_Score2Bin
  if (h.survfit && bd >= (double) (2 * h.nb - 1)) return h.survfit[2 * h.nb - 1];
  if (h.survfit && cov >= h.phi) {
    if (bd < -1.0) { *bad = 1; return NAN; }                            // the reference would read survfit[icov+1] out of bounds
    return h.survfit[(int) bd + 1];
  }
  if (cov >= h.xmax) return 1.0 / h.Nc;
  if (bd <= (double) h.imax) {
    const int icov = (bd < (double) h.imin) ? h.imin : (int) bd;
    if (icov >= h.imax - 1) return 1.0 / h.Nc;
    return (double) h.csum[icov] / h.Nc;
  }
  *bad = 1;
  return NAN;
}
