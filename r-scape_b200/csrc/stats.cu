// stats.cu -- probabilities, marginals and the per-pair covariation statistics from the count planes.
//
// Reference (src/correlators.c):
//   pp normalisation with the 1e-10 prior      :1713,1758      nseff / ngap            :1747-1753
//   corr_Marginals                             :1338-1375      corr_ValidateProbs      :1500-1545
//   CHI :93-181  OMES :227-314  GT :361-490  MI :536-614  MIr :660-745  MIg :789-874
//   RAF :877-933 (via the count-table identity)  RAFS :936-982  CCF :1011-1061
//
// All kernels walk the upper triangle in tiles of ST_TI rows (i) x ST_TJ columns (j): a thread owns one
// column j and loops over the tile's rows, so reads of the int64 count planes cnt[16][L][Lp] are
// coalesced along j.  Sums that the reference takes over all partners of a column (pm, APC row means)
// are produced as per-tile row partials (warp shuffle over j, fixed order) and column partials (a
// thread's own running sum over i) and combined in a fixed order by a finalize kernel, so every
// result is deterministic run to run; no floating-point atomics anywhere.
#include "rsb_common.cuh"
#include <math.h>
#pragma nv_diag_suppress 128      // template branches that return early leave the generic loop unreachable in some instantiations

namespace {

#ifdef RSB_BLOCKTRACE
__device__ unsigned long long *stats_trace_buf;
#endif
constexpr int ST_TI = RSB_TI;
constexpr int ST_TJ = RSB_TJ;

struct PairProbs { double pp[16]; double ne; double ng; };

// exact uint64 -> double without the slow 64-bit I2F path: both 32-bit halves are planted in the mantissa of a
// power of two and the offsets subtracted (each half exact, one rounding in the final add = the I2F result)
__device__ __forceinline__ double u64_to_f64(unsigned long long v)
{
  const double lo = __longlong_as_double(0x4330000000000000ULL | (v & 0xffffffffULL)) - 4503599627370496.0;              // 2^52
  const double hi = __longlong_as_double(0x4530000000000000ULL | (v >> 32))           - 19342813113834066795298816.0;   // 2^84
  return hi + lo;
}

// counts below 2^52 (the whole alignment's weight is: wtot < 2^52, checked by the caller) convert with one subtraction
__device__ __forceinline__ double u52_to_f64(unsigned long long v)
{
  return __longlong_as_double(0x4330000000000000ULL | v) - 4503599627370496.0;
}

// fixed-point counts of one pair -> nseff, ngap, pp (prior 1e-10 per cell).
// EXACT_DIV (values handed back to the host): Kahan-summed normaliser and a division per cell, as esl_vec_DNorm does
// (SURVEY 9.7).  Otherwise (statistic kernels, FP64-issue bound): pairwise-tree normaliser and a multiplication by the
// reciprocal; both differ from the exact form by a few ulp, far inside the 1e-9 score tolerance.
template <bool EXACT_DIV>
__device__ __forceinline__ void load_pair(const long long *__restrict__ cnt, size_t plane, size_t off,
                                          double scale, long long wtot, PairProbs &P)
{
  unsigned long long c[16];
  #pragma unroll
  for (int k = 0; k < 16; k++) c[k] = (unsigned long long) cnt[k * plane + off];
  unsigned long long ne = 0;
  #pragma unroll
  for (int k = 0; k < 16; k++) ne += c[k];
  if (EXACT_DIV) {
    P.ne = u64_to_f64(ne) * scale;
    P.ng = u64_to_f64((unsigned long long) wtot - ne) * scale;
    double sum = 0.0, comp = 0.0;
    #pragma unroll
    for (int k = 0; k < 16; k++) {
      P.pp[k] = 1e-10 + u64_to_f64(c[k]) * scale;
      const double y = P.pp[k] - comp, t = sum + y;      // esl_vec_DSum is Kahan-compensated
      comp = (t - sum) - y;
      sum  = t;
    }
    #pragma unroll
    for (int k = 0; k < 16; k++) P.pp[k] = P.pp[k] / sum;
  } else {
    const bool small = ((unsigned long long) wtot >> 52) == 0;       // uniform over the grid
    if (small) {
      P.ne = u52_to_f64(ne) * scale;
      P.ng = u52_to_f64((unsigned long long) wtot - ne) * scale;
      #pragma unroll
      for (int k = 0; k < 16; k++) P.pp[k] = fma(u52_to_f64(c[k]), scale, 1e-10);
    } else {
      P.ne = u64_to_f64(ne) * scale;
      P.ng = u64_to_f64((unsigned long long) wtot - ne) * scale;
      #pragma unroll
      for (int k = 0; k < 16; k++) P.pp[k] = fma(u64_to_f64(c[k]), scale, 1e-10);
    }
    double t[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) t[k] = P.pp[2 * k] + P.pp[2 * k + 1];
    const double sum = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
    const double inv = 1.0 / sum;
    #pragma unroll
    for (int k = 0; k < 16; k++) P.pp[k] = P.pp[k] * inv;
  }
}

// natural log of a positive normal double from a 128-entry table in shared memory: x = 2^e m, m = c (1 + r) with c the
// centre of m's 1/128 bin, so |r| <= 2^-8 and log1p(r) needs six terms (truncation 2^-58).  About 10 FP64 operations
// against ~40 for the library log; absolute error <= 4e-16 + 1 ulp(e ln 2).  tab[k] = { 1/c_k rounded, -log of that }.
constexpr int LOGTAB_N = 128;
__device__ __forceinline__ void logtab_init(double2 *tab)
{
  for (int k = threadIdx.x; k < LOGTAB_N; k += blockDim.x) {
    const double ic = 1.0 / (1.0 + (k + 0.5) * (1.0 / LOGTAB_N));
    tab[k] = make_double2(ic, -log(ic));
  }
}

__device__ __forceinline__ double fast_log(double x, const double2 *__restrict__ tab)
{
  const int hi = __double2hiint(x), lo = __double2loint(x);
  if ((unsigned) (hi - 0x00100000) >= 0x7fe00000u) return log(x);        // zero, subnormal, negative, inf, nan
  const double m  = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 t = tab[(hi >> 13) & (LOGTAB_N - 1)];
  const double r  = fma(m, t.x, -1.0);
  const double e  = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  return fma(e, 0.6931471805599453094, t.y) + fma(p, r * r, r);
}

__device__ __forceinline__ double warp_sum(double v)
{
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// marginal partials + nseff.  rowpart[r][jt][i][4], colpart[r][it][j][4]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_TJ)
marg_partial_kernel(const long long *__restrict__ cnt, int L, int Lp, double scale, long long wtot,
                    double *__restrict__ rowpart, double *__restrict__ colpart, double *__restrict__ nseff,
                    int nJT, int nIT, int sr, int sw)
{
  __shared__ double rowacc[ST_TJ / 32][ST_TI][4];
#ifdef RSB_BLOCKTRACE
  const unsigned long long t0 = rsb_gtime();
#endif
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t plane = (size_t) L * Lp;
  const long long *c = cnt + (size_t) r * 16 * plane;
  const bool tile_live = (it * ST_TI) < (jt * ST_TJ + ST_TJ - 1) && RSB_OWNED(it, sr, sw);      // some i < some j, rows owned by this rank
  double col[4] = { 0, 0, 0, 0 };

  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    double rs[4] = { 0, 0, 0, 0 };
    if (tile_live && i < L && j < L && i < j) {
      PairProbs P;
      load_pair<false>(c, plane, (size_t) i * Lp + j, scale, wtot, P);
      nseff[((size_t) r * L + i) * Lp + j] = P.ne;
      if (P.ne > 0) {                                                    // :1354
        #pragma unroll
        for (int a = 0; a < 4; a++)
          #pragma unroll
          for (int b = 0; b < 4; b++) { rs[a] += P.pp[a * 4 + b]; col[b] += P.pp[a * 4 + b]; }
      }
    }
    #pragma unroll
    for (int a = 0; a < 4; a++) { const double v = warp_sum(rs[a]); if (lane == 0) rowacc[warp][il][a] = v; }
  }
  if (j < L) {
    double *cp = colpart + (((size_t) r * nIT + it) * L + j) * 4;
    #pragma unroll
    for (int b = 0; b < 4; b++) cp[b] = col[b];
  }
  __syncthreads();
  if (threadIdx.x < ST_TI * 4) {
    const int il = threadIdx.x >> 2, a = threadIdx.x & 3, i = it * ST_TI + il;
    if (i < L) {
      double v = 0.0;
      #pragma unroll
      for (int w = 0; w < ST_TJ / 32; w++) v += rowacc[w][il][a];
      rowpart[(((size_t) r * nJT + jt) * L + i) * 4 + a] = v;
    }
  }
#ifdef RSB_BLOCKTRACE
  if (threadIdx.x == 0) rsb_trace_put(stats_trace_buf, 2, t0);
#endif
}

// unnormalised marginal sums msum[i][a] = sum of the tile partials in a fixed order: one warp per column, lanes stride
// over the partial blocks, fixed xor-tree over the lanes (deterministic).  With the pair grid sharded over ranks these are
// the vectors that are summed across ranks before normalisation.
__global__ void __launch_bounds__(256)
marg_sum_kernel(const double *__restrict__ rowpart, const double *__restrict__ colpart, int L,
                int nJT, int nIT, double *__restrict__ msum)
{
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), r = blockIdx.y;
  if (i >= L) return;
  double m[4] = { 0, 0, 0, 0 };
  for (int k = lane; k < nJT + nIT; k += 32) {
    const double *src = (k < nJT) ? rowpart + (((size_t) r * nJT + k) * L + i) * 4
                                  : colpart + (((size_t) r * nIT + (k - nJT)) * L + i) * 4;
    const double2 lo = *reinterpret_cast<const double2 *>(src), hi = *reinterpret_cast<const double2 *>(src + 2);
    m[0] += lo.x; m[1] += lo.y; m[2] += hi.x; m[3] += hi.y;
  }
  #pragma unroll
  for (int a = 0; a < 4; a++) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[a] += __shfl_xor_sync(0xffffffffu, m[a], o);
  }
  if (lane == 0) {
    #pragma unroll
    for (int a = 0; a < 4; a++) msum[((size_t) r * L + i) * 4 + a] = m[a];
  }
}

// pm[i] = normalise(msum[i]); validation flag as corr_Marginals' esl_vec_DValidate (:1363)
__global__ void marg_norm_kernel(const double *__restrict__ msum, int L, double tol, double *__restrict__ pm, int *__restrict__ flags)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (i >= L) return;
  double m[4];
  #pragma unroll
  for (int a = 0; a < 4; a++) m[a] = msum[((size_t) r * L + i) * 4 + a];
  double sum = 0.0, comp = 0.0;
  #pragma unroll
  for (int a = 0; a < 4; a++) { const double y = m[a] - comp, t = sum + y; comp = (t - sum) - y; sum = t; }
  double chk = 0.0;
  #pragma unroll
  for (int a = 0; a < 4; a++) { m[a] = (sum != 0.0) ? m[a] / sum : 0.25; chk += m[a]; pm[((size_t) r * L + i) * 4 + a] = m[a]; }
  if (!(fabs(chk - 1.0) <= tol)) atomicOr(flags, 1);
}

// ------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_allowed(unsigned mask, int x, int y) { return (mask >> (x * 4 + y)) & 1u; }

template <int STAT, int CLS>
__device__ __forceinline__ double pair_statistic(const PairProbs &P, const double *mi, const double *mj,
                                                 const double *lmi, const double *lmj, unsigned mask,
                                                 const double2 *__restrict__ tab)
{
  double v = 0.0, H = 0.0;
  if (CLS == RSB_C2) {
    double p_in = 0, p_out = 0, q_in = 0, q_out = 0;
    #pragma unroll
    for (int x = 0; x < 4; x++)
      #pragma unroll
      for (int y = 0; y < 4; y++) {
        if (cell_allowed(mask, x, y)) { p_in  += P.pp[x * 4 + y]; q_in  += mi[x] * mj[y]; }
        else                          { p_out += P.pp[x * 4 + y]; q_out += mi[x] * mj[y]; }
      }
    const double exp_in = P.ne * q_in, exp_out = P.ne * q_out, obs_in = P.ne * p_in, obs_out = P.ne * p_out;
    if (STAT == RSB_CHI) {
      v += (exp_in  > 0.) ? (obs_in  - exp_in)  * (obs_in  - exp_in)  / exp_in  : 0.0;
      v += (exp_out > 0.) ? (obs_out - exp_out) * (obs_out - exp_out) / exp_out : 0.0;
    } else if (STAT == RSB_OMES) {
      v += (exp_in  > 0.) ? (obs_in  - exp_in)  * (obs_in  - exp_in)  / P.ne : 0.0;
      v += (exp_out > 0.) ? (obs_out - exp_out) * (obs_out - exp_out) / P.ne : 0.0;
    } else if (STAT == RSB_GT) {
      v += (exp_in  > 0. && obs_in  > 0.) ? obs_in  * fast_log(obs_in  / exp_in, tab)  : 0.0;
      v += (exp_out > 0. && obs_out > 0.) ? obs_out * fast_log(obs_out / exp_out, tab) : 0.0;
      v *= 2.0;
    } else if (STAT == RSB_MI || STAT == RSB_MIg) {
      v += (p_in  > 0.) ? p_in  * (fast_log(p_in, tab)  - fast_log(q_in, tab))  : 0.0;
      v += (p_out > 0.) ? p_out * (fast_log(p_out, tab) - fast_log(q_out, tab)) : 0.0;
      if (STAT == RSB_MIg) v -= (P.ne > 0) ? P.ng / P.ne : 0.0;
    } else if (STAT == RSB_MIr) {
      H -= (p_in  > 0.) ? p_in  * fast_log(p_in, tab)  : 0.0;
      H -= (p_out > 0.) ? p_out * fast_log(p_out, tab) : 0.0;
      v += (p_in  > 0. && q_in  > 0.) ? p_in  * (fast_log(p_in, tab)  - fast_log(q_in, tab))  : 0.0;
      v += (p_out > 0. && q_out > 0.) ? p_out * (fast_log(p_out, tab) - fast_log(q_out, tab)) : 0.0;
      v = (H > 1e-2) ? v / H : 0.0;
    }
    return v;
  }
  // lmi / lmj: log of the marginals, taken once per column by the caller.
  if (STAT == RSB_GT && CLS == RSB_C16) {
    // G = 2 sum obs fast_log(obs/exp, tab) with obs = ne pp, exp = ne pm_i pm_j (:383-387): ne cancels inside the log, so
    // G = 2 ne sum pp (log pp - log pm_i - log pm_j), one log per cell and no division.  The terms kept are the
    // same (exp > 0 and obs > 0 <=> ne > 0, pm_i > 0, pm_j > 0, pp > 0); the rounding differs at the 1e-15 level.
    if (!(P.ne > 0.)) return 0.0;
    #pragma unroll
    for (int x = 0; x < 4; x++)
      #pragma unroll
      for (int y = 0; y < 4; y++) {
        const double pxy = P.pp[x * 4 + y];
        v += (pxy > 0.0 && mi[x] > 0.0 && mj[y] > 0.0) ? pxy * (fast_log(pxy, tab) - lmi[x] - lmj[y]) : 0.0;
      }
    return 2.0 * P.ne * v;
  }
  #pragma unroll
  for (int x = 0; x < 4; x++)
    #pragma unroll
    for (int y = 0; y < 4; y++) {
      if (CLS == RSB_CWC && !cell_allowed(mask, x, y)) continue;
      const double pxy = P.pp[x * 4 + y];
      const double ex  = P.ne * mi[x] * mj[y];
      const double ob  = P.ne * pxy;
      if      (STAT == RSB_CHI)  v += (ex > 0.) ? (ob - ex) * (ob - ex) / ex   : 0.0;
      else if (STAT == RSB_OMES) v += (ex > 0.) ? (ob - ex) * (ob - ex) / P.ne : 0.0;
      else if (STAT == RSB_GT)   v += (ex > 0. && ob > 0.) ? ob * fast_log(ob / ex, tab) : 0.0;
      else {
        const double lp = (pxy > 0.0) ? fast_log(pxy, tab) : 0.0;
        if (STAT == RSB_MIr) H -= (pxy > 0.0) ? pxy * lp : 0.0;
        v += (pxy > 0.0 && mi[x] > 0.0 && mj[y] > 0.0) ? pxy * (lp - lmi[x] - lmj[y]) : 0.0;
      }
    }
  if (STAT == RSB_GT)  v *= 2.0;
  if (STAT == RSB_MIg) v -= (P.ne > 0) ? P.ng / P.ne : 0.0;
  if (STAT == RSB_MIr) v  = (H > 1e-2) ? v / H : 0.0;
  return v;
}

// raw statistic for every pair of the tile; row/column partial sums for the background correction;
// per-block min/max.  rowpart[r][jt][i], colpart[r][it][j], mm[r][block][2]
template <int STAT, int CLS>
__global__ void __launch_bounds__(ST_TJ)
stat_kernel(const long long *__restrict__ cnt, const double *__restrict__ pm, int L, int Lp, double scale, long long wtot,
            unsigned mask, double *__restrict__ cov, double *__restrict__ rowpart, double *__restrict__ colpart,
            double *__restrict__ mm, int nJT, int nIT, int sr, int sw)
{
  __shared__ double rowacc[ST_TJ / 32][ST_TI];
  __shared__ double pmi[ST_TI][4], lpmi[ST_TI][4];
  __shared__ double smin[ST_TJ / 32], smax[ST_TJ / 32];
  __shared__ double2 tab[LOGTAB_N];
#ifdef RSB_BLOCKTRACE
  const unsigned long long t0 = rsb_gtime();
#endif
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t plane = (size_t) L * Lp;
  const long long *c = cnt + (size_t) r * 16 * plane;
  const bool tile_live = (it * ST_TI) < (jt * ST_TJ + ST_TJ - 1) && RSB_OWNED(it, sr, sw);
  double col = 0.0, vmin = INFINITY, vmax = -INFINITY;
  if (tile_live) logtab_init(tab);
  double mj[4] = { 0.25, 0.25, 0.25, 0.25 }, lmj[4];

  if (threadIdx.x < ST_TI * 4) {
    const int il = threadIdx.x >> 2, a = threadIdx.x & 3, i = it * ST_TI + il;
    pmi[il][a]  = (i < L) ? pm[((size_t) r * L + i) * 4 + a] : 0.25;
    lpmi[il][a] = log(pmi[il][a]);
  }
  if (j < L) {
    #pragma unroll
    for (int b = 0; b < 4; b++) mj[b] = pm[((size_t) r * L + j) * 4 + b];
  }
  #pragma unroll
  for (int b = 0; b < 4; b++) lmj[b] = log(mj[b]);
  __syncthreads();

  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    double v = 0.0;
    if (tile_live && i < L && j < L && i < j) {
      PairProbs P;
      load_pair<false>(c, plane, (size_t) i * Lp + j, scale, wtot, P);
      v = pair_statistic<STAT, CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
      cov[((size_t) r * L + i) * Lp + j] = v;
      col += v;
      vmin = fmin(vmin, v);
      vmax = fmax(vmax, v);
    }
    const double rs = warp_sum(v);
    if (lane == 0) rowacc[warp][il] = rs;
  }
  if (j < L) colpart[((size_t) r * nIT + it) * L + j] = col;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
  __syncthreads();
  if (threadIdx.x < ST_TI) {
    const int i = it * ST_TI + threadIdx.x;
    if (i < L) {
      double v = 0.0;
      #pragma unroll
      for (int w = 0; w < ST_TJ / 32; w++) v += rowacc[w][threadIdx.x];
      rowpart[((size_t) r * nJT + jt) * L + i] = v;
    }
  }
  if (threadIdx.x == 0) {
    double a = smin[0], b = smax[0];
    #pragma unroll
    for (int w = 1; w < ST_TJ / 32; w++) { a = fmin(a, smin[w]); b = fmax(b, smax[w]); }
    double *o = mm + (((size_t) r * nIT + it) * nJT + jt) * 2;
    o[0] = a; o[1] = b;
  }
#ifdef RSB_BLOCKTRACE
  if (threadIdx.x == 0) rsb_trace_put(stats_trace_buf, 3, t0);
#endif
}

// RAF from the UNWEIGHTED count table (planes built with wq = 1, S = 1): integer arithmetic up to the
// final divisions, so the result is bit-identical to the reference's O(N^2) loop (:903-918).
__global__ void __launch_bounds__(ST_TJ)
raf_kernel(const long long *__restrict__ cnt, int L, int Lp, int nseq, unsigned mask, double *__restrict__ out)
{
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const size_t plane = (size_t) L * Lp;
  const long long *c = cnt + (size_t) r * 16 * plane;
  if ((it * ST_TI) >= (jt * ST_TJ + ST_TJ - 1) || j >= L) return;
  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    if (i >= L || i >= j) continue;
    long long n[16], nallowed = 0, h = 0;
    #pragma unroll
    for (int k = 0; k < 16; k++) n[k] = c[k * plane + (size_t) i * Lp + j];
    #pragma unroll
    for (int c1 = 0; c1 < 16; c1++) {
      if (!((mask >> c1) & 1u)) continue;
      nallowed += n[c1];
      #pragma unroll
      for (int c2 = c1 + 1; c2 < 16; c2++) {
        if (!((mask >> c2) & 1u)) continue;
        h += n[c1] * n[c2] * (long long) (((c1 >> 2) != (c2 >> 2)) + ((c1 & 3) != (c2 & 3)));
      }
    }
    double qij = (double) (nseq - nallowed);
    qij /= nseq;
    double cij = (double) h;
    cij /= (nseq > 1) ? (double) nseq * ((double) nseq - 1.0) : 1.0;
    cij *= 2.0;
    out[((size_t) r * L + i) * Lp + j] = cij - 1.0 * qij;
  }
}

// RAFS 3-point anti-diagonal stencil (:954-962); min/max and correction partials come from reduce_cov_kernel
__global__ void rafs_kernel(const double *__restrict__ raf, int L, int Lp, double *__restrict__ out)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, r = blockIdx.z;
  if (j >= L || i >= j) return;
  const double *B = raf + (size_t) r * L * Lp;
  double v = 2.0 * B[(size_t) i * Lp + j];
  if (i > 0 && j < L - 1)              v += B[(size_t) (i - 1) * Lp + j + 1];
  if (j > 0 && i < L - 1 && i < j - 2) v += B[(size_t) (i + 1) * Lp + j - 1];
  out[((size_t) r * L + i) * Lp + j] = 0.25 * v;
}

// CCF (:1030-1058): pass 1 accumulates meanp[x] ~ sum_{i<j} nseff_ij pm_i[x] as per-block partials
__global__ void __launch_bounds__(ST_TJ)
ccf_meanp_kernel(const double *__restrict__ nseff, const double *__restrict__ pm, int L, int Lp, double *__restrict__ part, int nJT, int nIT)
{
  __shared__ double acc[ST_TJ / 32][4];
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s[4] = { 0, 0, 0, 0 };
  if ((it * ST_TI) < (jt * ST_TJ + ST_TJ - 1) && j < L)
    for (int il = 0; il < ST_TI; il++) {
      const int i = it * ST_TI + il;
      if (i >= L || i >= j) continue;
      const double ne = nseff[((size_t) r * L + i) * Lp + j];
      #pragma unroll
      for (int x = 0; x < 4; x++) s[x] += ne * pm[((size_t) r * L + i) * 4 + x];
    }
  #pragma unroll
  for (int x = 0; x < 4; x++) { const double v = warp_sum(s[x]); if (lane == 0) acc[warp][x] = v; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < ST_TJ / 32; w++) v += acc[w][threadIdx.x];
    part[((((size_t) r * nIT + it) * nJT) + jt) * 4 + threadIdx.x] = v;
  }
}

__global__ void ccf_meanp_final_kernel(const double *__restrict__ part, int nblocks, double *__restrict__ meanp)
{
  const int r = blockIdx.x;
  if (threadIdx.x != 0) return;
  double m[4] = { 0, 0, 0, 0 };
  for (int b = 0; b < nblocks; b++)
    for (int x = 0; x < 4; x++) m[x] += part[((size_t) r * nblocks + b) * 4 + x];
  double sum = 0.0, comp = 0.0;
  for (int x = 0; x < 4; x++) { const double y = m[x] - comp, t = sum + y; comp = (t - sum) - y; sum = t; }
  for (int x = 0; x < 4; x++) meanp[r * 4 + x] = (sum != 0.0) ? m[x] / sum : 0.25;
}

__global__ void ccf_kernel(const double *__restrict__ nseff, const double *__restrict__ pm, const double *__restrict__ meanp,
                           int L, int Lp, double *__restrict__ out)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, r = blockIdx.z;
  if (j >= L || i >= j) return;
  const double ne = nseff[((size_t) r * L + i) * Lp + j];
  double acc = 0.0;
  #pragma unroll
  for (int x = 0; x < 4; x++)
    #pragma unroll
    for (int y = 0; y < 4; y++) {
      const double cc = (ne * pm[((size_t) r * L + i) * 4 + x] - meanp[r * 4 + x]) * (ne * pm[((size_t) r * L + j) * 4 + y] - meanp[r * 4 + y]);
      acc += cc * cc;
    }
  out[((size_t) r * L + i) * Lp + j] = sqrt(acc);
}

// row/column partial sums + min/max of an already computed upper-triangle matrix (RAF, RAFS, CCF)
__global__ void __launch_bounds__(ST_TJ)
reduce_cov_kernel(const double *__restrict__ cov, int L, int Lp, double *__restrict__ rowpart, double *__restrict__ colpart,
                  double *__restrict__ mm, int nJT, int nIT)
{
  __shared__ double rowacc[ST_TJ / 32][ST_TI];
  __shared__ double smin[ST_TJ / 32], smax[ST_TJ / 32];
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tile_live = (it * ST_TI) < (jt * ST_TJ + ST_TJ - 1);
  double col = 0.0, vmin = INFINITY, vmax = -INFINITY;
  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    double v = 0.0;
    if (tile_live && i < L && j < L && i < j) {
      v = cov[((size_t) r * L + i) * Lp + j];
      col += v; vmin = fmin(vmin, v); vmax = fmax(vmax, v);
    }
    const double rs = warp_sum(v);
    if (lane == 0) rowacc[warp][il] = rs;
  }
  if (j < L) colpart[((size_t) r * nIT + it) * L + j] = col;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
  __syncthreads();
  if (threadIdx.x < ST_TI) {
    const int i = it * ST_TI + threadIdx.x;
    if (i < L) {
      double v = 0.0;
      for (int w = 0; w < ST_TJ / 32; w++) v += rowacc[w][threadIdx.x];
      rowpart[((size_t) r * nJT + jt) * L + i] = v;
    }
  }
  if (threadIdx.x == 0) {
    double a = smin[0], b = smax[0];
    for (int w = 1; w < ST_TJ / 32; w++) { a = fmin(a, smin[w]); b = fmax(b, smax[w]); }
    double *o = mm + (((size_t) r * nIT + it) * nJT + jt) * 2;
    o[0] = a; o[1] = b;
  }
}

// pp / nseff / ngap / ps in the reference's host layout for the real alignment (mutual_s fields):
// pp[i][j][16] both triangles (:1761-1763), nseff mirrored (:1765), ngap upper only (quirk Q4).
__global__ void export_probs_kernel(const long long *__restrict__ cnt, int L, int Lp, double scale, long long wtot,
                                    double *__restrict__ pp, double *__restrict__ nseff, double *__restrict__ ngap)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L) return;
  if (i == j) {
    for (int k = 0; k < 16; k++) pp[((size_t) i * L + j) * 16 + k] = 0.0;
    nseff[(size_t) i * L + j] = 0.0; ngap[(size_t) i * L + j] = 0.0;
    return;
  }
  if (i > j) { ngap[(size_t) i * L + j] = 0.0; return; }
  PairProbs P;
  load_pair<true>(cnt, (size_t) L * Lp, (size_t) i * Lp + j, scale, wtot, P);
  #pragma unroll
  for (int a = 0; a < 4; a++)
    #pragma unroll
    for (int b = 0; b < 4; b++) {
      pp[((size_t) i * L + j) * 16 + a * 4 + b] = P.pp[a * 4 + b];
      pp[((size_t) j * L + i) * 16 + b * 4 + a] = P.pp[a * 4 + b];
    }
  nseff[(size_t) i * L + j] = P.ne;
  nseff[(size_t) j * L + i] = P.ne;
  ngap[(size_t) i * L + j]  = P.ng;
}

// ps[i][a] = (1e-5 + colsum) / sum, a = 0..4 (:1792-1805)
__global__ void ps_kernel(const unsigned long long *__restrict__ colsum, int L, double scale, double *__restrict__ ps)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double p[5], sum = 0.0, comp = 0.0;
  for (int a = 0; a < 5; a++) {
    p[a] = 1e-5 + (double) colsum[(size_t) i * 5 + a] * scale;
    const double y = p[a] - comp, t = sum + y; comp = (t - sum) - y; sum = t;
  }
  for (int a = 0; a < 5; a++) ps[(size_t) i * 5 + a] = (sum != 0.0) ? p[a] / sum : 0.2;
}

} // namespace

// ---------------------------------------------------------------------------------------------- launchers
void rsb_stat_grid(int L, int *nJT, int *nIT) { *nJT = (L + ST_TJ - 1) / ST_TJ; *nIT = (L + ST_TI - 1) / ST_TI; }

// phase: 1 = partial sums only (-> msum), 2 = normalise msum -> pm, 3 = both
cudaError_t rsb_launch_marginals(const long long *cnt, int nrep, int L, int Lp, double scale, long long wtot, double tol,
                                 double *rowpart, double *colpart, double *nseff, double *msum, double *pm, int *flags,
                                 int sr, int sw, int phase, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  if (phase & 1) {
    rsb_coreside(marg_partial_kernel); marg_partial_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(cnt, L, Lp, scale, wtot, rowpart, colpart, nseff, nJT, nIT, sr, sw);
    rsb_coreside(marg_sum_kernel); marg_sum_kernel<<<dim3((L + 7) / 8, nrep), 256, 0, st>>>(rowpart, colpart, L, nJT, nIT, msum);
  }
  rsb_coreside(marg_norm_kernel);
  if (phase & 2) marg_norm_kernel<<<dim3((L + 127) / 128, nrep), 128, 0, st>>>(msum, L, tol, pm, flags);
  return cudaGetLastError();
}

#define RSB_STAT_CASE(STAT, CLS) \
  rsb_coreside(stat_kernel<STAT, CLS>); stat_kernel<STAT, CLS><<<grid, ST_TJ, 0, st>>>(cnt, pm, L, Lp, scale, wtot, mask, cov, rowpart, colpart, mm, nJT, nIT, sr, sw); break;

cudaError_t rsb_launch_statistic(int stat, int cls, const long long *cnt, const double *pm, int nrep, int L, int Lp, double scale,
                                 long long wtot, unsigned mask, double *cov, double *rowpart, double *colpart, double *mm, int sr, int sw, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  dim3 grid(nJT, nIT, nrep);
  const int key = stat * 4 + cls;
  switch (key) {
  case RSB_CHI  * 4 + RSB_C16: RSB_STAT_CASE(RSB_CHI,  RSB_C16)
  case RSB_CHI  * 4 + RSB_C2:  RSB_STAT_CASE(RSB_CHI,  RSB_C2)
  case RSB_OMES * 4 + RSB_C16: RSB_STAT_CASE(RSB_OMES, RSB_C16)
  case RSB_OMES * 4 + RSB_C2:  RSB_STAT_CASE(RSB_OMES, RSB_C2)
  case RSB_GT   * 4 + RSB_C16: RSB_STAT_CASE(RSB_GT,   RSB_C16)
  case RSB_GT   * 4 + RSB_C2:  RSB_STAT_CASE(RSB_GT,   RSB_C2)
  case RSB_GT   * 4 + RSB_CWC: RSB_STAT_CASE(RSB_GT,   RSB_CWC)
  case RSB_MI   * 4 + RSB_C16: RSB_STAT_CASE(RSB_MI,   RSB_C16)
  case RSB_MI   * 4 + RSB_C2:  RSB_STAT_CASE(RSB_MI,   RSB_C2)
  case RSB_MIr  * 4 + RSB_C16: RSB_STAT_CASE(RSB_MIr,  RSB_C16)
  case RSB_MIr  * 4 + RSB_C2:  RSB_STAT_CASE(RSB_MIr,  RSB_C2)
  case RSB_MIg  * 4 + RSB_C16: RSB_STAT_CASE(RSB_MIg,  RSB_C16)
  case RSB_MIg  * 4 + RSB_C2:  RSB_STAT_CASE(RSB_MIg,  RSB_C2)
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t rsb_launch_raf(const long long *cnt, int nrep, int L, int Lp, int nseq, unsigned mask, int smooth, double *tmp, double *cov,
                           double *rowpart, double *colpart, double *mm, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  raf_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(cnt, L, Lp, nseq, mask, smooth ? tmp : cov);
  if (smooth) rafs_kernel<<<dim3((L + 127) / 128, L, nrep), 128, 0, st>>>(tmp, L, Lp, cov);
  reduce_cov_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(cov, L, Lp, rowpart, colpart, mm, nJT, nIT);
  return cudaGetLastError();
}

cudaError_t rsb_launch_ccf(const double *nseff, const double *pm, int nrep, int L, int Lp, double *part, double *meanp, double *cov,
                           double *rowpart, double *colpart, double *mm, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  ccf_meanp_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(nseff, pm, L, Lp, part, nJT, nIT);
  ccf_meanp_final_kernel<<<nrep, 32, 0, st>>>(part, nJT * nIT, meanp);
  ccf_kernel<<<dim3((L + 127) / 128, L, nrep), 128, 0, st>>>(nseff, pm, meanp, L, Lp, cov);
  reduce_cov_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(cov, L, Lp, rowpart, colpart, mm, nJT, nIT);
  return cudaGetLastError();
}

cudaError_t rsb_launch_reduce_cov(const double *cov, int nrep, int L, int Lp, double *rowpart, double *colpart, double *mm, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  reduce_cov_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(cov, L, Lp, rowpart, colpart, mm, nJT, nIT);
  return cudaGetLastError();
}

cudaError_t rsb_launch_export_probs(const long long *cnt, int L, int Lp, double scale, long long wtot, double *pp, double *nseff,
                                    double *ngap, cudaStream_t st)
{
  export_probs_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(cnt, L, Lp, scale, wtot, pp, nseff, ngap);
  return cudaGetLastError();
}

cudaError_t rsb_launch_ps(const unsigned long long *colsum, int L, double scale, double *ps, cudaStream_t st)
{
  ps_kernel<<<(L + 127) / 128, 128, 0, st>>>(colsum, L, scale, ps);
  return cudaGetLastError();
}

#ifdef RSB_BLOCKTRACE
extern "C" void rsb_trace_set_stats(unsigned long long *buf) { cudaMemcpyToSymbol(stats_trace_buf, &buf, sizeof(buf)); }
#endif
