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

__global__ void logtab_kernel(double2 *tab)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= LOGTAB_N) return;
  const double ic = 1.0 / (1.0 + (k + 0.5) * (1.0 / LOGTAB_N));
  tab[k] = make_double2(ic, -log(ic));
}

__device__ __forceinline__ void logtab_load(double2 *tab, const double2 *__restrict__ gtab)
{
  for (int k = threadIdx.x; k < LOGTAB_N; k += blockDim.x) tab[k] = gtab[k];
}

// raw (unnormalised) pair table for the statistics that can work on it: x_k = 1e-10 + c_k scale, T = sum x, ne
__device__ __forceinline__ void load_pair_raw(const long long *__restrict__ cnt, size_t plane, size_t off,
                                              double scale, long long wtot, double *x, double &ne)
{
  unsigned long long c[16];
  #pragma unroll
  for (int k = 0; k < 16; k++) c[k] = (unsigned long long) cnt[k * plane + off];
  unsigned long long n = 0;
  #pragma unroll
  for (int k = 0; k < 16; k++) n += c[k];
  if (((unsigned long long) wtot >> 52) == 0) {                      // uniform over the grid
    ne = u52_to_f64(n) * scale;
    #pragma unroll
    for (int k = 0; k < 16; k++) x[k] = fma(u52_to_f64(c[k]), scale, 1e-10);
  } else {
    ne = u64_to_f64(n) * scale;
    #pragma unroll
    for (int k = 0; k < 16; k++) x[k] = fma(u64_to_f64(c[k]), scale, 1e-10);
  }
}

// G test, 16 classes (corr_CalculateGT_C16, :383-387): G = 2 sum obs log(obs/exp), obs = ne pp, exp = ne pm_i pm_j.
// ne cancels inside the log and sum pp = 1, so with the raw table x (pp = x / T), X_a = sum_b x_ab, Y_b = sum_a x_ab
//     G = 2 ne [ (sum x log x - sum_a X_a log pm_i[a] - sum_b Y_b log pm_j[b]) / T - log T ]
// : 17 logs, no division per cell, no branch.  The terms kept are the reference's (exp > 0 and obs > 0 <=> ne > 0 and
// pm > 0; pp > 0 always because of the prior); the rounding differs at the 1e-15 level.  x >= 1e-10 and T >= 1.6e-9 are
// positive normal numbers, so the log needs no clamp.
__device__ __forceinline__ double gt_c16_raw(const double *x, double ne, const double *lmi, const double *lmj,
                                             const double2 *__restrict__ tab)
{
  double X[4], Y[4];
  #pragma unroll
  for (int a = 0; a < 4; a++) X[a] = (x[a * 4] + x[a * 4 + 1]) + (x[a * 4 + 2] + x[a * 4 + 3]);
  #pragma unroll
  for (int b = 0; b < 4; b++) Y[b] = (x[b] + x[4 + b]) + (x[8 + b] + x[12 + b]);
  const double T = (X[0] + X[1]) + (X[2] + X[3]);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;                     // four independent chains
  #pragma unroll
  for (int k = 0; k < 16; k += 4) {
    s0 = fma(x[k],     fast_log<false>(x[k],     tab), s0);
    s1 = fma(x[k + 1], fast_log<false>(x[k + 1], tab), s1);
    s2 = fma(x[k + 2], fast_log<false>(x[k + 2], tab), s2);
    s3 = fma(x[k + 3], fast_log<false>(x[k + 3], tab), s3);
  }
  double m = 0.0;
  #pragma unroll
  for (int a = 0; a < 4; a++) m = fma(X[a], lmi[a], m);
  #pragma unroll
  for (int b = 0; b < 4; b++) m = fma(Y[b], lmj[b], m);
  const double v = (((s0 + s1) + (s2 + s3)) - m) / T - fast_log<false>(T, tab);
  return (ne > 0.0) ? 2.0 * ne * v : 0.0;
}

__device__ __forceinline__ double warp_sum(double v)
{
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// marginals.  The per-tile partial sums are produced by the epilogue of the tcgen05 kernel (gram_tcgen05.cu):
//   mrow[r][jb][i][4]          sum over the tile's columns j of sum_b pp_ij[a][b]     (column i as the left partner)
//   mcol[r][ib*4 + w][j][4]    sum over the 8 rows i of epilogue warp w of sum_a pp_ij[a][b]
// ------------------------------------------------------------------------------------------------
// unnormalised marginal sums msum[c][x] = sum of the partials of the tiles that exist, in a fixed order: one warp per
// column, lanes stride over the partial blocks, fixed xor-tree over the lanes (deterministic).  With the pair grid
// sharded over ranks these are the vectors that are summed across ranks before normalisation.
__global__ void __launch_bounds__(256)
marg_sum_kernel(const double *__restrict__ mrow, const double *__restrict__ mcol, int L, int CJ, int nJB, int nIB,
                int sr, int sw, double *__restrict__ msum, int E)
{
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), r = blockIdx.y;
  if (c >= L) return;
  const int ib_c = c / RSB_ICOLS, jb_c = c / CJ;
  const int nR = nJB * E;                                             // row-partial blocks: E per tile column block (one per epilogue group)
  double m[4] = { 0, 0, 0, 0 };
  for (int k = lane; k < nR + 4 * nIB; k += 32) {
    const double *src;
    if (k < nR) {
      if (!rsb_tile_exists(ib_c, k / E, CJ, L, sr, sw)) continue;
      src = mrow + (((size_t) r * nR + k) * L + c) * 4;
    } else {
      if (!rsb_tile_exists((k - nR) >> 2, jb_c, CJ, L, sr, sw)) continue;
      src = mcol + (((size_t) r * 4 * nIB + (k - nR)) * L + c) * 4;
    }
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
    for (int a = 0; a < 4; a++) msum[((size_t) r * L + c) * 4 + a] = m[a];
  }
}

// nseff[i][j] for the statistics that read it (CCF)
__global__ void nseff_kernel(const long long *__restrict__ cnt, int L, int Lp, double scale, double *__restrict__ nseff)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, r = blockIdx.z;
  if (j >= L || i >= j) return;
  const size_t plane = (size_t) L * Lp;
  const long long *c = cnt + (size_t) r * 16 * plane + (size_t) i * Lp + j;
  unsigned long long ne = 0;
  #pragma unroll
  for (int k = 0; k < 16; k++) ne += (unsigned long long) c[k * plane];
  nseff[((size_t) r * L + i) * Lp + j] = u64_to_f64(ne) * scale;
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
  // lmi / lmj: log of the marginals, taken once per column by the caller.  (GT x C16 has its own form, gt_c16_raw.)
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
      else if (STAT == RSB_GT) { const double t = ob * fast_log(ob / fmax(ex, 2.2250738585072014e-308), tab); v += (ex > 0. && ob > 0.) ? t : 0.0; }
      else {
        const double lp = fast_log(pxy, tab);
        const double t  = pxy * (lp - lmi[x] - lmj[y]);
        if (STAT == RSB_MIr) H -= (pxy > 0.0) ? pxy * lp : 0.0;
        v += (pxy > 0.0 && mi[x] > 0.0 && mj[y] > 0.0) ? t : 0.0;
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
stat_kernel(const long long *__restrict__ cnt, const double *__restrict__ pm, const double2 *__restrict__ gtab, int L, int Lp,
            double scale, long long wtot, unsigned mask, double *__restrict__ cov, double *__restrict__ rowpart, double *__restrict__ colpart,
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
  if (tile_live) logtab_load(tab, gtab);
  double mj[4] = { 0.25, 0.25, 0.25, 0.25 }, lmj[4];

  if (threadIdx.x < ST_TI * 4) {
    const int il = threadIdx.x >> 2, a = threadIdx.x & 3, i = it * ST_TI + il;
    pmi[il][a]  = (i < L) ? pm[((size_t) r * L + i) * 4 + a] : 0.25;
    lpmi[il][a] = (pmi[il][a] > 0.0) ? log(pmi[il][a]) : 0.0;      // pm > 0 always (prior); keep the logs finite regardless
  }
  if (j < L) {
    #pragma unroll
    for (int b = 0; b < 4; b++) mj[b] = pm[((size_t) r * L + j) * 4 + b];
  }
  #pragma unroll
  for (int b = 0; b < 4; b++) lmj[b] = (mj[b] > 0.0) ? log(mj[b]) : 0.0;
  __syncthreads();

  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    double v = 0.0;
    if (tile_live && i < L && j < L && i < j) {
      if (STAT == RSB_GT && CLS == RSB_C16) {
        double x[16], ne;
        load_pair_raw(c, plane, (size_t) i * Lp + j, scale, wtot, x, ne);
        v = gt_c16_raw(x, ne, lpmi[il], lmj, tab);
      } else {
        PairProbs P;
        load_pair<false>(c, plane, (size_t) i * Lp + j, scale, wtot, P);
        v = pair_statistic<STAT, CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
      }
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

// The six 16-class statistics of one pair from ONE set of 16 logs and no per-cell division (multi_stat_kernel).  With
// d = pp - pm_i pm_j per cell (obs - exp = ne d, exp = ne pm_i pm_j; src/correlators.c:150-159, 283-292, 383-387, 583-590, 712-722, 842-849):
//     CHI = ne sum d^2 / (pm_i pm_j)    OMES = ne sum d^2    MI = sum pp (log pp - log pm_i - log pm_j)    GT = 2 ne MI
//     MIr = MI / H (H = -sum pp log pp > 1e-2, else 0)      MIg = MI - ngap / ne
// rmi / rmj are the reciprocals of the marginals, taken once per column like their logs.  The guards of the reference (exp > 0,
// obs > 0, pp > 0, pm > 0) reduce to ne > 0: pp and pm are positive because of the 1e-10 prior.  The values differ from
// pair_statistic's (the reference's operation order) by rounding only, ~1e-15 relative; tests/test_gpu_multi.py bounds it.
__device__ __forceinline__ void multi_c16(const PairProbs &P, const double *mi, const double *mj, const double *rmi, const double *rmj,
                                          const double *lmi, const double *lmj, const double2 *__restrict__ tab, double *o)
{
  double chi = 0.0, om0 = 0.0, om1 = 0.0, m0 = 0.0, m1 = 0.0, h0 = 0.0, h1 = 0.0;
  #pragma unroll
  for (int x = 0; x < 4; x++) {
    double cx = 0.0;
    #pragma unroll
    for (int y = 0; y < 4; y++) {
      const double pxy = P.pp[x * 4 + y];
      const double d   = fma(-mi[x], mj[y], pxy);
      const double d2  = d * d;
      cx = fma(d2, rmj[y], cx);
      const double lp = fast_log<false>(pxy, tab);
      const double t  = (lp - lmi[x]) - lmj[y];
      if (y & 1) { om1 += d2; m1 = fma(pxy, t, m1); h1 = fma(pxy, lp, h1); }
      else       { om0 += d2; m0 = fma(pxy, t, m0); h0 = fma(pxy, lp, h0); }
    }
    chi = fma(cx, rmi[x], chi);
  }
  const double mi_v = m0 + m1, H = -(h0 + h1);
  const bool live = P.ne > 0.0;
  o[0] = live ? P.ne * chi : 0.0;
  o[1] = live ? P.ne * (om0 + om1) : 0.0;
  o[2] = live ? 2.0 * P.ne * mi_v : 0.0;
  o[3] = mi_v;
  o[4] = (H > 1e-2) ? mi_v / H : 0.0;
  o[5] = mi_v - (live ? P.ng / P.ne : 0.0);
}

// Several statistics of one (weighted) count table in one pass -- BASELINE config 5 sweeps MI, MIr, MIg, CHI, OMES and GT over the
// same alignments, and cov_Calculate's dispatch (src/covariation.c:100-258) differs only in which corr_Calculate* it calls on
// the probabilities of one corr_Probs.  The 128 B of counts per pair are read once and every requested raw statistic is written
// to its own matrix.  16 classes: multi_c16 (one set of logs for all six).  2 classes: the very pair_statistic code of the
// single-statistic kernel.  Row/column partials and the score range come from reduce_cov_kernel afterwards (same tiling, same
// summation order as stat_kernel's own).
struct MultiOut { double *cov[6]; size_t rep_stride; };             // CHI, OMES, GT, MI, MIr, MIg (NULL = not requested); doubles between replicates
template <int CLS>
__global__ void __launch_bounds__(ST_TJ)
multi_stat_kernel(const long long *__restrict__ cnt, const double *__restrict__ pm, const double2 *__restrict__ gtab, int L, int Lp,
                  double scale, long long wtot, unsigned mask, MultiOut out, int sr, int sw)
{
  __shared__ double pmi[ST_TI][4], lpmi[ST_TI][4], rpmi[ST_TI][4];
  __shared__ double2 tab[LOGTAB_N];
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const size_t plane = (size_t) L * Lp;
  const long long *c = cnt + (size_t) r * 16 * plane;
  const bool tile_live = (it * ST_TI) < (jt * ST_TJ + ST_TJ - 1) && RSB_OWNED(it, sr, sw);
  if (!tile_live) return;
  logtab_load(tab, gtab);
  double mj[4] = { 0.25, 0.25, 0.25, 0.25 }, lmj[4], rmj[4];
  if (threadIdx.x < ST_TI * 4) {
    const int il = threadIdx.x >> 2, a = threadIdx.x & 3, i = it * ST_TI + il;
    pmi[il][a]  = (i < L) ? pm[((size_t) r * L + i) * 4 + a] : 0.25;
    lpmi[il][a] = (pmi[il][a] > 0.0) ? log(pmi[il][a]) : 0.0;
    rpmi[il][a] = (pmi[il][a] > 0.0) ? 1.0 / pmi[il][a] : 0.0;
  }
  if (j < L) {
    #pragma unroll
    for (int b = 0; b < 4; b++) mj[b] = pm[((size_t) r * L + j) * 4 + b];
  }
  #pragma unroll
  for (int b = 0; b < 4; b++) { lmj[b] = (mj[b] > 0.0) ? log(mj[b]) : 0.0; rmj[b] = (mj[b] > 0.0) ? 1.0 / mj[b] : 0.0; }
  __syncthreads();
  if (j >= L) return;
  #pragma unroll 1
  for (int il = 0; il < ST_TI; il++) {
    const int i = it * ST_TI + il;
    if (i >= L || i >= j) continue;
    const size_t off = (size_t) i * Lp + j, o = (size_t) r * out.rep_stride + off;
    PairProbs P;
    load_pair<false>(c, plane, off, scale, wtot, P);
    if (CLS == RSB_C16) {                                 // all six from one set of logs
      double v[6];
      multi_c16(P, pmi[il], mj, rpmi[il], rmj, lpmi[il], lmj, tab, v);
      #pragma unroll
      for (int k = 0; k < 6; k++) if (out.cov[k]) out.cov[k][o] = v[k];
      continue;
    }
    if (out.cov[0]) out.cov[0][o] = pair_statistic<RSB_CHI,  CLS == RSB_CWC ? RSB_C16 : CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
    if (out.cov[1]) out.cov[1][o] = pair_statistic<RSB_OMES, CLS == RSB_CWC ? RSB_C16 : CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
    if (out.cov[2]) {
      if (CLS == RSB_C16) {
        double x[16], ne;
        load_pair_raw(c, plane, off, scale, wtot, x, ne);
        out.cov[2][o] = gt_c16_raw(x, ne, lpmi[il], lmj, tab);
      } else out.cov[2][o] = pair_statistic<RSB_GT, CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
    }
    if (out.cov[3]) out.cov[3][o] = pair_statistic<RSB_MI,  CLS == RSB_CWC ? RSB_C16 : CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
    if (out.cov[4]) out.cov[4][o] = pair_statistic<RSB_MIr, CLS == RSB_CWC ? RSB_C16 : CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
    if (out.cov[5]) out.cov[5][o] = pair_statistic<RSB_MIg, CLS == RSB_CWC ? RSB_C16 : CLS>(P, pmi[il], mj, lpmi[il], lmj, mask, tab);
  }
}

// G test on 16 classes from the per-pair records left by the record epilogue of the tcgen05 kernel (gram_tcgen05.cu): with
// l_i[a] = log pm_i[a],
//     G_ij = A - [ N2 l_i[3] + sum_{a<3} U_a (l_i[a] - l_i[3]) ] - [ N2 l_j[3] + sum_{b<3} V_b (l_j[b] - l_j[3]) ]
// (corr_CalculateGT_C16, src/correlators.c:383-387; pairs with nseff = 0 carry an all-zero record and score 0, as the
// reference's `exp > 0 && obs > 0` guard leaves them).  64 B read + 8 B written per pair and 14 multiply-adds: HBM-bound, so
// unlike stat_kernel it can run beside the contraction.  Same tiling, partial sums and min/max as stat_kernel.
#ifndef RSB_FIN_U
#define RSB_FIN_U 2             // rows whose 8 record planes a thread loads together; 2 -> <= 64 registers, 8 blocks per SM: the ~960 live tiles
#define RSB_FIN_BLOCKS 8        // of the SSU shape fit in ONE wave (4 rows / 5 blocks: 1.3 waves, the second nearly empty)
#endif
__global__ void __launch_bounds__(ST_TJ, RSB_FIN_BLOCKS)
gt_finish_kernel(const double *__restrict__ rec, const double *__restrict__ pm, int L, int Lp, double *__restrict__ cov,
                 double *__restrict__ rowpart, double *__restrict__ colpart, double *__restrict__ mm, int nJT, int nIT, int sr, int sw,
                 size_t slot_stride)
{
  __shared__ double rowacc[ST_TJ / 32][ST_TI];
  __shared__ double li[ST_TI][4];                                    // { l_i[0]-l_i[3], l_i[1]-l_i[3], l_i[2]-l_i[3], l_i[3] }
  __shared__ double smin[ST_TJ / 32], smax[ST_TJ / 32];
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * ST_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t plane = (size_t) L * Lp;
  const double *R = rec + (size_t) r * slot_stride;
  const bool tile_live = (it * ST_TI) < (jt * ST_TJ + ST_TJ - 1) && RSB_OWNED(it, sr, sw);
  double col = 0.0, vmin = INFINITY, vmax = -INFINITY;

  if (threadIdx.x < ST_TI) {
    const int i = it * ST_TI + threadIdx.x;
    double l[4];
    #pragma unroll
    for (int a = 0; a < 4; a++) { const double p = (i < L) ? pm[((size_t) r * L + i) * 4 + a] : 0.25; l[a] = (p > 0.0) ? log(p) : 0.0; }
    li[threadIdx.x][0] = l[0] - l[3]; li[threadIdx.x][1] = l[1] - l[3]; li[threadIdx.x][2] = l[2] - l[3]; li[threadIdx.x][3] = l[3];
  }
  double lj[4] = { 0.0, 0.0, 0.0, 0.0 };
  if (tile_live && j < L) {
    double l[4];
    #pragma unroll
    for (int b = 0; b < 4; b++) { const double p = pm[((size_t) r * L + j) * 4 + b]; l[b] = (p > 0.0) ? log(p) : 0.0; }
    lj[0] = l[0] - l[3]; lj[1] = l[1] - l[3]; lj[2] = l[2] - l[3]; lj[3] = l[3];
  }
  __syncthreads();

  // the kernel runs beside the persistent tcgen05 kernel, at a few blocks per SM: the loads of FIN_U rows (8 planes each) are
  // issued together so that the few resident warps keep enough bytes in flight
  constexpr int FIN_U = RSB_FIN_U;
  #pragma unroll 1
  for (int il0 = 0; il0 < ST_TI; il0 += FIN_U) {
    double q[FIN_U][8];
    bool ok[FIN_U];
    #pragma unroll
    for (int u = 0; u < FIN_U; u++) {
      const int i = it * ST_TI + il0 + u;
      ok[u] = tile_live && i < L && j < L && i < j;
      const double *p = R + (size_t) i * Lp + j;
      #pragma unroll
      for (int k = 0; k < 8; k++) q[u][k] = ok[u] ? __ldcs(p + k * plane) : 0.0;
    }
    #pragma unroll
    for (int u = 0; u < FIN_U; u++) {
      const int il = il0 + u, i = it * ST_TI + il;
      double v = 0.0;
      if (ok[u]) {
        double m = q[u][1] * (li[il][3] + lj[3]);
        m = fma(q[u][2], li[il][0], m); m = fma(q[u][3], li[il][1], m); m = fma(q[u][4], li[il][2], m);
        m = fma(q[u][5], lj[0], m);     m = fma(q[u][6], lj[1], m);     m = fma(q[u][7], lj[2], m);
        v = q[u][0] - m;
        cov[((size_t) r * L + i) * Lp + j] = v;
        col += v;
        vmin = fmin(vmin, v);
        vmax = fmax(vmax, v);
      }
      const double rs = warp_sum(v);
      if (lane == 0) rowacc[warp][il] = rs;
    }
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
}

// nseff / ngap in the host layout from a slot that holds records instead of counts (quirk Q3, rsb_last_nseff): nseff = N2 / 2
// (exact), ngap = wtot scale - nseff
__global__ void export_nseff_rec_kernel(const double *__restrict__ rec, int L, int Lp, double wtot_scaled, double *__restrict__ nseff, double *__restrict__ ngap)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L) return;
  if (i >= j) { ngap[(size_t) i * L + j] = 0.0; if (i == j) nseff[(size_t) i * L + j] = 0.0; return; }
  const double ne = 0.5 * rec[(size_t) L * Lp + (size_t) i * Lp + j];
  nseff[(size_t) i * L + j] = ne;
  nseff[(size_t) j * L + i] = ne;
  ngap[(size_t) i * L + j]  = wtot_scaled - ne;
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

// phase: 1 = sum the tile partials (-> msum), 2 = normalise msum -> pm, 3 = both
cudaError_t rsb_launch_marginals(const double *mrow, const double *mcol, int nrep, int L, int CJ, int nJB, int nIB, double tol,
                                 double *msum, double *pm, int *flags, int sr, int sw, int phase, int mrow_blocks, cudaStream_t st)
{
  if (phase & 1) {
    rsb_coreside(marg_sum_kernel); marg_sum_kernel<<<dim3((L + 7) / 8, nrep), 256, 0, st>>>(mrow, mcol, L, CJ, nJB, nIB, sr, sw, msum, mrow_blocks);
  }
  rsb_coreside(marg_norm_kernel);
  if (phase & 2) marg_norm_kernel<<<dim3((L + 127) / 128, nrep), 128, 0, st>>>(msum, L, tol, pm, flags);
  return cudaGetLastError();
}

cudaError_t rsb_launch_nseff(const long long *cnt, int nrep, int L, int Lp, double scale, double *nseff, cudaStream_t st)
{
  nseff_kernel<<<dim3((L + 127) / 128, L, nrep), 128, 0, st>>>(cnt, L, Lp, scale, nseff);
  return cudaGetLastError();
}

#define RSB_STAT_CASE(STAT, CLS) \
  rsb_coreside(stat_kernel<STAT, CLS>); stat_kernel<STAT, CLS><<<grid, ST_TJ, 0, st>>>(cnt, pm, logtab, L, Lp, scale, wtot, mask, cov, rowpart, colpart, mm, nJT, nIT, sr, sw); break;

cudaError_t rsb_launch_logtab(void *tab, cudaStream_t st)
{
  logtab_kernel<<<(LOGTAB_N + 127) / 128, 128, 0, st>>>((double2 *) tab);
  return cudaGetLastError();
}
size_t rsb_logtab_bytes() { return sizeof(double2) * LOGTAB_N; }

cudaError_t rsb_launch_statistic(int stat, int cls, const long long *cnt, const double *pm, const void *logtab_, int nrep, int L, int Lp, double scale,
                                 long long wtot, unsigned mask, double *cov, double *rowpart, double *colpart, double *mm, int sr, int sw, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  dim3 grid(nJT, nIT, nrep);
  const double2 *logtab = (const double2 *) logtab_;
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

// cov6[k] (k = CHI, OMES, GT, MI, MIr, MIg; NULL = skip): raw statistic matrices [L][Lp] of the counts in cnt, rep_stride doubles apart per replicate
cudaError_t rsb_launch_multi_statistic(int cls, const long long *cnt, const double *pm, const void *logtab_, int nrep, int L, int Lp, double scale,
                                       long long wtot, unsigned mask, double *const *cov6, size_t rep_stride, int sr, int sw, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  dim3 grid(nJT, nIT, nrep);
  const double2 *logtab = (const double2 *) logtab_;
  MultiOut out;
  for (int k = 0; k < 6; k++) out.cov[k] = cov6[k];
  out.rep_stride = rep_stride;
  switch (cls) {
  case RSB_C16: rsb_coreside(multi_stat_kernel<RSB_C16>); multi_stat_kernel<RSB_C16><<<grid, ST_TJ, 0, st>>>(cnt, pm, logtab, L, Lp, scale, wtot, mask, out, sr, sw); break;
  case RSB_C2:  rsb_coreside(multi_stat_kernel<RSB_C2>);  multi_stat_kernel<RSB_C2><<<grid, ST_TJ, 0, st>>>(cnt, pm, logtab, L, Lp, scale, wtot, mask, out, sr, sw); break;
  case RSB_CWC:                                                       // only the G test is defined on the Watson-Crick cells (correlators.c:72)
    for (int k = 0; k < 6; k++) if (k != 2 && cov6[k]) return cudaErrorInvalidValue;
    rsb_coreside(multi_stat_kernel<RSB_CWC>); multi_stat_kernel<RSB_CWC><<<grid, ST_TJ, 0, st>>>(cnt, pm, logtab, L, Lp, scale, wtot, mask, out, sr, sw); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t rsb_launch_gt_finish(const double *rec, size_t slot_stride, const double *pm, int nrep, int L, int Lp, double *cov,
                                 double *rowpart, double *colpart, double *mm, int sr, int sw, cudaStream_t st)
{
  int nJT, nIT; rsb_stat_grid(L, &nJT, &nIT);
  rsb_coreside(gt_finish_kernel);
  gt_finish_kernel<<<dim3(nJT, nIT, nrep), ST_TJ, 0, st>>>(rec, pm, L, Lp, cov, rowpart, colpart, mm, nJT, nIT, sr, sw, slot_stride);
  return cudaGetLastError();
}

cudaError_t rsb_launch_export_nseff_rec(const double *rec, int L, int Lp, double wtot_scaled, double *nseff, double *ngap, cudaStream_t st)
{
  export_nseff_rec_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(rec, L, Lp, wtot_scaled, nseff, ngap);
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
