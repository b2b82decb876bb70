// correct_hist.cu -- background correction (APC / ASC), score range, and the score histogram.
//
// Reference: corr_CalculateCOVCorrected (src/correlators.c:1064-1157); histogram fill of
// cov_SignificantPairs_Ranking (src/covariation.c:415-432) with Easel's bin rule
// b = ceil((x - bmin)/w - 1), value clamped to max(x, bmin + w); histogram width
// from the first null, calculate_width_histo (src/R-scape.c:1357-1360); cumulative null histogram,
// null_add2cumranklist (src/R-scape.c:1565-1612): with a common (bmin, w) every replicate's bin b
// is the cumulative histogram's bin b, so all replicates add into one device array.
#include "rsb_common.cuh"
#include <math.h>

namespace {

constexpr int CH_TI = RSB_TI;
constexpr int CH_TJ = RSB_TJ;
constexpr int CH_SMEM_BINS = 2048;        // 8 KB: several blocks fit beside the tcgen05 kernel's ring

// covsum[i] = sum_{j != i} COV[i][j] from the tile partials (one thread per column, fixed summation order); each block
// also leaves the sum of its row partials (= its share of sum_{i<j} COV) in blocksum[r][block].
__global__ void __launch_bounds__(128)
covsum_kernel(const double *__restrict__ rowpart, const double *__restrict__ colpart, int L, int nJT, int nIT,
              double *__restrict__ covsum, double *__restrict__ blocksum)
{
  __shared__ double red[128];
  const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  double rs = 0.0, cs = 0.0;
  if (i < L) {
    for (int jt = 0; jt < nJT; jt++) rs += rowpart[((size_t) r * nJT + jt) * L + i];
    for (int it = 0; it < nIT; it++) cs += colpart[((size_t) r * nIT + it) * L + i];
    covsum[(size_t) r * (L + 4) + i] = rs + cs;
  }
  red[threadIdx.x] = rs;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) blocksum[(size_t) r * gridDim.x + blockIdx.x] = red[0];
}

// covsum[r][L..L+2] = { sum_{i<j} COV, raw min, raw max } from the per-block / per-tile partials.  Together with covsum[0..L)
// this is the vector that is reduced across ranks when the pair grid is sharded (sum, sum, min, max).
__global__ void __launch_bounds__(256)
covtot_kernel(const double *__restrict__ blocksum, int nblk, const double *__restrict__ mm, int L, int ntiles, double *__restrict__ covsum)
{
  __shared__ double rmin[256], rmax[256];
  const int r = blockIdx.x;
  double a = INFINITY, b = -INFINITY;
  for (int k = threadIdx.x; k < ntiles; k += blockDim.x) {
    a = fmin(a, mm[((size_t) r * ntiles + k) * 2]);
    b = fmax(b, mm[((size_t) r * ntiles + k) * 2 + 1]);
  }
  rmin[threadIdx.x] = a; rmax[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      rmin[threadIdx.x] = fmin(rmin[threadIdx.x], rmin[threadIdx.x + o]);
      rmax[threadIdx.x] = fmax(rmax[threadIdx.x], rmax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < nblk; k++) tot += blocksum[(size_t) r * nblk + k];        // fixed order
    double *o = covsum + (size_t) r * (L + 4) + L;
    o[0] = tot; o[1] = rmin[0]; o[2] = rmax[0]; o[3] = 0.0;
  }
}

// COVx[i] = covsum[i] / (L-1) (:1101-1108); COVavg = 2/(L(L-1)) sum_{i<j} COV (:1093-1098).  scal[r] = { COVavg, raw min, raw max, - }
__global__ void covx_kernel(const double *__restrict__ covsum, int L, double *__restrict__ covx, double *__restrict__ scal)
{
  const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const double *cs = covsum + (size_t) r * (L + 4);
  if (i < L) {
    double x = cs[i];
    if (L > 1) x /= (double) L - 1.;
    covx[(size_t) r * L + i] = x;
  }
  if (i == 0) {
    double avg = cs[L];
    if (L > 1) avg /= (double) L * ((double) L - 1.);
    avg *= 2.;
    scal[r * 4 + 0] = avg; scal[r * 4 + 1] = cs[L + 1]; scal[r * 4 + 2] = cs[L + 2]; scal[r * 4 + 3] = 0.0;
  }
}

__device__ __forceinline__ double corrected(int actype, double raw, double xi, double xj, double avg)
{
  if (actype == RSB_APC) return (avg != 0.0) ? raw - xi * xj / avg : 0.0;      // :1118
  if (actype == RSB_ASC) return raw - (xi + xj - avg);                         // :1120
  return raw;
}

// One pass over the upper triangle of the raw statistic.
//   mode bit 0: write the corrected score back (upper triangle; symmetrize_kernel mirrors it and sets the diagonal)
//   mode bit 1: add max(x, bmin+w) into the histogram (w read from *wptr so that it can come from the
//               device-side width computation without a host round trip)
// Always: per-block min/max of the corrected score, NaN flag (:1124).
__global__ void __launch_bounds__(CH_TJ, 6)
correct_hist_kernel(double *__restrict__ cov, const double *__restrict__ covx, const double *__restrict__ scal, int L, int Lp,
                    int actype, int mode, double bmin, const double *__restrict__ wptr, unsigned long long *__restrict__ hist,
                    int nbins, double *__restrict__ mm, int *__restrict__ flags, int nJT, int nIT, int sr, int sw,
                    const int *__restrict__ m2p, int mind)
{
  __shared__ unsigned int sh[CH_SMEM_BINS];
  __shared__ double smin[CH_TJ / 32], smax[CH_TJ / 32];
  const int jt = blockIdx.x, it = blockIdx.y, r = blockIdx.z;
  const int j  = jt * CH_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tile_live = (it * CH_TI) < (jt * CH_TJ + CH_TJ - 1) && RSB_OWNED(it, sr, sw);
  const bool do_hist = (mode & 2) != 0;
  const double avg = scal[r * 4];
  const double w   = do_hist ? *wptr : 1.0;
  double vmin = INFINITY, vmax = -INFINITY;

  // shared-memory bins cover [b0, b0 + CH_SMEM_BINS), centred on the bin of score 0 (corrected null scores pile up
  // there, and with a small w -- MI-type statistics -- that bin lies far above bin 0); the rest goes to global atomics
  int b0 = 0;
  if (do_hist && w > 0.0) {
    const double c = ceil((-bmin) / w - 1.0) - (double) (CH_SMEM_BINS / 2);
    b0 = (c > 0.0 && c < (double) nbins) ? (int) c : 0;
  }
  __shared__ double sxi[CH_TI];                                       // COVx of the tile's rows
  if (do_hist) { for (int k = threadIdx.x; k < CH_SMEM_BINS; k += CH_TJ) sh[k] = 0; }
  if (threadIdx.x < CH_TI) { const int i = it * CH_TI + threadIdx.x; sxi[threadIdx.x] = (i < L) ? covx[(size_t) r * L + i] : 0.0; }
  __syncthreads();

  if (tile_live && j < L) {
    const double xj = covx[(size_t) r * L + j];
    const int mpj = m2p ? m2p[j] : -1;
    double *C = cov + (size_t) r * L * Lp;
    #pragma unroll 1
    for (int il0 = 0; il0 < CH_TI; il0 += 8) {
    double raw[8];                                                    // the loads of 8 rows first (few resident warps beside the contraction)
    #pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = it * CH_TI + il0 + u;
      raw[u] = (i < L && i < j) ? C[(size_t) i * Lp + j] : 0.0;
    }
    #pragma unroll
    for (int u = 0; u < 8; u++) {
      const int il = il0 + u, i = it * CH_TI + il;
      if (i >= L || i >= j) continue;
      const double v = corrected(actype, raw[u], sxi[il], xj, avg);
      if (isnan(v)) atomicOr(flags, 2);
      vmin = fmin(vmin, v);
      vmax = fmax(vmax, v);
      if (mode & 1) C[(size_t) i * Lp + j] = v;                     // mirrored by symmetrize_kernel
      // pairs closer than `mind` in the PDB sequence stay out of the histogram (covariation.c:421-427); min/max keep them
      const bool excl = m2p && m2p[i] >= 0 && mpj >= 0 && mpj - m2p[i] < mind;
      if (do_hist && w > 0.0 && !excl) {
        const double x = fmax(v, bmin + w);                           // ESL_MAX(cov, bmin+w), covariation.c:431
        const double bd = ceil(((x - bmin) / w) - 1.);                // esl_histogram_Score2Bin
        if (bd >= 0.0 && bd < (double) nbins) {
          const int b = (int) bd;
          if (b >= b0 && b < b0 + CH_SMEM_BINS) atomicAdd(&sh[b - b0], 1u);
          else                                  atomicAdd(&hist[b], 1ull);
        } else atomicOr(flags, 4);                                    // histogram capacity exceeded
      }
    }
    }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
  __syncthreads();
  if (do_hist)
    for (int k = threadIdx.x; k < CH_SMEM_BINS && b0 + k < nbins; k += CH_TJ)
      if (sh[k]) atomicAdd(&hist[b0 + k], (unsigned long long) sh[k]);
  if (threadIdx.x == 0) {
    double a = smin[0], b = smax[0];
    for (int q = 1; q < CH_TJ / 32; q++) { a = fmin(a, smin[q]); b = fmax(b, smax[q]); }
    double *o = mm + (((size_t) r * nIT + it) * nJT + jt) * 2;
    o[0] = a; o[1] = b;
  }
}

// Several (statistic, correction) combinations in ONE launch (rsb_null_hist_multi): blockIdx.z = r * ncombo + k.  Combination k
// reads the raw matrix of its statistic, virtual replicate q = r * nw + kidx[k] of the stacked buffers (cov, covx, scal), corrects with
// act[k], and adds into its own histogram hist + k * nbins with its own width w[k] (<= 0: score range only).  Body = correct_hist_kernel's.
struct ComboTab { int kidx[64]; int act[64]; };
__global__ void __launch_bounds__(CH_TJ, 6)
correct_hist_multi_kernel(const double *__restrict__ cov, const double *__restrict__ covx, const double *__restrict__ scal, int L, int Lp,
                          ComboTab tab, int ncombo, int nw, double bmin, const double *__restrict__ wptr, unsigned long long *__restrict__ hist,
                          int nbins, double *__restrict__ mm, int *__restrict__ flags, int nJT, int nIT, const int *__restrict__ m2p, int mind)
{
  __shared__ unsigned int sh[CH_SMEM_BINS];
  __shared__ double smin[CH_TJ / 32], smax[CH_TJ / 32];
  __shared__ double sxi[CH_TI];
  const int jt = blockIdx.x, it = blockIdx.y, z = blockIdx.z;
  const int r = z / ncombo, k = z % ncombo, q = r * nw + tab.kidx[k], actype = tab.act[k];
  const int j  = jt * CH_TJ + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tile_live = (it * CH_TI) < (jt * CH_TJ + CH_TJ - 1);
  const double w = wptr[k];
  const bool do_hist = w > 0.0;
  const double avg = scal[q * 4];
  unsigned long long *H = hist + (size_t) k * nbins;
  double vmin = INFINITY, vmax = -INFINITY;
  int b0 = 0;
  if (do_hist) {
    const double c = ceil((-bmin) / w - 1.0) - (double) (CH_SMEM_BINS / 2);
    b0 = (c > 0.0 && c < (double) nbins) ? (int) c : 0;
    for (int t = threadIdx.x; t < CH_SMEM_BINS; t += CH_TJ) sh[t] = 0;
  }
  if (threadIdx.x < CH_TI) { const int i = it * CH_TI + threadIdx.x; sxi[threadIdx.x] = (i < L) ? covx[(size_t) q * L + i] : 0.0; }
  __syncthreads();
  if (tile_live && j < L) {
    const double xj = covx[(size_t) q * L + j];
    const int mpj = m2p ? m2p[j] : -1;
    const double *C = cov + (size_t) q * L * Lp;
    #pragma unroll 1
    for (int il0 = 0; il0 < CH_TI; il0 += 8) {
      double raw[8];
      #pragma unroll
      for (int u = 0; u < 8; u++) { const int i = it * CH_TI + il0 + u; raw[u] = (i < L && i < j) ? C[(size_t) i * Lp + j] : 0.0; }
      #pragma unroll
      for (int u = 0; u < 8; u++) {
        const int il = il0 + u, i = it * CH_TI + il;
        if (i >= L || i >= j) continue;
        const double v = corrected(actype, raw[u], sxi[il], xj, avg);
        if (isnan(v)) atomicOr(flags, 2);
        vmin = fmin(vmin, v);
        vmax = fmax(vmax, v);
        const bool excl = m2p && m2p[i] >= 0 && mpj >= 0 && mpj - m2p[i] < mind;
        if (do_hist && !excl) {
          const double x = fmax(v, bmin + w);
          const double bd = ceil(((x - bmin) / w) - 1.);
          if (bd >= 0.0 && bd < (double) nbins) {
            const int b = (int) bd;
            if (b >= b0 && b < b0 + CH_SMEM_BINS) atomicAdd(&sh[b - b0], 1u);
            else                                  atomicAdd(&H[b], 1ull);
          } else atomicOr(flags, 4);
        }
      }
    }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
  __syncthreads();
  if (do_hist)
    for (int t = threadIdx.x; t < CH_SMEM_BINS && b0 + t < nbins; t += CH_TJ)
      if (sh[t]) atomicAdd(&H[b0 + t], (unsigned long long) sh[t]);
  if (threadIdx.x == 0) {
    double a = smin[0], b = smax[0];
    for (int t = 1; t < CH_TJ / 32; t++) { a = fmin(a, smin[t]); b = fmax(b, smax[t]); }
    double *o = mm + (((size_t) z * nIT + it) * nJT + jt) * 2;
    o[0] = a; o[1] = b;
  }
}

// out[r][0..1] = min/max over the block partials
__global__ void __launch_bounds__(256)
minmax_final_kernel(const double *__restrict__ mm, int nblocks, double *__restrict__ out)
{
  __shared__ double rmin[256], rmax[256];
  const int r = blockIdx.x;
  double a = INFINITY, b = -INFINITY;
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
    a = fmin(a, mm[((size_t) r * nblocks + k) * 2]);
    b = fmax(b, mm[((size_t) r * nblocks + k) * 2 + 1]);
  }
  rmin[threadIdx.x] = a; rmax[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { rmin[threadIdx.x] = fmin(rmin[threadIdx.x], rmin[threadIdx.x + o]); rmax[threadIdx.x] = fmax(rmax[threadIdx.x], rmax[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[r * 2] = rmin[0]; out[r * 2 + 1] = rmax[0]; }
}

// (min, max) pairs <-> (-min, max): one MAX all-reduce then serves both ends of the range
__global__ void negate_min_kernel(double *__restrict__ minmax, int n)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) minmax[2 * k] = -minmax[2 * k];
}

// scores of rows owned by other ranks -> 0, so that a SUM all-reduce assembles the whole upper triangle
__global__ void zero_unowned_rows_kernel(double *__restrict__ cov, int L, int Lp, int sr, int sw)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= Lp) return;
  if ((i / RSB_ICOLS) % sw != sr || j <= i) cov[(size_t) i * Lp + j] = 0.0;
}

// w = min(w_old, (maxCOV - max(bmin, minCOV)) / hpts), zeroed below tol: src/R-scape.c:1357-1360
__global__ void width_kernel(const double *__restrict__ minmax, double w_old, double bmin, int hpts, double tol, double *__restrict__ wout)
{
  if (threadIdx.x || blockIdx.x) return;
  const double lo = fmax(bmin, minmax[0]);
  const double w_new = (minmax[1] - lo) / (double) hpts;
  double w = fmin(w_old, w_new);
  if (w < tol) w = 0.0;
  *wout = w;
}

// copy raw -> lower triangle mirror and -inf diagonal for a host-visible symmetric matrix (no correction case)
__global__ void symmetrize_kernel(double *__restrict__ cov, int L, int Lp)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L) return;
  if (i == j) cov[(size_t) i * Lp + j] = -INFINITY;
  else if (i < j) cov[(size_t) j * Lp + i] = cov[(size_t) i * Lp + j];
}

// The three histograms of the input alignment's scan (src/covariation.c:420-457): ha gets every pair, hb the pairs
// flagged in pairmask (the structure's contacts / base pairs, by data->samplesize), ht the others.
__global__ void hist3_kernel(const double *__restrict__ cov, int L, int Lp, const uint8_t *__restrict__ pairmask, double bmin, double w,
                             int nb, unsigned long long *__restrict__ ha, unsigned long long *__restrict__ hb, unsigned long long *__restrict__ ht,
                             int *__restrict__ flags, const int *__restrict__ m2p, int mind)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L || i >= j) return;
  if (m2p && m2p[i] >= 0 && m2p[j] >= 0 && m2p[j] - m2p[i] < mind) return;          // covariation.c:421-427
  const double x  = fmax(cov[(size_t) i * Lp + j], bmin + w);
  const double bd = ceil(((x - bmin) / w) - 1.);
  if (!(bd >= 0.0 && bd < (double) nb)) { atomicOr(flags, 4); return; }
  const int b = (int) bd;
  atomicAdd(&ha[b], 1ull);
  if (pairmask) atomicAdd(pairmask[(size_t) i * L + j] ? &hb[b] : &ht[b], 1ull);
}

} // namespace

void rsb_corr_grid(int L, int *nJT, int *nIT) { *nJT = (L + CH_TJ - 1) / CH_TJ; *nIT = (L + CH_TI - 1) / CH_TI; }

// phase: 1 = reduce the tile partials -> covsum[r][L+4], 2 = covsum -> COVx, COVavg, 3 = both
cudaError_t rsb_launch_correct_final(const double *rowpart, const double *colpart, const double *mm, int nrep, int L,
                                     double *covx, double *scal, double *blocksum, double *covsum, int phase, cudaStream_t st)
{
  int nJT, nIT; rsb_corr_grid(L, &nJT, &nIT);
  const int nblk = (L + 127) / 128;
  if (phase & 1) {
    rsb_coreside(covsum_kernel); covsum_kernel<<<dim3(nblk, nrep), 128, 0, st>>>(rowpart, colpart, L, nJT, nIT, covsum, blocksum);
    rsb_coreside(covtot_kernel); covtot_kernel<<<nrep, 256, 0, st>>>(blocksum, nblk, mm, L, nJT * nIT, covsum);
  }
  rsb_coreside(covx_kernel);
  if (phase & 2) covx_kernel<<<dim3(nblk, nrep), 128, 0, st>>>(covsum, L, covx, scal);
  return cudaGetLastError();
}

cudaError_t rsb_launch_correct_hist(double *cov, const double *covx, const double *scal, int nrep, int L, int Lp, int actype, int mode,
                                    double bmin, const double *wptr, unsigned long long *hist, int nbins, double *mm, double *minmax_out,
                                    int *flags, int sr, int sw, const int *m2p, int mind, cudaStream_t st)
{
  int nJT, nIT; rsb_corr_grid(L, &nJT, &nIT);
  rsb_coreside(correct_hist_kernel); correct_hist_kernel<<<dim3(nJT, nIT, nrep), CH_TJ, 0, st>>>(cov, covx, scal, L, Lp, actype, mode, bmin, wptr, hist, nbins, mm, flags, nJT, nIT, sr, sw, m2p, mind);
  rsb_coreside(minmax_final_kernel); minmax_final_kernel<<<nrep, 256, 0, st>>>(mm, nJT * nIT, minmax_out);
  return cudaGetLastError();
}

// nrep x ncombo corrections + histograms in one launch; minmax_out[(r * ncombo + k) * 2]; mm scratch: nrep * ncombo * tiles * 2 doubles
cudaError_t rsb_launch_correct_hist_multi(const double *cov, const double *covx, const double *scal, int nrep, int L, int Lp, int ncombo, int nw,
                                          const int *kidx, const int *act, double bmin, const double *wptr, unsigned long long *hist, int nbins,
                                          double *mm, double *minmax_out, int *flags, const int *m2p, int mind, cudaStream_t st)
{
  int nJT, nIT; rsb_corr_grid(L, &nJT, &nIT);
  ComboTab tab;
  for (int k = 0; k < 64; k++) { tab.kidx[k] = k < ncombo ? kidx[k] : 0; tab.act[k] = k < ncombo ? act[k] : RSB_NOCORR; }
  rsb_coreside(correct_hist_multi_kernel);
  correct_hist_multi_kernel<<<dim3(nJT, nIT, nrep * ncombo), CH_TJ, 0, st>>>(cov, covx, scal, L, Lp, tab, ncombo, nw, bmin, wptr, hist, nbins, mm, flags, nJT, nIT, m2p, mind);
  rsb_coreside(minmax_final_kernel); minmax_final_kernel<<<nrep * ncombo, 256, 0, st>>>(mm, nJT * nIT, minmax_out);
  return cudaGetLastError();
}

cudaError_t rsb_launch_width(const double *minmax, double w_old, double bmin, int hpts, double tol, double *wout, cudaStream_t st)
{
  width_kernel<<<1, 32, 0, st>>>(minmax, w_old, bmin, hpts, tol, wout);
  return cudaGetLastError();
}

cudaError_t rsb_launch_negate_min(double *minmax, int n, cudaStream_t st)
{
  negate_min_kernel<<<(n + 127) / 128, 128, 0, st>>>(minmax, n);
  return cudaGetLastError();
}

cudaError_t rsb_launch_zero_unowned_rows(double *cov, int L, int Lp, int sr, int sw, cudaStream_t st)
{
  zero_unowned_rows_kernel<<<dim3((Lp + 127) / 128, L), 128, 0, st>>>(cov, L, Lp, sr, sw);
  return cudaGetLastError();
}

cudaError_t rsb_launch_symmetrize(double *cov, int L, int Lp, cudaStream_t st)
{
  symmetrize_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(cov, L, Lp);
  return cudaGetLastError();
}

cudaError_t rsb_launch_hist3(const double *cov, int L, int Lp, const uint8_t *pairmask, double bmin, double w, int nb,
                             unsigned long long *ha, unsigned long long *hb, unsigned long long *ht, int *flags, const int *m2p, int mind, cudaStream_t st)
{
  hist3_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(cov, L, Lp, pairmask, bmin, w, nb, ha, hb, ht, flags, m2p, mind);
  return cudaGetLastError();
}
