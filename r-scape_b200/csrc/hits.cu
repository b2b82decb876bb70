// hits.cu -- E-values of every pair and the significant-pair list of the input alignment's scan:
// the per-pair loop of cov_CreateHitList, src/covariation.c:828-910, over cov2evalue (:2370-2400).
//
// One thread per pair (i, j), i < j, in 32 x 32 tiles of the upper triangle: p-value from the cumulative null histogram
// (rsb_evalue.cuh), E-value = p x the number of tests of the pair's set (Nb for pairs of the given structure, Nt for the
// others; :850-853), written to both triangles of mi->Eval (:855) -- the mirror goes through a shared-memory transpose so
// that both stores are row-contiguous -- and appended to the hit list when E < thresh (or when every pair is reported,
// thresh > MAX_EVAL, :859).  The list is compacted with one atomic per warp; the host layer sorts the few entries back
// into the reference's row-major order.  HBM-bound: 8 B read + 16 B written per pair (+1 B of the structure mask).
#include "rsb_common.cuh"
#include "rsb_evalue.cuh"

namespace {

// `h < expBP` rule of :852 (only with --structured): the reference multiplies the p-value of a pair outside the structure
// by expBP while fewer than expBP hits have been listed, by Nt afterwards.  Hits are listed in row-major order, so the
// rule is "pairs with linear index n <= switch_n use expBP" where switch_n is the index of the expBP-th hit; the caller
// finds it with one pass at switch_n = (all pairs) and, if the list got that long, a second pass (capi.cu).
__global__ void __launch_bounds__(256)
evalue_hits_kernel(const double *__restrict__ cov, int L, int Lp, rsb_nullview nv, const uint8_t *__restrict__ pairmask,
                   double Nb, double Nt, double expBP, long long switch_n, double thresh, int report_all, int sr, int sw,
                   double *__restrict__ eval, long long cap, long long *__restrict__ hit_ij, double *__restrict__ hit_sc,
                   double *__restrict__ hit_eval, double *__restrict__ hit_pval, unsigned long long *__restrict__ nhit,
                   int *__restrict__ flags)
{
  // a block owns the 32 x 32 tile (ti, tj) of the upper triangle; warp w handles rows w, w+8, w+16, w+24 of it, a lane a column
  __shared__ double tile[32][33];
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (tj < ti) return;                                                             // whole block: nothing of the tile has i < j
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = tj * 32 + tx;
  #pragma unroll
  for (int r = 0; r < 4; r++) {
    const int il = ty + 8 * r, i = ti * 32 + il;
    const bool inside = (i < L && j < L);
    const bool mine = inside && i < j && (sw <= 1 || (i / RSB_ICOLS) % sw == sr);
    bool   hit = false;
    double sc = 0.0, ev = 0.0, pv = 0.0;
    if (mine) {
      int bad = 0;
      sc = cov[(size_t) i * Lp + j];
      pv = rsb_cov2pval(sc, nv, &bad);
      if (bad) atomicOr(flags, 8);
      const long long n = (long long) i * L - (long long) i * (i + 1) / 2 + (j - i - 1);    // index of the pair in the reference's loop
      const bool isbp = pairmask && pairmask[(size_t) i * L + j];
      ev = pv * (isbp ? Nb : (n <= switch_n ? expBP : Nt));
      hit = report_all || ev < thresh;
    }
    if (eval) {
      tile[il][tx] = ev;
      if (mine)                 eval[(size_t) i * L + j] = ev;                      // row-contiguous store
      else if (inside && i == j) eval[(size_t) i * L + i] = INFINITY;               // corr_ReuseCOV leaves +inf there, :1261
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);                            // every lane of the warp gets here, 4 times
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned long long base = 0;
      if (tx == leader) base = atomicAdd(nhit, (unsigned long long) __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (hit) {
        const unsigned long long k = base + (unsigned long long) __popc(m & ((1u << tx) - 1u));
        if (k < (unsigned long long) cap) {
          hit_ij[k] = ((long long) i << 32) | (long long) j;
          hit_sc[k] = sc; hit_eval[k] = ev; hit_pval[k] = pv;
        }
      }
    }
  }
  if (eval) {                                                                      // the mirror mi->Eval[j][i], transposed through shared memory
    __syncthreads();
    const int ii = ti * 32 + tx;
    #pragma unroll
    for (int r = 0; r < 4; r++) {
      const int jl = ty + 8 * r, jj = tj * 32 + jl;
      if (ii < L && jj < L && ii < jj && (sw <= 1 || (ii / RSB_ICOLS) % sw == sr)) eval[(size_t) jj * L + ii] = tile[tx][jl];
    }
  }
}

} // namespace

cudaError_t rsb_launch_evalue_hits(const double *cov, int L, int Lp, const rsb_nullview &nv, const uint8_t *pairmask, double Nb, double Nt,
                                   double expBP, long long switch_n, double thresh, int report_all, int sr, int sw, double *eval, long long cap,
                                   long long *hit_ij, double *hit_sc, double *hit_eval, double *hit_pval, unsigned long long *nhit, int *flags,
                                   cudaStream_t st)
{
  const int nT = (L + 31) / 32;
  evalue_hits_kernel<<<dim3(nT, nT), 256, 0, st>>>(cov, L, Lp, nv, pairmask, Nb, Nt, expBP, switch_n, thresh, report_all, sr, sw,
                                                               eval, cap, hit_ij, hit_sc, hit_eval, hit_pval, nhit, flags);
  return cudaGetLastError();
}
