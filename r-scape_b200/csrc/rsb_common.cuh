// rsb_common.cuh -- shared declarations of the sm_100a covariation kernels.
//
// Data layout in HBM (per context; R = replicates in flight, L = alignment length, N = sequences,
// S = number of 8-bit weight slices, Kpad = N rounded up to 128, K = 4 residues):
//
//   res      u8   [R][N][L]            digital residues as ESL_MSA ax (A0 C1 G2 U3, gap 4, N 15)
//   wdig     u8   [S][Kpad]            k-th base-256 digit of round(w_s * 2^q)   (pack.cu)
//   planeA   u8   [R][MA][Kpad]        one-hot rows, row = 4*i + a, MA = 4*roundup(L,32)
//   planeB   u8   [R][NB][Kpad]        weighted one-hot rows, row = (j*S + k)*4 + b, value = digit_k(w_s)[x_sj == b]
//   cnt      i64  [R][16][L][L]        fixed-point pair counts, plane = a*4+b, upper triangle i<j only
//   pm       f64  [R][L][4]            partner-averaged marginals (corr_Marginals)
//   cov      f64  [R][L][L]            raw statistic, then corrected in place for the real MSA
//   hist     u64  [NB_HIST]            cumulative score histogram of the null batch
//
// Reference arithmetic: src/correlators.c; each kernel cites its lines.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#define RSB_K      4
#define RSB_K2     16
#define RSB_MTILE  128          // UMMA M = 128 rows of planeA = 32 alignment columns i
#define RSB_ICOLS  32
#define RSB_KSTAGE 128          // bytes of K (sequences) per pipeline stage = one 128B swizzle atom
#define RSB_MAX_SLICES 6
#ifndef RSB_EPI_GROUPS
#define RSB_EPI_GROUPS 2        // groups of 4 epilogue warps in the tcgen05 kernel (gram_tcgen05.cu)
#endif
#ifndef RSB_GRAM_BOUND
#define RSB_GRAM_BOUND 512      // thread count declared in the tcgen05 kernel's __launch_bounds__ when larger than its real one: caps its
#endif                          // registers (65536 / bound = 128, no spills) so that blocks of the statistics chain fit beside it.  Uncapped,
                                // the pair-per-thread record epilogue takes 154 registers x 384 threads = 90 % of the SM's file, no block of
                                // gt_finish / correct_hist can be co-resident and the chain only runs between two contractions
// pair-tile of the HBM-bound passes (marginals, statistic, correction, histogram): a block owns
// RSB_TI rows x RSB_TJ columns of the upper triangle, one thread per column j
#define RSB_TI     16
#define RSB_TJ     128
// row-block sharding of the pair grid across ranks (LSU-scale L): rank sr of sw owns the 32-column row blocks ib with
// ib % sw == sr (cyclic, because row i has L-1-i pairs); a tile row `it` of RSB_TI rows belongs to block it*RSB_TI/32
#define RSB_OWNED(it, sr, sw) ((sw) <= 1 || (((it) * RSB_TI / RSB_ICOLS) % (sw)) == (sr))

#define RSB_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rsb_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 1; } } while (0)

// columns j per gram tile for a given slice count: N tile = 4*S*CJ <= 256 and a multiple of 16
__host__ __device__ constexpr int rsb_cj_for(int S) { return S == 1 ? 64 : S == 2 ? 32 : S == 3 ? 20 : S == 4 ? 16 : S == 5 ? 12 : 10; }

// statistic / class / correction codes = the reference's enums (src/correlators.h:36-92)
enum { RSB_CHI = 0, RSB_GT = 3, RSB_MI = 6, RSB_MIr = 9, RSB_MIg = 12, RSB_OMES = 15, RSB_RAF = 18, RSB_RAFS = 21, RSB_CCF = 24 };
enum { RSB_C16 = 0, RSB_C2 = 1, RSB_CWC = 2, RSB_CSELECT = 3 };
enum { RSB_APC = 0, RSB_ASC = 1, RSB_NOCORR = 2 };

// The kernels of the statistics chain run beside the persistent tcgen05 kernel, which holds every SM's shared memory at
// the maximum carve-out: they ask for the same carve-out so that an SM never has to drain to switch configuration.
template <typename K> static inline void rsb_coreside(K kernel)
{
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
}

// exact uint64 -> double without the slow 64-bit I2F path: both 32-bit halves are planted in the mantissa of a
// power of two and the offsets subtracted (each half exact, one rounding in the final add = the I2F result)
__device__ __forceinline__ double u64_to_f64(unsigned long long v)
{
  const double lo = __longlong_as_double(0x4330000000000000ULL | (v & 0xffffffffULL)) - 4503599627370496.0;              // 2^52
  const double hi = __longlong_as_double(0x4530000000000000ULL | (v >> 32))           - 19342813113834066795298816.0;   // 2^84
  return hi + lo;
}

// counts below 2^52 (true of every count when the alignment's total weight wtot < 2^52) convert with one subtraction
__device__ __forceinline__ double u52_to_f64(unsigned long long v)
{
  return __longlong_as_double(0x4330000000000000ULL | v) - 4503599627370496.0;
}

// natural log from a 512-entry table in shared memory: x = 2^e m, m = c (1 + r) with c the centre of m's 1/512 bin, so
// |r| <= 2^-10 and log1p(r) = r - r^2/2 + r^3/3 - r^4/4 (truncation 2^-52).  8 FP64 operations and no branch, against
// ~40 and a branchy special-case path for the library log; absolute error <= 4e-16 + 1 ulp(e ln 2).  Arguments are
// clamped to the smallest normal double: every call site multiplies the log of a zero/subnormal probability by that
// probability (or discards it under a `> 0` guard), so the clamp never changes a result.  tab[k] = { 1/c_k rounded,
// -log of that }, built once per context by logtab_kernel and copied to shared memory by each block.
constexpr int LOGTAB_N = 512;
template <bool CLAMP = true>
__device__ __forceinline__ double fast_log(double x, const double2 *__restrict__ tab)
{
  if (CLAMP) x = fmax(x, 2.2250738585072014e-308);                  // (not needed where the argument is known to be a positive normal)
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const double m  = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 t = tab[(hi >> 11) & (LOGTAB_N - 1)];
  const double r  = fma(m, t.x, -1.0);
  const double e  = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
  double p = fma(r, -0.25, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  return fma(e, 0.6931471805599453094, t.y) + fma(p, r * r, r);
}

// does the gram tile (ib, jb) exist?  Same rule as build_geo (capi.cu): it holds some pair i < j and its row block is
// owned by this rank
__host__ __device__ __forceinline__ bool rsb_tile_exists(int ib, int jb, int CJ, int L, int sr, int sw)
{
  const int maxj = (jb * CJ + CJ - 1 < L - 1) ? jb * CJ + CJ - 1 : L - 1;
  return ib * RSB_ICOLS < maxj && (sw <= 1 || ib % sw == sr);
}

#ifdef RSB_BLOCKTRACE
// experiment: per-block (kernel id, SM, start, end) records, to see which kernels really share an SM
__device__ __forceinline__ unsigned long long rsb_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned rsb_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
__device__ __forceinline__ void rsb_trace_put(unsigned long long *buf, unsigned kid, unsigned long long t0)
{
  if (!buf) return;
  const unsigned long long t1 = rsb_gtime();
  const unsigned long long k = atomicAdd(buf, 1ULL);
  if (k < (1ULL << 20)) { buf[1 + 3 * k] = ((unsigned long long) kid << 32) | rsb_smid(); buf[2 + 3 * k] = t0; buf[3 + 3 * k] = t1; }
}
#endif

struct rsb_ctx;
void rsb_set_error(rsb_ctx *ctx, const char *fmt, ...);
