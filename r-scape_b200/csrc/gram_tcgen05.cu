// gram_tcgen05.cu -- the pair-count contraction on the 5th-generation tensor cores.
//
// Replaces the accumulate loop of mutual_naive_ppij (reference src/correlators.c:1724-1755), run
// over all i<j by corr_NaivePP (:1301-1317).  The reference walks every pair of columns and every
// sequence; here the whole table set is one integer Gram product over one-hot residue planes,
//
//     cnt[a*4+b][i][j] = sum_s wq_s [x_si = a][x_sj = b]          wq_s = round(w_s 2^q)  (pack.cu)
//                      = sum_k 256^k  sum_s planeA[4i+a][s] * planeB[(j S + k) 4 + b][s]
//
// with unsigned 8-bit operands and int32 accumulation (exact), one accumulator column block per
// weight slice k, recombined to int64 in the epilogue.  S = 1 with wq = 1 is the unweighted count.
//
// Kernel shape (one CTA per SM, persistent over a host-built list of upper-triangle tiles):
//   warp 0      TMA producer : cp.async.bulk.tensor 128B-swizzled boxes of planeA (128 rows) and
//                              planeB (4*S*CJ rows) x 128 sequences into a 4-stage smem ring
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1.kind::i8, M=128, N=4*S*CJ, K=32, D in TMEM
//   warp 2      TMEM allocator (512 columns = two accumulator buffers)
//   warps 4-7   epilogue     : tcgen05.ld 32x32b, recombine slices, store int64 count planes
// Accumulators are double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
#include "rsb_common.cuh"

namespace {

#ifndef RSB_NSTAGE
#define RSB_NSTAGE 4                          // TMA ring depth (48 KB per stage at S = 4)
#endif
constexpr int NSTAGE       = RSB_NSTAGE;
constexpr int GRAM_THREADS = 256;              // CTA-pair kernel: warps 0-3 roles, warps 4-7 epilogue
// single-CTA kernel: RSB_EPI_GROUPS groups of 4 epilogue warps; group g drains the columns [g, g+1) * cpg of every tile, so that
// the latency of one tile's epilogue -- what the MMAs of the tile after next wait for -- is divided by the number of groups
constexpr int GRAM_EPI_GROUPS = RSB_EPI_GROUPS;
constexpr int GRAM1_THREADS   = 128 + 128 * GRAM_EPI_GROUPS;
__host__ __device__ constexpr int gram_cols_per_group(int CJ, int E) { return ((CJ / 2 + E - 1) / E) * 2; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;"  ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2.  Fields per cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return  (uint64_t) ((saddr >> 4) & 0x3FFFu)
        | ((uint64_t) 1u  << 16)
        | ((uint64_t) 64u << 32)
        | ((uint64_t) 1u  << 46)
        | ((uint64_t) 2u  << 61);
}

// Epilogue of one tile, shared by the single-CTA and the CTA-pair kernels: this warp's 32 TMEM lanes (rows (i, a)) of the
// accumulator at `trow` -> int64 count planes, and (part) the tile's marginal partial sums.  Returns the row partial.
//
// REC (the G test on 16 classes, corr_CalculateGT_C16, src/correlators.c:361-395, for alignments whose counts nobody reads:
// the nulls): instead of the 16 int64 counts (128 B) the epilogue leaves a 64-byte record per pair from which the statistic
// follows with 14 multiply-adds once the marginals pm are known (gt_finish_kernel, stats.cu).  With the raw table
// x_ab = 1e-10 + c_ab scale, X_a = sum_b x_ab, Y_b = sum_a x_ab, T = sum x, ne = nseff_ij (stats.cu, gt_c16_raw):
//     G = 2 ne [ (sum x log x)/T - log T ]  -  sum_a (2 ne X_a/T) log pm_i[a]  -  sum_b (2 ne Y_b/T) log pm_j[b]
//       =           A                       -  sum_a      U_a     log pm_i[a]  -  sum_b      V_b     log pm_j[b]
// and, because sum_a U_a = sum_b V_b = 2 ne, only U_0..2, V_0..2 and N2 = 2 ne are stored:
//     record planes  0: A   1: N2   2..4: U_0..U_2   5..7: V_0..V_2        (double [8][L][Lp], i < j; aliases the count planes)
// The 17 logs per pair -- the FP64 work that needed a serial window of its own as a separate kernel -- are spread over the
// 4 lanes that hold the pair's 16 cells and run here under the MMAs of the next tile.
// timing experiments only (tools/build_variant.sh): parts of the record epilogue switched off -- results are then wrong
#ifdef RSB_EXP_NOLOG
#define GLOG(x) (x)
#else
#define GLOG(x) fast_log<false>((x), ltab)
#endif
template <int S, bool REC>
__device__ __forceinline__ double gram_epilogue_tile(uint32_t trow, int jb, int i, int a, int lane, int ew, int ib, int r, int L, int Lp,
                                                     long long *__restrict__ base, size_t plane, double scale, bool part,
                                                     double *__restrict__ mcol, int nIB, bool small52, const double2 *__restrict__ ltab,
                                                     int jl_begin = 0, int jl_end = rsb_cj_for(S))
{
  constexpr int CJ = rsb_cj_for(S);
  (void) Lp; (void) CJ;
  double racc = 0.0;                                               // marginal partial of row (i, a) over this tile's columns
  // REC: this lane's record planes (row i): lanes a < 3 write U_a and V_a, lane 3 writes N2, lane 0 also A
  double *const recrow = reinterpret_cast<double *>(base) - (size_t) a * 4 * plane;          // plane 0 of this slot, row i
  double *const pU = recrow + (size_t) (a < 3 ? 2 + a : 1) * plane;
  double *const pV = recrow + (size_t) (5 + (a < 3 ? a : 0)) * plane;
  #pragma unroll 1
  for (int jl = jl_begin; jl < jl_end; jl += 2) {
    uint32_t d[2][S][4];
    #pragma unroll
    for (int u = 0; u < 2; u++)
      #pragma unroll
      for (int k = 0; k < S; k++)
        tmem_ld4(trow + (uint32_t) (((jl + u) * S + k) * 4), d[u][k][0], d[u][k][1], d[u][k][2], d[u][k][3]);
    tmem_ld_wait();

    const int j0 = jb * CJ + jl;
    unsigned long long c[2][4];
    #pragma unroll
    for (int u = 0; u < 2; u++)
      #pragma unroll
      for (int b = 0; b < 4; b++) {
        unsigned long long v = 0;
        #pragma unroll
        for (int k = S - 1; k >= 0; k--) v = (v << 8) + (unsigned long long) d[u][k][b];
        c[u][b] = v;
      }
    const bool ok0 = (i < L) && (j0 < L)     && (i < j0);
    const bool ok1 = (i < L) && (j0 + 1 < L) && (i < j0 + 1);
    if (REC) {
    } else if (ok0 && ok1) {
      #pragma unroll
      for (int b = 0; b < 4; b++)
        *reinterpret_cast<ulonglong2 *>(base + b * plane + j0) = make_ulonglong2(c[0][b], c[1][b]);
    } else {
      #pragma unroll
      for (int b = 0; b < 4; b++) {
        if (ok0) base[b * plane + j0]     = (long long) c[0][b];
        if (ok1) base[b * plane + j0 + 1] = (long long) c[1][b];
      }
    }

    if (part) {
      // marginal partials (corr_Probs :1713-1758 + corr_Marginals :1335-1370, the per-pair part): the 16 cells of a
      // pair sit in 4 adjacent lanes (a = lane & 3) x 4 registers (b).  pp = (1e-10 + c scale) / sum; pairs with
      // nseff = 0 are skipped (:1354).  Fixed shuffle trees => deterministic.
      double x[2][4];
      double rA[2], rU[2], rV[2];                                    // REC: this lane's record values of the two pairs
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        #pragma unroll
        for (int b = 0; b < 4; b++) x[u][b] = fma(small52 ? u52_to_f64(c[u][b]) : u64_to_f64(c[u][b]), scale, 1e-10);
        const double rs = (x[u][0] + x[u][1]) + (x[u][2] + x[u][3]);
        double sum = rs;
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const unsigned nz  = __ballot_sync(0xffffffffu, (c[u][0] | c[u][1] | c[u][2] | c[u][3]) != 0ULL);
        const bool     use = (u ? ok1 : ok0) && ((nz >> (lane & ~3)) & 0xFu) != 0u;
        if (REC) {
          // sum x log x over the pair's 16 cells: 4 logs per lane (x >= 1e-10 and T >= 1.6e-9 are positive normals: no clamp)
          double sl = x[u][0] * GLOG(x[u][0]);
          sl = fma(x[u][1], GLOG(x[u][1]), sl);
          double s2 = x[u][2] * GLOG(x[u][2]);
          s2 = fma(x[u][3], GLOG(x[u][3]), s2);
          sl += s2;
          sl += __shfl_xor_sync(0xffffffffu, sl, 1);
          sl += __shfl_xor_sync(0xffffffffu, sl, 2);
          // Y_b = sum over the 4 lanes (a) of x[b], transposed so that lane a ends up with Y_a: 3 shuffles
          const bool q2 = lane & 2, q1 = lane & 1;
          const double k0 = q2 ? x[u][2] : x[u][0], k1 = q2 ? x[u][3] : x[u][1];
          const double t0 = q2 ? x[u][0] : x[u][2], t1 = q2 ? x[u][1] : x[u][3];
          const double h0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 2), h1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 2);
          const double yb = (q1 ? h1 : h0) + __shfl_xor_sync(0xffffffffu, q1 ? h0 : h1, 1);
          // nseff exactly: integer sum of the 16 counts
          const unsigned long long nl = (c[u][0] + c[u][1]) + (c[u][2] + c[u][3]);
          double ne;
          if (small52) {                                              // every partial sum < 2^52: the double adds are exact
            ne = u52_to_f64(nl);
            ne += __shfl_xor_sync(0xffffffffu, ne, 1);
            ne += __shfl_xor_sync(0xffffffffu, ne, 2);
          } else {
            unsigned long long nn = nl;
            nn += __shfl_xor_sync(0xffffffffu, nn, 1);
            nn += __shfl_xor_sync(0xffffffffu, nn, 2);
            ne = u64_to_f64(nn);
          }
          const double n2   = 2.0 * (ne * scale);
          const double invT = 1.0 / sum;
          const double coef = n2 * invT;
          rA[u] = fma(coef, sl, -n2 * GLOG(sum));
          rU[u] = (a < 3) ? coef * rs : n2;
          rV[u] = coef * yb;
          const double inv = use ? invT : 0.0;
          racc = fma(rs, inv, racc);
          #pragma unroll
          for (int b = 0; b < 4; b++) x[u][b] *= inv;
        } else {
          const double   inv = use ? 1.0 / sum : 0.0;
          racc = fma(rs, inv, racc);
          #pragma unroll
          for (int b = 0; b < 4; b++) x[u][b] *= inv;
        }
      }
#ifdef RSB_EXP_NORECSTORE
      if (REC && rA[0] + rU[0] + rV[0] + rA[1] + rU[1] + rV[1] == -1.2345) {
#else
      if (REC) {
#endif
        if (ok0 && ok1) {
          *reinterpret_cast<double2 *>(pU + j0) = make_double2(rU[0], rU[1]);
          if (a < 3)  *reinterpret_cast<double2 *>(pV + j0) = make_double2(rV[0], rV[1]);
          if (a == 0) *reinterpret_cast<double2 *>(recrow + j0) = make_double2(rA[0], rA[1]);
        } else {
          if (ok0) { pU[j0] = rU[0];     if (a < 3) pV[j0] = rV[0];     if (a == 0) recrow[j0] = rA[0]; }
          if (ok1) { pU[j0 + 1] = rU[1]; if (a < 3) pV[j0 + 1] = rV[1]; if (a == 0) recrow[j0 + 1] = rA[1]; }
        }
      }
#ifndef RSB_EXP_NOCOLPART
      // column partials: 8 values (u, b) summed over the warp's 32 lanes by a halving butterfly -- after the three
      // halving steps a lane keeps (u, b) = (bit 4, bits 3:2 of its lane id), then two full steps sum over a
      const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
      double y[4], z[2], v1;
      #pragma unroll
      for (int b = 0; b < 4; b++) {
        const double keep = h4 ? x[1][b] : x[0][b], send = h4 ? x[0][b] : x[1][b];
        y[b] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      #pragma unroll
      for (int q = 0; q < 2; q++) {
        const double keep = h3 ? y[2 + q] : y[q], send = h3 ? y[q] : y[2 + q];
        z[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const double keep = h2 ? z[1] : z[0], send = h2 ? z[0] : z[1];
        v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      const int jw = j0 + (h4 ? 1 : 0);
      if ((lane & 3) == 0 && jw < L)
        mcol[(((size_t) r * 4 * nIB + (size_t) ib * 4 + ew) * L + jw) * 4 + ((h3 ? 2 : 0) + (h2 ? 1 : 0))] = v1;
#endif
    }
  }
  return racc;
}

// ------------------------------------------------------------------------------------------------------------------
// Record epilogue, pair-per-thread form (S = 1, 2, 4).  The accumulator hands a thread ONE residue row (i, a) of every pair, so
// the form above pays ~30 shuffles and as many selects per pair to bring the 4 rows of a pair together -- and every ALU
// instruction issued beside the running MMAs is throttled (ncu: "math pipe throttle" is the top stall of this kernel).  Here a
// warp moves the digits of 32 pairs (its 8 columns i x 4 columns j) through 2 KB of shared memory, 16 bytes per lane and store,
// bank-conflict free by an XOR swizzle, after which every lane owns the whole 4 x 4 table of one pair: sums, the 17 logs, the
// one division and the record are thread-local; what stays across lanes are 4 double shuffles per pair for the column marginal
// partials and two reductions per tile for the row partials.  Same record, same arithmetic per value as the form above.
template <int S>
__device__ __forceinline__ void gram_epilogue_rec_pairs(uint32_t trow, int jb, int i0w, int lane, int ew, int ib, int r, int L, int Lp,
                                                        double *__restrict__ rec0, size_t plane, double scale, double *__restrict__ mrow_blk,
                                                        double *__restrict__ mcol, int nIB, bool small52, const double2 *__restrict__ ltab,
                                                        uint4 *__restrict__ xbuf, int jl_begin, int jl_end)
{
  constexpr int CJ = rsb_cj_for(S);
  const int il_w = lane >> 2, a_w = lane & 3;                       // what this lane's TMEM row is: column i0w + il_w, residue a_w
  const int il_t = lane & 7, jj = lane >> 3;                        // the pair this lane owns after the exchange: column i0w + il_t, j0 + jj
  const int i = i0w + il_t;
  const int wslot = a_w ^ ((il_w >> 1) & 3);                        // swizzled 16-byte slot of row a_w inside the pair's 64 bytes
  const int rsw = (lane >> 1) & 3;
  double racc[4] = { 0.0, 0.0, 0.0, 0.0 };                          // row marginal partials of (i, a) over this lane's columns
  #pragma unroll 1
  for (int jl = jl_begin; jl < jl_end; jl += 4) {
    unsigned long long c[4][4];
    #pragma unroll
    for (int k = 0; k < S; k++) {
      uint32_t d[4][4];
      #pragma unroll
      for (int u = 0; u < 4; u++) tmem_ld4(trow + (uint32_t) (((jl + u) * S + k) * 4), d[u][0], d[u][1], d[u][2], d[u][3]);
      tmem_ld_wait();
      #pragma unroll
      for (int u = 0; u < 4; u++) xbuf[(u * 8 + il_w) * 4 + wslot] = make_uint4(d[u][0], d[u][1], d[u][2], d[u][3]);
      __syncwarp();
      #pragma unroll
      for (int a = 0; a < 4; a++) {
        const uint4 v = xbuf[lane * 4 + (a ^ rsw)];
        if (k == 0) { c[a][0] = v.x; c[a][1] = v.y; c[a][2] = v.z; c[a][3] = v.w; }
        else {
          c[a][0] += (unsigned long long) v.x << (8 * k); c[a][1] += (unsigned long long) v.y << (8 * k);
          c[a][2] += (unsigned long long) v.z << (8 * k); c[a][3] += (unsigned long long) v.w << (8 * k);
        }
      }
      __syncwarp();
    }
    const int  j  = jb * CJ + jl + jj;
    const bool ok = (i < L) && (j < L) && (i < j);

    double x[4][4], X[4], Y[4], sl[4];
    unsigned long long nl = 0;
    #pragma unroll
    for (int a = 0; a < 4; a++) {
      #pragma unroll
      for (int b = 0; b < 4; b++) x[a][b] = fma(small52 ? u52_to_f64(c[a][b]) : u64_to_f64(c[a][b]), scale, 1e-10);
      X[a] = (x[a][0] + x[a][1]) + (x[a][2] + x[a][3]);
      nl += (c[a][0] + c[a][1]) + (c[a][2] + c[a][3]);
      double s1 = x[a][0] * GLOG(x[a][0]);
      s1 = fma(x[a][1], GLOG(x[a][1]), s1);
      double s2 = x[a][2] * GLOG(x[a][2]);
      s2 = fma(x[a][3], GLOG(x[a][3]), s2);
      sl[a] = s1 + s2;
    }
    #pragma unroll
    for (int b = 0; b < 4; b++) Y[b] = (x[b][b] + x[b ^ 2][b]) + (x[b ^ 1][b] + x[b ^ 3][b]);
    const double sum  = (X[0] + X[1]) + (X[2] + X[3]);
    const double slog = (sl[0] + sl[1]) + (sl[2] + sl[3]);
    const double ne   = small52 ? u52_to_f64(nl) : u64_to_f64(nl);
    const double n2   = 2.0 * (ne * scale);
    const double invT = 1.0 / sum;
    const double coef = n2 * invT;
    const double rA   = fma(coef, slog, -n2 * GLOG(sum));
    if (ok) {
      double *q = rec0 + (size_t) i * Lp + j;
      q[0] = rA;
      q[plane] = n2;
      q[2 * plane] = coef * X[0]; q[3 * plane] = coef * X[1]; q[4 * plane] = coef * X[2];
      q[5 * plane] = coef * Y[0]; q[6 * plane] = coef * Y[1]; q[7 * plane] = coef * Y[2];
    }
    // marginal partials (corr_Marginals :1335-1370, the per-pair part); pairs with nseff = 0 are skipped (:1354)
    const double inv = (ok && nl != 0ULL) ? invT : 0.0;
    #pragma unroll
    for (int a = 0; a < 4; a++) racc[a] = fma(X[a], inv, racc[a]);
    // column partials of (j, b): sum over the warp's 8 columns i (lanes jj * 8 .. jj * 8 + 7), halving butterfly
    {
      const bool h4 = lane & 4, h2 = lane & 2;
      const double y0 = Y[0] * inv, y1 = Y[1] * inv, y2 = Y[2] * inv, y3 = Y[3] * inv;
      const double k0 = h4 ? y2 : y0, k1 = h4 ? y3 : y1, t0 = h4 ? y0 : y2, t1 = h4 ? y1 : y3;
      const double z0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 4), z1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 4);
      double v = (h2 ? z1 : z0) + __shfl_xor_sync(0xffffffffu, h2 ? z0 : z1, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if ((lane & 1) == 0 && j < L)
        mcol[(((size_t) r * 4 * nIB + (size_t) ib * 4 + ew) * L + j) * 4 + ((h4 ? 2 : 0) + (h2 ? 1 : 0))] = v;
    }
  }
  // row partials: sum over the 4 lanes (jj) that share a column i, then one lane per column writes its 4 values
  #pragma unroll
  for (int a = 0; a < 4; a++) {
    racc[a] += __shfl_xor_sync(0xffffffffu, racc[a], 8);
    racc[a] += __shfl_xor_sync(0xffffffffu, racc[a], 16);
  }
  if (lane < 8 && i < L) {
    double *o = mrow_blk + (size_t) i * 4;
    *reinterpret_cast<double2 *>(o)     = make_double2(racc[0], racc[1]);
    *reinterpret_cast<double2 *>(o + 2) = make_double2(racc[2], racc[3]);
  }
}

#ifndef RSB_REC_PAIRS
#define RSB_REC_PAIRS 1         // 1: pair-per-thread record epilogue for S = 1, 2, 4 (0: the row-per-thread form everywhere)
#endif
template <int S> __host__ __device__ constexpr bool rec_pairs_form() { return RSB_REC_PAIRS && (S == 1 || S == 2 || S == 4) && (gram_cols_per_group(rsb_cj_for(S), GRAM_EPI_GROUPS) % 4 == 0) && (rsb_cj_for(S) % 4 == 0); }

#ifdef RSB_BLOCKTRACE
__device__ unsigned long long *gram_trace_buf;
#endif

template <int S, bool REC>
__global__ void __launch_bounds__(RSB_GRAM_BOUND > GRAM1_THREADS ? RSB_GRAM_BOUND : GRAM1_THREADS, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
               const int2 *__restrict__ tiles, int ntiles, int rep0, int nrep, int L, int Lp, int kstages,
               long long *__restrict__ cnt, double scale, double *__restrict__ mrow, double *__restrict__ mcol, int nJB, int nIB, int small52,
               const double2 *__restrict__ glogtab)
{
  constexpr int      CJ          = rsb_cj_for(S);
  constexpr int      NT          = 4 * S * CJ;                 // UMMA N
  constexpr uint32_t A_BYTES     = RSB_MTILE * RSB_KSTAGE;
  constexpr uint32_t B_BYTES     = NT * RSB_KSTAGE;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // instruction descriptor: D = s32 (2 << 4), A/B unsigned 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t IDESC       = (2u << 4) | ((uint32_t) (NT >> 3) << 17) | ((uint32_t) (RSB_MTILE >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
#ifdef RSB_BLOCKTRACE
  const unsigned long long t0 = rsb_gtime();
#endif
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle-128B tiles need 1024 B alignment
  const uint32_t bar_base  = smem_base + NSTAGE * STAGE_BYTES;
  auto full_bar   = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar  = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
  auto tfull_bar  = [&](int a) { return bar_base + 8u * (2 * NSTAGE + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 4);
  volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *) (smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // REC: the log table of the statistic (rsb_common.cuh, fast_log), 8 KB behind the barriers
  const double2 *ltab = reinterpret_cast<const double2 *>(smem_raw + (bar_base + 128u - smem_u32(smem_raw)));
  // REC, pair-per-thread form: 2 KB of exchange space per epilogue warp behind the table
  uint4 *xbuf_all = reinterpret_cast<uint4 *>(smem_raw + (bar_base + 128u + (uint32_t) (sizeof(double2) * LOGTAB_N) - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++)      { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128 * GRAM_EPI_GROUPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (REC && warp >= 4) {
    double2 *dst = const_cast<double2 *>(ltab);
    for (int k = threadIdx.x - 128; k < LOGTAB_N; k += 128 * GRAM_EPI_GROUPS) dst[k] = glogtab[k];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const long long nwork = (long long) ntiles * nrep;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int  r = rep0 + (int) (w / ntiles);
        const int2 t = tiles[(int) (w % ntiles)];
        for (int ks = 0; ks < kstages; ks++) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          tma_load_3d(sA,           &tmapA, full_bar(stage), ks * RSB_KSTAGE, t.x * RSB_MTILE, r);
          tma_load_3d(sA + A_BYTES, &tmapB, full_bar(stage), ks * RSB_KSTAGE, t.y * NT,        r);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0;   uint32_t acc_phase = 0;
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);               // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t) acc * 256u;
        for (int ks = 0; ks < kstages; ks++) {
          mbar_wait(full_bar(stage), phase);                       // TMA bytes have landed
          tc_fence_after();
          const uint32_t sA    = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = smem_desc_sw128(sA);
          const uint64_t bdesc = smem_desc_sw128(sA + A_BYTES);
#ifndef RSB_EXP_NOMMA
          #pragma unroll
          for (int k = 0; k < RSB_KSTAGE / 32; k++)                // UMMA K = 32 bytes: +32 B = +2 in the address field
            umma_i8(d_tmem, adesc + 2u * k, bdesc + 2u * k, IDESC, (uint32_t) ((ks | k) != 0));
#endif
          umma_commit(empty_bar(stage));                           // frees the smem slot when these MMAs retire
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));                               // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> int64 count planes
    const int ew = (warp - 4) & 3;                                 // TMEM lane quarter owned by this warp (= warp id % 4)
    const int eg = (warp - 4) >> 2;                                // epilogue group: its share of every tile's columns
    constexpr int CPG = gram_cols_per_group(rsb_cj_for(S), GRAM_EPI_GROUPS);
    const int jl0 = eg * CPG, jl1 = (jl0 + CPG < rsb_cj_for(S)) ? jl0 + CPG : rsb_cj_for(S);
    const int m  = ew * 32 + lane;                                 // accumulator row = planeA row within the tile
    const int il = m >> 2, a = m & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int  r = rep0 + (int) (w / ntiles);
      const int2 t = tiles[(int) (w % ntiles)];
      const int  i = t.x * RSB_ICOLS + il;
      long long *base = cnt + ((size_t) r * 16 + (size_t) a * 4) * (size_t) L * Lp + (size_t) i * Lp;
      const size_t plane = (size_t) L * Lp;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) acc * 256u;

#ifdef RSB_EXP_NOEPI
      const double racc = 0.0;
      if (mrow != nullptr && i < L && trow == 0xffffffffu) mrow[(size_t) 0] = 0.0;
#else
      if (REC && rec_pairs_form<S>()) {
        double *rec0 = reinterpret_cast<double *>(cnt) + (size_t) r * 16 * plane;
        double *mrow_blk = mrow + (((size_t) r * nJB + t.y) * GRAM_EPI_GROUPS + eg) * (size_t) L * 4;
        gram_epilogue_rec_pairs<S>(trow, t.y, t.x * RSB_ICOLS + ew * 8, lane, ew, t.x, r, L, Lp, rec0, plane, scale, mrow_blk, mcol, nIB, small52 != 0, ltab,
                                   xbuf_all + (size_t) (warp - 4) * 128, jl0, jl1);
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      const double racc = gram_epilogue_tile<S, REC>(trow, t.y, i, a, lane, ew, t.x, r, L, Lp, base, plane, scale, mrow != nullptr, mcol, nIB, small52 != 0, ltab, jl0, jl1);
      if (mrow != nullptr && i < L) mrow[((((size_t) r * nJB + t.y) * GRAM_EPI_GROUPS + eg) * L + i) * 4 + a] = racc;
#endif
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
#ifdef RSB_BLOCKTRACE
    if ((threadIdx.x & 31) == 0) rsb_trace_put(gram_trace_buf, 1, t0);
#endif
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster, on the two SMs of one TPC, contract a 256-row x
// N-column tile together.  CTA r owns the accumulator rows of row block ib = 2 ibp + r (its own 128 planeA rows) and
// loads only HALF of the planeB tile; the tensor cores of both SMs read both halves.  Per CTA and stage that is 32 KB
// of operands instead of 48 KB -- a third less L2 -> SM traffic, the resource this kernel is bound by -- and six ring
// stages instead of four in the same shared memory.  Synchronisation:
//   full[s]    leader CTA's barrier; the TMA loads of BOTH CTAs complete their bytes on it (cp.async.bulk.tensor
//              .cta_group::2 with the barrier address' peer bit cleared); the leader's producer arms it with the pair's bytes
//   empty[s]   one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a]   one per CTA (commit multicast): accumulator a complete, each CTA's epilogue drains its own 128 TMEM lanes
//   tempty[a]  leader's barrier, 2 x 128 arrivals: the epilogue threads of both CTAs (the peer's arrive remotely)
// Only the leader's warp 1 issues MMAs.  The epilogue is the single-CTA kernel's.
constexpr int      NSTAGE2        = 6;
constexpr uint32_t PEER_BIT_MASK  = 0xFEFFFFFFu;        // shared::cluster address of the same location in the pair's even CTA

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"(tmap), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(bar), "h"((unsigned short) 3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GRAM_THREADS, 1)
gram_i8_pair_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapBh,
                    const int2 *__restrict__ tiles, int ntiles, int rep0, int nrep, int L, int Lp, int kstages,
                    long long *__restrict__ cnt, double scale, double *__restrict__ mrow, double *__restrict__ mcol, int nJB, int nIB, int small52)
{
  constexpr int      CJ          = rsb_cj_for(S);
  constexpr int      NT          = 4 * S * CJ;                 // UMMA N (whole planeB tile; each CTA stages NT / 2 rows)
  constexpr uint32_t A_BYTES     = RSB_MTILE * RSB_KSTAGE;
  constexpr uint32_t B_BYTES     = (NT / 2) * RSB_KSTAGE;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // instruction descriptor as in the single-CTA kernel with M = 256 (the pair's rows)
  constexpr uint32_t IDESC       = (2u << 4) | ((uint32_t) (NT >> 3) << 17) | ((uint32_t) (256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base  = smem_base + NSTAGE2 * STAGE_BYTES;
  auto full_bar   = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar  = [&](int s) { return bar_base + 8u * (NSTAGE2 + s); };
  auto tfull_bar  = [&](int a) { return bar_base + 8u * (2 * NSTAGE2 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE2 + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE2 + 4);
  volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *) (smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = (cta_rank == 0);
  const int  cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE2; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++)       { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                               // both CTAs' barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const long long nwork = (long long) ntiles * nrep;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long w = cluster_id; w < nwork; w += nclusters) {
        const int  r = rep0 + (int) (w / ntiles);
        const int2 t = tiles[(int) (w % ntiles)];                  // (ibp, jb)
        const int  ib = 2 * t.x + (int) cta_rank;
        for (int ks = 0; ks < kstages; ks++) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) mbar_expect_tx(full_bar(stage), 2u * STAGE_BYTES);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          tma_load_3d_pair(sA,           &tmapA,  full_bar(stage), ks * RSB_KSTAGE, ib * RSB_MTILE,                        r);
          tma_load_3d_pair(sA + A_BYTES, &tmapBh, full_bar(stage), ks * RSB_KSTAGE, t.y * NT + (int) cta_rank * (NT / 2), r);
          if (++stage == NSTAGE2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (lane == 0 && leader) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0;   uint32_t acc_phase = 0;
      for (long long w = cluster_id; w < nwork; w += nclusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);               // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t) acc * 256u;
        for (int ks = 0; ks < kstages; ks++) {
          mbar_wait(full_bar(stage), phase);                       // the bytes of both CTAs have landed
          tc_fence_after();
          const uint32_t sA    = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = smem_desc_sw128(sA);
          const uint64_t bdesc = smem_desc_sw128(sA + A_BYTES);
          #pragma unroll
          for (int k = 0; k < RSB_KSTAGE / 32; k++)
            umma_i8_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, IDESC, (uint32_t) ((ks | k) != 0));
          umma_commit_pair(empty_bar(stage));                      // frees the slot in both CTAs
          if (++stage == NSTAGE2) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(tfull_bar(acc));                          // accumulator complete -> both epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: this CTA's 128 TMEM lanes -> int64 count planes
    const int ew = warp - 4;
    const int m  = ew * 32 + lane;
    const int il = m >> 2, a = m & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long w = cluster_id; w < nwork; w += nclusters) {
      const int  r = rep0 + (int) (w / ntiles);
      const int2 t = tiles[(int) (w % ntiles)];
      const int  ib = 2 * t.x + (int) cta_rank;
      const int  i = ib * RSB_ICOLS + il;
      long long *base = cnt + ((size_t) r * 16 + (size_t) a * 4) * (size_t) L * Lp + (size_t) i * Lp;
      const size_t plane = (size_t) L * Lp;
      const bool   part  = (mrow != nullptr) && (ib < nIB);        // marginal partials of this row block

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) acc * 256u;
      const double racc = gram_epilogue_tile<S, false>(trow, t.y, i, a, lane, ew, ib, r, L, Lp, base, plane, scale, part, mcol, nIB, small52 != 0, nullptr);
      if (part && i < L) mrow[(((size_t) r * nJB + t.y) * L + i) * 4 + a] = racc;
      tc_fence_before();
      mbar_arrive_leader(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                               // the peer may still be using this CTA's barriers / TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int S> constexpr size_t gram_pair_smem_bytes() {
  return (size_t) NSTAGE2 * (RSB_MTILE * RSB_KSTAGE + 2 * S * rsb_cj_for(S) * RSB_KSTAGE) + 8 * (2 * NSTAGE2 + 4) + 16 + 1024;
}

template <int S, bool REC> constexpr size_t gram_smem_bytes() {
  return (size_t) NSTAGE * (RSB_MTILE * RSB_KSTAGE + 4 * S * rsb_cj_for(S) * RSB_KSTAGE) + 128 +
         (REC ? sizeof(double2) * LOGTAB_N + (rec_pairs_form<S>() ? (size_t) 2048 * 4 * GRAM_EPI_GROUPS : 0) : 0) + 1024;
}

template <int S, bool REC>
cudaError_t launch_gram(const CUtensorMap &tmA, const CUtensorMap &tmB, const int2 *tiles, int ntiles, int rep0, int nrep,
                        int L, int Lp, int kstages, long long *cnt, double scale, double *mrow, double *mcol, int nJB, int nIB,
                        int small52, int grid, const double2 *logtab, cudaStream_t st)
{
  constexpr size_t smem = gram_smem_bytes<S, REC>();
  cudaError_t e = cudaFuncSetAttribute(gram_i8_kernel<S, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  // the ring takes ~190 KB; with the SM's carve-out at its maximum (228 KB) the rest is left for the blocks of the
  // statistics chain, which run beside this kernel (a 196 KB carve-out would leave them no shared memory at all)
  rsb_coreside(gram_i8_kernel<S, REC>);
  gram_i8_kernel<S, REC><<<grid, GRAM1_THREADS, smem, st>>>(tmA, tmB, tiles, ntiles, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, logtab);
  return cudaGetLastError();
}

template <int S>
cudaError_t launch_gram_pair(const CUtensorMap &tmA, const CUtensorMap &tmBh, const int2 *tiles2, int ntiles2, int rep0, int nrep,
                             int L, int Lp, int kstages, long long *cnt, double scale, double *mrow, double *mcol, int nJB, int nIB,
                             int small52, int max_clusters, cudaStream_t st)
{
  constexpr size_t smem = gram_pair_smem_bytes<S>();
  cudaError_t e = cudaFuncSetAttribute(gram_i8_pair_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  rsb_coreside(gram_i8_pair_kernel<S>);
  const long long work = (long long) ntiles2 * nrep;
  const int nclusters = (int) (work < max_clusters ? work : max_clusters);
  gram_i8_pair_kernel<S><<<2 * nclusters, GRAM_THREADS, smem, st>>>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale,
                                                                      mrow, mcol, nJB, nIB, small52);
  return cudaGetLastError();
}

// how many CTA pairs of the pair kernel can be resident at once (SMs that cannot be paired inside their GPC stay idle)
template <int S>
int pair_clusters()
{
  constexpr size_t smem = gram_pair_smem_bytes<S>();
  if (cudaFuncSetAttribute(gram_i8_pair_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 148); cfg.blockDim = dim3(GRAM_THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gram_i8_pair_kernel<S>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

} // namespace

// Host entry used by capi.cu.  tmA/tmB are 3-D tensor maps {Kpad, rows, replicate} with 128B swizzle
// and boxes {128, 128, 1} / {128, 4*S*CJ, 1}.  logtab != NULL selects the record epilogue (G test, 16 classes): the slot's
// count planes then hold double [8][L][Lp] records instead of counts, and mrow / mcol must be given.
cudaError_t rsb_launch_gram_i8(int S, const CUtensorMap &tmA, const CUtensorMap &tmB, const int2 *tiles, int ntiles,
                               int rep0, int nrep, int L, int Lp, int kstages, long long *cnt, double scale, double *mrow,
                               double *mcol, int nJB, int nIB, int small52, int grid, const void *logtab, cudaStream_t st)
{
  const double2 *lt = (const double2 *) logtab;
  if (lt && !(mrow && mcol)) return cudaErrorInvalidValue;
#define RSB_GRAM_CASE(SS) \
  case SS: return lt ? launch_gram<SS, true>(tmA, tmB, tiles, ntiles, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, grid, lt, st) \
                     : launch_gram<SS, false>(tmA, tmB, tiles, ntiles, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, grid, nullptr, st);
  switch (S) {
  RSB_GRAM_CASE(1) RSB_GRAM_CASE(2) RSB_GRAM_CASE(3) RSB_GRAM_CASE(4) RSB_GRAM_CASE(5) RSB_GRAM_CASE(6)
  }
#undef RSB_GRAM_CASE
  return cudaErrorInvalidValue;
}

// CTA-pair variant: tmBh has boxes {128, 2*S*CJ, 1} (half a planeB tile), tiles2 lists (ibp, jb) = row-block pairs.
cudaError_t rsb_launch_gram_i8_pair(int S, const CUtensorMap &tmA, const CUtensorMap &tmBh, const int2 *tiles2, int ntiles2,
                                    int rep0, int nrep, int L, int Lp, int kstages, long long *cnt, double scale, double *mrow,
                                    double *mcol, int nJB, int nIB, int small52, int max_clusters, cudaStream_t st)
{
  switch (S) {
  case 1: return launch_gram_pair<1>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  case 2: return launch_gram_pair<2>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  case 3: return launch_gram_pair<3>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  case 4: return launch_gram_pair<4>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  case 5: return launch_gram_pair<5>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  case 6: return launch_gram_pair<6>(tmA, tmBh, tiles2, ntiles2, rep0, nrep, L, Lp, kstages, cnt, scale, mrow, mcol, nJB, nIB, small52, max_clusters, st);
  default: return cudaErrorInvalidValue;
  }
}

// row-partial blocks per tile column block jb in mrow: one per epilogue group (single-CTA kernel), one (CTA-pair kernel)
int rsb_gram_mrow_blocks(int pair) { return pair ? 1 : GRAM_EPI_GROUPS; }

int rsb_gram_pair_clusters(int S)
{
  switch (S) {
  case 1: return pair_clusters<1>(); case 2: return pair_clusters<2>(); case 3: return pair_clusters<3>();
  case 4: return pair_clusters<4>(); case 5: return pair_clusters<5>(); case 6: return pair_clusters<6>();
  default: return 0;
  }
}

#ifdef RSB_BLOCKTRACE
extern "C" void rsb_trace_set_gram(unsigned long long *buf) { cudaMemcpyToSymbol(gram_trace_buf, &buf, sizeof(buf)); }
#endif
