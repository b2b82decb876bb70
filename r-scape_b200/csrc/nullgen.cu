// nullgen.cu -- null-alignment generators on the device.
//
// (B) null_simulate_level_kernel: cov_GenerateAlignment's ungapped, structure-free path
//     (src/cov_simulate.c:289-324 tree walk, :585-631 emission, :724-773 inverse-CDF draw) with
//     P(t) = exp(tQ) per branch (src/ratematrix.c:185-233; matrices are built on the host in capi.cu,
//     4x4 per branch, and shipped as cumulative integer thresholds).  Every (replicate, column) is an independent chain
//     down the tree; the tree is walked one level per launch (nodes of a level are independent given their parents), a
//     thread owns 4 columns of one (replicate, node).
//     Randomness: Philox4x32-10 keyed by (seed, replicate), counter (node, column pair): one 128-bit block
//     serves both children of a node in two columns.  The reference consumes one Mersenne-Twister stream in branch-major
//     order; the streams differ, the per-draw distribution is the same (validated distributionally).
//
// (A) Fitch + shuffle, R-scape's default null (src/R-scape.c:1653-1668):
//     fitch_up/down_level tree_fitch_column (src/msatree.c:1700-1831): sets as 5-bit masks; post-order = one launch per
//                         tree level from the deepest up, pre-order = from the root down; a thread owns 4 columns of one
//                         (replicate, node); the up pass is shared by all replicates when no residue is unknown.
//     permutation_kernel + permute_root_kernel msamanip_ShuffleColumns (src/msamanip.c:1164-1233).  Only the root's permuted row is
//                         ever read (every other row is overwritten by its parent's row at msamanip.c:1645).
//     replay_level_row_kernel shuffle_tree_substitutions + shuffle_tree_substitute_all (src/msamanip.c:1597-1780):
//                         one warp per (replicate, branch) of one tree level; counts the 5x5 substitutions of the
//                         branch on the Fitch rows, copies the shuffled parent row to the child and re-places
//                         the substitutions at positions drawn uniformly without replacement among the columns
//                         holding the source residue (rejection-sampled ranks instead of the reference's
//                         Fisher-Yates over an index list: same distribution, exact counts).
//     replay_level_kernel the same with one thread per branch and one-pass selection sampling; fallback for alignments
//                         whose rank table does not fit in shared memory.
#include "rsb_common.cuh"
#include <stdlib.h>

namespace {

struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ static void round1(uint32_t *c, const uint32_t *k) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) const {
    uint32_t c[4] = { c0, c1, c2, c3 }, k[2] = { key[0], key[1] };
    #pragma unroll
    for (int r = 0; r < 10; r++) { round1(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};


// ------------------------------------------------------------------------------------------------ generator B
// One launch per tree level (nodes of a level are independent given their parents); a thread owns W columns (W = 4: rows
// moved as 32-bit words, L % 4 == 0; W = 1 otherwise) of one (replicate, node) and emits both children of the node.
// grid = (columns/(128 W), nodes of the level, replicates).  One Philox block serves two columns (4 draws: 2 children x 2
// columns; block (node, column >> 1), element 2 (column & 1) + side).  The branch's cumulative matrices arrive as integer
// thresholds thr = ceil(cdf 2^32): "first k with cdf_k > x", x = rnd / 2^32 (cov_addres, :757-773) is exactly
// "first k with rnd < thr_k", with no floating point in the kernel.
template <int W>
__global__ void __launch_bounds__(128)
null_simulate_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order,
                           int lvl_begin, const unsigned long long *__restrict__ pthr, int N, int L, const uint8_t *__restrict__ root,
                           const uint8_t *__restrict__ gapmask, unsigned long long seed, unsigned long long id0,
                           int first_rep, uint8_t *__restrict__ res, uint8_t *__restrict__ scratch)
{
  __shared__ unsigned long long thr[32];                         // [side][parent residue][k]
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (id0 + blockIdx.z);            // global replicate id: the stream does not depend on where the replicate is stored
  if (threadIdx.x < 32) thr[threadIdx.x] = pthr[(size_t) v * 32 + threadIdx.x];
  __syncthreads();
  if (c0 >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ (0x9E3779B9u * (rid + 1u)); rng.key[1] = (uint32_t) (seed >> 32) + rid;
  uint8_t *anc  = scratch + (size_t) r * (N - 1) * L;        // internal node states [N-1][L]
  uint8_t *leaf = res + (size_t) r * N * L;
  const uint8_t *prow = (v == 0) ? root : anc + (size_t) v * L;             // cov_add_root for the root
  const uint32_t pw = (W == 4) ? *reinterpret_cast<const uint32_t *>(prow + c0) : (uint32_t) prow[c0];
  const int kl = left[v], kr = right[v];
  uint32_t ol = 0, orr = 0;
  uint32_t rnd[4];
  #pragma unroll
  for (int q = 0; q < W; q++) {
    const int c = c0 + q;
    if ((q & 1) == 0 || W == 1) rng.block((uint32_t) v, (uint32_t) (c >> 1), 0x5eedu, 0u, rnd);
    const int par = (pw >> (8 * q)) & 3;
    #pragma unroll
    for (int side = 0; side < 2; side++) {
      const unsigned long long x = rnd[2 * (c & 1) + side];
      const unsigned long long *t = thr + side * 16 + par * 4;
      const uint32_t k = (uint32_t) !(x < t[0]) + (uint32_t) !(x < t[1]) + (uint32_t) !(x < t[2]);      // first k with rnd < thr_k, else 3
      if (side) orr |= k << (8 * q); else ol |= k << (8 * q);
    }
  }
  #pragma unroll
  for (int side = 0; side < 2; side++) {
    const int child = side ? kr : kl;
    uint32_t o = side ? orr : ol;
    if (child > 0) {
      if (W == 4) *reinterpret_cast<uint32_t *>(anc + (size_t) child * L + c0) = o;
      else anc[(size_t) child * L + c0] = (uint8_t) o;
    } else {
      if (gapmask) {                                              // gaps and unknowns of the input alignment are stamped on the leaves
        const uint32_t g = (W == 4) ? *reinterpret_cast<const uint32_t *>(gapmask + (size_t) (-child) * L + c0) : (uint32_t) gapmask[(size_t) (-child) * L + c0];
        #pragma unroll
        for (int q = 0; q < W; q++) { const uint32_t gq = (g >> (8 * q)) & 0xFFu; if (gq >= 4u) o = (o & ~(0xFFu << (8 * q))) | (gq << (8 * q)); }
      }
      if (W == 4) *reinterpret_cast<uint32_t *>(leaf + (size_t) (-child) * L + c0) = o;
      else leaf[(size_t) (-child) * L + c0] = (uint8_t) o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ generator A
__device__ __forceinline__ int pick_member(unsigned set, uint32_t rnd)      // uniform member of a non-empty 5-bit set
{
  const int n = __popc(set);
  const int k = (int) (((unsigned long long) rnd * (unsigned) n) >> 32);
  return __fns(set, 0, k + 1);
}

// Fitch sets, one launch per tree level from the deepest up (:1758-1777, :1869-1905): a thread owns W columns (W = 4:
// rows moved as 32-bit words, L % 4 == 0; W = 1 otherwise) of one (replicate, node).
// grid = (columns/(128 W), nodes of the level, replicates).
template <int W>
__global__ void fitch_up_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin,
                                      int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0,
                                      const unsigned long long *__restrict__ ids, int first_rep, uint8_t *__restrict__ ancbuf)
{
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (ids ? ids[blockIdx.z] : id0 + blockIdx.z);
  if (c0 >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ 0xF17C4u; rng.key[1] = (uint32_t) (seed >> 32) ^ (0x85EBCA6Bu * (rid + 1u));
  uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  auto row_sets = [&](int n) -> uint32_t {                                              // sets of W columns of child n, one per byte
    if (n > 0) return (W == 4) ? *reinterpret_cast<const uint32_t *>(anc + (size_t) n * L + c0) : (uint32_t) anc[(size_t) n * L + c0];
    const uint32_t xw = (W == 4) ? *reinterpret_cast<const uint32_t *>(msa + (size_t) (-n) * L + c0) : (uint32_t) msa[(size_t) (-n) * L + c0];
    uint32_t out = 0;
    #pragma unroll
    for (int q = 0; q < W; q++) {
      const int x = (xw >> (8 * q)) & 0xFF;
      unsigned S;
      if (x <= 4) S = 1u << x;
      else {                                                                            // unknown -> one of the 5 at random (:1735)
        uint32_t rnd[4]; rng.block((uint32_t) (-n), (uint32_t) (c0 + q), 0x1eafu, 0u, rnd);
        S = 1u << (int) (((unsigned long long) rnd[0] * 5u) >> 32);
      }
      out |= S << (8 * q);
    }
    return out;
  };
  const uint32_t Sl = row_sets(left[v]), Sr = row_sets(right[v]);
  uint32_t S = Sl & Sr;                                                                 // per byte: intersection ...
  const uint32_t un = (Sl | Sr) & 0x0F0F0F0Fu;                                          // ... else the union of residues; a gap never joins
  const uint32_t empty = __vcmpeq4(S, 0u);                                              // 0xFF in the bytes whose intersection is empty
  S |= un & empty;
  if (W == 4) *reinterpret_cast<uint32_t *>(anc + (size_t) v * L + c0) = S;
  else anc[(size_t) v * L + c0] = (uint8_t) S;
}

// Traceback, one launch per level from the root down (:1779-1815): each internal child gets the parent's residue if that
// is in the child's Fitch set, else a uniform member of the set.  A thread owns W columns (W = 4: rows moved as 32-bit
// words, L % 4 == 0; W = 1 otherwise) of one (replicate, node).  The sets come from `sets` (replicate stride
// sets_stride; 0 when the up pass is shared by all replicates, or the residue buffer itself when it ran per replicate);
// the Philox block of a (node, column) is only computed where a random choice is really needed -- most branches carry
// no substitution in most columns.
template <int W>
__global__ void fitch_down_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin,
                                        int N, int L, unsigned long long seed, unsigned long long id0, const unsigned long long *__restrict__ ids, int first_rep,
                                        const uint8_t *sets, size_t sets_stride, uint8_t *ancbuf)
{
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (ids ? ids[blockIdx.z] : id0 + blockIdx.z);
  if (c0 >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ 0xF17C4u; rng.key[1] = (uint32_t) (seed >> 32) ^ (0x85EBCA6Bu * (rid + 1u));
  uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  const uint8_t *st = sets + (size_t) blockIdx.z * sets_stride;
  uint32_t axw = 0;
  if (v == 0) {                                                                          // root: uniform member of its set (:1779)
    #pragma unroll
    for (int q = 0; q < W; q++) {
      uint32_t rnd[4]; rng.block(0xffffffffu, (uint32_t) (c0 + q), 0x600du, 0u, rnd);
      axw |= (uint32_t) pick_member(st[c0 + q], rnd[0]) << (8 * q);
    }
    if (W == 4) *reinterpret_cast<uint32_t *>(anc + c0) = axw;
    else anc[c0] = (uint8_t) axw;
  } else axw = (W == 4) ? *reinterpret_cast<const uint32_t *>(anc + (size_t) v * L + c0) : (uint32_t) anc[(size_t) v * L + c0];
  const int kl = left[v], kr = right[v];
  const uint32_t Sl = (kl <= 0) ? 0u : (W == 4) ? *reinterpret_cast<const uint32_t *>(st + (size_t) kl * L + c0) : (uint32_t) st[(size_t) kl * L + c0];
  const uint32_t Sr = (kr <= 0) ? 0u : (W == 4) ? *reinterpret_cast<const uint32_t *>(st + (size_t) kr * L + c0) : (uint32_t) st[(size_t) kr * L + c0];
  uint32_t ol = 0, orr = 0;
  #pragma unroll
  for (int q = 0; q < W; q++) {
    const int ax = (axw >> (8 * q)) & 0xFF;
    const unsigned sl = (Sl >> (8 * q)) & 0xFF, sr = (Sr >> (8 * q)) & 0xFF;
    int xl = ax, xr = ax;
    const bool needl = (kl > 0) && !((sl >> ax) & 1u), needr = (kr > 0) && !((sr >> ax) & 1u);
    if (needl || needr) {
      uint32_t rnd[4]; rng.block((uint32_t) v, (uint32_t) (c0 + q), 0xd0c0u, 0u, rnd);
      if (needl) xl = pick_member(sl, rnd[0]);
      if (needr) xr = pick_member(sr, rnd[1]);
    }
    ol |= (uint32_t) xl << (8 * q); orr |= (uint32_t) xr << (8 * q);
  }
  if (kl > 0) { if (W == 4) *reinterpret_cast<uint32_t *>(anc + (size_t) kl * L + c0) = ol;  else anc[(size_t) kl * L + c0] = (uint8_t) ol; }
  if (kr > 0) { if (W == 4) *reinterpret_cast<uint32_t *>(anc + (size_t) kr * L + c0) = orr; else anc[(size_t) kr * L + c0] = (uint8_t) orr; }
}

// does the alignment hold residues other than A C G U and the gap?  (they are resolved at random per replicate, :1735,
// which makes the Fitch sets replicate-specific)
__global__ void unknown_flag_kernel(const uint8_t *__restrict__ msa, size_t n, int *__restrict__ flag)
{
  // bit 0: unknown residues (N = 15) present; bit 1: codes the reference's Fitch pass cannot set up (degenerate symbols, '*', '~':
  // only esl_abc_XIsUnknown gets the uniform set, anything else fails with "S not set up properly", src/msatree.c:1731-1740)
  bool any = false, bad = false;
  for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t) gridDim.x * blockDim.x) {
    const uint8_t c = msa[k];
    any |= c > 4;
    bad |= c > 4 && c != 15;
  }
  const int a1 = __syncthreads_or(any), b1 = __syncthreads_or(bad);
  if (threadIdx.x == 0 && (a1 || b1)) atomicOr(flag, (a1 ? 1 : 0) | (b1 ? 2 : 0));
}

// one random permutation per replicate (Fisher-Yates by one thread; L is a few thousand).  It depends on (seed, replicate id)
// only, so all replicates of a generator call are done by ONE launch before the per-chunk work.
__global__ void permutation_kernel(int L, unsigned long long seed, unsigned long long id0, const unsigned long long *__restrict__ ids, int first_rep,
                                   int *__restrict__ permbuf)
{
  const int r = first_rep + blockIdx.x;
  const uint32_t rid = (uint32_t) (ids ? ids[blockIdx.x] : id0 + blockIdx.x);
  int *perm = permbuf + (size_t) r * L;
  for (int c = threadIdx.x; c < L; c += blockDim.x) perm[c] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    Philox rng; rng.key[0] = (uint32_t) seed ^ 0x5AFF1Eu; rng.key[1] = (uint32_t) (seed >> 32) + 0x27D4EB2Fu * (rid + 1u);
    uint32_t rnd[4]; int have = 0;
    for (int n = L; n > 1; n--) {                                                          // esl_vec_IShuffle
      if (!have) { rng.block((uint32_t) n, 0u, 0x9e37u, 0u, rnd); have = 4; }
      const int w = (int) (((unsigned long long) rnd[--have] * (unsigned) n) >> 32);
      const int t = perm[w]; perm[w] = perm[n - 1]; perm[n - 1] = t;
    }
  }
}

// the permuted root row (msamanip_ShuffleColumns, src/msamanip.c:1164-1233; only the root's permuted row is ever read)
__global__ void permute_root_kernel(int N, int L, int first_rep, const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf,
                                    const int *__restrict__ permbuf)
{
  const int r = first_rep + blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L) return;
  shancbuf[(size_t) r * (N - 1) * L + c] = ancbuf[(size_t) r * (N - 1) * L + permbuf[(size_t) r * L + c]];
}

constexpr int RP_THREADS = 128;
constexpr int RP_BATCH   = 8;        // 32-bit words (4 alignment columns each) loaded together

// PCG32 (XSH-RR): the cheap sequential stream of one (replicate, branch), seeded from a Philox block
struct Pcg32 {
  unsigned long long state, inc;
  __device__ __forceinline__ uint32_t next() {
    const unsigned long long old = state;
    state = old * 6364136223846793005ULL + inc;
    const uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u), rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
  }
};

// One thread per (replicate, branch of this tree level).  Pass 1 counts the branch's 5x5 substitutions on the Fitch rows
// and the residue classes of the parent row (the shuffled parent row has the same composition: the root row is a
// permutation, and re-placing a branch's substitutions changes the composition exactly as the branch did); pass 2 walks
// the shuffled parent row once and re-places the substitutions by sequential selection sampling (Knuth's Algorithm S):
// a position currently holding class a is picked with probability k_a / m_a (substitutions out of a still to place /
// a-positions still to come), a picked position takes target d with probability n_{a->d} / k_a.  Every subset of positions
// and every assignment of targets is equally likely, exactly as after the reference's Fisher-Yates shuffle of the
// position list (src/msamanip.c:1718-1757), and the substitution counts are reproduced exactly.
// The per-class counters live in registers (selected by predication): a thread is alone on its rows, so the length of
// the dependent instruction chain per position is what bounds a tree level's latency.
// WORD: rows are read and written as 32-bit words (L % 4 == 0), otherwise byte by byte.
#define RP_SEL5(c, v0, v1, v2, v3, v4) ((c) == 0 ? (v0) : (c) == 1 ? (v1) : (c) == 2 ? (v2) : (c) == 3 ? (v3) : (v4))

template <bool WORD>
__global__ void __launch_bounds__(RP_THREADS)
replay_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin, int lvl_count,
                    int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0, const unsigned long long *__restrict__ ids,
                    int first_rep, int nrep, const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf, uint8_t *__restrict__ res)
{
  __shared__ unsigned short nsub[25][RP_THREADS];     // substitutions a -> d still to place (touched on differences / picks only)
  const int t = threadIdx.x;
  const int L4 = L / 4;
  const long long task = (long long) blockIdx.x * RP_THREADS + t;
  if (task >= 2LL * lvl_count * nrep) return;
  const int side = (int) (task & 1);
  const long long nt = task >> 1;
  const int rr = (int) (nt / lvl_count);
  const int r = first_rep + rr;
  const uint32_t rid = (uint32_t) (ids ? ids[rr] : id0 + (unsigned long long) rr);
  const int v = order[lvl_begin + (int) (nt % lvl_count)];
  const uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  uint8_t *shanc = shancbuf + (size_t) r * (N - 1) * L;
  uint8_t *leaves = res + (size_t) r * N * L;
  const uint8_t *par_o = anc + (size_t) v * L;
  const uint8_t *par_s = shanc + (size_t) v * L;
  const int ch = side ? right[v] : left[v];
  const uint8_t *kid_o = (ch > 0) ? anc + (size_t) ch * L : msa + (size_t) (-ch) * L;
  uint8_t *kid_s = (ch > 0) ? shanc + (size_t) ch * L : leaves + (size_t) (-ch) * L;

  Philox ph; ph.key[0] = (uint32_t) seed ^ (0xC2B2AE35u * (rid + 1u)); ph.key[1] = (uint32_t) (seed >> 32) ^ 0x5bd1e995u;
  uint32_t sd[4]; ph.block((uint32_t) v, 0x7ee1u + (uint32_t) side, 0u, 0u, sd);
  Pcg32 rng; rng.state = ((unsigned long long) sd[0] << 32) | sd[1]; rng.inc = ((((unsigned long long) sd[2] << 32) | sd[3]) << 1) | 1ULL;
  rng.next();

  #pragma unroll
  for (int k = 0; k < 25; k++) nsub[k][t] = 0;
  int m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;

  // pass 1: composition of the parent row; substitutions of this branch on the Fitch rows (msamanip.c:1634-1643)
  auto count1 = [&](int pa, int kd) {
    m0 += (pa == 0); m1 += (pa == 1); m2 += (pa == 2); m3 += (pa == 3); m4 += (pa == 4);
    if (pa != kd && pa <= 4 && kd <= 4) nsub[pa * 5 + kd][t]++;
  };
  if (WORD) {
    const uint32_t *po4 = reinterpret_cast<const uint32_t *>(par_o), *ko4 = reinterpret_cast<const uint32_t *>(kid_o);
    for (int c4 = 0; c4 < L4; c4 += RP_BATCH) {          // loads of a batch are issued together (latency hiding within the thread)
      uint32_t wp[RP_BATCH], wk[RP_BATCH];
      #pragma unroll
      for (int u = 0; u < RP_BATCH; u++) { const bool in = c4 + u < L4; wp[u] = in ? __ldcg(po4 + c4 + u) : 0x05050505u; wk[u] = in ? __ldcg(ko4 + c4 + u) : 0x05050505u; }
      #pragma unroll
      for (int u = 0; u < RP_BATCH; u++)
        #pragma unroll
        for (int q = 0; q < 4; q++) count1((wp[u] >> (8 * q)) & 0xff, (wk[u] >> (8 * q)) & 0xff);
    }
  } else {
    for (int c = 0; c < L; c++) count1(par_o[c], kid_o[c]);
  }
  int k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0;
  #pragma unroll
  for (int d = 0; d < 5; d++) { k0 += nsub[d][t]; k1 += nsub[5 + d][t]; k2 += nsub[10 + d][t]; k3 += nsub[15 + d][t]; k4 += nsub[20 + d][t]; }
  int ktot = k0 + k1 + k2 + k3 + k4;

  // pass 2: copy the shuffled parent row (:1645) and re-place the substitutions
  auto place = [&](int cls) -> int {
    if (cls > 4) return cls;
    const int k = RP_SEL5(cls, k0, k1, k2, k3, k4);
    const int m = RP_SEL5(cls, m0, m1, m2, m3, m4);
    int out = cls;
    if (k > 0 && (int) __umulhi(rng.next(), (uint32_t) (m > 0 ? m : 1)) < k) {
      int pick = (int) __umulhi(rng.next(), (uint32_t) k);
      int d = 0;
      for (; d < 4; d++) { const int n = nsub[cls * 5 + d][t]; if (pick < n) break; pick -= n; }
      nsub[cls * 5 + d][t]--;
      k0 -= (cls == 0); k1 -= (cls == 1); k2 -= (cls == 2); k3 -= (cls == 3); k4 -= (cls == 4);
      ktot--;
      out = d;
    }
    m0 -= (cls == 0); m1 -= (cls == 1); m2 -= (cls == 2); m3 -= (cls == 3); m4 -= (cls == 4);
    return out;
  };
  if (WORD) {
    const uint32_t *ps4 = reinterpret_cast<const uint32_t *>(par_s);
    uint32_t *ks4 = reinterpret_cast<uint32_t *>(kid_s);
    for (int c4 = 0; c4 < L4; c4 += RP_BATCH) {
      uint32_t wb[RP_BATCH];
      #pragma unroll
      for (int u = 0; u < RP_BATCH; u++) wb[u] = (c4 + u < L4) ? __ldcg(ps4 + c4 + u) : 0u;
      #pragma unroll
      for (int u = 0; u < RP_BATCH; u++) {
        if (c4 + u >= L4) break;
        uint32_t w = wb[u];
        if (ktot > 0) {
          uint32_t o = 0;
          #pragma unroll
          for (int q = 0; q < 4; q++) o |= (uint32_t) place((w >> (8 * q)) & 0xff) << (8 * q);
          w = o;
        }
        ks4[c4 + u] = w;
      }
    }
  } else {
    for (int c = 0; c < L; c++) { const int cls = par_s[c]; kid_s[c] = (uint8_t) (ktot > 0 ? place(cls) : cls); }
  }
}

// Row variant of the replay: one WARP per (replicate, branch), rows moved as coalesced 32-bit words (W = 4 columns per
// lane; W = 1 when L % 4 != 0).
//   sweep 1  (Fitch rows of parent and child): composition m_a of the parent row and the substitution counts n[a->d]
//            (msamanip.c:1634-1643); the shuffled parent row has the same composition (see replay_level_kernel).
//   draw     for every source class a, k_a = sum_d n[a->d] distinct RANKS in [0, m_a) by parallel rejection sampling:
//            each lane draws a candidate, duplicates inside the round are resolved towards the lowest lane
//            (__match_any_sync, deterministic), a 4-bit code table in shared memory (claimed bit + target) says which ranks
//            are taken; the i-th accepted draw takes the i-th target of the multiset {d x n[a->d]}.  The accepted ranks
//            are a uniformly random ordered sample without replacement, so every subset of positions and every assignment
//            of targets is equally likely -- the distribution of the reference's Fisher-Yates shuffle (:1718-1757).
//   sweep 2  (shuffled parent row -> child row): a column of class a is the (running count of a)-th of its class; the
//            code table says whether that rank was drawn and what it becomes (:1645-1757).
// Work per branch is a few coalesced passes over its rows, independent of how the substitutions fall.
constexpr int RPR_WARPS = 2;        // 2 warps x 3.2 KB (packed code table + staged row at L = 1800)

template <int W, bool SEG>
__global__ void __launch_bounds__(RPR_WARPS * 32)
replay_level_row_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin, int lvl_count,
                        int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0, const unsigned long long *__restrict__ ids,
                        int first_rep, int nrep, const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf, uint8_t *__restrict__ res, int code_words, int ws_log2)
{
  extern __shared__ unsigned rpr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int WS = 1 << ws_log2, WSP = WS + 1;                                    // SEG: words per lane segment, padded stride (odd: no bank conflicts)
  const int warp_words = code_words + 32 + (SEG ? 32 * WSP : 0);
  // one nibble per column: bit 3 = rank drawn, bits 2:0 = target.  The five classes' tables are packed back to back -- class a
  // starts at nibble m_0 + ... + m_{a-1}, the compositions sum to <= L -- so the table is L nibbles instead of 5 L: half the
  // shared memory per warp, twice the resident warps of this latency-bound kernel
  unsigned *code = rpr_smem + (size_t) warp * warp_words;
  int *nsub = reinterpret_cast<int *>(code + code_words);                      // [25]
  const long long task = (long long) blockIdx.x * RPR_WARPS + warp;
  if (task >= 2LL * lvl_count * nrep) return;                                   // whole warp
  const int side = (int) (task & 1);
  const long long nt = task >> 1;
  const int rr = (int) (nt / lvl_count);
  const int r = first_rep + rr;
  const uint32_t rid = (uint32_t) (ids ? ids[rr] : id0 + (unsigned long long) rr);
  const int v = order[lvl_begin + (int) (nt % lvl_count)];
  const uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  uint8_t *shanc = shancbuf + (size_t) r * (N - 1) * L;
  uint8_t *leaves = res + (size_t) r * N * L;
  const uint8_t *par_o = anc + (size_t) v * L;
  const uint8_t *par_s = shanc + (size_t) v * L;
  const int ch = side ? right[v] : left[v];
  const uint8_t *kid_o = (ch > 0) ? anc + (size_t) ch * L : msa + (size_t) (-ch) * L;
  uint8_t *kid_s = (ch > 0) ? shanc + (size_t) ch * L : leaves + (size_t) (-ch) * L;
  const int LW = (L + W - 1) / W;                                               // row length in lane units
  const unsigned lt = (1u << lane) - 1u;

  auto load = [&](const uint8_t *row, int u) -> uint32_t {
    if (W == 4) return __ldcg(reinterpret_cast<const uint32_t *>(row) + u);
    return (uint32_t) row[u] | 0xFFFFFF00u;
  };

  for (int k = lane; k < code_words + 32; k += 32) code[k] = 0u;               // (nsub included)
  __syncwarp();

  // ---- sweep 1
  int m[5] = { 0, 0, 0, 0, 0 };
  for (int u0 = 0; u0 < LW; u0 += 64) {                                         // two units per lane in flight
    uint32_t wp[2], wk[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const int u = u0 + h * 32 + lane;
      wp[h] = (u < LW) ? load(par_o, u) : 0xFFFFFFFFu;
      wk[h] = (u < LW) ? load(kid_o, u) : 0xFFFFFFFFu;
    }
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      #pragma unroll
      for (int a = 0; a < 5; a++) m[a] += __popc(__vcmpeq4(wp[h], 0x01010101u * a) & 0x01010101u);
      if (wp[h] != wk[h]) {
        #pragma unroll
        for (int q = 0; q < W; q++) {
          const int pa = (wp[h] >> (8 * q)) & 0xFF, kd = (wk[h] >> (8 * q)) & 0xFF;
          if (pa != kd && pa <= 4 && kd <= 4) atomicAdd(&nsub[pa * 5 + kd], 1);
        }
      }
    }
  }
  #pragma unroll
  for (int a = 0; a < 5; a++) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[a] += __shfl_xor_sync(0xffffffffu, m[a], o);
  }
  __syncwarp();
  int ka[5], ktot = 0;
  #pragma unroll
  for (int a = 0; a < 5; a++) { ka[a] = nsub[a * 5] + nsub[a * 5 + 1] + nsub[a * 5 + 2] + nsub[a * 5 + 3] + nsub[a * 5 + 4]; ktot += ka[a]; }

  // ---- draw
  if (ktot > 0) {
    Philox ph; ph.key[0] = (uint32_t) seed ^ (0xC2B2AE35u * (rid + 1u)); ph.key[1] = (uint32_t) (seed >> 32) ^ 0x5bd1e995u;
    uint32_t sd[4]; ph.block((uint32_t) v, 0x7ee1u + (uint32_t) side, 0x3041u + (uint32_t) lane, 0u, sd);
    Pcg32 rng; rng.state = ((unsigned long long) sd[0] << 32) | sd[1]; rng.inc = ((((unsigned long long) sd[2] << 32) | sd[3]) << 1) | 1ULL;
    rng.next();
    unsigned offa = 0;                                                          // first nibble of class a's table
    #pragma unroll 1
    for (int a = 0; a < 5; a++) {
      const unsigned o = offa;
      offa += (unsigned) m[a];
      if (ka[a] == 0) continue;                                                 // uniform
      int c0 = nsub[a * 5], c1 = c0 + nsub[a * 5 + 1], c2 = c1 + nsub[a * 5 + 2], c3 = c2 + nsub[a * 5 + 3];   // cumulative targets
      int succ = 0;
      while (succ < ka[a]) {                                                    // uniform
        const unsigned cand = __umulhi(rng.next(), (uint32_t) m[a]);
        const unsigned same = __match_any_sync(0xffffffffu, cand);
        bool ok = false;
        const unsigned nibble = o + cand, sh = (nibble & 7u) * 4u;
        unsigned *slot = code + (nibble >> 3);
        if ((same & lt) == 0u) ok = !((atomicOr(slot, 8u << sh) >> sh) & 8u);  // lowest lane of its value claims the rank
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int idx = succ + __popc(bal & lt);
          if (idx < ka[a]) atomicOr(slot, (unsigned) ((idx >= c0) + (idx >= c1) + (idx >= c2) + (idx >= c3)) << sh);
          else             atomicAnd(slot, ~(0xFu << sh));                         // more accepted than needed: give the rank back
        }
        succ += __popc(bal);
      }
    }
  }
  __syncwarp();

  // ---- sweep 2, segment form (SEG): the row is staged in shared memory, every lane owns a contiguous segment of it,
  // counts its classes (five 12-bit fields of one 64-bit word), one warp scan gives the rank at which its segment starts
  // in every class, and the lane then walks its segment alone -- no cross-lane traffic per column.
  if (SEG) {
    const uint32_t *ps4 = reinterpret_cast<const uint32_t *>(par_s);
    uint32_t *ks4 = reinterpret_cast<uint32_t *>(kid_s);
    if (ktot == 0) {                                                            // uniform: nothing to place, plain copy (:1645)
      for (int u = lane; u < LW; u += 32) ks4[u] = __ldcg(ps4 + u);
      return;
    }
    unsigned *rowbuf = code + code_words + 32;
    for (int u = lane; u < LW; u += 32) rowbuf[(u >> ws_log2) * WSP + (u & (WS - 1))] = __ldcg(ps4 + u);
    __syncwarp();
    unsigned *seg = rowbuf + lane * WSP;
    int nw = LW - lane * WS; nw = nw < 0 ? 0 : (nw > WS ? WS : nw);
    unsigned kmask = 0;
    #pragma unroll
    for (int a = 0; a < 5; a++) kmask |= (ka[a] > 0 ? 1u : 0u) << a;
    unsigned long long cnt = 0;
    for (int t = 0; t < nw; t++) {
      const uint32_t w = seg[t];
      #pragma unroll
      for (int q = 0; q < 4; q++) { const unsigned x = (w >> (8 * q)) & 0xFFu; cnt += 1ULL << (12u * (x < 5u ? x : 5u)); }
    }
    unsigned long long incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    unsigned long long runp = incl - cnt;                                       // ranks at which this lane's segment starts ...
    {                                                                           // ... as nibble indices into the packed table
      unsigned long long offs = 0; unsigned o = 0;
      #pragma unroll
      for (int a = 0; a < 5; a++) { offs |= (unsigned long long) o << (12 * a); o += (unsigned) m[a]; }
      runp += offs;                                                             // (first nibble + rank < L <= 4092: the 12-bit fields hold)
    }
    for (int t = 0; t < nw; t++) {
      const uint32_t w = seg[t];
      uint32_t o = w;
      #pragma unroll
      for (int q = 0; q < 4; q++) {
        const unsigned x = (w >> (8 * q)) & 0xFFu, xs = 12u * (x < 5u ? x : 5u);
        const unsigned rank = (unsigned) (runp >> xs) & 0xFFFu;
        runp += 1ULL << xs;
        if (x < 5u && ((kmask >> x) & 1u)) {
          const unsigned nib = (code[rank >> 3] >> ((rank & 7u) * 4u)) & 0xFu;
          if (nib & 8u) o = (o & ~(0xFFu << (8 * q))) | ((nib & 7u) << (8 * q));
        }
      }
      seg[t] = o;
    }
    __syncwarp();
    for (int u = lane; u < LW; u += 32) ks4[u] = rowbuf[(u >> ws_log2) * WSP + (u & (WS - 1))];
    return;
  }

  // ---- sweep 2, ballot form
  int run[5] = { 0, m[0], m[0] + m[1], m[0] + m[1] + m[2], m[0] + m[1] + m[2] + m[3] };   // running nibble index of every class in the packed table
  for (int u0 = 0; u0 < LW; u0 += 64) {
    uint32_t w[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) { const int u = u0 + h * 32 + lane; w[h] = (u < LW) ? load(par_s, u) : 0xFFFFFFFFu; }
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const int u = u0 + h * 32 + lane;
      uint32_t o = w[h];
      if (ktot > 0) {                                                           // uniform
        #pragma unroll
        for (int q = 0; q < W; q++) {
          const int x = (w[h] >> (8 * q)) & 0xFF;
          #pragma unroll
          for (int a = 0; a < 5; a++) {
            if (ka[a] == 0) continue;                                           // uniform
            const unsigned bal = __ballot_sync(0xffffffffu, x == a);
            if (x == a) {
              const int rank = run[a] + __popc(bal & lt);
              const unsigned nib = (code[rank >> 3] >> ((rank & 7) * 4)) & 0xFu;
              if (nib & 8u) o = (o & ~(0xFFu << (8 * q))) | ((nib & 7u) << (8 * q));
            }
            run[a] += __popc(bal);
          }
        }
      }
      if (u < LW) {
        if (W == 4) reinterpret_cast<uint32_t *>(kid_s)[u] = o;
        else kid_s[u] = (uint8_t) o;
      }
    }
  }
}

// Position form of the replay (RSCAPE_B200_REPLAY=position): one warp per (replicate, branch) of one tree level.
//   pass 1   one coalesced pass over three rows: the shuffled parent row is copied to the child (msamanip.c:1645) while the Fitch
//            rows of parent and child are compared word by word; the (rare) differences are counted per (source, target) class,
//            n[a->d] (:1634-1643).
//   place    substitution number s of the branch (in the canonical order a-major, d) belongs to lane s % 32; the lane draws columns
//            uniformly until it hits one that holds its source class in the shuffled parent row and has not been substituted yet
//            on this branch (child byte still equal to the parent's), all lanes of a round in parallel, two lanes on the same
//            column resolved towards the lower lane.  Each acceptance picks uniformly among the still-free columns of the class,
//            so the placement is a uniformly random injection of the substitutions into the columns of their source class --
//            the distribution of the reference's Fisher-Yates shuffle of the position list (:1718-1757) -- and the counts are
//            reproduced exactly.  A lane still without a column after RPP_ROUNDS rounds (a source class that is rare in the row)
//            selects exactly: the warp counts the free columns of the class and takes the one of a uniformly drawn rank.
// Work per branch: 3 rows read + 1 written, a handful of instructions per word, plus ~L/m_a probes per substitution: a few
// times less issue work than the rank form above, and 256 bytes of shared memory per warp instead of 7 KB.
constexpr int RPP_WARPS  = 8;
constexpr int RPP_ROUNDS = 48;

template <int W>
__global__ void __launch_bounds__(RPP_WARPS * 32)
replay_level_pos_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin, int lvl_count,
                        int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0, const unsigned long long *__restrict__ ids,
                        int first_rep, int nrep, const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf, uint8_t *__restrict__ res)
{
  __shared__ int nsub_all[RPP_WARPS][64];                                       // [25] n[a->d], then [26] their exclusive prefix sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int *nsub = nsub_all[warp], *cum = nsub_all[warp] + 32;
  const long long task = (long long) blockIdx.x * RPP_WARPS + warp;
  if (task >= 2LL * lvl_count * nrep) return;                                   // whole warp
  const int side = (int) (task & 1);
  const long long nt = task >> 1;
  const int rr = (int) (nt / lvl_count);
  const int r = first_rep + rr;
  const uint32_t rid = (uint32_t) (ids ? ids[rr] : id0 + (unsigned long long) rr);
  const int v = order[lvl_begin + (int) (nt % lvl_count)];
  const uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  uint8_t *shanc = shancbuf + (size_t) r * (N - 1) * L;
  uint8_t *leaves = res + (size_t) r * N * L;
  const uint8_t *par_o = anc + (size_t) v * L;
  const uint8_t *par_s = shanc + (size_t) v * L;
  const int ch = side ? right[v] : left[v];
  const uint8_t *kid_o = (ch > 0) ? anc + (size_t) ch * L : msa + (size_t) (-ch) * L;
  uint8_t *kid_s = (ch > 0) ? shanc + (size_t) ch * L : leaves + (size_t) (-ch) * L;
  const int LW = (L + W - 1) / W;
  const unsigned lt = (1u << lane) - 1u;

  nsub[lane] = 0;
  __syncwarp();
  // ---- pass 1: copy + compare
  auto load = [&](const uint8_t *row, int u) -> uint32_t {
    if (W == 4) return __ldcg(reinterpret_cast<const uint32_t *>(row) + u);
    return (uint32_t) row[u] | 0xFFFFFF00u;
  };
  for (int u0 = 0; u0 < LW; u0 += 64) {                                         // two units per lane in flight
    uint32_t wp[2], wk[2], ws[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const int u = u0 + h * 32 + lane;
      const bool in = u < LW;
      wp[h] = in ? load(par_o, u) : 0xFFFFFFFFu;
      wk[h] = in ? load(kid_o, u) : 0xFFFFFFFFu;
      ws[h] = in ? load(par_s, u) : 0u;
    }
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const int u = u0 + h * 32 + lane;
      if (u < LW) {
        if (W == 4) reinterpret_cast<uint32_t *>(kid_s)[u] = ws[h];
        else kid_s[u] = (uint8_t) ws[h];
      }
      if (wp[h] != wk[h]) {
        #pragma unroll
        for (int q = 0; q < W; q++) {
          const int pa = (wp[h] >> (8 * q)) & 0xFF, kd = (wk[h] >> (8 * q)) & 0xFF;
          if (pa != kd && pa <= 4 && kd <= 4) atomicAdd(&nsub[pa * 5 + kd], 1);
        }
      }
    }
  }
  __syncwarp();
  // exclusive prefix sums of the 25 counts (lane k holds n_k)
  {
    const int nk = (lane < 25) ? nsub[lane] : 0;
    int incl = nk;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    cum[lane] = incl - nk;                                                      // cum[25..31] = total
  }
  __syncwarp();
  const int ktot = cum[25];
  if (ktot == 0) return;                                                        // uniform: most branches of a shallow tree

  // ---- place
  Philox ph; ph.key[0] = (uint32_t) seed ^ (0xC2B2AE35u * (rid + 1u)); ph.key[1] = (uint32_t) (seed >> 32) ^ 0x5bd1e995u;
  uint32_t sd[4]; ph.block((uint32_t) v, 0x7ee1u + (uint32_t) side, 0x9051u + (uint32_t) lane, 0u, sd);
  Pcg32 rng; rng.state = ((unsigned long long) sd[0] << 32) | sd[1]; rng.inc = ((((unsigned long long) sd[2] << 32) | sd[3]) << 1) | 1ULL;
  rng.next();
  volatile uint8_t *kid_v = kid_s;
  #pragma unroll 1
  for (int base = 0; base < ktot; base += 32) {
    const int sidx = base + lane;
    bool pending = sidx < ktot;
    int a = 0, d = 0;
    if (pending) {
      int k = 0;
      while (k < 24 && cum[k + 1] <= sidx) k++;                                 // the (a, d) this substitution belongs to
      a = k / 5; d = k % 5;
    }
    #pragma unroll 1
    for (int round = 0; round < RPP_ROUNDS && __any_sync(0xffffffffu, pending); round++) {
      unsigned key = 0x80000000u | (unsigned) lane;                            // lanes without a valid candidate never match anyone
      int c = 0;
      if (pending) {
        c = (int) __umulhi(rng.next(), (uint32_t) L);
        const int xp = par_s[c], xk = kid_v[c];
        if (xp == a && xk == xp) key = (unsigned) c;
      }
      const unsigned same = __match_any_sync(0xffffffffu, key);
      if (pending && !(key & 0x80000000u) && (same & lt) == 0u) { kid_v[c] = (uint8_t) d; pending = false; }
      __syncwarp();
    }
    // exact selection for whoever is still waiting (rare source class): one lane at a time, the whole warp scans the row
    unsigned waiting = __ballot_sync(0xffffffffu, pending);
    while (waiting) {
      const int src = __ffs(waiting) - 1;
      waiting &= waiting - 1;
      const int      aa = __shfl_sync(0xffffffffu, a, src), dd = __shfl_sync(0xffffffffu, d, src);
      const uint32_t rn = __shfl_sync(0xffffffffu, rng.next(), src);
      int nfree = 0;
      for (int c0 = 0; c0 < L; c0 += 32) {
        const int c = c0 + lane;
        const bool fr = c < L && par_s[c] == aa && kid_v[c] == aa;
        nfree += __popc(__ballot_sync(0xffffffffu, fr));
      }
      if (nfree == 0) continue;                                                 // cannot happen: the parent row holds >= k_a columns of class a
      int target = (int) __umulhi(rn, (uint32_t) nfree);
      for (int c0 = 0; c0 < L; c0 += 32) {
        const int c = c0 + lane;
        const bool fr = c < L && par_s[c] == aa && kid_v[c] == aa;
        const unsigned bal = __ballot_sync(0xffffffffu, fr);
        const int n = __popc(bal);
        if (target < n) {
          if (fr && __popc(bal & lt) == target) kid_v[c] = (uint8_t) dd;
          break;
        }
        target -= n;
      }
      __syncwarp();
    }
  }
}

} // namespace

cudaError_t rsb_launch_null_simulate(const int *left, const int *right, const int *order, const int *level_start_host, int nlevels,
                                     const unsigned long long *pthr, int N, int L, const uint8_t *root,
                                     const uint8_t *gapmask, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *scratch, cudaStream_t st)
{
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    if (L % 4 == 0) null_simulate_level_kernel<4><<<dim3((L / 4 + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, pthr, N, L, root, gapmask, seed, id0, first_rep, res, scratch);
    else            null_simulate_level_kernel<1><<<dim3((L + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, pthr, N, L, root, gapmask, seed, id0, first_rep, res, scratch);
  }
  return cudaGetLastError();
}

// the column permutations of replicates [first_rep, first_rep + nrep) with ids id0 ... (once per generator call)
cudaError_t rsb_launch_permutations(int L, unsigned long long seed, unsigned long long id0, const unsigned long long *ids, int first_rep, int nrep,
                                    int *perm, cudaStream_t st)
{
  permutation_kernel<<<nrep, 256, 0, st>>>(L, seed, id0, ids, first_rep, perm);
  return cudaGetLastError();
}

// does the alignment hold residues other than A C G U and the gap?  Synchronises st.
cudaError_t rsb_launch_unknown_check(const uint8_t *msa, size_t n, int *d_flag, int *unknown, cudaStream_t st)
{
  cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  unknown_flag_kernel<<<296, 256, 0, st>>>(msa, n, d_flag);
  cudaMemcpyAsync(unknown, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  return cudaStreamSynchronize(st);
}

cudaError_t rsb_launch_fitch_shuffle(const int *left, const int *right, const int *parent, const int *order, const int *level_start_host,
                                     int nlevels, int N, int L, const uint8_t *msa, unsigned long long seed, unsigned long long id0,
                                     const unsigned long long *ids, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *anc, uint8_t *shanc, int *perm, uint8_t *sets_shared, int build_sets, cudaStream_t st)
{
  (void) parent;
  if (L > 65535) return cudaErrorInvalidValue;            // per-class counters are 16 bit
  // The Fitch sets do not depend on the replicate unless the alignment holds unknown residues (rsb_launch_unknown_check):
  // then one up pass into `sets_shared` serves every replicate (build_sets: run it now; otherwise it is already there)
  // and only the traceback is per replicate.  sets_shared == NULL: sets per replicate, in the residue buffer itself.
  const bool shared = sets_shared != nullptr;
  if (!shared || build_sets)
  for (int lv = nlevels - 1; lv >= 0; lv--) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    uint8_t *dst = shared ? sets_shared : anc;
    const int z = shared ? 1 : nrep, fr = shared ? 0 : first_rep;
    if (L % 4 == 0) fitch_up_level_kernel<4><<<dim3((L / 4 + 127) / 128, cnt, z), 128, 0, st>>>(left, right, order, b, N, L, msa, seed, id0, ids, fr, dst);
    else            fitch_up_level_kernel<1><<<dim3((L + 127) / 128, cnt, z), 128, 0, st>>>(left, right, order, b, N, L, msa, seed, id0, ids, fr, dst);
  }
  const uint8_t *sets = shared ? sets_shared : anc + (size_t) first_rep * (N - 1) * L;
  const size_t sets_stride = shared ? 0 : (size_t) (N - 1) * L;
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    if (L % 4 == 0) fitch_down_level_kernel<4><<<dim3((L / 4 + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, N, L, seed, id0, ids, first_rep, sets, sets_stride, anc);
    else            fitch_down_level_kernel<1><<<dim3((L + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, N, L, seed, id0, ids, first_rep, sets, sets_stride, anc);
  }
  permute_root_kernel<<<dim3((L + 255) / 256, nrep), 256, 0, st>>>(N, L, first_rep, anc, shanc, perm);
  // RSCAPE_B200_REPLAY=position selects the position-form kernel (same distribution, tested side by side; faster when a branch
  // carries few substitutions, slower on the bench family, whose leaf branches carry ~200 each: 52 vs 27 ms per 100 SSU replicates)
  const char *form = getenv("RSCAPE_B200_REPLAY");
  const bool rank_form = !(form && form[0] == 'p');
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    const long long tasks = 2LL * cnt * nrep;                    // branches of this level over all replicates
    if (!rank_form) {
      const unsigned grid = (unsigned) ((tasks + RPP_WARPS - 1) / RPP_WARPS);
      if (L % 4 == 0) replay_level_pos_kernel<4><<<grid, RPP_WARPS * 32, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res);
      else            replay_level_pos_kernel<1><<<grid, RPP_WARPS * 32, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res);
      continue;
    }
    const int code_words = (L + 7) / 8 + 1;                       // one nibble per column (the classes' tables packed back to back)
    const size_t smem = (size_t) RPR_WARPS * (code_words + 32) * sizeof(unsigned);
    int ws_log2 = 0; while ((32 << ws_log2) < L / 4) ws_log2++;   // words per lane segment (power of two)
    const bool seg = (L % 4 == 0) && L <= 4092;                   // segment form of the second sweep: 12-bit class ranks
    const size_t smem_seg = (size_t) RPR_WARPS * (code_words + 32 + 32 * ((1 << ws_log2) + 1)) * sizeof(unsigned);
    if (seg && smem_seg <= 200 * 1024) {
      const unsigned grid = (unsigned) ((tasks + RPR_WARPS - 1) / RPR_WARPS);
      if (smem_seg > 48 * 1024) cudaFuncSetAttribute(replay_level_row_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_seg);
      replay_level_row_kernel<4, true><<<grid, RPR_WARPS * 32, smem_seg, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res, code_words, ws_log2);
      continue;
    }
    if (smem <= 200 * 1024) {                                    // ballot form: any L
      const unsigned grid = (unsigned) ((tasks + RPR_WARPS - 1) / RPR_WARPS);
      if (L % 4 == 0) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(replay_level_row_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        replay_level_row_kernel<4, false><<<grid, RPR_WARPS * 32, smem, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res, code_words, 0);
      } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(replay_level_row_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        replay_level_row_kernel<1, false><<<grid, RPR_WARPS * 32, smem, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res, code_words, 0);
      }
      continue;
    }
    // alignments too long for the code table in shared memory (L > ~80 000): one thread per branch, selection sampling
    const unsigned grid = (unsigned) ((tasks + RP_THREADS - 1) / RP_THREADS);       // many branches: throughput variant, one thread per branch
    if (L % 4 == 0) replay_level_kernel<true><<<grid, RP_THREADS, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res);
    else            replay_level_kernel<false><<<grid, RP_THREADS, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, ids, first_rep, nrep, anc, shanc, res);
  }
  return cudaGetLastError();
}
