// nullgen.cu -- null-alignment generators on the device.
//
// (B) null_simulate_kernel: cov_GenerateAlignment's ungapped, structure-free path
//     (src/cov_simulate.c:289-324 tree walk, :585-631 emission, :724-773 inverse-CDF draw) with
//     P(t) = exp(tQ) per branch (src/ratematrix.c:185-233; matrices are built on the host in capi.cu,
//     4x4 per branch).  Every (replicate, column) is an independent chain down the tree; the tree is walked one
//     level per launch (nodes of a level are independent given their parents), one thread per (replicate, node, column).
//     Randomness: Philox4x32-10 keyed by (seed, replicate), counter (node, column): one 128-bit block
//     serves both children of a node.  The reference consumes one Mersenne-Twister stream in branch-major
//     order; the streams differ, the per-draw distribution is the same (validated distributionally).
//
// (A) Fitch + shuffle, R-scape's default null (src/R-scape.c:1653-1668):
//     fitch_up/down_level tree_fitch_column (src/msatree.c:1700-1831): sets as 5-bit masks; post-order = one launch per
//                         tree level from the deepest up, pre-order = from the root down; thread per (replicate, node, column).
//     permute_root_kernel msamanip_ShuffleColumns (src/msamanip.c:1164-1233).  Only the root's permuted row is
//                         ever read (every other row is overwritten by its parent's row at msamanip.c:1645).
//     replay_level_kernel shuffle_tree_substitutions + shuffle_tree_substitute_all (src/msamanip.c:1597-1780):
//                         one thread per (replicate, branch) of one tree level; counts the 5x5 substitutions of the
//                         branch on the Fitch rows, copies the shuffled parent row to the child and re-places
//                         the substitutions at positions drawn uniformly without replacement among the columns
//                         holding the source residue (one-pass selection sampling instead of the reference's
//                         Fisher-Yates over an index list: same distribution, exact counts).
#include "rsb_common.cuh"

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

__device__ __forceinline__ double u01(uint32_t x) { return (double) x * (1.0 / 4294967296.0); }   // as esl_random: x / 2^32

// ------------------------------------------------------------------------------------------------ generator B
// One launch per tree level (nodes of a level are independent given their parents); a thread owns one
// (replicate, node, column) and emits both children of the node.  grid = (columns/128, nodes of the level, replicates).
__global__ void null_simulate_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order,
                                           int lvl_begin, const double *__restrict__ pcdf, int N, int L, const uint8_t *__restrict__ root,
                                           const uint8_t *__restrict__ gapmask, unsigned long long seed, unsigned long long id0,
                                           int first_rep, uint8_t *__restrict__ res, uint8_t *__restrict__ scratch)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (id0 + blockIdx.z);            // global replicate id: the stream does not depend on where the replicate is stored
  if (c >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ (0x9E3779B9u * (rid + 1u)); rng.key[1] = (uint32_t) (seed >> 32) + rid;
  uint8_t *anc  = scratch + (size_t) r * (N - 1) * L;        // internal node states [N-1][L]
  uint8_t *leaf = res + (size_t) r * N * L;
  const int par = (v == 0 ? root[c] : anc[(size_t) v * L + c]) & 3;       // cov_add_root for the root
  uint32_t rnd[4];
  rng.block((uint32_t) v, (uint32_t) c, 0x5eedu, 0u, rnd);
  #pragma unroll
  for (int side = 0; side < 2; side++) {
    const int child = side ? right[v] : left[v];
    const double *cdf = pcdf + ((size_t) v * 2 + side) * 16 + par * 4;
    const double x = u01(rnd[side]);
    int k = 0;                                              // cov_addres: first k with cdf_k > x, else K-1
    while (k < 3 && !(cdf[k] > x)) k++;
    if (child > 0) anc[(size_t) child * L + c] = (uint8_t) k;
    else {
      uint8_t out = (uint8_t) k;
      if (gapmask) { const uint8_t g = gapmask[(size_t) (-child) * L + c]; if (g >= 4) out = g; }
      leaf[(size_t) (-child) * L + c] = out;
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

// Fitch sets, one launch per tree level from the deepest up (:1758-1777, :1869-1905): a thread owns one
// (replicate, node, column).  grid = (columns/128, nodes of the level, replicates).
__global__ void fitch_up_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin,
                                      int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0,
                                      int first_rep, uint8_t *__restrict__ ancbuf)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (id0 + blockIdx.z);
  if (c >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ 0xF17C4u; rng.key[1] = (uint32_t) (seed >> 32) ^ (0x85EBCA6Bu * (rid + 1u));
  uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  auto leaf_set = [&](int n) -> unsigned {
    const int x = msa[(size_t) n * L + c];
    if (x <= 4) return 1u << x;
    uint32_t rnd[4]; rng.block((uint32_t) n, (uint32_t) c, 0x1eafu, 0u, rnd);          // unknown -> one of the 5 at random (:1735)
    return 1u << (int) (((unsigned long long) rnd[0] * 5u) >> 32);
  };
  const int l = left[v], rr = right[v];
  const unsigned Sl = (l > 0) ? anc[(size_t) l * L + c] : leaf_set(-l);
  const unsigned Sr = (rr > 0) ? anc[(size_t) rr * L + c] : leaf_set(-rr);
  unsigned S = Sl & Sr;
  if (!S) S = (Sl | Sr) & 0xFu;                                                         // union of residues; a gap never joins
  anc[(size_t) v * L + c] = (uint8_t) S;
}

// Traceback, one launch per level from the root down (:1779-1815): the stored set of each internal child is replaced by
// the chosen residue (the parent's residue if it is in the child's set, else a uniform member).
__global__ void fitch_down_level_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin,
                                        int N, int L, unsigned long long seed, unsigned long long id0, int first_rep, uint8_t *__restrict__ ancbuf)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = order[lvl_begin + blockIdx.y];
  const int r = first_rep + blockIdx.z;
  const uint32_t rid = (uint32_t) (id0 + blockIdx.z);
  if (c >= L) return;
  Philox rng; rng.key[0] = (uint32_t) seed ^ 0xF17C4u; rng.key[1] = (uint32_t) (seed >> 32) ^ (0x85EBCA6Bu * (rid + 1u));
  uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  int ax;
  if (v == 0) {                                                                          // root: uniform member of its set (:1779)
    uint32_t rnd[4]; rng.block(0xffffffffu, (uint32_t) c, 0x600du, 0u, rnd);
    ax = pick_member(anc[c], rnd[0]);
    anc[c] = (uint8_t) ax;
  } else ax = anc[(size_t) v * L + c];
  uint32_t rnd[4]; rng.block((uint32_t) v, (uint32_t) c, 0xd0c0u, 0u, rnd);
  const int kids[2] = { left[v], right[v] };
  #pragma unroll
  for (int side = 0; side < 2; side++) {
    const int ch = kids[side];
    if (ch <= 0) continue;
    const unsigned S = anc[(size_t) ch * L + c];
    anc[(size_t) ch * L + c] = (uint8_t) (((S >> ax) & 1u) ? ax : pick_member(S, rnd[side]));
  }
}

// one random permutation per replicate (Fisher-Yates by one thread; L is a few thousand) and the permuted root row
__global__ void permute_root_kernel(int N, int L, unsigned long long seed, unsigned long long id0, int first_rep, const uint8_t *__restrict__ ancbuf,
                                    uint8_t *__restrict__ shancbuf, int *__restrict__ permbuf)
{
  const int r = first_rep + blockIdx.x;
  const uint32_t rid = (uint32_t) (id0 + blockIdx.x);
  int *perm = permbuf + (size_t) r * L;
  const uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  uint8_t *sh = shancbuf + (size_t) r * (N - 1) * L;
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
  __syncthreads();
  for (int c = threadIdx.x; c < L; c += blockDim.x) sh[c] = anc[perm[c]];
}

constexpr int RP_THREADS = 128;
constexpr int RP_BATCH   = 8;        // 32-bit words (4 alignment columns each) loaded together
constexpr long long RP_WARP_TASKS = 12000;   // levels with at most this many (replicate, branch) tasks use the warp-per-branch variant

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
                    int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                    const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf, uint8_t *__restrict__ res)
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
  const uint32_t rid = (uint32_t) (id0 + (unsigned long long) rr);
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

constexpr int RPW_WARPS = 4;
constexpr int RPW_MAXWORDS = 128;      // 32-column words per class bitmask (L <= 4096)

// Latency variant of the replay for tree levels with few branches (near the root): one WARP per (replicate, branch).
// The classes of the shuffled parent row are kept as bitmasks; every substitution picks the k-th remaining candidate of
// its source class (k uniform) by a warp-wide rank-select and clears that bit -- uniform sampling without replacement,
// the same distribution as the thread-per-branch kernel and as the reference's Fisher-Yates shuffle.  The work per branch
// is larger, but a level finishes in tens of microseconds instead of waiting for one thread to walk the whole row.
__global__ void __launch_bounds__(RPW_WARPS * 32)
replay_level_warp_kernel(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ order, int lvl_begin, int lvl_count,
                         int N, int L, const uint8_t *__restrict__ msa, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                         const uint8_t *__restrict__ ancbuf, uint8_t *__restrict__ shancbuf, uint8_t *__restrict__ res)
{
  __shared__ unsigned masks[RPW_WARPS][5][RPW_MAXWORDS];
  __shared__ int nsub[RPW_WARPS][25];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long task = (long long) blockIdx.x * RPW_WARPS + warp;
  if (task >= 2LL * lvl_count * nrep) return;
  const int side = (int) (task & 1);
  const long long nt = task >> 1;
  const int rr = (int) (nt / lvl_count);
  const int r = first_rep + rr;
  const uint32_t rid = (uint32_t) (id0 + (unsigned long long) rr);
  const int v = order[lvl_begin + (int) (nt % lvl_count)];
  const int nwords = (L + 31) >> 5;
  const uint8_t *anc = ancbuf + (size_t) r * (N - 1) * L;
  uint8_t *shanc = shancbuf + (size_t) r * (N - 1) * L;
  uint8_t *leaves = res + (size_t) r * N * L;
  const uint8_t *par_o = anc + (size_t) v * L;
  const uint8_t *par_s = shanc + (size_t) v * L;
  const int ch = side ? right[v] : left[v];
  const uint8_t *kid_o = (ch > 0) ? anc + (size_t) ch * L : msa + (size_t) (-ch) * L;
  uint8_t *kid_s = (ch > 0) ? shanc + (size_t) ch * L : leaves + (size_t) (-ch) * L;

  Philox ph; ph.key[0] = (uint32_t) seed ^ (0xC2B2AE35u * (rid + 1u)); ph.key[1] = (uint32_t) (seed >> 32) ^ 0x5bd1e995u;
  uint32_t sd[4]; ph.block((uint32_t) v, 0x7ee1u + (uint32_t) side, 0u, 0u, sd);          // same stream as the thread variant's seed
  Pcg32 rng; rng.state = ((unsigned long long) sd[0] << 32) | sd[1]; rng.inc = ((((unsigned long long) sd[2] << 32) | sd[3]) << 1) | 1ULL;
  rng.next();

  if (lane < 25) nsub[warp][lane] = 0;
  __syncwarp();
  // one sweep: class bitmasks of the shuffled parent row, copy of that row to the child (:1645), substitution counts (:1634-1643)
  for (int c0 = 0; c0 < L; c0 += 32) {
    const int c = c0 + lane;
    const bool in = c < L;
    const int xs = in ? par_s[c] : 255;
    #pragma unroll
    for (int a = 0; a < 5; a++) { const unsigned b = __ballot_sync(0xffffffffu, xs == a); if (lane == 0) masks[warp][a][c0 >> 5] = b; }
    if (in) {
      kid_s[c] = (uint8_t) xs;
      const int pa = par_o[c], kd = kid_o[c];
      if (pa != kd && pa <= 4 && kd <= 4) atomicAdd(&nsub[warp][pa * 5 + kd], 1);
    }
  }
  __syncwarp();
  for (int a = 0; a < 5; a++) {
    // working copy of the class-a mask in registers: lane owns words lane, lane+32, ...
    unsigned wv[RPW_MAXWORDS / 32]; int pc = 0;
    #pragma unroll
    for (int q = 0; q < RPW_MAXWORDS / 32; q++) { const int wd = lane + 32 * q; wv[q] = (wd < nwords) ? masks[warp][a][wd] : 0u; pc += __popc(wv[q]); }
    int remaining = pc;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) remaining += __shfl_xor_sync(0xffffffffu, remaining, o);
    for (int d = 0; d < 5; d++) {
      int s = nsub[warp][a * 5 + d];
      while (s > 0 && remaining > 0) {
        const int k = (int) __umulhi(rng.next(), (uint32_t) remaining);                  // k-th remaining candidate, lane-major order
        int incl = pc;                                                                     // inclusive scan of the per-lane counts
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int owner = __ffs(__ballot_sync(0xffffffffu, incl > k)) - 1;
        const int before = __shfl_sync(0xffffffffu, incl - pc, owner);
        if (lane == owner) {
          int kk = k - before;
          #pragma unroll
          for (int q = 0; q < RPW_MAXWORDS / 32; q++) {
            const int n = __popc(wv[q]);
            if (kk >= 0 && kk < n) {
              const int bit = __fns(wv[q], 0, kk + 1);
              wv[q] &= ~(1u << bit);
              kid_s[(lane + 32 * q) * 32 + bit] = (uint8_t) d;
              kk = -1;
            } else if (kk >= 0) kk -= n;
          }
          pc--;
        }
        remaining--; s--;
      }
    }
  }
}

} // namespace

cudaError_t rsb_launch_null_simulate(const int *left, const int *right, const int *order, const int *level_start_host, int nlevels,
                                     const double *pcdf, int N, int L, const uint8_t *root,
                                     const uint8_t *gapmask, long long gap_stride, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *scratch, cudaStream_t st)
{
  (void) gap_stride;
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    null_simulate_level_kernel<<<dim3((L + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, pcdf, N, L, root, gapmask, seed, id0,
                                                                                 first_rep, res, scratch);
  }
  return cudaGetLastError();
}

cudaError_t rsb_launch_fitch_shuffle(const int *left, const int *right, const int *parent, const int *order, const int *level_start_host,
                                     int nlevels, int N, int L, const uint8_t *msa, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *anc, uint8_t *shanc, int *perm, cudaStream_t st)
{
  (void) parent;
  if (L > 65535) return cudaErrorInvalidValue;            // per-class counters are 16 bit
  for (int lv = nlevels - 1; lv >= 0; lv--) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    fitch_up_level_kernel<<<dim3((L + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, N, L, msa, seed, id0, first_rep, anc);
  }
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    fitch_down_level_kernel<<<dim3((L + 127) / 128, cnt, nrep), 128, 0, st>>>(left, right, order, b, N, L, seed, id0, first_rep, anc);
  }
  permute_root_kernel<<<nrep, 256, 0, st>>>(N, L, seed, id0, first_rep, anc, shanc, perm);
  for (int lv = 0; lv < nlevels; lv++) {
    const int b = level_start_host[lv], cnt = level_start_host[lv + 1] - b;
    const long long tasks = 2LL * cnt * nrep;                    // branches of this level over all replicates
    if (tasks <= RP_WARP_TASKS && L <= RPW_MAXWORDS * 32) {        // few branches: latency variant, one warp per branch
      const unsigned grid = (unsigned) ((tasks + RPW_WARPS - 1) / RPW_WARPS);
      replay_level_warp_kernel<<<grid, RPW_WARPS * 32, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, first_rep, nrep, anc, shanc, res);
      continue;
    }
    const unsigned grid = (unsigned) ((tasks + RP_THREADS - 1) / RP_THREADS);       // many branches: throughput variant, one thread per branch
    if (L % 4 == 0) replay_level_kernel<true><<<grid, RP_THREADS, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, first_rep, nrep, anc, shanc, res);
    else            replay_level_kernel<false><<<grid, RP_THREADS, 0, st>>>(left, right, order, b, cnt, N, L, msa, seed, id0, first_rep, nrep, anc, shanc, res);
  }
  return cudaGetLastError();
}
