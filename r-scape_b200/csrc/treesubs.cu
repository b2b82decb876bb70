// treesubs.cu -- Tree_Substitutions, src/msatree.c:1455-1540: substitution counts over the branches of the tree, the
// input of R-scape's power calculation (src/R-scape.c:2782-2868).
//
// The reference walks, for every pair of columns, all 2(N-1) branches and counts those on which both columns change
// (ndouble) or at least one does (njoin): O(L^2 N), the same contraction shape as the pair counts of the scan.  Here each
// branch becomes one row of a "branch alignment" whose residues say what the branch did in a column,
//     0 = substitution (child != parent),  1 = no substitution,  4 (gap) = the branch does not count in this column
// (without includegaps: parent or child not canonical, :1469-1474), and the unweighted count table of a column pair over
// those rows -- the very tcgen05 contraction of the scan with unit weights -- holds the answers:
//     ndouble[i][j] = C[0,0]                    both columns substituted on the branch (:1496-1500)
//     njoin[i][j]   = C[0,0] + C[0,1] + C[1,0]  counted in both columns and at least one substituted (:1522-1526)
// Counts are integers, so the result is exact.  The two small kernels here build the rows (counting the single-column
// substitutions, nsubs, :1462-1476, on the way) and pick the two tables out of the count planes.
#include "rsb_common.cuh"

namespace {

// rows[e][c] for branch e = 2 v + side (side 0 = left child of internal node v, 1 = right child), and nsubs[c] = number of
// rows with code 0 in column c (:1462-1476).  leaves [N][L], internal [N-1][L] (row v = ancestral sequence of node v from the
// Fitch pass).  A thread owns 4 consecutive columns (one 32-bit word when VEC) of `rpb` consecutive branches, keeps the four
// substitution counts in registers and adds them to nsubs once.
constexpr int BR_THREADS = 128;

__device__ __forceinline__ unsigned branch_code(unsigned p, unsigned x, int includegaps)
{
  if (includegaps) return (x != p) ? 0u : 1u;
  return (p < RSB_K && x < RSB_K) ? ((x != p) ? 0u : 1u) : 4u;
}

template <bool VEC>
__global__ void __launch_bounds__(BR_THREADS)
branch_rows_kernel(const uint8_t *__restrict__ leaves, const uint8_t *__restrict__ internal, const int *__restrict__ left,
                   const int *__restrict__ right, int L, int nrows, int rpb, int includegaps, uint8_t *__restrict__ rows,
                   int *__restrict__ nsubs)
{
  const int c0 = (blockIdx.x * BR_THREADS + threadIdx.x) * 4;
  if (c0 >= L) return;
  const int e0 = blockIdx.y * rpb, e1 = (e0 + rpb < nrows) ? e0 + rpb : nrows;
  int cnt[4] = { 0, 0, 0, 0 };
  for (int e = e0; e < e1; e++) {
    const int v = e >> 1, kid = (e & 1) ? right[v] : left[v];                      // the same for the whole block
    const uint8_t *pp = internal + (size_t) v * L + c0;
    const uint8_t *xp = ((kid > 0) ? internal + (size_t) kid * L : leaves + (size_t) (-kid) * L) + c0;   // Easel: child <= 0 is leaf -child
    uint8_t *out = rows + (size_t) e * L + c0;
    if (VEC) {                                                                     // L % 4 == 0: rows are word-aligned
      const unsigned p = *reinterpret_cast<const unsigned *>(pp), x = *reinterpret_cast<const unsigned *>(xp);
      unsigned w = 0;
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned code = branch_code((p >> (8 * k)) & 0xffu, (x >> (8 * k)) & 0xffu, includegaps);
        w |= code << (8 * k);
        cnt[k] += (code == 0u);
      }
      *reinterpret_cast<unsigned *>(out) = w;
    } else {
      #pragma unroll
      for (int k = 0; k < 4; k++)
        if (c0 + k < L) {
          const unsigned code = branch_code(pp[k], xp[k], includegaps);
          out[k] = (uint8_t) code;
          cnt[k] += (code == 0u);
        }
    }
  }
  #pragma unroll
  for (int k = 0; k < 4; k++)
    if (cnt[k] && c0 + k < L) atomicAdd(&nsubs[c0 + k], cnt[k]);
}

// the two pair tables out of the count planes (plane a*4+b, upper triangle): int [L][L], entries i<j, the rest 0 (:1484-1485)
__global__ void __launch_bounds__(128)
subs_tables_kernel(const long long *__restrict__ cnt, int L, int Lp, int *__restrict__ ndouble, int *__restrict__ njoin)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L) return;
  const size_t plane = (size_t) L * Lp, at = (size_t) i * Lp + j, out = (size_t) i * L + j;
  int d = 0, jn = 0;
  if (i < j) {
    const long long c00 = cnt[at], c01 = cnt[plane + at], c10 = cnt[4 * plane + at];
    d = (int) c00; jn = (int) (c00 + c01 + c10);
  }
  if (ndouble) ndouble[out] = d;
  if (njoin)   njoin[out]   = jn;
}

} // namespace

// rows + nsubs (int [L], zeroed here) in one pass over the reconstruction
cudaError_t rsb_launch_branch_rows(const uint8_t *leaves, const uint8_t *internal, const int *left, const int *right, int ntaxa, int L,
                                   int includegaps, uint8_t *rows, int *nsubs, cudaStream_t st)
{
  const int nrows = 2 * (ntaxa - 1);
  cudaError_t e = cudaMemsetAsync(nsubs, 0, sizeof(int) * (size_t) L, st);
  if (e != cudaSuccess) return e;
  // rows per block: 64 when that still gives the 148 SMs a few blocks each, fewer (down to 8) for short alignments
  const int nbx = (L + 4 * BR_THREADS - 1) / (4 * BR_THREADS);
  int rpb = (int) (((long long) nrows * nbx) / (148 * 4));
  rpb = rpb > 64 ? 64 : (rpb < 8 ? 8 : rpb);
  const dim3 grid(nbx, (nrows + rpb - 1) / rpb);
  if (L % 4 == 0) branch_rows_kernel<true><<<grid, BR_THREADS, 0, st>>>(leaves, internal, left, right, L, nrows, rpb, includegaps, rows, nsubs);
  else            branch_rows_kernel<false><<<grid, BR_THREADS, 0, st>>>(leaves, internal, left, right, L, nrows, rpb, includegaps, rows, nsubs);
  return cudaGetLastError();
}

cudaError_t rsb_launch_subs_tables(const long long *cnt, int L, int Lp, int *ndouble, int *njoin, cudaStream_t st)
{
  subs_tables_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(cnt, L, Lp, ndouble, njoin);
  return cudaGetLastError();
}
