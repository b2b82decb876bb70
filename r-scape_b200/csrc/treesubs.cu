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
// Counts are integers, so the result is exact.  The three small kernels here build the rows, count the single-column
// substitutions (nsubs, :1462-1476) and pick the two tables out of the count planes.
#include "rsb_common.cuh"

namespace {

// rows[e][c] for branch e = 2 v + side (side 0 = left child of internal node v, 1 = right child).
// leaves [N][L], internal [N-1][L] (row v = ancestral sequence of node v from the Fitch pass).
__global__ void __launch_bounds__(256)
branch_rows_kernel(const uint8_t *__restrict__ leaves, const uint8_t *__restrict__ internal, const int *__restrict__ left,
                   const int *__restrict__ right, int L, int includegaps, uint8_t *__restrict__ rows)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
  if (c >= L) return;
  const int v = e >> 1, kid = (e & 1) ? right[v] : left[v];
  const uint8_t p = internal[(size_t) v * L + c];
  const uint8_t x = (kid > 0) ? internal[(size_t) kid * L + c] : leaves[(size_t) (-kid) * L + c];     // Easel: child <= 0 is leaf -child
  uint8_t code;
  if (includegaps) code = (x != p) ? 0 : 1;
  else             code = (p < RSB_K && x < RSB_K) ? ((x != p) ? 0 : 1) : 4;
  rows[(size_t) e * L + c] = code;
}

// nsubs[c] = number of rows with code 0 in column c; blockIdx.y strides over the rows
__global__ void __launch_bounds__(128)
nsubs_kernel(const uint8_t *__restrict__ rows, int nrows, int L, int *__restrict__ nsubs)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L) return;
  int n = 0;
  for (int e = blockIdx.y; e < nrows; e += gridDim.y) n += (rows[(size_t) e * L + c] == 0);
  if (n) atomicAdd(&nsubs[c], n);
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

cudaError_t rsb_launch_branch_rows(const uint8_t *leaves, const uint8_t *internal, const int *left, const int *right, int ntaxa, int L,
                                   int includegaps, uint8_t *rows, cudaStream_t st)
{
  branch_rows_kernel<<<dim3((L + 255) / 256, 2 * (ntaxa - 1)), 256, 0, st>>>(leaves, internal, left, right, L, includegaps, rows);
  return cudaGetLastError();
}

cudaError_t rsb_launch_nsubs(const uint8_t *rows, int nrows, int L, int *nsubs, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(nsubs, 0, sizeof(int) * (size_t) L, st);
  if (e != cudaSuccess) return e;
  // few columns, many rows: split the rows over enough blocks to fill the chip (148 SMs x 8 blocks of 128 threads)
  const int nbx = (L + 127) / 128;
  int ny = (148 * 8 + nbx - 1) / nbx;
  if (ny > nrows) ny = nrows;
  if (ny < 1) ny = 1;
  nsubs_kernel<<<dim3(nbx, ny), 128, 0, st>>>(rows, nrows, L, nsubs);
  return cudaGetLastError();
}

cudaError_t rsb_launch_subs_tables(const long long *cnt, int L, int Lp, int *ndouble, int *njoin, cudaStream_t st)
{
  subs_tables_kernel<<<dim3((L + 127) / 128, L), 128, 0, st>>>(cnt, L, Lp, ndouble, njoin);
  return cudaGetLastError();
}
