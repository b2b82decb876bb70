// msaprep.cu -- the O(N L) / O(N^2 L) preprocessing that defines the scanned alignment and its weights:
//
//   gap-column filter   msamanip_RemoveGapColumns, src/msamanip.c:486-500: keep a column iff the weighted residue fraction
//                       r / (r + gap) is >= 1 - gapthresh and r > 0 (missing data '~' and '*' count on neither side)
//   PB weights          esl_msaweight_PB (Henikoff position-based; R-scape uses it for nseq > 1000, src/R-scape.c:1555-1556):
//                       per column a sequence holding canonical residue x gets 1 / (r n_x), r = number of distinct canonical
//                       residues in the column, n_x = sequences holding x; summed over the columns, divided by the sequence's
//                       number of canonical residues, normalised to sum N
//   pairwise identity   esl_dst_XPairId: identical canonical positions / min(len_a, len_b), len = canonical residues
//                       (esl_dst_XAverageId, src/msamanip.c:1967; the distance matrix 1 - pid of esl_msaweight_GSC, nseq <= 1000)
//
// Easel is not part of the reference tree: these follow SURVEY 9.7's restatement (pinned only through the tutorial transcript:
// GSC weights -> scores, "avgid 65.82").  All of it is HBM/L2-bound byte work: integer counts in registers and shared memory,
// fixed-order reductions (deterministic), no tensor cores.
#include "rsb_common.cuh"

namespace {

__device__ __forceinline__ bool is_canonical(unsigned x) { return x < 4u; }
__device__ __forceinline__ bool is_residue(unsigned x)   { return x < 4u || (x > 4u && x < 16u); }     // esl_abc_XIsResidue: x < K or K < x < Kp-2
__device__ __forceinline__ bool is_gap(unsigned x)       { return x == 4u; }

constexpr int PC_COLS = 32, PC_ROWS = 8;          // block = 32 columns x 8 row lanes

// per column: r = sum of the weights of the sequences holding a residue, tot = r + those holding a gap; unit weights are
// counted as integers (exact); useme[col] = r > 0 && r / tot >= idthresh
__global__ void __launch_bounds__(PC_COLS * PC_ROWS)
gap_columns_kernel(const uint8_t *__restrict__ msa, int N, int L, long long row_stride, const double *__restrict__ wgt, double idthresh,
                   uint8_t *__restrict__ useme)
{
  __shared__ double sr[PC_ROWS][PC_COLS], st[PC_ROWS][PC_COLS];
  const int tx = threadIdx.x % PC_COLS, ty = threadIdx.x / PC_COLS;
  const int col = blockIdx.x * PC_COLS + tx;
  double r = 0.0, tot = 0.0;
  long long ri = 0, ti = 0;
  if (col < L)
    for (int s = ty; s < N; s += PC_ROWS) {
      const unsigned x = msa[(size_t) s * row_stride + col];
      if (wgt) { const double w = wgt[s]; if (is_residue(x)) { r += w; tot += w; } else if (is_gap(x)) tot += w; }
      else     { ri += is_residue(x); ti += is_residue(x) || is_gap(x); }
    }
  if (!wgt) { r = (double) ri; tot = (double) ti; }
  sr[ty][tx] = r; st[ty][tx] = tot;
  __syncthreads();
  if (ty == 0 && col < L) {
    for (int k = 1; k < PC_ROWS; k++) { r += sr[k][tx]; tot += st[k][tx]; }          // fixed order
    useme[col] = (r > 0.0 && r / tot >= idthresh) ? 1 : 0;
  }
}

// per column: n_x for the 4 canonical residues -> coef[col][x] = 1 / (r n_x) (0 where n_x = 0)
__global__ void __launch_bounds__(PC_COLS * PC_ROWS)
pb_column_kernel(const uint8_t *__restrict__ msa, int N, int L, long long row_stride, double *__restrict__ coef)
{
  __shared__ int sn[PC_ROWS][PC_COLS][4];
  const int tx = threadIdx.x % PC_COLS, ty = threadIdx.x / PC_COLS;
  const int col = blockIdx.x * PC_COLS + tx;
  int n[4] = { 0, 0, 0, 0 };
  if (col < L)
    for (int s = ty; s < N; s += PC_ROWS) {
      const unsigned x = msa[(size_t) s * row_stride + col];
      n[0] += (x == 0u); n[1] += (x == 1u); n[2] += (x == 2u); n[3] += (x == 3u);
    }
  #pragma unroll
  for (int a = 0; a < 4; a++) sn[ty][tx][a] = n[a];
  __syncthreads();
  if (ty == 0 && col < L) {
    for (int k = 1; k < PC_ROWS; k++)
      #pragma unroll
      for (int a = 0; a < 4; a++) n[a] += sn[k][tx][a];
    const int r = (n[0] > 0) + (n[1] > 0) + (n[2] > 0) + (n[3] > 0);
    #pragma unroll
    for (int a = 0; a < 4; a++) coef[(size_t) col * 4 + a] = (n[a] > 0) ? 1.0 / ((double) r * (double) n[a]) : 0.0;
  }
}

// one warp per sequence: w_s = (sum over its canonical columns of coef[col][x]) / (number of canonical residues)
__global__ void __launch_bounds__(256)
pb_sequence_kernel(const uint8_t *__restrict__ msa, int N, int L, long long row_stride, const double *__restrict__ coef, double *__restrict__ w)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= N) return;
  double acc = 0.0; int rlen = 0;
  for (int c = lane; c < L; c += 32) {
    const unsigned x = msa[(size_t) s * row_stride + c];
    if (is_canonical(x)) { acc += coef[(size_t) c * 4 + x]; rlen++; }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); rlen += __shfl_xor_sync(0xffffffffu, rlen, o); }
  if (lane == 0) w[s] = (rlen > 0) ? acc / (double) rlen : 0.0;
}

// w *= N / sum(w) (all 1 if the sum is 0); one block, fixed summation order
__global__ void __launch_bounds__(1024)
normalise_weights_kernel(double *__restrict__ w, int N)
{
  __shared__ double red[1024];
  double a = 0.0;
  for (int s = threadIdx.x; s < N; s += 1024) a += w[s];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const double sum = red[0];
  for (int s = threadIdx.x; s < N; s += 1024) w[s] = (sum > 0.0) ? w[s] * ((double) N / sum) : 1.0;
}

// one warp per pair (a, b): identical canonical positions, canonical lengths -> pid = same / min(len_a, len_b) (0 if that is 0).
// pairs == NULL: all pairs a < b in row-major order, results into the full symmetric matrix out[N][N] as DISTANCES 1 - pid
// (diagonal 0); else out[k] = pid of pair k.
__global__ void __launch_bounds__(256)
pair_identity_kernel(const uint8_t *__restrict__ msa, int N, int L, long long row_stride, const int *__restrict__ pairs, long long npairs,
                     double *__restrict__ out)
{
  const int lane = threadIdx.x & 31;
  const long long k = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= npairs) return;
  int a, b;
  if (pairs) { a = pairs[2 * k]; b = pairs[2 * k + 1]; }
  else {                                                              // k-th pair of the upper triangle, row-major
    const double Nd = (double) N;
    a = (int) floor(((2.0 * Nd - 1.0) - sqrt((2.0 * Nd - 1.0) * (2.0 * Nd - 1.0) - 8.0 * (double) k)) * 0.5);
    if (a < 0) a = 0;
    while ((long long) a * N - (long long) a * (a + 1) / 2 > k) a--;
    while ((long long) (a + 1) * N - (long long) (a + 1) * (a + 2) / 2 <= k) a++;
    b = (int) (k - ((long long) a * N - (long long) a * (a + 1) / 2)) + a + 1;
  }
  const uint8_t *ra = msa + (size_t) a * row_stride, *rb = msa + (size_t) b * row_stride;
  int same = 0, la = 0, lb = 0;
  for (int c = lane; c < L; c += 32) {
    const unsigned x = ra[c], y = rb[c];
    la += is_canonical(x); lb += is_canonical(y);
    same += (is_canonical(x) && x == y);
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    same += __shfl_xor_sync(0xffffffffu, same, o); la += __shfl_xor_sync(0xffffffffu, la, o); lb += __shfl_xor_sync(0xffffffffu, lb, o);
  }
  if (lane == 0) {
    const int len = la < lb ? la : lb;
    const double pid = (len > 0) ? (double) same / (double) len : 0.0;
    if (pairs) out[k] = pid;
    else { out[(size_t) a * N + b] = 1.0 - pid; out[(size_t) b * N + a] = 1.0 - pid; }
  }
}

__global__ void zero_diagonal_kernel(double *__restrict__ D, int N)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < N) D[(size_t) a * N + a] = 0.0;
}

// gather the kept columns: out[s][k] = msa[s][cols[k]]
__global__ void column_subset_kernel(const uint8_t *__restrict__ msa, int N, long long row_stride, const int *__restrict__ cols, int nkeep,
                                     uint8_t *__restrict__ out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (k < nkeep) out[(size_t) s * nkeep + k] = msa[(size_t) s * row_stride + cols[k]];
}

} // namespace

cudaError_t rsb_launch_gap_columns(const uint8_t *msa, int N, int L, long long row_stride, const double *wgt, double idthresh, uint8_t *useme, cudaStream_t st)
{
  gap_columns_kernel<<<(L + PC_COLS - 1) / PC_COLS, PC_COLS * PC_ROWS, 0, st>>>(msa, N, L, row_stride, wgt, idthresh, useme);
  return cudaGetLastError();
}

cudaError_t rsb_launch_pb_weights(const uint8_t *msa, int N, int L, long long row_stride, double *coef, double *w, cudaStream_t st)
{
  pb_column_kernel<<<(L + PC_COLS - 1) / PC_COLS, PC_COLS * PC_ROWS, 0, st>>>(msa, N, L, row_stride, coef);
  pb_sequence_kernel<<<(N + 7) / 8, 256, 0, st>>>(msa, N, L, row_stride, coef, w);
  normalise_weights_kernel<<<1, 1024, 0, st>>>(w, N);
  return cudaGetLastError();
}

cudaError_t rsb_launch_pair_identity(const uint8_t *msa, int N, int L, long long row_stride, const int *pairs, long long npairs, double *out, cudaStream_t st)
{
  if (npairs > 0) pair_identity_kernel<<<(unsigned) ((npairs + 7) / 8), 256, 0, st>>>(msa, N, L, row_stride, pairs, npairs, out);
  if (!pairs) zero_diagonal_kernel<<<(N + 255) / 256, 256, 0, st>>>(out, N);
  return cudaGetLastError();
}

cudaError_t rsb_launch_column_subset(const uint8_t *msa, int N, long long row_stride, const int *cols, int nkeep, uint8_t *out, cudaStream_t st)
{
  if (nkeep > 0) column_subset_kernel<<<dim3((nkeep + 255) / 256, N), 256, 0, st>>>(msa, N, row_stride, cols, nkeep, out);
  return cudaGetLastError();
}
