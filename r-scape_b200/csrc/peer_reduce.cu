// peer_reduce.cu -- one-shot all-reduce of the small per-scan vectors over NVLink peer memory.
//
// With the L x L pair grid of a scan dealt to the GPUs of the box, three vectors per scan are summed over the ranks: the marginal
// partial sums [L][4], the APC row sums [L+4] and the score range (src/correlators.c:1338-1375, 1064-1157 are the loops whose
// totals they are).  They are a few KB to ~100 KB: latency-bound, and they sit inside the statistics chain of every replicate,
// which runs beside the persistent tcgen05 contraction (gram_tcgen05.cu; ~10 KB of shared memory left per SM).  The chains of
// consecutive replicates overlap on their own streams (capi.cu), so the reductions need (a) no shared memory, (b) an order that
// is the same on every rank PER STREAM rather than globally, (c) results that are bit-identical on all ranks.  NCCL gives (c) only
// by luck of its algorithm choice and wants all collectives of a communicator in one global order, i.e. all chains on one stream
// (measured on 2 GPUs at the SSU shape: 51.8 ms per step against 43.7 with this kernel and one channel per stream; with a single
// stream the two are equal).  The kernel:
//
//   every rank owns an exchange block  x[channel][2][W][cap] doubles + flag[channel][2][W][PEER_CTAS]  mapped into all ranks
//   (cudaIpc handles between processes, direct peer access inside one process).  All-reduce number `seq` of a channel uses
//   half seq & 1 of that channel:
//     push   each CTA copies its chunk of the vector into x[half][my rank] of EVERY rank (stores over NVLink), fences, and
//            sets flag[half][my rank][cta] = seq on every rank;
//     wait   it spins until its own flag[half][q][cta] >= seq for every rank q;
//     sum    it adds the W copies in rank order 0..W-1 -- the same order on every rank, so all ranks hold bit-identical
//            results and the result does not depend on arrival order (no floating-point atomics).
//   Two halves suffice because the all-reduces of one channel are serialised on every rank (an event chain in capi.cu): a rank
//   can start number seq + 2 only after it has finished seq + 1, for which every peer must have pushed seq + 1, i.e. finished
//   reading seq.  Channels are independent sequences (own halves, flags and event chain): the statistics chains of consecutive
//   replicates run on different streams and overlap, each with its own channel; every rank issues the same calls in the same
//   order per channel.
#include "rsb_common.cuh"
#include "peer_reduce.h"

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int OP>     // 0 sum, 1 max
__global__ void __launch_bounds__(RSB_PEER_THREADS)
peer_allreduce_kernel(double *__restrict__ buf, int count, int chunk, RsbPeerView pv, int channel, unsigned long long seq)
{
  const int W = pv.W, me = pv.rank, half = channel * 2 + (int) (seq & 1ULL), cta = blockIdx.x;
  const int i0 = cta * chunk, i1 = min(count, i0 + chunk);
  // push my chunk to every rank (my own block included)
  for (int q = 0; q < W; q++) {
    double *dst = pv.x[(me + q) % W] + ((size_t) half * W + me) * pv.cap;          // start with myself, then round the ring: spreads the links
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < W) st_release_sys(pv.flag[threadIdx.x] + ((size_t) half * W + me) * RSB_PEER_CTAS + cta, seq);
  // wait for every rank's chunk
  if (threadIdx.x < W) {
    const unsigned long long *f = pv.flag[me] + ((size_t) half * W + threadIdx.x) * RSB_PEER_CTAS + cta;
    while (ld_acquire_sys(f) < seq) __nanosleep(40);
  }
  __syncthreads();
  // combine in rank order (L1 is not coherent with peer writes: read through L2)
  const double *x0 = pv.x[me] + (size_t) half * W * pv.cap;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    double acc = __ldcg(x0 + i);
    for (int q = 1; q < W; q++) {
      const double v = __ldcg(x0 + (size_t) q * pv.cap + i);
      acc = OP ? fmax(acc, v) : acc + v;
    }
    buf[i] = acc;
  }
}

} // namespace

cudaError_t rsb_launch_peer_allreduce(double *buf, size_t count, int op_max, const RsbPeerView &pv, int channel, unsigned long long seq, cudaStream_t st)
{
  int ctas = (int) ((count + 2047) / 2048);
  ctas = ctas < 1 ? 1 : (ctas > RSB_PEER_CTAS ? RSB_PEER_CTAS : ctas);
  const int chunk = (int) ((count + ctas - 1) / ctas);
  if (op_max) peer_allreduce_kernel<1><<<ctas, RSB_PEER_THREADS, 0, st>>>(buf, (int) count, chunk, pv, channel, seq);
  else        peer_allreduce_kernel<0><<<ctas, RSB_PEER_THREADS, 0, st>>>(buf, (int) count, chunk, pv, channel, seq);
  return cudaGetLastError();
}
