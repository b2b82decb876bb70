// peer_reduce.h -- view of the ranks' exchange blocks handed to the one-shot all-reduce kernel (peer_reduce.cu)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define RSB_PEER_MAX     8        // GPUs of one NVSwitch box
#define RSB_PEER_CTAS    8        // chunks (CTAs) a vector is split into at most; one flag per (half, source rank, chunk)
#define RSB_PEER_THREADS 256
#define RSB_PEER_CHANNELS 4       // independent sequences of all-reduces (one per statistics stream of the pipelined null loop)

struct RsbPeerView {
  double             *x[RSB_PEER_MAX];        // x[q]: exchange block of rank q as mapped into THIS process, [channel][2][W][cap] doubles
  unsigned long long *flag[RSB_PEER_MAX];     // flag[q]: its sequence flags, [channel][2][W][RSB_PEER_CTAS]
  size_t cap;                                 // doubles per (half, source rank)
  int W, rank;
};

// bytes of one rank's exchange block: the doubles, then the flags
static inline size_t rsb_peer_doubles(int W, size_t cap) { return (size_t) RSB_PEER_CHANNELS * 2 * W * cap; }
static inline size_t rsb_peer_block_bytes(int W, size_t cap) { return rsb_peer_doubles(W, cap) * sizeof(double) + (size_t) RSB_PEER_CHANNELS * 2 * W * RSB_PEER_CTAS * sizeof(unsigned long long); }

cudaError_t rsb_launch_peer_allreduce(double *buf, size_t count, int op_max, const RsbPeerView &pv, int channel, unsigned long long seq, cudaStream_t st);
