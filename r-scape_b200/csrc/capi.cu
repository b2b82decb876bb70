// capi.cu -- the C-ABI of include/rscape_b200.h: context, buffers in HBM, TMA descriptors, kernel sequencing.
//
// One context = one CUDA device + one stream.  All work is enqueued asynchronously on that stream;
// the only host synchronisations are the D2H reads the caller asked for.  There is no CPU path:
// rsb_create fails unless a compute-capability-10.x device is present.
#include "rsb_common.cuh"
#include "peer_reduce.h"
#include "rsb_evalue.cuh"
#include "nccl_dyn.h"
#include "../../include/rscape_b200.h"
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>

// ---- kernel launchers defined in the other translation units
cudaError_t rsb_launch_gram_i8(int S, const CUtensorMap &tmA, const CUtensorMap &tmB, const int2 *tiles, int ntiles,
                               int rep0, int nrep, int L, int Lp, int kstages, long long *cnt, double scale, double *mrow,
                               double *mcol, int nJB, int nIB, int small52, int grid, const void *logtab, cudaStream_t st);
cudaError_t rsb_launch_gt_finish(const double *rec, size_t slot_stride, const double *pm, int nrep, int L, int Lp, double *cov,
                                 double *rowpart, double *colpart, double *mm, int sr, int sw, cudaStream_t st);
cudaError_t rsb_launch_export_nseff_rec(const double *rec, int L, int Lp, double wtot_scaled, double *nseff, double *ngap, cudaStream_t st);
cudaError_t rsb_launch_gram_i8_pair(int S, const CUtensorMap &tmA, const CUtensorMap &tmBh, const int2 *tiles2, int ntiles2,
                                    int rep0, int nrep, int L, int Lp, int kstages, long long *cnt, double scale, double *mrow,
                                    double *mcol, int nJB, int nIB, int small52, int max_clusters, cudaStream_t st);
int rsb_gram_pair_clusters(int S);
cudaError_t rsb_launch_quantise(const double *w, int N, int q, int umax, double vmax, uint8_t *mul, long long *V, double *err, cudaStream_t st);
cudaError_t rsb_launch_pack(int S, const uint8_t *res, int nrep, int N, int L, long long rep_stride_res, const uint8_t *wdig,
                            int Kpad, uint8_t *planeA, int MA, uint8_t *planeB, int NBrows, int Lcover, cudaStream_t st);
cudaError_t rsb_launch_colsum(const uint8_t *res, int N, int L, const unsigned long long *wq, unsigned long long *colsum, cudaStream_t st);
cudaError_t rsb_launch_counts_direct(const uint8_t *res, int N, int L, int Lp, const unsigned long long *wq, long long *cnt, cudaStream_t st);
void        rsb_stat_grid(int L, int *nJT, int *nIT);
cudaError_t rsb_launch_marginals(const double *mrow, const double *mcol, int nrep, int L, int CJ, int nJB, int nIB, double tol,
                                 double *msum, double *pm, int *flags, int sr, int sw, int phase, int mrow_blocks, cudaStream_t st);
int rsb_gram_mrow_blocks(int pair);
cudaError_t rsb_launch_nseff(const long long *cnt, int nrep, int L, int Lp, double scale, double *nseff, cudaStream_t st);
cudaError_t rsb_launch_logtab(void *tab, cudaStream_t st);
size_t rsb_logtab_bytes();
cudaError_t rsb_launch_statistic(int stat, int cls, const long long *cnt, const double *pm, const void *logtab, int nrep, int L, int Lp, double scale,
                                 long long wtot, unsigned mask, double *cov, double *rowpart, double *colpart, double *mm, int sr, int sw, cudaStream_t st);
cudaError_t rsb_launch_multi_statistic(int cls, const long long *cnt, const double *pm, const void *logtab, int nrep, int L, int Lp, double scale,
                                       long long wtot, unsigned mask, double *const *cov6, size_t rep_stride, int sr, int sw, cudaStream_t st);
cudaError_t rsb_launch_correct_hist_multi(const double *cov, const double *covx, const double *scal, int nrep, int L, int Lp, int ncombo, int nw,
                                          const int *kidx, const int *act, double bmin, const double *wptr, unsigned long long *hist, int nbins,
                                          double *mm, double *minmax_out, int *flags, const int *m2p, int mind, cudaStream_t st);
cudaError_t rsb_launch_raf(const long long *cnt, int nrep, int L, int Lp, int nseq, unsigned mask, int smooth, double *tmp, double *cov,
                           double *rowpart, double *colpart, double *mm, cudaStream_t st);
cudaError_t rsb_launch_ccf(const double *nseff, const double *pm, int nrep, int L, int Lp, double *part, double *meanp, double *cov,
                           double *rowpart, double *colpart, double *mm, cudaStream_t st);
cudaError_t rsb_launch_reduce_cov(const double *cov, int nrep, int L, int Lp, double *rowpart, double *colpart, double *mm, cudaStream_t st);
cudaError_t rsb_launch_export_probs(const long long *cnt, int L, int Lp, double scale, long long wtot, double *pp, double *nseff,
                                    double *ngap, cudaStream_t st);
cudaError_t rsb_launch_ps(const unsigned long long *colsum, int L, double scale, double *ps, cudaStream_t st);
cudaError_t rsb_launch_correct_final(const double *rowpart, const double *colpart, const double *mm, int nrep, int L,
                                     double *covx, double *scal, double *blocksum, double *covsum, int phase, cudaStream_t st);
cudaError_t rsb_launch_correct_hist(double *cov, const double *covx, const double *scal, int nrep, int L, int Lp, int actype, int mode,
                                    double bmin, const double *wptr, unsigned long long *hist, int nbins, double *mm, double *minmax_out,
                                    int *flags, int sr, int sw, const int *m2p, int mind, cudaStream_t st);
cudaError_t rsb_launch_width(const double *minmax, double w_old, double bmin, int hpts, double tol, double *wout, cudaStream_t st);
cudaError_t rsb_launch_symmetrize(double *cov, int L, int Lp, cudaStream_t st);
cudaError_t rsb_launch_negate_min(double *minmax, int n, cudaStream_t st);
cudaError_t rsb_launch_zero_unowned_rows(double *cov, int L, int Lp, int sr, int sw, cudaStream_t st);
cudaError_t rsb_launch_hist3(const double *cov, int L, int Lp, const uint8_t *pairmask, double bmin, double w, int nb,
                             unsigned long long *ha, unsigned long long *hb, unsigned long long *ht, int *flags, const int *m2p, int mind, cudaStream_t st);
cudaError_t rsb_launch_evalue_hits(const double *cov, int L, int Lp, const rsb_nullview &nv, const uint8_t *pairmask, double Nb, double Nt,
                                   double expBP, long long switch_n, double thresh, int report_all, int sr, int sw, double *eval, long long cap,
                                   long long *hit_ij, double *hit_sc, double *hit_eval, double *hit_pval, unsigned long long *nhit, int *flags,
                                   cudaStream_t st);
cudaError_t rsb_launch_branch_rows(const uint8_t *leaves, const uint8_t *internal, const int *left, const int *right, int ntaxa, int L,
                                   int includegaps, uint8_t *rows, int *nsubs, cudaStream_t st);
cudaError_t rsb_launch_subs_tables(const long long *cnt, int L, int Lp, int *ndouble, int *njoin, cudaStream_t st);
cudaError_t rsb_launch_null_simulate(const int *left, const int *right, const int *order, const int *level_start_host, int nlevels,
                                     const unsigned long long *pthr, int N, int L, const uint8_t *root,
                                     const uint8_t *gapmask, unsigned long long seed, unsigned long long id0, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *scratch, cudaStream_t st);
cudaError_t rsb_launch_fitch_shuffle(const int *left, const int *right, const int *parent, const int *order, const int *level_start,
                                     int nlevels, int N, int L, const uint8_t *msa, unsigned long long seed, unsigned long long id0,
                                     const unsigned long long *ids, int first_rep, int nrep,
                                     uint8_t *res, uint8_t *anc, uint8_t *shanc, int *perm, uint8_t *sets_shared, int build_sets, cudaStream_t st);
cudaError_t rsb_launch_unknown_check(const uint8_t *msa, size_t n, int *d_flag, int *unknown, cudaStream_t st);
cudaError_t rsb_launch_permutations(int L, unsigned long long seed, unsigned long long id0, const unsigned long long *ids, int first_rep, int nrep,
                                    int *perm, cudaStream_t st);

cudaError_t rsb_launch_gap_columns(const uint8_t *msa, int N, int L, long long row_stride, const double *wgt, double idthresh, uint8_t *useme, cudaStream_t st);
cudaError_t rsb_launch_pb_weights(const uint8_t *msa, int N, int L, long long row_stride, double *coef, double *w, cudaStream_t st);
cudaError_t rsb_launch_pair_identity(const uint8_t *msa, int N, int L, long long row_stride, const int *pairs, long long npairs, double *out, cudaStream_t st);
cudaError_t rsb_launch_column_subset(const uint8_t *msa, int N, long long row_stride, const int *cols, int nkeep, uint8_t *out, cudaStream_t st);

#include <nvtx3/nvToolsExt.h>
namespace {
// NVTX ranges (tracing): one per C-ABI phase on the calling thread, so that a timeline tool shows which host call
// enqueued which kernels; header-only NVTX v3, a no-op unless a tool is attached
struct NvtxRange { explicit NvtxRange(const char *name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };
#define RSB_RANGE(name) NvtxRange nvtx_range_(name)
constexpr int HIST_BINS = 1 << 22;

// operand geometry for one slice count
struct Geo {
  int S = 0, CJ = 0, NT = 0, nJB = 0, NBrows = 0, ntiles = 0, q = 0;
  CUtensorMap tmB, tmBh;                  // planeB boxes of a whole tile / of half a tile (CTA-pair kernel)
  int2 *d_tiles = nullptr, *d_tiles2 = nullptr;      // upper-triangle tiles (ib, jb) / row-block pairs (ibp, jb)
  int ntiles2 = 0, pair_clusters = 0;     // pair_clusters: resident CTA pairs of the pair kernel (0 = use the single-CTA kernel)
  uint8_t *d_wdig = nullptr;
  unsigned long long *d_wq = nullptr;
  std::vector<long long> wq;
  double scale = 1.0;
  double qerr_abs = 0.0, maxw = 1.0;      // largest |wq 2^-q - w| over the sequences, largest weight
  long long wtot = 0;
  bool ready = false;
};
} // namespace

constexpr int RSB_GROUPS = 4;          // slot groups of the pipelined null loop at most (alignments in flight: being packed / contracted / scored)

struct rsb_ctx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  char err[512] = { 0 };

  int N = 0, L = 0, Lp = 0, Kpad = 0, MA = 0, nIB = 0, Rcap = 0, Sreq = 0, Lcover = 0;
  size_t planeB_rows_cap = 0;
  Geo geo[3];                         // [0] weighted, [1] unit weights (RAF/RAFS), [2] weighted with Snull digit slices: the null alignments
  int Snull = 0;                      // digit slices of the null alignments' weights (0 = those of the input alignment), rsb_set_null_slices
  int cur_geo = 0;                    // geometry of the counts currently in d_cnt
  bool cur_rec = false;               // ... which hold per-pair G-test records instead of counts (record epilogue, gram_tcgen05.cu)
  bool fused_gt = true;               // nulls scored with GT x C16 use the record epilogue (RSCAPE_B200_FUSED_GT=0: counts + stat_kernel)
  int last_slot = 0;                  // replicate slot scanned last (quirk Q3)
  CUtensorMap tmA;
  std::vector<double> wgt;

  uint8_t *d_res = nullptr, *d_planeA = nullptr, *d_planeB = nullptr;
  long long *d_cnt = nullptr;
  double *d_nseff = nullptr, *d_pm = nullptr, *d_cov = nullptr, *d_tmp = nullptr;
  double *d_mrow = nullptr, *d_mcol = nullptr; size_t mrow_stride = 0, mcol_stride = 0;   // marginal partials of the gram tiles, per slot
  double *d_rowpart = nullptr, *d_colpart = nullptr, *d_mm = nullptr, *d_scal = nullptr, *d_covx = nullptr, *d_minmax = nullptr;
  // generation stream: the null generators run beside the scans; pool_ready[rep] is the event after which pool entry
  // `rep` is complete (NULL: nothing pending), taken from a ring of events
  cudaStream_t stream_gen = nullptr, stream_hi = nullptr;
  cudaEvent_t ev_exit = nullptr;
  std::vector<cudaEvent_t> pool_ready, gen_ring;
  size_t gen_next = 0;
  uint8_t *d_qscratch = nullptr;                     // quantise(): w[N] f64 | V[N] i64 | err[N] f64 | u[N] u8
  void *d_logtab = nullptr;                          // table of the statistic kernels' log (stats.cu), built once
  double *h_mm = nullptr; size_t h_mm_cap = 0;      // pinned staging of the per-replicate min/max: a pageable target would block the enqueuing thread
  double *d_meanp = nullptr, *d_w = nullptr, *d_blocksum = nullptr, *d_msum = nullptr, *d_covsum = nullptr;
  int shard_rank = 0, shard_world = 1;  // row-block sharding of the pair grid across ranks
  rsb_nccl::ncclComm_t comm = nullptr; int comm_rank = 0, comm_size = 1;   // NCCL communicator over the ranks / devices of the job (rsb_comm_*)
  // one-shot all-reduce of the per-scan vectors over peer memory (peer_reduce.cu): this rank's exchange block, the peers' blocks as
  // mapped here, the sequence number of the last all-reduce and the event that serialises them
  void *peer_block = nullptr; void *peer_map[RSB_PEER_MAX] = { nullptr }; bool peer_ipc = false;
  RsbPeerView peer_view; int peer_state = 0;       // 0 not tried, 1 in use, -1 unavailable (NCCL all-reduce instead)
  unsigned long long peer_seq[RSB_PEER_CHANNELS] = { 0 }, peer_reductions = 0; cudaEvent_t ev_peer[RSB_PEER_CHANNELS] = { nullptr };
  cudaStream_t stream_aux = nullptr, stream_aux2 = nullptr, stream_aux3 = nullptr, stream_aux4 = nullptr, stream_copy = nullptr;   // statistics (one per slot group) / uploads of the pipelined null loop
  cudaEvent_t ev_entry = nullptr, ev_up[RSB_GROUPS] = { nullptr }, ev_counts[RSB_GROUPS] = { nullptr }, ev_stats[RSB_GROUPS] = { nullptr },
              ev_marg[RSB_GROUPS] = { nullptr }, ev_statk[RSB_GROUPS] = { nullptr };
  unsigned long long *d_hist = nullptr, *d_colsum = nullptr;
  int *d_flags = nullptr;
  double *d_ps = nullptr, *d_pp_out = nullptr, *d_nseff_out = nullptr, *d_ngap_out = nullptr;
  // tree + simulators
  int *d_left = nullptr, *d_right = nullptr, *d_parent = nullptr, *d_order = nullptr, *d_level_start = nullptr, *d_perm = nullptr;
  int nlevels = 0;
  std::vector<int> h_left, h_right, h_parent;
  std::vector<double> h_ld, h_rd;
  double *d_pcdf = nullptr;
  std::vector<int> h_level_start;
  uint8_t *d_root = nullptr, *d_gapmask = nullptr, *d_simscratch = nullptr, *d_msa0 = nullptr, *d_anc = nullptr, *d_shanc = nullptr, *d_sets = nullptr;
  int *d_genflag = nullptr;
  unsigned long long *d_pthr = nullptr;                   // generator B: branch matrices as cumulative thresholds
  double sim_Q[16] = { 0 }; bool sim_valid = false;       // ... valid for this rate matrix and the current tree
  unsigned long long *d_ids = nullptr; size_t ids_cap = 0;    // explicit replicate ids of a generator call
  bool have_tree = false;
  uint8_t *d_pool = nullptr;          // device-resident null alignments [Rpool][N][L] (output of the generators)
  int Rpool = 0;

  unsigned long long hist_n = 0;
  // several statistics per contraction (rsb_null_hist_multi): raw matrices [6][Rcap][L][Lp], one histogram / width / range per combination
  double *d_covm = nullptr, *d_wm = nullptr, *d_minmaxm = nullptr; unsigned long long *d_histm = nullptr;
  double *d_rowpart_m = nullptr, *d_colpart_m = nullptr, *d_mmr_m = nullptr, *d_covx_m = nullptr, *d_scal_m = nullptr, *d_blocksum_m = nullptr,
         *d_covsum_m = nullptr, *d_mmc_m = nullptr;                  // the statistics chain's buffers with (slot, statistic) as the replicate index
  int multi_cap = 0; std::vector<unsigned long long> hist_n_m;
  int *d_m2p = nullptr; int mind = 1;                 // PDB positions of the columns + minimum distance: pairs kept out of the histograms
  unsigned long long pairs_in_hist = 0;               // pairs i<j that are not excluded by that rule
  long long launches = 0, gram_launches = 0;
  double gram_ms = 0.0;
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  std::vector<int> pending_geo;                                   // geometry of each timed contraction
  double gram_ms_geo[3] = { 0.0, 0.0, 0.0 }; long long gram_launches_geo[3] = { 0, 0, 0 };
  double last_gram_ms_geo[3] = { 0.0, 0.0, 0.0 }; long long last_gram_launches_geo[3] = { 0, 0, 0 };   // as of the last rsb_counters call
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_aux;   // statistics chain of the pipelined null loop
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_stage; // its stage boundaries (after marginals, after statistic)
  double stage_ms[3] = { 0.0, 0.0, 0.0 };
  double aux_ms = 0.0; long long aux_chains = 0;
};

void rsb_set_error(rsb_ctx *ctx, const char *fmt, ...)
{
  if (!ctx) return;
  va_list ap; va_start(ap, fmt);
  vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
  va_end(ap);
}

namespace {
char g_create_err[512] = "";

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn) p;
  }
  return fn;
}

// 3-D u8 tensor map {Kpad (contiguous), rows, replicates}, box {128, box_rows, 1}, 128-byte swizzle
int make_plane_map(rsb_ctx *ctx, CUtensorMap *tm, void *base, int Kpad, size_t rows, int reps, int box_rows)
{
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { rsb_set_error(ctx, "cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t dims[3]    = { (cuuint64_t) Kpad, (cuuint64_t) rows, (cuuint64_t) reps };
  cuuint64_t strides[2] = { (cuuint64_t) Kpad, (cuuint64_t) Kpad * rows };
  cuuint32_t box[3]     = { (cuuint32_t) RSB_KSTAGE, (cuuint32_t) box_rows, 1 };
  cuuint32_t estr[3]    = { 1, 1, 1 };
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { rsb_set_error(ctx, "cuTensorMapEncodeTiled failed with CUresult %d", (int) rc); return 1; }
  return 0;
}

template <class T> void dfree(T *&p) { if (p) { cudaFree(p); p = nullptr; } }

void free_geo(Geo &g) { dfree(g.d_tiles); dfree(g.d_tiles2); dfree(g.d_wdig); dfree(g.d_wq); g.ready = false; }

void free_plan(rsb_ctx *c)
{
  if (c->stream_gen) cudaStreamSynchronize(c->stream_gen);
  c->pool_ready.clear();
  free_geo(c->geo[0]); free_geo(c->geo[1]); free_geo(c->geo[2]);
  dfree(c->d_qscratch);
  dfree(c->d_res); dfree(c->d_planeA); dfree(c->d_planeB); dfree(c->d_cnt); dfree(c->d_nseff); dfree(c->d_pm); dfree(c->d_cov);
  dfree(c->d_mrow); dfree(c->d_mcol); dfree(c->d_tmp); dfree(c->d_rowpart); dfree(c->d_colpart); dfree(c->d_mm); dfree(c->d_scal); dfree(c->d_covx); dfree(c->d_minmax);
  if (c->h_mm) { cudaFreeHost(c->h_mm); c->h_mm = nullptr; c->h_mm_cap = 0; }
  dfree(c->d_meanp); dfree(c->d_w); dfree(c->d_blocksum); dfree(c->d_msum); dfree(c->d_covsum); dfree(c->d_hist); dfree(c->d_colsum); dfree(c->d_flags); dfree(c->d_ps); dfree(c->d_pp_out);
  dfree(c->d_nseff_out); dfree(c->d_ngap_out); dfree(c->d_left); dfree(c->d_right); dfree(c->d_parent); dfree(c->d_order);
  dfree(c->d_level_start); dfree(c->d_perm); dfree(c->d_pcdf); dfree(c->d_root); dfree(c->d_gapmask); dfree(c->d_simscratch);
  dfree(c->d_m2p); c->mind = 1;
  dfree(c->d_covm); dfree(c->d_wm); dfree(c->d_minmaxm); dfree(c->d_histm); c->multi_cap = 0; c->hist_n_m.clear();
  dfree(c->d_rowpart_m); dfree(c->d_colpart_m); dfree(c->d_mmr_m); dfree(c->d_covx_m); dfree(c->d_scal_m); dfree(c->d_blocksum_m); dfree(c->d_covsum_m); dfree(c->d_mmc_m);
  dfree(c->d_msa0); dfree(c->d_anc); dfree(c->d_shanc); dfree(c->d_pool); dfree(c->d_sets); dfree(c->d_genflag); dfree(c->d_ids); c->ids_cap = 0; dfree(c->d_pthr); c->sim_valid = false;
  c->Rpool = 0;
  c->have_tree = false;
}

// tiles (ib, jb) of the upper triangle, jb-major so that concurrently running CTAs share planeB rows in L2
int build_geo(rsb_ctx *ctx, Geo &g, int S)
{
  g.S  = S;
  g.CJ = rsb_cj_for(S);
  g.NT = 4 * S * g.CJ;
  g.nJB = (ctx->L + g.CJ - 1) / g.CJ;
  g.NBrows = g.nJB * g.NT;
  if ((size_t) g.NBrows > ctx->planeB_rows_cap) { rsb_set_error(ctx, "internal: planeB capacity"); return 1; }
  std::vector<int2> tiles;
  for (int jb = 0; jb < g.nJB; jb++) {
    const int maxj = std::min(jb * g.CJ + g.CJ - 1, ctx->L - 1);
    for (int ib = 0; ib < ctx->nIB; ib++)
      if (ib * RSB_ICOLS < maxj && (ctx->shard_world <= 1 || ib % ctx->shard_world == ctx->shard_rank)) tiles.push_back(make_int2(ib, jb));
  }
  g.ntiles = (int) tiles.size();
  dfree(g.d_tiles);
  if (g.ntiles > 0) {
    RSB_CUDA_OK(cudaMalloc(&g.d_tiles, sizeof(int2) * tiles.size()));
    RSB_CUDA_OK(cudaMemcpyAsync(g.d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  if (make_plane_map(ctx, &g.tmB, ctx->d_planeB, ctx->Kpad, (size_t) g.NBrows, ctx->Rcap, g.NT)) return 1;
  // CTA-pair kernel: row blocks taken two at a time (not with a sharded pair grid, whose ownership is per row block)
  g.ntiles2 = 0; g.pair_clusters = 0;
  dfree(g.d_tiles2);
  // Measured on the SSU shape (S = 4): 0.714 ms per contraction against 0.700 ms for the single-CTA kernel -- the
  // contraction already runs at the rate of the cuBLAS bf16 reference x2, so halving the planeB traffic buys nothing.
  // Kept selectable (RSCAPE_B200_GRAM_PAIR=1, read when the tile lists are built) and covered by the count tests.
  const char *pair_env = getenv("RSCAPE_B200_GRAM_PAIR");
  const bool use_pair = pair_env && atoi(pair_env) != 0;
  if (ctx->shard_world <= 1 && use_pair && g.ntiles > 0) {
    std::vector<int2> t2;
    for (int jb = 0; jb < g.nJB; jb++) {
      const int maxj = std::min(jb * g.CJ + g.CJ - 1, ctx->L - 1);
      for (int ibp = 0; 2 * ibp < ctx->nIB; ibp++) if (2 * ibp * RSB_ICOLS < maxj) t2.push_back(make_int2(ibp, jb));
    }
    const int nc = rsb_gram_pair_clusters(S);
    if (nc > 0 && !t2.empty()) {
      if (make_plane_map(ctx, &g.tmBh, ctx->d_planeB, ctx->Kpad, (size_t) g.NBrows, ctx->Rcap, g.NT / 2)) return 1;
      RSB_CUDA_OK(cudaMalloc(&g.d_tiles2, sizeof(int2) * t2.size()));
      RSB_CUDA_OK(cudaMemcpyAsync(g.d_tiles2, t2.data(), sizeof(int2) * t2.size(), cudaMemcpyHostToDevice, ctx->stream));
      RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      g.ntiles2 = (int) t2.size();
      g.pair_clusters = std::min(nc, ctx->sm_count / 2);
      if (getenv("RSCAPE_B200_TRACE")) fprintf(stderr, "[rsb] CTA-pair contraction: %d resident pairs on %d SMs, %d pair tiles\n", g.pair_clusters, ctx->sm_count, g.ntiles2);
    }
  }
  dfree(g.d_wdig); dfree(g.d_wq);
  RSB_CUDA_OK(cudaMalloc(&g.d_wdig, (size_t) (S + 1) * ctx->Kpad));
  RSB_CUDA_OK(cudaMalloc(&g.d_wq, sizeof(unsigned long long) * ctx->N));
  return 0;
}

// fixed-point weights: wq = round(w 2^q) < 256^S, digits base 256
// Fixed-point weights for the u8 x u8 -> s32 tensor-core contraction.  A weight is represented as
//     wq_s = u_s * V_s,   u_s in [1, 255] (carried by planeA's one-hot rows),  V_s < 256^S (S base-256 digits in planeB)
// and approximates w_s 2^q.  All arithmetic on wq is exact (int32 accumulators: N * 255 * 255 < 2^31; the digit sums are
// recombined in int64), so results are bit-reproducible functions of wq; (u_s, V_s) is the pair minimising
// |w_s 2^q - u V| over u, which buys ~6 more bits than V alone for the same number of slices.  Integer weights <= 255
// with S = 1 (and the unit weights of the RAF tables) are exact with u = 1, q = 0.
int quantise(rsb_ctx *ctx, Geo &g, const std::vector<double> &w, bool unit)
{
  const int N = ctx->N, S = g.S;
  g.wq.assign(N, 1);
  g.q = 0;
  g.qerr_abs = 0.0; g.maxw = 1.0;
  std::vector<uint8_t> mul(N, 1);
  std::vector<long long> V(N, 1);
  if (!unit) {
    double maxw = 0.0;
    for (int s = 0; s < N; s++) {
      if (!(w[s] >= 0.0) || !std::isfinite(w[s])) { rsb_set_error(ctx, "sequence weight %d is negative or not finite", s); return 1; }
      maxw = std::max(maxw, w[s]);
    }
    g.maxw = maxw;
    bool small_int = true;
    for (int s = 0; s < N && small_int; s++) small_int = (w[s] == std::floor(w[s]) && w[s] <= 255.0);
    const long long vlim = (S >= 8) ? (1LL << 62) : (1LL << (8 * S));                    // V < vlim
    if (small_int && S == 1) {
      for (int s = 0; s < N; s++) V[s] = (long long) w[s];
    } else if (maxw > 0.0) {
      const long long acc_cap = 2147483647LL / (255LL * std::max(ctx->Kpad, 1));        // keeps sum_s u_s d_s inside int32
      int umax = (int) std::max(1LL, std::min(255LL, acc_cap));
      for (;; umax = std::max(1, umax / 2)) {
        int e; std::frexp(maxw, &e);                                                     // maxw < 2^e
        int eu = 0; while ((2 << eu) <= umax) eu++;                                      // 2^eu <= umax
        g.q = 8 * S + eu - e;
        while (std::ldexp((long double) maxw, g.q) > (long double) umax * (long double) (vlim - 1)) g.q--;
        long double tot = 0.0L;
        g.qerr_abs = 0.0;
        const bool fast = std::ldexp(maxw, g.q) < 1125899906842624.0;                     // W < 2^50: doubles are exact enough
        if (fast) {
          // the search over u runs on the device (quantise_kernel, pack.cu)
          if (!ctx->d_qscratch) RSB_CUDA_OK(cudaMalloc(&ctx->d_qscratch, (size_t) N * 25));
          double *d_w = (double *) ctx->d_qscratch; long long *d_V = (long long *) (ctx->d_qscratch + (size_t) N * 8);
          double *d_err = (double *) (ctx->d_qscratch + (size_t) N * 16); uint8_t *d_mul = ctx->d_qscratch + (size_t) N * 24;
          std::vector<double> herr(N);
          RSB_CUDA_OK(cudaMemcpyAsync(d_w, w.data(), sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
          RSB_CUDA_OK(rsb_launch_quantise(d_w, N, g.q, umax, (double) (vlim - 1), d_mul, d_V, d_err, ctx->stream));
          RSB_CUDA_OK(cudaMemcpyAsync(V.data(), d_V, sizeof(long long) * N, cudaMemcpyDeviceToHost, ctx->stream));
          RSB_CUDA_OK(cudaMemcpyAsync(mul.data(), d_mul, N, cudaMemcpyDeviceToHost, ctx->stream));
          RSB_CUDA_OK(cudaMemcpyAsync(herr.data(), d_err, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
          RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
          for (int s = 0; s < N; s++) {
            if (V[s] < 0) { rsb_set_error(ctx, "internal: weight %d not representable", s); return 1; }
            g.qerr_abs = std::max(g.qerr_abs, std::ldexp(herr[s], -g.q));
            tot += (long double) mul[s] * (long double) V[s];
          }
        } else
        for (int s = 0; s < N; s++) {
          int ub = 1; long long vb = 0; long double best = -1.0L;
          {
            const long double W = std::ldexp((long double) w[s], g.q);
            for (int u = 1; u <= umax; u++) {
              const long double v = std::nearbyint(W / u);
              if (v >= (long double) vlim) continue;
              const long double err = std::fabs(W - u * v);
              if (best < 0.0L || err < best) { best = err; ub = u; vb = (long long) v; }
            }
          }
          if (best < 0.0L) { rsb_set_error(ctx, "internal: weight %d not representable", s); return 1; }
          mul[s] = (uint8_t) ub; V[s] = vb;
          g.qerr_abs = std::max(g.qerr_abs, (double) std::ldexp(best, -g.q));
          tot += (long double) mul[s] * (long double) V[s];
        }
        if (tot < 4.0e18L) break;                                                        // every count <= wtot < 2^62
        if (umax == 1) { rsb_set_error(ctx, "%d weight slices with %d sequences overflow the 63-bit count (use fewer slices)", S, N); return 1; }
      }
    }
    for (int s = 0; s < N; s++) g.wq[s] = (long long) mul[s] * V[s];
  }
  g.scale = std::ldexp(1.0, -g.q);
  g.wtot = 0;
  std::vector<uint8_t> dig((size_t) (S + 1) * ctx->Kpad, 0);                             // rows 0..S-1: digits of V, row S: u
  for (int s = 0; s < N; s++) {
    g.wtot += g.wq[s];
    for (int k = 0; k < S; k++) dig[(size_t) k * ctx->Kpad + s] = (uint8_t) ((V[s] >> (8 * k)) & 0xFF);
    dig[(size_t) S * ctx->Kpad + s] = mul[s];
  }
  RSB_CUDA_OK(cudaMemcpyAsync(g.d_wdig, dig.data(), dig.size(), cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(g.d_wq, g.wq.data(), sizeof(long long) * N, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  g.ready = true;
  return 0;
}

int ensure_geo(rsb_ctx *ctx, int which)
{
  Geo &g = ctx->geo[which];
  if (g.ready) return 0;
  if (which == 1) {
    if (build_geo(ctx, g, 1)) return 1;
    return quantise(ctx, g, ctx->wgt, true);
  }
  rsb_set_error(ctx, "rsb_set_weights has not been called");
  return 1;
}

// geometry that scores the null alignments: their own (coarser) fixed-point weights when rsb_set_null_slices asked for them
int null_geo(rsb_ctx *ctx) { return (ctx->Snull > 0 && ctx->geo[2].ready && ctx->geo[2].S != ctx->geo[0].S) ? 2 : 0; }

unsigned allow_mask(const double *allowpair)
{
  // default WC + GU (src/R-scape.c:883-887)
  static const double dflt[16] = { 0, 0, 0, 1,  0, 0, 1, 0,  0, 1, 0, 1,  1, 0, 1, 0 };
  const double *ap = allowpair ? allowpair : dflt;
  unsigned m = 0;
  for (int k = 0; k < 16; k++) if (ap[k] > 0.0) m |= 1u << k;
  return m;
}

// alignments -> replicate slots [first_slot, first_slot + nrep) on stream st
int upload_msa(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int64_t rep_stride, int nrep, int first_slot, int on_device, cudaStream_t st)
{
  const size_t repbytes = (size_t) ctx->N * ctx->L;
  if (first_slot + nrep > ctx->Rcap) { rsb_set_error(ctx, "%d replicates exceed the configured %d slots", first_slot + nrep, ctx->Rcap); return 1; }
  for (int r = 0; r < nrep; r++) {
    const uint8_t *src = msa + (size_t) r * rep_stride;
    RSB_CUDA_OK(cudaMemcpy2DAsync(ctx->d_res + (size_t) (first_slot + r) * repbytes, ctx->L, src, (size_t) row_stride, ctx->L, ctx->N,
                                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  }
  return 0;
}

// residues -> operand planes of replicate slots [s0, s0 + nrep) with geometry `which`, on stream st.
// src: [nrep][N][L] contiguous (the slots themselves, or a caller's device buffer read in place).
int enqueue_pack(rsb_ctx *ctx, int which, int s0, int nrep, const uint8_t *src, cudaStream_t st)
{
  RSB_RANGE("rsb:pack_planes");
  if (ensure_geo(ctx, which)) return 1;
  Geo &g = ctx->geo[which];
  RSB_CUDA_OK(rsb_launch_pack(g.S, src, nrep, ctx->N, ctx->L, (long long) ctx->N * ctx->L, g.d_wdig, ctx->Kpad,
                              ctx->d_planeA + (size_t) s0 * ctx->MA * ctx->Kpad, ctx->MA,
                              ctx->d_planeB + (size_t) s0 * g.NBrows * ctx->Kpad, g.NBrows, ctx->Lcover, st));
  ctx->launches++;
  return 0;
}

// tcgen05 contraction of the planes of slots [s0, s0 + nrep) -> count planes, on stream st
int enqueue_gram(rsb_ctx *ctx, int which, int s0, int nrep, cudaStream_t st, bool rec = false)
{
  RSB_RANGE("rsb:gram_tcgen05");
  Geo &g = ctx->geo[which];
  if (rec && (which == 1 || g.pair_clusters > 0)) { rsb_set_error(ctx, "internal: record epilogue not available here"); return 1; }
  if (g.ntiles > 0) {
    const long long work = (long long) g.ntiles * nrep;
    // with in-library collectives a few SMs stay free: the NCCL kernels cannot share an SM with the persistent contraction
    // (its ring fills the shared memory) and would otherwise wait for the end of every launch
    const int sms = (ctx->shard_world > 1 && ctx->comm && ctx->comm_size == ctx->shard_world) ? std::max(1, ctx->sm_count - 4) : ctx->sm_count;
    const int grid = (int) std::min<long long>(work, sms);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profile) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    // the weighted geometry also emits the marginal partial sums of every tile (slot-indexed like the counts)
    double *mrow = (which != 1) ? ctx->d_mrow : nullptr, *mcol = (which != 1) ? ctx->d_mcol : nullptr;
    if (which != 1 && ((size_t) g.nJB * rsb_gram_mrow_blocks(0) * ctx->L * 4 > ctx->mrow_stride)) { rsb_set_error(ctx, "internal: marginal partial capacity"); return 1; }
    if (g.pair_clusters > 0)
      RSB_CUDA_OK(rsb_launch_gram_i8_pair(g.S, ctx->tmA, g.tmBh, g.d_tiles2, g.ntiles2, s0, nrep, ctx->L, ctx->Lp, ctx->Kpad / RSB_KSTAGE,
                                          ctx->d_cnt, g.scale, mrow, mcol, g.nJB, ctx->nIB, ((unsigned long long) g.wtot >> 52) == 0, g.pair_clusters, st));
    else
      RSB_CUDA_OK(rsb_launch_gram_i8(g.S, ctx->tmA, g.tmB, g.d_tiles, g.ntiles, s0, nrep, ctx->L, ctx->Lp, ctx->Kpad / RSB_KSTAGE,
                                     ctx->d_cnt, g.scale, mrow, mcol, g.nJB, ctx->nIB, ((unsigned long long) g.wtot >> 52) == 0, grid,
                                     rec ? ctx->d_logtab : nullptr, st));
    if (ctx->profile) { cudaEventRecord(e1, st); ctx->pending.push_back({ e0, e1 }); ctx->pending_geo.push_back(which); }
    ctx->launches++;
    ctx->gram_launches++;
    ctx->gram_launches_geo[which]++;
  }
  ctx->cur_geo = which;
  ctx->cur_rec = rec;
  ctx->last_slot = s0 + nrep - 1;
  return 0;
}

int enqueue_counts(rsb_ctx *ctx, int which, int s0, int nrep, const uint8_t *src, cudaStream_t st, bool rec = false)
{
  if (enqueue_pack(ctx, which, s0, nrep, src, st)) return 1;
  return enqueue_gram(ctx, which, s0, nrep, st, rec);
}

// may the nulls of this (statistic, class) be scored through the record epilogue?
bool use_record(rsb_ctx *ctx, int stat, int covclass)
{
  return ctx->fused_gt && stat == RSB_GT && covclass == RSB_C16 && ctx->geo[0].pair_clusters == 0 && ctx->geo[2].pair_clusters == 0;
}

int check_flags(rsb_ctx *ctx, const char *what)
{
  int f = 0;
  RSB_CUDA_OK(cudaMemcpyAsync(&f, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (f) {
    RSB_CUDA_OK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    if (f & 1) { rsb_set_error(ctx, "%s: pm validation failed", what); return 1; }           // corr_Marginals / corr_ValidateProbs
    if (f & 2) { rsb_set_error(ctx, "%s: bad covariation (NaN)", what); return 1; }          // correlators.c:1124
    if (f & 4) { rsb_set_error(ctx, "%s: score histogram capacity (%d bins) exceeded", what, HIST_BINS); return 1; }
    if (f & 8) { rsb_set_error(ctx, "%s: cannot find evalue for a covariation score", what); return 1; }      // covariation.c:2394
  }
  return 0;
}

int resolve_stat(rsb_ctx *ctx, int stat, int covclass)
{
  if (covclass == RSB_CSELECT) { rsb_set_error(ctx, "covclass must be resolved by the caller (CSELECT rule, correlators.c:336)"); return 1; }
  switch (stat) {
  case RSB_GT: if (covclass == RSB_C16 || covclass == RSB_C2 || covclass == RSB_CWC) return 0; break;
  case RSB_CHI: case RSB_OMES: case RSB_MI: case RSB_MIr: case RSB_MIg:
    if (covclass == RSB_C16 || covclass == RSB_C2) return 0;
    rsb_set_error(ctx, "CWC not implemented for this statistic"); return 1;                   // e.g. correlators.c:72
  case RSB_RAF: case RSB_RAFS: case RSB_CCF: return 0;
  }
  rsb_set_error(ctx, "wrong covariation type %d / class %d", stat, covclass);
  return 1;
}

// per-slot views of the replicate-indexed buffers
struct SlotPtrs {
  long long *cnt; double *nseff, *pm, *cov, *tmp, *scal, *covx, *minmax, *meanp, *blocksum, *mm, *msum, *covsum;
};
SlotPtrs slot_ptrs(rsb_ctx *c, int s0)
{
  const size_t L = c->L, Lp = c->Lp;
  int nJT, nIT; rsb_stat_grid(c->L, &nJT, &nIT);
  SlotPtrs p;
  p.cnt = c->d_cnt + (size_t) s0 * 16 * L * Lp;
  p.nseff = c->d_nseff + (size_t) s0 * L * Lp;
  p.pm = c->d_pm + (size_t) s0 * L * 4;
  p.cov = c->d_cov + (size_t) s0 * L * Lp;
  p.tmp = c->d_tmp + (size_t) s0 * std::max(L * Lp, (size_t) nJT * nIT * 4);
  p.scal = c->d_scal + (size_t) s0 * 4;
  p.covx = c->d_covx + (size_t) s0 * L;
  p.minmax = c->d_minmax + (size_t) s0 * 2;
  p.meanp = c->d_meanp + (size_t) s0 * 4;
  p.blocksum = c->d_blocksum + (size_t) s0 * ((L + 127) / 128);
  p.mm = c->d_mm + (size_t) s0 * nJT * nIT * 2;
  p.msum = c->d_msum + (size_t) s0 * L * 4;
  p.covsum = c->d_covsum + (size_t) s0 * (L + 4);
  return p;
}

// ---- exchange blocks of the one-shot all-reduce (peer_reduce.cu)
void peer_teardown(rsb_ctx *ctx)
{
  if (ctx->peer_ipc) for (int q = 0; q < RSB_PEER_MAX; q++) if (ctx->peer_map[q]) cudaIpcCloseMemHandle(ctx->peer_map[q]);
  for (int q = 0; q < RSB_PEER_MAX; q++) ctx->peer_map[q] = nullptr;
  if (ctx->peer_block) cudaFree(ctx->peer_block);
  ctx->peer_block = nullptr; ctx->peer_ipc = false; ctx->peer_state = 0;
  for (int k = 0; k < RSB_PEER_CHANNELS; k++) ctx->peer_seq[k] = 0;
}

void peer_fill_view(rsb_ctx *ctx, void *const *blocks, size_t cap)
{
  RsbPeerView &v = ctx->peer_view;
  v.W = ctx->comm_size; v.rank = ctx->comm_rank; v.cap = cap;
  for (int q = 0; q < ctx->comm_size; q++) {
    v.x[q]    = (double *) blocks[q];
    v.flag[q] = (unsigned long long *) ((char *) blocks[q] + rsb_peer_doubles(ctx->comm_size, cap) * sizeof(double));
  }
}

// One process per GPU: every rank allocates its block, the cudaIpc handles travel by an NCCL all-gather, every rank maps the
// others' blocks, and an NCCL all-reduce of a flag makes the ranks agree on whether all of them succeeded (a rank that fell back
// to NCCL alone would deadlock the others).  Collective: every rank calls it at the same point (the first small all-reduce).
int peer_setup_ipc(rsb_ctx *ctx, size_t need)
{
  rsb_nccl::Api &a = rsb_nccl::api();
  const int W = ctx->comm_size;
  if (ctx->peer_block) RSB_CUDA_OK(cudaDeviceSynchronize());          // (re-sizing: nothing of this rank may still use the old block)
  peer_teardown(ctx);
  ctx->peer_state = -1;
  if (getenv("RSCAPE_B200_PEER_REDUCE") && atoi(getenv("RSCAPE_B200_PEER_REDUCE")) == 0) return 0;
  if (W > RSB_PEER_MAX || !a.AllGather) return 0;
  const size_t cap = std::max(need, (size_t) 4 * ctx->L * std::max(1, ctx->Rcap) + 64);
  const size_t bytes = rsb_peer_block_bytes(W, cap);
  long long ok = 1;
  cudaIpcMemHandle_t mine;
  char *d_h = nullptr; long long *d_ok = nullptr;
  std::vector<cudaIpcMemHandle_t> all(W);
  void *blocks[RSB_PEER_MAX] = { nullptr };
  if (cudaMalloc(&ctx->peer_block, bytes) != cudaSuccess) { cudaGetLastError(); ctx->peer_block = nullptr; ok = 0; }
  if (ok && cudaMemsetAsync(ctx->peer_block, 0, bytes, ctx->stream) != cudaSuccess) ok = 0;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, ctx->peer_block) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  RSB_CUDA_OK(cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * (W + 1)));
  RSB_CUDA_OK(cudaMalloc(&d_ok, sizeof(long long)));
  RSB_CUDA_OK(cudaMemcpyAsync(d_h + sizeof(mine) * W, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  int rc = a.AllGather(d_h + sizeof(mine) * W, d_h, sizeof(mine), 0 /* ncclInt8 */, ctx->comm, ctx->stream);
  if (rc == rsb_nccl::ncclSuccess) {
    RSB_CUDA_OK(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * W, cudaMemcpyDeviceToHost, ctx->stream));
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < W && ok; q++) {
      if (q == ctx->comm_rank) { blocks[q] = ctx->peer_block; continue; }
      if (cudaIpcOpenMemHandle(&ctx->peer_map[q], all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ctx->peer_map[q] = nullptr; ok = 0; }
      blocks[q] = ctx->peer_map[q];
    }
    ctx->peer_ipc = true;
  } else ok = 0;
  // agreement (and a barrier: every rank's block is zeroed before anyone pushes into it)
  RSB_CUDA_OK(cudaMemcpyAsync(d_ok, &ok, sizeof(ok), cudaMemcpyHostToDevice, ctx->stream));
  rc = a.AllReduce(d_ok, d_ok, 1, rsb_nccl::ncclInt64, rsb_nccl::ncclMin, ctx->comm, ctx->stream);
  if (rc != rsb_nccl::ncclSuccess) { cudaFree(d_h); cudaFree(d_ok); rsb_set_error(ctx, "ncclAllReduce: %s", a.GetErrorString(rc)); return 1; }
  RSB_CUDA_OK(cudaMemcpyAsync(&ok, d_ok, sizeof(ok), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_h); cudaFree(d_ok);
  if (!ok) { peer_teardown(ctx); ctx->peer_state = -1; return 0; }
  peer_fill_view(ctx, blocks, cap);
  ctx->peer_state = 1;
  return 0;
}

// in-place all-reduce over the job's ranks on stream st (no-op without a communicator of more than one rank).  Small fp64 vectors
// (sum / max) go through the one-shot kernel over peer memory when the ranks could map each other's exchange blocks.
int comm_allreduce(rsb_ctx *ctx, void *buf, size_t count, int dtype, int op, cudaStream_t st)
{
  if (!ctx->comm || ctx->comm_size <= 1) return 0;
  if (dtype == rsb_nccl::ncclFloat64 && (op == rsb_nccl::ncclSum || op == rsb_nccl::ncclMax) && count <= ((size_t) 1 << 20)) {
    if (ctx->peer_state == 0 && peer_setup_ipc(ctx, count)) return 1;                   // (collective: the first small all-reduce of every rank)
    if (ctx->peer_state == 1 && count <= ctx->peer_view.cap) {                            // (a vector beyond the block's capacity: NCCL, on every rank alike)
      // one channel per statistics stream (the chains of consecutive replicates overlap), channel 0 for everything else
      const int ch = (st == ctx->stream_aux2) ? 1 : (st == ctx->stream_aux3) ? 2 : (st == ctx->stream_aux4) ? 3 : 0;
      if (!ctx->ev_peer[ch]) RSB_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_peer[ch], cudaEventDisableTiming));
      if (ctx->peer_seq[ch] > 0) RSB_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev_peer[ch], 0));   // the all-reduces of a channel are serialised
      RSB_CUDA_OK(rsb_launch_peer_allreduce((double *) buf, count, op == rsb_nccl::ncclMax, ctx->peer_view, ch, ++ctx->peer_seq[ch], st));
      RSB_CUDA_OK(cudaEventRecord(ctx->ev_peer[ch], st));
      ctx->launches++; ctx->peer_reductions++;
      return 0;
    }
  }
  const int rc = rsb_nccl::api().AllReduce(buf, buf, count, dtype, op, ctx->comm, st);
  if (rc != rsb_nccl::ncclSuccess) { rsb_set_error(ctx, "ncclAllReduce: %s", rsb_nccl::api().GetErrorString(rc)); return 1; }
  ctx->launches++;
  return 0;
}
// is the pair grid sharded over the communicator's ranks (the data-path all-reduces are then done inside the library)?
bool grid_comm(rsb_ctx *ctx) { return ctx->shard_world > 1 && ctx->comm && ctx->comm_size == ctx->shard_world; }

// per-replicate (min, max) pairs -> global: MIN on [0], MAX on [1] as one MAX all-reduce of (-min, max)
int comm_reduce_minmax(rsb_ctx *ctx, double *minmax, int n, cudaStream_t st)
{
  if (!ctx->comm || ctx->comm_size <= 1) return 0;
  RSB_CUDA_OK(rsb_launch_negate_min(minmax, n, st));
  if (comm_allreduce(ctx, minmax, (size_t) 2 * n, rsb_nccl::ncclFloat64, rsb_nccl::ncclMax, st)) return 1;
  RSB_CUDA_OK(rsb_launch_negate_min(minmax, n, st));
  ctx->launches += 2;
  return 0;
}

// marginals (corr_Marginals) for slots [s0, s0+nrep) on stream st.  phase 1 = partial sums -> msum, 2 = normalise, 3 = both
int enqueue_marginals(rsb_ctx *ctx, int s0, int nrep, double tol, cudaStream_t st, int phase = 3)
{
  RSB_RANGE("rsb:marginals");
  Geo &g = ctx->geo[ctx->cur_geo];                                  // (0 or 2: the weighted geometry the contraction just ran with)
  SlotPtrs p = slot_ptrs(ctx, s0);
  // partials are addressed [r][block][L][4] with r the absolute slot, as the gram kernel wrote them
  const int E = rsb_gram_mrow_blocks(g.pair_clusters > 0);
  if (phase == 3 && grid_comm(ctx)) {                               // partial sums of the owned tiles, summed over the ranks, then normalised
    if (enqueue_marginals(ctx, s0, nrep, tol, st, 1)) return 1;
    if (comm_allreduce(ctx, p.msum, (size_t) nrep * ctx->L * 4, rsb_nccl::ncclFloat64, rsb_nccl::ncclSum, st)) return 1;
    return enqueue_marginals(ctx, s0, nrep, tol, st, 2);
  }
  RSB_CUDA_OK(rsb_launch_marginals(ctx->d_mrow + (size_t) s0 * g.nJB * E * ctx->L * 4, ctx->d_mcol + (size_t) s0 * 4 * ctx->nIB * ctx->L * 4,
                                   nrep, ctx->L, g.CJ, g.nJB, ctx->nIB, tol, p.msum, p.pm, ctx->d_flags,
                                   ctx->shard_rank, ctx->shard_world, phase, E, st));
  ctx->launches += (phase == 3) ? 2 : 1;
  return 0;
}

// statistic on the counts of slots [s0, s0+nrep); leaves raw cov and (phase bit 1) the reduced sums covsum, (phase bit 2)
// COVx, COVavg and the raw min/max
int enqueue_statistic(rsb_ctx *ctx, int s0, int nrep, int stat, int covclass, unsigned mask, cudaStream_t st, int phase = 3, int part = 3)
{
  RSB_RANGE("rsb:statistic");
  // part: 1 = the statistic kernel(s) only, 2 = the reductions of its partial sums only, 3 = both
  Geo &g = ctx->geo[ctx->cur_geo];
  int nJT, nIT; rsb_stat_grid(ctx->L, &nJT, &nIT);
  SlotPtrs p = slot_ptrs(ctx, s0);
  double *rowpart = ctx->d_rowpart + (size_t) s0 * nJT * ctx->L, *colpart = ctx->d_colpart + (size_t) s0 * nIT * ctx->L;
  if ((phase & 1) && (part & 1)) {
    if (stat == RSB_RAF || stat == RSB_RAFS) {
      if (ctx->cur_geo != 1) { rsb_set_error(ctx, "internal: RAF needs unit-weight counts"); return 1; }
      if (ctx->shard_world > 1) { rsb_set_error(ctx, "RAF/RAFS are not available with a sharded pair grid"); return 1; }
      RSB_CUDA_OK(rsb_launch_raf(p.cnt, nrep, ctx->L, ctx->Lp, ctx->N, mask, stat == RSB_RAFS, p.tmp, p.cov, rowpart, colpart, p.mm, st));
      ctx->launches += (stat == RSB_RAFS) ? 3 : 2;
    } else if (stat == RSB_CCF) {
      if (ctx->shard_world > 1) { rsb_set_error(ctx, "CCF is not available with a sharded pair grid"); return 1; }
      RSB_CUDA_OK(rsb_launch_nseff(p.cnt, nrep, ctx->L, ctx->Lp, g.scale, p.nseff, st));
      RSB_CUDA_OK(rsb_launch_ccf(p.nseff, p.pm, nrep, ctx->L, ctx->Lp, p.tmp, p.meanp, p.cov, rowpart, colpart, p.mm, st));
      ctx->launches += 5;
    } else if (ctx->cur_rec) {
      if (!(stat == RSB_GT && covclass == RSB_C16)) { rsb_set_error(ctx, "internal: the slots hold G-test records, not counts"); return 1; }
      RSB_CUDA_OK(rsb_launch_gt_finish((const double *) p.cnt, (size_t) 16 * ctx->L * ctx->Lp, p.pm, nrep, ctx->L, ctx->Lp, p.cov, rowpart, colpart, p.mm,
                                       ctx->shard_rank, ctx->shard_world, st));
      ctx->launches++;
    } else {
      RSB_CUDA_OK(rsb_launch_statistic(stat, covclass, p.cnt, p.pm, ctx->d_logtab, nrep, ctx->L, ctx->Lp, g.scale, g.wtot, mask, p.cov, rowpart, colpart, p.mm,
                                       ctx->shard_rank, ctx->shard_world, st));
      ctx->launches++;
    }
  }
  if (part & 2) {
    if (phase == 3 && grid_comm(ctx)) {                             // row sums + total of the owned rows, summed over the ranks
      RSB_CUDA_OK(rsb_launch_correct_final(rowpart, colpart, p.mm, nrep, ctx->L, p.covx, p.scal, p.blocksum, p.covsum, 1, st));
      // (entries L+1, L+2 -- the raw min / max of the owned rows -- are summed too and mean nothing afterwards; nobody reads
      // them on this path: the corrected range is reduced separately)
      if (comm_allreduce(ctx, p.covsum, (size_t) nrep * (ctx->L + 4), rsb_nccl::ncclFloat64, rsb_nccl::ncclSum, st)) return 1;
      RSB_CUDA_OK(rsb_launch_correct_final(rowpart, colpart, p.mm, nrep, ctx->L, p.covx, p.scal, p.blocksum, p.covsum, 2, st));
      ctx->launches += 3;
    } else {
    RSB_CUDA_OK(rsb_launch_correct_final(rowpart, colpart, p.mm, nrep, ctx->L, p.covx, p.scal, p.blocksum, p.covsum, phase, st));
    ctx->launches += (phase == 3) ? 3 : (phase == 1 ? 2 : 1);
    }
  }
  return 0;
}

// correction + min/max (+ optional write-back / histogram) for slots [s0, s0+nrep)
int enqueue_correct(rsb_ctx *ctx, int s0, int nrep, int actype, int mode, double bmin, cudaStream_t st)
{
  RSB_RANGE("rsb:correct_hist");
  SlotPtrs p = slot_ptrs(ctx, s0);
  RSB_CUDA_OK(rsb_launch_correct_hist(p.cov, p.covx, p.scal, nrep, ctx->L, ctx->Lp, actype, mode, bmin, ctx->d_w, ctx->d_hist, HIST_BINS,
                                      p.mm, p.minmax, ctx->d_flags, ctx->shard_rank, ctx->shard_world, ctx->d_m2p, ctx->mind, st));
  ctx->launches += 2;
  if (grid_comm(ctx) && comm_reduce_minmax(ctx, p.minmax, nrep, st)) return 1;
  return 0;
}

// pairs i<j whose row this rank owns (all of them without a sharded pair grid), minus those kept out of the histograms
unsigned long long owned_pairs_in_hist(rsb_ctx *ctx)
{
  if (ctx->shard_world <= 1) return ctx->pairs_in_hist;
  unsigned long long mine = 0;
  for (int i = 0; i < ctx->L; i++) if ((i / RSB_ICOLS) % ctx->shard_world == ctx->shard_rank) mine += (unsigned long long) (ctx->L - 1 - i);
  return mine;
}

// serial pipeline on the main stream, slots [0,nrep): counts -> (marginals) -> statistic
int run_pipeline(rsb_ctx *ctx, int nrep, int stat, int covclass, unsigned mask, double tol, int weighted_geo = 0)
{
  const bool raf = (stat == RSB_RAF || stat == RSB_RAFS);
  if (enqueue_counts(ctx, raf ? 1 : weighted_geo, 0, nrep, ctx->d_res, ctx->stream, use_record(ctx, stat, covclass))) return 1;
  if (!raf && enqueue_marginals(ctx, 0, nrep, tol, ctx->stream)) return 1;
  return enqueue_statistic(ctx, 0, nrep, stat, covclass, mask, ctx->stream);
}

int copy_matrix_out(rsb_ctx *ctx, const double *dsrc, double *hdst)   // [L][Lp] device -> [L][L] host
{
  RSB_CUDA_OK(cudaMemcpy2DAsync(hdst, sizeof(double) * ctx->L, dsrc, sizeof(double) * ctx->Lp, sizeof(double) * ctx->L, ctx->L,
                                cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

} // namespace

// =============================================================================================== C-ABI
static int pool_range_ok(rsb_ctx *ctx, int first_rep, int nrep)
{
  if (first_rep < 0 || nrep < 1 || first_rep + nrep > ctx->Rpool) { rsb_set_error(ctx, "pool replicates [%d,%d) out of range (reserved %d)", first_rep, first_rep + nrep, ctx->Rpool); return 1; }
  return 0;
}

// make `st` wait until the generators have finished pool entries [first_rep, first_rep + nrep)
static int pool_wait(rsb_ctx *ctx, int first_rep, int nrep, cudaStream_t st)
{
  cudaEvent_t last = nullptr;
  for (int r = first_rep; r < first_rep + nrep && r < (int) ctx->pool_ready.size(); r++) {
    cudaEvent_t e = ctx->pool_ready[r];
    if (e && e != last) { RSB_CUDA_OK(cudaStreamWaitEvent(st, e, 0)); last = e; }
  }
  return 0;
}

extern "C" {

const char *rsb_create_error(void) { return g_create_err; }
const char *rsb_error(const rsb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int rsb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int rsb_create(int device, void *stream, rsb_ctx **out)
{
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    snprintf(g_create_err, sizeof(g_create_err), "no CUDA device available (%s); librscape_b200 has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 1;
  }
  if (device < 0 || device >= ndev) { snprintf(g_create_err, sizeof(g_create_err), "device %d out of range (%d devices)", device, ndev); return 1; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "%s", cudaGetErrorString(e)); return 1; }
  if (prop.major != 10) {
    snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return 1;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "%s", cudaGetErrorString(e)); return 1; }
  rsb_ctx *c = new rsb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (stream) c->stream = (cudaStream_t) stream;
  else { cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); c->own_stream = true; }
  cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->stream_aux2, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->stream_aux3, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->stream_aux4, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);                     // lo = least, hi = greatest priority
    cudaStreamCreateWithPriority(&c->stream_gen, cudaStreamNonBlocking, lo);
    cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, hi);
    cudaStreamDestroy(c->stream_copy);
    cudaStreamCreateWithPriority(&c->stream_copy, cudaStreamNonBlocking, hi);
  }
  cudaEventCreateWithFlags(&c->ev_exit, cudaEventDisableTiming);
  c->gen_ring.resize(64);
  for (auto &e : c->gen_ring) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming);
  for (int g = 0; g < RSB_GROUPS; g++) {
    cudaEventCreateWithFlags(&c->ev_up[g], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_counts[g], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_stats[g], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_marg[g], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_statk[g], cudaEventDisableTiming);
  }
  if ((e = cudaGetLastError()) != cudaSuccess || !c->stream || !c->stream_aux || !c->stream_aux2 || !c->stream_aux3 || !c->stream_aux4 || !c->stream_copy || !c->stream_gen || !c->stream_hi ||
      !c->ev_exit || !c->ev_entry) {
    snprintf(g_create_err, sizeof(g_create_err), "rsb_create: streams / events: %s", cudaGetErrorString(e));
    delete c; return 1;
  }
  if (cudaMalloc(&c->d_logtab, rsb_logtab_bytes()) != cudaSuccess || rsb_launch_logtab(c->d_logtab, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess) {
    snprintf(g_create_err, sizeof(g_create_err), "rsb_create: log table: %s", cudaGetErrorString(cudaGetLastError()));
    delete c; return 1;
  }
  if (const char *e = getenv("RSCAPE_B200_FUSED_GT")) c->fused_gt = atoi(e) != 0;
  *out = c;
  return 0;
}

void rsb_destroy(rsb_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream_gen);
  for (auto &p : ctx->pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto &p : ctx->pending_aux) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto &p : ctx->pending_stage) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  if (ctx->comm) { cudaDeviceSynchronize(); peer_teardown(ctx); rsb_nccl::api().CommDestroy(ctx->comm); ctx->comm = nullptr; }
  for (int k = 0; k < RSB_PEER_CHANNELS; k++) if (ctx->ev_peer[k]) cudaEventDestroy(ctx->ev_peer[k]);
  free_plan(ctx);
  if (ctx->d_logtab) cudaFree(ctx->d_logtab);
  cudaStreamSynchronize(ctx->stream_aux); cudaStreamSynchronize(ctx->stream_aux2); cudaStreamSynchronize(ctx->stream_aux3); cudaStreamSynchronize(ctx->stream_aux4);
  cudaStreamSynchronize(ctx->stream_copy); cudaStreamSynchronize(ctx->stream_gen);
  cudaStreamDestroy(ctx->stream_aux2); cudaStreamDestroy(ctx->stream_aux3); cudaStreamDestroy(ctx->stream_aux4);
  cudaStreamDestroy(ctx->stream_aux); cudaStreamDestroy(ctx->stream_copy); cudaStreamDestroy(ctx->stream_gen); cudaStreamDestroy(ctx->stream_hi); cudaEventDestroy(ctx->ev_exit);
  for (auto &e : ctx->gen_ring) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev_entry);
  for (int g = 0; g < 2; g++) { cudaEventDestroy(ctx->ev_up[g]); cudaEventDestroy(ctx->ev_counts[g]); cudaEventDestroy(ctx->ev_stats[g]); cudaEventDestroy(ctx->ev_marg[g]); cudaEventDestroy(ctx->ev_statk[g]); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int configure_impl(rsb_ctx *ctx, int nseq, int alen, int max_replicates, int nslices);
int rsb_configure(rsb_ctx *ctx, int nseq, int alen, int max_replicates, int nslices)
{
  const int rc = configure_impl(ctx, nseq, alen, max_replicates, nslices);
  if (rc != 0 && ctx) {                                              // never leave a half-allocated plan that later calls would accept
    cudaGetLastError();
    free_plan(ctx);
    ctx->N = ctx->L = ctx->Rcap = 0;
  }
  return rc;
}

static int configure_impl(rsb_ctx *ctx, int nseq, int alen, int max_replicates, int nslices)
{
  if (nseq < 1 || alen < 1 || max_replicates < 1 || nslices < 0 || nslices > RSB_MAX_SLICES) { rsb_set_error(ctx, "bad configuration"); return 1; }
  if (nslices && ctx->Snull > nslices) { rsb_set_error(ctx, "the nulls are set to %d weight slices, more than the %d of the input alignment", ctx->Snull, nslices); return 1; }
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  free_plan(ctx);
  ctx->N = nseq; ctx->L = alen; ctx->Rcap = max_replicates; ctx->Sreq = nslices;
  ctx->Lp   = (alen + 3) & ~3;
  ctx->Kpad = (nseq + RSB_KSTAGE - 1) / RSB_KSTAGE * RSB_KSTAGE;
  ctx->nIB  = (alen + RSB_ICOLS - 1) / RSB_ICOLS;
  ctx->MA   = ctx->nIB * RSB_MTILE;
  // planeB must hold the widest geometry (any S in 1..6); rows = nJB * NT
  size_t rows_cap = 0; int lcover = ctx->nIB * RSB_ICOLS, njb_cap = 0;
  for (int S = 1; S <= RSB_MAX_SLICES; S++) {
    if (nslices && S > nslices) continue;                          // (fewer slices: the unit-weight tables, the nulls' own weights)
    const int CJ = rsb_cj_for(S), nJB = (alen + CJ - 1) / CJ;
    rows_cap = std::max(rows_cap, (size_t) nJB * 4 * S * CJ);
    njb_cap  = std::max(njb_cap, nJB);
    lcover = std::max(lcover, nJB * CJ);
  }
  ctx->planeB_rows_cap = rows_cap;
  ctx->Lcover = (lcover + 31) & ~31;

  const size_t R = max_replicates, L = alen, Lp = ctx->Lp, N = nseq;
  int nJT, nIT; rsb_stat_grid(alen, &nJT, &nIT);
  RSB_CUDA_OK(cudaMalloc(&ctx->d_res, R * N * L));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_planeA, R * ctx->MA * ctx->Kpad));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_planeB, R * rows_cap * ctx->Kpad));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_cnt, R * 16 * L * Lp * sizeof(long long)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_nseff, R * L * Lp * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_cov, R * L * Lp * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_tmp, std::max(R * L * Lp, R * (size_t) nJT * nIT * 4) * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_pm, R * L * 4 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_rowpart, R * nJT * L * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_colpart, R * nIT * L * sizeof(double)));
  ctx->mrow_stride = (size_t) njb_cap * rsb_gram_mrow_blocks(0) * L * 4; ctx->mcol_stride = (size_t) 4 * ctx->nIB * L * 4;
  RSB_CUDA_OK(cudaMalloc(&ctx->d_mrow, R * ctx->mrow_stride * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_mcol, R * ctx->mcol_stride * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_mm, R * nJT * nIT * 2 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_scal, R * 4 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_covx, R * L * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_minmax, R * 2 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_meanp, R * 4 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_blocksum, R * ((L + 127) / 128) * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_msum, R * L * 4 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_covsum, R * (L + 4) * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_w, sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_hist, sizeof(unsigned long long) * HIST_BINS));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_colsum, L * 5 * sizeof(unsigned long long)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_ps, L * 5 * sizeof(double)));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_flags, sizeof(int)));
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_hist, 0, sizeof(unsigned long long) * HIST_BINS, ctx->stream));
  // counts outside the computed upper-triangle tiles are never read, but keep the buffer defined
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_cnt, 0, R * 16 * L * Lp * sizeof(long long), ctx->stream));
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_cov, 0, R * L * Lp * sizeof(double), ctx->stream));
  ctx->hist_n = 0;
  ctx->pairs_in_hist = (unsigned long long) L * (L - 1) / 2;
  if (make_plane_map(ctx, &ctx->tmA, ctx->d_planeA, ctx->Kpad, (size_t) ctx->MA, ctx->Rcap, RSB_MTILE)) return 1;
  ctx->wgt.assign(nseq, 1.0);
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rsb_set_weights(rsb_ctx *ctx, const double *wgt)
{
  if (ctx->N == 0) { rsb_set_error(ctx, "rsb_configure first"); return 1; }
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (wgt) ctx->wgt.assign(wgt, wgt + ctx->N); else ctx->wgt.assign(ctx->N, 1.0);
  int S = ctx->Sreq;
  if (S == 0) {
    bool small_int = true;
    for (int s = 0; s < ctx->N && small_int; s++) small_int = (ctx->wgt[s] >= 0 && ctx->wgt[s] == std::floor(ctx->wgt[s]) && ctx->wgt[s] <= 255.0);
    S = small_int ? 1 : 4;
  }
  Geo &g = ctx->geo[0];
  if (g.S != S || !g.d_tiles) { if (build_geo(ctx, g, S)) return 1; }
  if (quantise(ctx, g, ctx->wgt, false)) return 1;
  // the null alignments' own, coarser representation of the same weights (rsb_set_null_slices)
  Geo &gn = ctx->geo[2];
  gn.ready = false;
  if (ctx->Snull > 0 && ctx->Snull < S) {
    if (gn.S != ctx->Snull || !gn.d_tiles) { if (build_geo(ctx, gn, ctx->Snull)) return 1; }
    if (quantise(ctx, gn, ctx->wgt, false)) return 1;
  }
  return 0;
}

/* Mixed precision (north_star: "a stated bound (split path)"): the null alignments are contracted with nslices base-256 digits
 * of the weights instead of the input alignment's.  0 = the same weights everywhere (default).  Takes effect at the next
 * rsb_set_weights. */
int rsb_set_null_slices(rsb_ctx *ctx, int nslices)
{
  if (nslices < 0 || nslices > RSB_MAX_SLICES) { rsb_set_error(ctx, "bad slice count %d", nslices); return 1; }
  if (ctx->N && ctx->Sreq && nslices > ctx->Sreq) { rsb_set_error(ctx, "the nulls cannot carry more weight slices (%d) than the input alignment (%d)", nslices, ctx->Sreq); return 1; }
  ctx->Snull = nslices;
  ctx->geo[2].ready = false;
  return 0;
}

int rsb_get_null_quantisation(rsb_ctx *ctx, int64_t *wq, int *q, int *nslices, double *max_abs_err, double *effective_bits)
{
  if (!ctx->geo[0].ready) { rsb_set_error(ctx, "rsb_set_weights has not been called"); return 1; }
  Geo &g = ctx->geo[null_geo(ctx)];
  if (wq) for (int s = 0; s < ctx->N; s++) wq[s] = g.wq[s];
  if (q) *q = g.q;
  if (nslices) *nslices = g.S;
  if (max_abs_err) *max_abs_err = g.qerr_abs;
  if (effective_bits) *effective_bits = (g.qerr_abs > 0.0) ? std::log2(g.maxw / g.qerr_abs) : 64.0;
  return 0;
}

int rsb_get_quantisation(rsb_ctx *ctx, int64_t *wq, int *q, int *nslices)
{
  Geo &g = ctx->geo[0];
  if (!g.ready) { rsb_set_error(ctx, "rsb_set_weights has not been called"); return 1; }
  if (wq) for (int s = 0; s < ctx->N; s++) wq[s] = g.wq[s];
  if (q) *q = g.q;
  if (nslices) *nslices = g.S;
  return 0;
}

int rsb_get_quantisation_error(rsb_ctx *ctx, double *max_abs_err, double *effective_bits)
{
  Geo &g = ctx->geo[0];
  if (!g.ready) { rsb_set_error(ctx, "rsb_set_weights has not been called"); return 1; }
  if (max_abs_err) *max_abs_err = g.qerr_abs;
  if (effective_bits) *effective_bits = (g.qerr_abs > 0.0) ? std::log2(g.maxw / g.qerr_abs) : 64.0;
  return 0;
}

int rsb_fetch_probs(rsb_ctx *ctx, double *pp, double *pm, double *ps, double *nseff, double *ngap)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  Geo &g = ctx->geo[0];
  if (!g.ready || ctx->cur_geo != 0 || ctx->cur_rec) { rsb_set_error(ctx, "no weighted counts resident (call rsb_probs)"); return 1; }
  const size_t L = ctx->L;
  if (pp || nseff || ngap) {
    if (!ctx->d_pp_out) {
      RSB_CUDA_OK(cudaMalloc(&ctx->d_pp_out, L * L * 16 * sizeof(double)));
      RSB_CUDA_OK(cudaMalloc(&ctx->d_nseff_out, L * L * sizeof(double)));
      RSB_CUDA_OK(cudaMalloc(&ctx->d_ngap_out, L * L * sizeof(double)));
    }
    RSB_CUDA_OK(rsb_launch_export_probs(ctx->d_cnt, ctx->L, ctx->Lp, g.scale, g.wtot, ctx->d_pp_out, ctx->d_nseff_out, ctx->d_ngap_out, ctx->stream));
    ctx->launches++;
    if (pp)    RSB_CUDA_OK(cudaMemcpyAsync(pp, ctx->d_pp_out, L * L * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (nseff) RSB_CUDA_OK(cudaMemcpyAsync(nseff, ctx->d_nseff_out, L * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (ngap)  RSB_CUDA_OK(cudaMemcpyAsync(ngap, ctx->d_ngap_out, L * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (ps) {
    RSB_CUDA_OK(rsb_launch_colsum(ctx->d_res, ctx->N, ctx->L, g.d_wq, ctx->d_colsum, ctx->stream));
    RSB_CUDA_OK(rsb_launch_ps(ctx->d_colsum, ctx->L, g.scale, ctx->d_ps, ctx->stream));
    ctx->launches += 2;
    RSB_CUDA_OK(cudaMemcpyAsync(ps, ctx->d_ps, L * 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (pm) RSB_CUDA_OK(cudaMemcpyAsync(pm, ctx->d_pm, L * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rsb_probs(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, double tol,
              double *pp, double *pm, double *ps, double *nseff, double *ngap)
{
  RSB_RANGE("rsb_probs");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ensure_geo(ctx, 0)) return 1;
  if (upload_msa(ctx, msa, row_stride, 0, 1, 0, on_device, ctx->stream)) return 1;
  if (enqueue_counts(ctx, 0, 0, 1, ctx->d_res, ctx->stream)) return 1;
  if (enqueue_marginals(ctx, 0, 1, tol, ctx->stream)) return 1;
  if (pp || pm || ps || nseff || ngap) { if (rsb_fetch_probs(ctx, pp, pm, ps, nseff, ngap)) return 1; }
  return check_flags(ctx, "corr_Probs");
}

/* corr_CalculateCOVCorrected on a host-supplied raw matrix (the reference corrects whatever mi->COV holds,
 * e.g. Potts scores filled on the host): upload, reduce, correct. */
int rsb_correct_host(rsb_ctx *ctx, int actype, double *cov, double *mincov, double *maxcov)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (actype != RSB_APC && actype != RSB_ASC) { rsb_set_error(ctx, "wrong correction type"); return 1; }
  RSB_CUDA_OK(cudaMemcpy2DAsync(ctx->d_cov, sizeof(double) * ctx->Lp, cov, sizeof(double) * ctx->L, sizeof(double) * ctx->L, ctx->L,
                                cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(rsb_launch_reduce_cov(ctx->d_cov, 1, ctx->L, ctx->Lp, ctx->d_rowpart, ctx->d_colpart, ctx->d_mm, ctx->stream));
  RSB_CUDA_OK(rsb_launch_correct_final(ctx->d_rowpart, ctx->d_colpart, ctx->d_mm, 1, ctx->L, ctx->d_covx, ctx->d_scal, ctx->d_blocksum, ctx->d_covsum, 3, ctx->stream));
  ctx->launches += 4;
  return rsb_correct(ctx, actype, cov, mincov, maxcov);
}

int rsb_statistic(rsb_ctx *ctx, int stat, int covclass, const double *allowpair, const uint8_t *msa, int64_t row_stride, int on_device,
                  double *cov, double *mincov, double *maxcov)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  if (stat == RSB_RAF || stat == RSB_RAFS) {
    if (!msa) { rsb_set_error(ctx, "RAF/RAFS need the alignment"); return 1; }
    if (upload_msa(ctx, msa, row_stride, 0, 1, 0, on_device, ctx->stream)) return 1;
    if (enqueue_counts(ctx, 1, 0, 1, ctx->d_res, ctx->stream)) return 1;
  } else if (ctx->cur_geo != 0 || ctx->cur_rec) { rsb_set_error(ctx, "rsb_probs must precede this statistic"); return 1; }
  if (enqueue_statistic(ctx, 0, 1, stat, covclass, allow_mask(allowpair), ctx->stream)) return 1;
  double sc[4];
  if (cov) {
    RSB_CUDA_OK(rsb_launch_symmetrize(ctx->d_cov, ctx->L, ctx->Lp, ctx->stream));
    ctx->launches++;
    if (copy_matrix_out(ctx, ctx->d_cov, cov)) return 1;
  }
  RSB_CUDA_OK(cudaMemcpyAsync(sc, ctx->d_scal, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (mincov) *mincov = sc[1];
  if (maxcov) *maxcov = sc[2];
  return 0;
}

int rsb_correct(rsb_ctx *ctx, int actype, double *cov, double *mincov, double *maxcov)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (actype != RSB_APC && actype != RSB_ASC) { rsb_set_error(ctx, "wrong correction type"); return 1; }
  if (enqueue_correct(ctx, 0, 1, actype, 1, 0.0, ctx->stream)) return 1;
  RSB_CUDA_OK(rsb_launch_symmetrize(ctx->d_cov, ctx->L, ctx->Lp, ctx->stream));
  ctx->launches++;
  double mmx[2];
  if (cov && copy_matrix_out(ctx, ctx->d_cov, cov)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(mmx, ctx->d_minmax, sizeof(mmx), cudaMemcpyDeviceToHost, ctx->stream));
  if (check_flags(ctx, "corr_CalculateCOVCorrected")) return 1;
  if (mincov) *mincov = mmx[0];
  if (maxcov) *maxcov = mmx[1];
  return 0;
}

int rsb_scan(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, int stat, int covclass, int actype,
             const double *allowpair, double tol, double *cov, double *mincov, double *maxcov,
             double *pp, double *pm, double *ps, double *nseff, double *ngap)
{
  RSB_RANGE("rsb_scan");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  const bool raf = (stat == RSB_RAF || stat == RSB_RAFS);
  if (!raf) { if (rsb_probs(ctx, msa, row_stride, on_device, tol, pp, pm, ps, nseff, ngap)) return 1; }
  else {
    const size_t L = ctx->L;            // corr_Probs is skipped for RAF*: the probability fields stay zero (covariation.c:82-84)
    if (pp) memset(pp, 0, L * L * 16 * sizeof(double));
    if (pm) memset(pm, 0, L * 4 * sizeof(double));
    if (ps) memset(ps, 0, L * 5 * sizeof(double));
    if (nseff) memset(nseff, 0, L * L * sizeof(double));
    if (ngap)  memset(ngap, 0, L * L * sizeof(double));
  }
  const bool corr = (actype == RSB_APC || actype == RSB_ASC);
  if (rsb_statistic(ctx, stat, covclass, allowpair, msa, row_stride, on_device, corr ? nullptr : cov, mincov, maxcov)) return 1;
  if (corr) return rsb_correct(ctx, actype, cov, mincov, maxcov);
  return 0;
}

int rsb_get_counts(rsb_ctx *ctx, int64_t *counts)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->cur_rec) { rsb_set_error(ctx, "no counts resident: the last scan was a null scored through the record epilogue"); return 1; }
  const size_t L = ctx->L;
  RSB_CUDA_OK(cudaMemcpy2DAsync(counts, sizeof(int64_t) * L, ctx->d_cnt, sizeof(int64_t) * ctx->Lp, sizeof(int64_t) * L, 16 * L,
                                cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // only the upper triangle is defined: clear the rest so that comparisons are well defined
  for (size_t p = 0; p < 16; p++)
    for (size_t i = 0; i < L; i++)
      for (size_t j = 0; j <= i && j < L; j++) counts[(p * L + i) * L + j] = 0;
  return 0;
}

int rsb_get_counts_direct(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int64_t *counts)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  Geo &g = ctx->geo[ctx->cur_geo];
  if (!g.ready) { rsb_set_error(ctx, "no weights set"); return 1; }
  if (upload_msa(ctx, msa, row_stride, 0, 1, 0, 0, ctx->stream)) return 1;
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_cnt, 0, (size_t) 16 * ctx->L * ctx->Lp * sizeof(long long), ctx->stream));
  RSB_CUDA_OK(rsb_launch_counts_direct(ctx->d_res, ctx->N, ctx->L, ctx->Lp, g.d_wq, ctx->d_cnt, ctx->stream));
  ctx->launches++;
  return rsb_get_counts(ctx, counts);
}

// ---------------------------------------------------------------------------------------------- nulls
int rsb_null_width(rsb_ctx *ctx, const uint8_t *null0, int64_t row_stride, int on_device, int stat, int covclass, int actype,
                   const double *allowpair, double tol, double w_old, double bmin, int hpts, double *w_out, double *mincov, double *maxcov)
{
  RSB_RANGE("rsb_null_width");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  if (ctx->shard_world > 1 && !grid_comm(ctx)) { rsb_set_error(ctx, "calculate_width_histo on a sharded pair grid needs a communicator (rsb_comm_init)"); return 1; }
  if (null0 && upload_msa(ctx, null0, row_stride, 0, 1, 0, on_device, ctx->stream)) return 1;
  if (run_pipeline(ctx, 1, stat, covclass, allow_mask(allowpair), tol, null_geo(ctx))) return 1;     // a null: scored as the nulls are
  if (enqueue_correct(ctx, 0, 1, actype, 0, bmin, ctx->stream)) return 1;
  RSB_CUDA_OK(rsb_launch_width(ctx->d_minmax, w_old, bmin, hpts, tol, ctx->d_w, ctx->stream));
  ctx->launches++;
  double mmx[2], w;
  RSB_CUDA_OK(cudaMemcpyAsync(mmx, ctx->d_minmax, sizeof(mmx), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(&w, ctx->d_w, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream));
  if (check_flags(ctx, "calculate_width_histo")) return 1;
  if (!(mmx[1] > bmin)) { rsb_set_error(ctx, "bmin %f should be larger than maxCOV %f", bmin, mmx[1]); return 1; }   // R-scape.c:1355
  if (w_out) *w_out = w;
  if (mincov) *mincov = mmx[0];
  if (maxcov) *maxcov = mmx[1];
  return 0;
}

// The null loop, software-pipelined over up to four groups of replicate slots (chunk c lives in group c mod G):
//   copy stream      : alignments of chunk c+1 -> slots (host buffers only), then pack into operand planes   (HBM)
//   main stream      : tcgen05 gram of chunk c                    (tensor pipe)
//   aux stream c%G   : marginals, statistic, correction, histogram of chunk c   (latency-bound small kernels beside the contractions)
// so the statistics chain of one chunk hides under the tensor-core contractions of the next G - 1.  Events order the
// reuse of each slot group; the caller's stream (main) waits for everything before the call returns.
// src_dev: device-resident nulls [nrep][N][L] read in place (no copy), else host/strided input that is uploaded.
static int null_hist_pipelined(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device,
                               int stat, int covclass, int actype, unsigned mask, double tol, double w, double bmin,
                               double *minmax, int pool_first = -1)
{
  RSB_RANGE("rsb:null_loop");
  const bool raf = (stat == RSB_RAF || stat == RSB_RAFS);
  const int  wgeo = raf ? 1 : null_geo(ctx);
  // GT x C16: the contraction's epilogue leaves per-pair records and the statistic finishes with an HBM-bound kernel on the
  // aux stream -- nothing but contractions on the main stream
  const bool rec = !raf && use_record(ctx, stat, covclass);
  if (ctx->shard_world > 1 && !grid_comm(ctx)) {
    rsb_set_error(ctx, "the null loop on a sharded pair grid needs a communicator over the %d shards (rsb_comm_init); "
                       "without one use rsb_sharded_counts / _statistic / _correct and reduce on the host", ctx->shard_world);
    return 1;
  }
  if (ctx->shard_world > 1 && ctx->d_m2p) { rsb_set_error(ctx, "a pair exclusion (rsb_set_pair_exclusion) is not offered on a sharded pair grid"); return 1; }
  const bool in_place = (on_device && row_stride == ctx->L && rep_stride == (int64_t) ctx->N * ctx->L);
  // Slot groups: chunk c is contracted in group c % G, and the contraction of chunk c + G waits for the statistics chain of chunk c.
  // The chain runs beside the contractions at a few blocks per SM, bound by latency (0.3 - 0.5 ms at the SSU shape whatever the share
  // of the grid a rank owns): with two groups the loop's period is max(contraction, chain); with three or four groups (slots that
  // are a multiple of 3 or 4) a chain has G - 1 contractions to finish in and consecutive chains overlap on their own streams.
  if (grid_comm(ctx) && ctx->peer_state == 0 && peer_setup_ipc(ctx, (size_t) 4 * ctx->L * std::max(1, ctx->Rcap) + 64)) return 1;   // (collective)
  static const int gmax = getenv("RSCAPE_B200_GROUPS") ? std::min(RSB_GROUPS, std::max(1, atoi(getenv("RSCAPE_B200_GROUPS")))) : RSB_GROUPS;
  const int  G = std::min(gmax, (ctx->Rcap % 4 == 0 && ctx->Rcap >= 4) ? 4 : (ctx->Rcap % 3 == 0 && ctx->Rcap >= 3) ? 3 : (ctx->Rcap >= 2) ? 2 : 1);
  const int  chunk = std::max(1, ctx->Rcap / G);
  const size_t repbytes = (size_t) ctx->N * ctx->L;
  static const int serial = getenv("RSCAPE_B200_SERIAL") ? atoi(getenv("RSCAPE_B200_SERIAL")) : 0;      // experiments: 1 = statistics on the main stream, 2 = pack too
  // contraction + statistic run on an internal high-priority stream, so that their blocks are placed ahead of those of the
  // null generators working on later replicates (generation stream, lowest priority)
  cudaStream_t sm = ctx->stream_hi;
  cudaStream_t st_copy = (serial & 2) ? sm : ctx->stream_copy;
  // one statistics stream per slot group: the chains of consecutive chunks overlap each other (they run beside the
  // contraction at low occupancy, bound by latency rather than by a pipe)
  // (sharded grid: the one-shot all-reduce kernel has one ordered channel per statistics stream, so the chains keep their own
  // streams; if the ranks could not map each other's memory everything that talks to NCCL stays on ONE stream -- same order of
  // collectives on every rank)
  const bool one_aux = grid_comm(ctx) && !(ctx->peer_state == 1 && ctx->peer_view.cap >= (size_t) 4 * ctx->L * chunk);   // (a block sized for an earlier, smaller plan: NCCL)
  cudaStream_t aux_streams[RSB_GROUPS] = { ctx->stream_aux, ctx->stream_aux2, ctx->stream_aux3, ctx->stream_aux4 };
  auto aux_of = [&](int g) -> cudaStream_t { return (serial & 1) ? sm : one_aux ? ctx->stream_aux : aux_streams[g % RSB_GROUPS]; };

  if (minmax && ctx->h_mm_cap < (size_t) nrep) {
    if (ctx->h_mm) cudaFreeHost(ctx->h_mm);
    ctx->h_mm = nullptr; ctx->h_mm_cap = 0;
    RSB_CUDA_OK(cudaMallocHost(&ctx->h_mm, sizeof(double) * 2 * (size_t) nrep));
    ctx->h_mm_cap = (size_t) nrep;
  }
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));
  RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_entry, 0));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_w, &w, sizeof(double), cudaMemcpyHostToDevice, sm));
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, sm));
  for (int g = 0; g < RSB_GROUPS; g++) RSB_CUDA_OK(cudaStreamWaitEvent(aux_streams[g], ctx->ev_entry, 0));
  RSB_CUDA_OK(cudaStreamWaitEvent(st_copy, ctx->ev_entry, 0));
  bool used[RSB_GROUPS] = { false, false, false, false };

  // Stream plan.  The statistic kernel is FP64-bound, and FP64 issue collapses (~6x, measured) while the tensor pipe is
  // busy, so it runs ALONE on the main stream between two contractions; everything light (operand packing on the copy
  // stream; marginal sums, reductions, correction + histogram on the aux stream) runs beside the contraction.
  //   main:  G(0) G(1) S(0) G(2) S(1) ...      aux:  M(c) after G(c);  R(c) + C(c) after S(c)
  auto tail = [&](int pc, int pr0) -> int {                         // S(pc) on main, then its reductions/correction on aux
    const int pg = pc % G, ps0 = pg * chunk, pn = std::min(chunk, nrep - pr0);
    cudaStream_t st_aux = aux_of(pg);
    cudaEvent_t a0 = nullptr, a1 = nullptr, am = nullptr, as = nullptr;
    if (ctx->profile) { cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&am); cudaEventCreate(&as); }
    if (rec) {                                                       // the whole chain on the aux stream, behind M(pc)
      if (ctx->profile) cudaEventRecord(a0, st_aux);
      if (enqueue_statistic(ctx, ps0, pn, stat, covclass, mask, st_aux, 3, 1)) return 1;
      if (ctx->profile) { cudaEventRecord(am, st_aux); cudaEventRecord(as, st_aux); }
    } else {
    RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_marg[pg], 0));
    if (pc >= G) RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_stats[pg], 0));   // the group's scores and partial sums have been consumed
    if (ctx->profile) cudaEventRecord(a0, sm);
    if (enqueue_statistic(ctx, ps0, pn, stat, covclass, mask, sm, 3, 1)) return 1;
    if (ctx->profile) cudaEventRecord(am, sm);
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_statk[pg], sm));
    RSB_CUDA_OK(cudaStreamWaitEvent(st_aux, ctx->ev_statk[pg], 0));
    if (ctx->profile) cudaEventRecord(as, st_aux);
    }
    if (enqueue_statistic(ctx, ps0, pn, stat, covclass, mask, st_aux, 3, 2)) return 1;
    if (enqueue_correct(ctx, ps0, pn, actype, 2, bmin, st_aux)) return 1;
    if (minmax) RSB_CUDA_OK(cudaMemcpyAsync(ctx->h_mm + 2 * (size_t) pr0, ctx->d_minmax + 2 * (size_t) ps0, sizeof(double) * 2 * pn,
                                            cudaMemcpyDeviceToHost, st_aux));
    if (ctx->profile) { cudaEventRecord(a1, st_aux); ctx->pending_aux.push_back({ a0, a1 }); ctx->pending_stage.push_back({ am, as }); ctx->aux_chains++; }
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_stats[pg], st_aux));
    if (w > 0.0) ctx->hist_n += (unsigned long long) pn * owned_pairs_in_hist(ctx);
    return 0;
  };

  int c = 0;
  for (int r0 = 0; r0 < nrep; r0 += chunk, c++) {
    const int g = c % G, s0 = g * chunk, n = std::min(chunk, nrep - r0);
    const uint8_t *src;
    if (used[g]) RSB_CUDA_OK(cudaStreamWaitEvent(st_copy, ctx->ev_counts[g], 0));       // the group's slots and planes have been consumed
    if (pool_first >= 0 && pool_wait(ctx, pool_first + r0, n, st_copy)) return 1;    // entries still being generated
    if (in_place) src = nulls + (size_t) r0 * repbytes;
    else {
      if (upload_msa(ctx, nulls + (size_t) r0 * rep_stride, row_stride, rep_stride, n, s0, on_device, st_copy)) return 1;
      src = ctx->d_res + (size_t) s0 * repbytes;
    }
    if (enqueue_pack(ctx, wgeo, s0, n, src, st_copy)) return 1;                   // HBM-bound transpose: hides under the previous gram
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_up[g], st_copy));
    RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_up[g], 0));
    // (the counts and marginal partials of this group were consumed by S(c - G), earlier on this stream; with the record
    // epilogue their consumers run on the aux stream)
    if (rec && used[g]) RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_stats[g], 0));
    if (enqueue_gram(ctx, wgeo, s0, n, sm, rec)) return 1;
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_counts[g], sm));

    RSB_CUDA_OK(cudaStreamWaitEvent(aux_of(g), ctx->ev_counts[g], 0));
    if (!raf && enqueue_marginals(ctx, s0, n, tol, aux_of(g))) return 1;
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_marg[g], aux_of(g)));
    used[g] = true;
    if (G == 1 || rec) { if (tail(c, r0)) return 1; }
    else if (c >= 1) { if (tail(c - 1, r0 - chunk)) return 1; }
  }
  if (G > 1 && c >= 1 && !rec) { if (tail(c - 1, (c - 1) * chunk)) return 1; }
  for (int g = 0; g < G; g++) if (used[g]) RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_stats[g], 0));
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_exit, sm));                      // back to the caller's stream
  RSB_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_exit, 0));
  return 0;
}


// ---------------------------------------------------------------------------------------------- several statistics per contraction
// BASELINE config 5 (statistic sweep): cov_Calculate (src/covariation.c:100-258) runs corr_Probs once per alignment and then ONE
// corr_Calculate*; a sweep over statistics and corrections repeats the whole scan.  Here a null is contracted once, every
// requested statistic is evaluated from the same count planes in one pass (multi_stat_kernel), and each (statistic, correction)
// combination gets its own reductions, correction and histogram on the aux stream, beside the next contraction.
static int stat_slot(int stat)
{
  switch (stat) { case RSB_CHI: return 0; case RSB_OMES: return 1; case RSB_GT: return 2; case RSB_MI: return 3; case RSB_MIr: return 4; case RSB_MIg: return 5; }
  return -1;
}
// the unweighted statistics share a contraction of their own (unit weights, one digit slice): matrix slots 0 = RAF, 1 = RAFS
static int unit_stat_slot(int stat) { return stat == RSB_RAF ? 0 : stat == RSB_RAFS ? 1 : -1; }

static int multi_reserve(rsb_ctx *ctx, int ncombo)
{
  const size_t L = ctx->L, Lp = ctx->Lp;
  int nJT, nIT; rsb_stat_grid(ctx->L, &nJT, &nIT);
  const size_t V = 6 * (size_t) ctx->Rcap, tiles = (size_t) nJT * nIT;       // virtual replicates (slot, statistic)
  if (!ctx->d_covm) {
    RSB_CUDA_OK(cudaMalloc(&ctx->d_covm, sizeof(double) * V * L * Lp));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_rowpart_m, sizeof(double) * V * nJT * L));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_colpart_m, sizeof(double) * V * nIT * L));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_mmr_m, sizeof(double) * V * tiles * 2));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_covx_m, sizeof(double) * V * L));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_scal_m, sizeof(double) * V * 4));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_blocksum_m, sizeof(double) * V * ((L + 127) / 128)));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_covsum_m, sizeof(double) * V * (L + 4)));
  }
  if (ncombo > ctx->multi_cap) {
    dfree(ctx->d_wm); dfree(ctx->d_minmaxm); dfree(ctx->d_histm); dfree(ctx->d_mmc_m);
    RSB_CUDA_OK(cudaMalloc(&ctx->d_mmc_m, sizeof(double) * (size_t) ncombo * ctx->Rcap * tiles * 2));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_wm, sizeof(double) * ncombo));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_minmaxm, sizeof(double) * 2 * (size_t) ncombo * ctx->Rcap));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_histm, sizeof(unsigned long long) * (size_t) ncombo * HIST_BINS));
    RSB_CUDA_OK(cudaMemsetAsync(ctx->d_histm, 0, sizeof(unsigned long long) * (size_t) ncombo * HIST_BINS, ctx->stream));
    ctx->multi_cap = ncombo;
    ctx->hist_n_m.assign(ncombo, 0);
  }
  return 0;
}

static int null_hist_multi(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device, int pool_first,
                           int ncombo, const int *stat, const int *actype, int covclass, unsigned mask, double tol, const double *w, double bmin,
                           double *minmax)
{
  RSB_RANGE("rsb:null_loop_multi");
  if (ncombo < 1 || ncombo > 64) { rsb_set_error(ctx, "bad number of (statistic, correction) combinations %d", ncombo); return 1; }
  if (ctx->shard_world > 1) { rsb_set_error(ctx, "rsb_null_hist_multi is not offered on a sharded pair grid"); return 1; }
  if (covclass != RSB_C16 && covclass != RSB_C2 && covclass != RSB_CWC) { rsb_set_error(ctx, "covclass must be resolved by the caller"); return 1; }
  bool want[6] = { false, false, false, false, false, false };
  const bool unit = unit_stat_slot(stat[0]) >= 0;                    // RAF / RAFS combinations: the unit-weight contraction, shared by their corrections
  auto slot_of = [&](int st) { return unit ? unit_stat_slot(st) : stat_slot(st); };
  for (int k = 0; k < ncombo; k++) {
    const int sl = slot_of(stat[k]);
    if (sl < 0) { rsb_set_error(ctx, "statistic %d cannot share a contraction with statistic %d (weighted: CHI OMES GT MI MIr MIg; unit weights: RAF RAFS; "
                                     "CCF: rsb_null_hist)", stat[k], stat[0]); return 1; }
    if (resolve_stat(ctx, stat[k], covclass)) return 1;
    if (actype[k] != RSB_APC && actype[k] != RSB_ASC && actype[k] != RSB_NOCORR) { rsb_set_error(ctx, "wrong correction type %d", actype[k]); return 1; }
    want[sl] = true;
  }
  if (multi_reserve(ctx, ncombo)) return 1;
  const int  wgeo = unit ? 1 : null_geo(ctx);
  if (ensure_geo(ctx, wgeo)) return 1;
  Geo &g = ctx->geo[wgeo];
  const bool in_place = (on_device && row_stride == ctx->L && rep_stride == (int64_t) ctx->N * ctx->L);
  // slot groups as in null_hist_pipelined: the reductions + correction/histogram chain of a chunk (one launch each for all its
  // statistics and combinations) runs beside the following contractions and has G - 1 of them to finish in
  static const int gmax = getenv("RSCAPE_B200_GROUPS") ? std::min(RSB_GROUPS, std::max(1, atoi(getenv("RSCAPE_B200_GROUPS")))) : RSB_GROUPS;
  const int  G = std::min(gmax, (ctx->Rcap % 4 == 0 && ctx->Rcap >= 4) ? 4 : (ctx->Rcap % 3 == 0 && ctx->Rcap >= 3) ? 3 : (ctx->Rcap >= 2) ? 2 : 1);
  const int  chunk = std::max(1, ctx->Rcap / G);
  const size_t repbytes = (size_t) ctx->N * ctx->L, LL = (size_t) ctx->L * ctx->Lp;
  cudaStream_t sm = ctx->stream_hi, st_copy = ctx->stream_copy;
  cudaStream_t aux_streams[RSB_GROUPS] = { ctx->stream_aux, ctx->stream_aux2, ctx->stream_aux3, ctx->stream_aux4 };
  auto aux_of = [&](int gi) -> cudaStream_t { return aux_streams[gi % RSB_GROUPS]; };
  const size_t need = (size_t) nrep * ncombo;
  if (ctx->h_mm_cap < need) {
    if (ctx->h_mm) cudaFreeHost(ctx->h_mm);
    ctx->h_mm = nullptr; ctx->h_mm_cap = 0;
    RSB_CUDA_OK(cudaMallocHost(&ctx->h_mm, sizeof(double) * 2 * need));
    ctx->h_mm_cap = need;
  }
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));
  RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_entry, 0));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_wm, w, sizeof(double) * ncombo, cudaMemcpyHostToDevice, sm));
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, sm));
  for (int g = 0; g < RSB_GROUPS; g++) RSB_CUDA_OK(cudaStreamWaitEvent(aux_streams[g], ctx->ev_entry, 0));
  RSB_CUDA_OK(cudaStreamWaitEvent(st_copy, ctx->ev_entry, 0));
  bool used[RSB_GROUPS] = { false, false, false, false };
  int nJT, nIT; rsb_stat_grid(ctx->L, &nJT, &nIT);

  // S(pc): all statistics of chunk pc in one pass on the main stream (FP64: alone between two contractions), then on the aux
  // stream ONE reduction over the stacked matrices (virtual replicate q = slot * nw + index of the statistic), one set of row-mean
  // kernels, and ONE launch that corrects and histograms every (replicate, combination)
  int nw = 0, kidx_of[6], kidx[64], act[64];
  for (int sl = 0; sl < 6; sl++) kidx_of[sl] = want[sl] ? nw++ : -1;
  for (int k = 0; k < ncombo; k++) { kidx[k] = kidx_of[slot_of(stat[k])]; act[k] = actype[k]; }
  const size_t tiles = (size_t) nJT * nIT, nblk = (size_t) (ctx->L + 127) / 128;
  auto tail = [&](int pc, int pr0) -> int {
    const int pg = pc % G, ps0 = pg * chunk, pn = std::min(chunk, nrep - pr0);
    cudaStream_t st_aux = aux_of(pg);
    SlotPtrs p = slot_ptrs(ctx, ps0);
    const size_t q0 = (size_t) ps0 * nw;
    double *covq = ctx->d_covm + q0 * LL;
    double *cov6[6];
    for (int sl = 0; sl < 6; sl++) cov6[sl] = want[sl] ? covq + (size_t) kidx_of[sl] * LL : nullptr;
    RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_marg[pg], 0));
    if (pc >= G) RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_stats[pg], 0));   // the group's matrices have been consumed
    if (unit) {                                                      // RAF from the count-table identity, RAFS = its 3-point stencil (stats.cu)
      double *rp = ctx->d_rowpart + (size_t) ps0 * nJT * ctx->L, *cp = ctx->d_colpart + (size_t) ps0 * nIT * ctx->L;
      for (int r = 0; r < pn; r++) {                                 // (the RAF kernels index replicates L * Lp apart: one replicate per launch)
        SlotPtrs pr = slot_ptrs(ctx, ps0 + r);
        if (want[0]) { RSB_CUDA_OK(rsb_launch_raf(pr.cnt, 1, ctx->L, ctx->Lp, ctx->N, mask, 0, pr.tmp, cov6[0] + (size_t) r * nw * LL, rp, cp, pr.mm, sm)); ctx->launches += 2; }
        if (want[1]) { RSB_CUDA_OK(rsb_launch_raf(pr.cnt, 1, ctx->L, ctx->Lp, ctx->N, mask, 1, pr.tmp, cov6[1] + (size_t) r * nw * LL, rp, cp, pr.mm, sm)); ctx->launches += 3; }
      }
    } else {
      RSB_CUDA_OK(rsb_launch_multi_statistic(covclass, p.cnt, p.pm, ctx->d_logtab, pn, ctx->L, ctx->Lp, g.scale, g.wtot, mask, cov6, (size_t) nw * LL, 0, 1, sm));
      ctx->launches++;
    }
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_statk[pg], sm));
    RSB_CUDA_OK(cudaStreamWaitEvent(st_aux, ctx->ev_statk[pg], 0));
    double *rowpart = ctx->d_rowpart_m + q0 * nJT * ctx->L, *colpart = ctx->d_colpart_m + q0 * nIT * ctx->L, *mmr = ctx->d_mmr_m + q0 * tiles * 2;
    double *covx = ctx->d_covx_m + q0 * ctx->L, *scal = ctx->d_scal_m + q0 * 4;
    RSB_CUDA_OK(rsb_launch_reduce_cov(covq, pn * nw, ctx->L, ctx->Lp, rowpart, colpart, mmr, st_aux));
    RSB_CUDA_OK(rsb_launch_correct_final(rowpart, colpart, mmr, pn * nw, ctx->L, covx, scal, ctx->d_blocksum_m + q0 * nblk, ctx->d_covsum_m + q0 * (ctx->L + 4), 3, st_aux));
    double *mmk = ctx->d_minmaxm + (size_t) ps0 * ncombo * 2;
    RSB_CUDA_OK(rsb_launch_correct_hist_multi(covq, covx, scal, pn, ctx->L, ctx->Lp, ncombo, nw, kidx, act, bmin, ctx->d_wm, ctx->d_histm, HIST_BINS,
                                              ctx->d_mmc_m + (size_t) ps0 * ncombo * tiles * 2, mmk, ctx->d_flags, ctx->d_m2p, ctx->mind, st_aux));
    ctx->launches += 6;
    // (replicate-major here; transposed to [combination][replicate] on the host at the end)
    if (minmax) RSB_CUDA_OK(cudaMemcpyAsync(ctx->h_mm + 2 * (size_t) pr0 * ncombo, mmk, sizeof(double) * 2 * (size_t) pn * ncombo, cudaMemcpyDeviceToHost, st_aux));
    for (int k = 0; k < ncombo; k++) if (w[k] > 0.0) ctx->hist_n_m[k] += (unsigned long long) pn * ctx->pairs_in_hist;
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_stats[pg], st_aux));
    return 0;
  };

  int c = 0;
  for (int r0 = 0; r0 < nrep; r0 += chunk, c++) {
    const int gi = c % G, s0 = gi * chunk, n = std::min(chunk, nrep - r0);
    const uint8_t *src;
    if (used[gi]) RSB_CUDA_OK(cudaStreamWaitEvent(st_copy, ctx->ev_counts[gi], 0));
    if (pool_first >= 0 && pool_wait(ctx, pool_first + r0, n, st_copy)) return 1;
    if (in_place) src = nulls + (size_t) r0 * repbytes;
    else {
      if (upload_msa(ctx, nulls + (size_t) r0 * rep_stride, row_stride, rep_stride, n, s0, on_device, st_copy)) return 1;
      src = ctx->d_res + (size_t) s0 * repbytes;
    }
    if (enqueue_pack(ctx, wgeo, s0, n, src, st_copy)) return 1;
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_up[gi], st_copy));
    RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_up[gi], 0));
    if (enqueue_gram(ctx, wgeo, s0, n, sm, false)) return 1;           // (its counts were consumed by S(c - G), earlier on this stream)
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_counts[gi], sm));
    RSB_CUDA_OK(cudaStreamWaitEvent(aux_of(gi), ctx->ev_counts[gi], 0));
    if (!unit && enqueue_marginals(ctx, s0, n, tol, aux_of(gi))) return 1;
    RSB_CUDA_OK(cudaEventRecord(ctx->ev_marg[gi], aux_of(gi)));
    used[gi] = true;
    if (G == 1) { if (tail(c, r0)) return 1; }
    else if (c >= 1) { if (tail(c - 1, r0 - chunk)) return 1; }
  }
  if (G > 1 && c >= 1) { if (tail(c - 1, (c - 1) * chunk)) return 1; }
  for (int gi = 0; gi < G; gi++) if (used[gi]) RSB_CUDA_OK(cudaStreamWaitEvent(sm, ctx->ev_stats[gi], 0));
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_exit, sm));
  RSB_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_exit, 0));
  if (check_flags(ctx, "null_rscape")) return 1;
  if (minmax)
    for (int r = 0; r < nrep; r++)
      for (int k = 0; k < ncombo; k++) {
        minmax[2 * ((size_t) k * nrep + r)]     = ctx->h_mm[2 * ((size_t) r * ncombo + k)];
        minmax[2 * ((size_t) k * nrep + r) + 1] = ctx->h_mm[2 * ((size_t) r * ncombo + k) + 1];
      }
  return 0;
}

int rsb_null_hist_multi(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device, int ncombo,
                        const int *stat, const int *actype, int covclass, const double *allowpair, double tol, const double *w, double bmin,
                        double *minmax)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  return null_hist_multi(ctx, nulls, nrep, row_stride, rep_stride, on_device, -1, ncombo, stat, actype, covclass, allow_mask(allowpair), tol, w, bmin, minmax);
}

int rsb_null_hist_multi_pool(rsb_ctx *ctx, int first_rep, int nrep, int ncombo, const int *stat, const int *actype, int covclass,
                             const double *allowpair, double tol, const double *w, double bmin, double *minmax)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const size_t rb = (size_t) ctx->N * ctx->L;
  return null_hist_multi(ctx, ctx->d_pool + (size_t) first_rep * rb, nrep, ctx->L, (int64_t) rb, 1, first_rep, ncombo, stat, actype, covclass,
                         allow_mask(allowpair), tol, w, bmin, minmax);
}

int rsb_hist_read_multi(rsb_ctx *ctx, int combo, uint64_t *bins, int nb_cap, uint64_t *n_out, int *imax_out)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (combo < 0 || combo >= ctx->multi_cap) { rsb_set_error(ctx, "combination %d out of range (%d histograms)", combo, ctx->multi_cap); return 1; }
  const int nb = std::min(nb_cap, HIST_BINS);
  RSB_CUDA_OK(cudaMemcpyAsync(bins, ctx->d_histm + (size_t) combo * HIST_BINS, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (n_out) *n_out = ctx->hist_n_m[combo];
  if (imax_out) { int im = -1; for (int b = nb - 1; b >= 0; b--) if (bins[b]) { im = b; break; } *imax_out = im; }
  return 0;
}

int rsb_hist_reset_multi(rsb_ctx *ctx)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->d_histm) RSB_CUDA_OK(cudaMemsetAsync(ctx->d_histm, 0, sizeof(unsigned long long) * (size_t) ctx->multi_cap * HIST_BINS, ctx->stream));
  std::fill(ctx->hist_n_m.begin(), ctx->hist_n_m.end(), 0ULL);
  return 0;
}

int rsb_null_hist(rsb_ctx *ctx, const uint8_t *nulls, int nrep, int64_t row_stride, int64_t rep_stride, int on_device, int stat, int covclass,
                  int actype, const double *allowpair, double tol, double w, double bmin, double *minmax)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  if (null_hist_pipelined(ctx, nulls, nrep, row_stride, rep_stride, on_device, stat, covclass, actype, allow_mask(allowpair),
                          tol, w, bmin, minmax)) return 1;
  if (check_flags(ctx, "null_rscape")) return 1;
  if (minmax) memcpy(minmax, ctx->h_mm, sizeof(double) * 2 * (size_t) nrep);
  return 0;
}

int rsb_null_hist_pool(rsb_ctx *ctx, int first_rep, int nrep, int stat, int covclass, int actype, const double *allowpair,
                       double tol, double w, double bmin, double *minmax)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  if (resolve_stat(ctx, stat, covclass)) return 1;
  const size_t rb = (size_t) ctx->N * ctx->L;
  if (null_hist_pipelined(ctx, ctx->d_pool + (size_t) first_rep * rb, nrep, ctx->L, (int64_t) rb, 1, stat, covclass, actype,
                          allow_mask(allowpair), tol, w, bmin, minmax, first_rep)) return 1;
  if (check_flags(ctx, "null_rscape")) return 1;
  if (minmax) memcpy(minmax, ctx->h_mm, sizeof(double) * 2 * (size_t) nrep);
  return 0;
}

int rsb_null_width_pool(rsb_ctx *ctx, int rep, int stat, int covclass, int actype, const double *allowpair, double tol,
                        double w_old, double bmin, int hpts, double *w_out, double *mincov, double *maxcov)
{
  if (pool_range_ok(ctx, rep, 1)) return 1;
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_wait(ctx, rep, 1, ctx->stream)) return 1;
  return rsb_null_width(ctx, ctx->d_pool + (size_t) rep * ctx->N * ctx->L, ctx->L, 1, stat, covclass, actype, allowpair, tol,
                        w_old, bmin, hpts, w_out, mincov, maxcov);
}

// ---------------------------------------------------------------------------------------------- sharded pair grid
int rsb_set_shard(rsb_ctx *ctx, int rank, int world)
{
  if (world < 1 || rank < 0 || rank >= world) { rsb_set_error(ctx, "bad shard %d of %d", rank, world); return 1; }
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->shard_rank = rank; ctx->shard_world = world;
  free_geo(ctx->geo[0]); free_geo(ctx->geo[1]); free_geo(ctx->geo[2]);          // tile lists depend on the shard: rsb_set_weights rebuilds them
  ctx->geo[0].S = 0; ctx->geo[1].S = 0; ctx->geo[2].S = 0;
  return 0;
}

int rsb_sharded_counts(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, double tol, double *marg_sums)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ensure_geo(ctx, 0)) return 1;
  if (upload_msa(ctx, msa, row_stride, 0, 1, 0, on_device, ctx->stream)) return 1;
  if (enqueue_counts(ctx, 0, 0, 1, ctx->d_res, ctx->stream)) return 1;
  if (enqueue_marginals(ctx, 0, 1, tol, ctx->stream, 1)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(marg_sums, ctx->d_msum, sizeof(double) * 4 * ctx->L, cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rsb_sharded_counts_pool(rsb_ctx *ctx, int rep, double tol, double *marg_sums)
{
  if (pool_range_ok(ctx, rep, 1)) return 1;
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_wait(ctx, rep, 1, ctx->stream)) return 1;
  return rsb_sharded_counts(ctx, ctx->d_pool + (size_t) rep * ctx->N * ctx->L, ctx->L, 1, tol, marg_sums);
}

int rsb_sharded_statistic(rsb_ctx *ctx, const double *marg_sums, double tol, int stat, int covclass, const double *allowpair, double *cov_sums)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  if (stat == RSB_RAF || stat == RSB_RAFS || stat == RSB_CCF) { rsb_set_error(ctx, "statistic not available with a sharded pair grid"); return 1; }
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_msum, marg_sums, sizeof(double) * 4 * ctx->L, cudaMemcpyHostToDevice, ctx->stream));
  if (enqueue_marginals(ctx, 0, 1, tol, ctx->stream, 2)) return 1;
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_cov, 0, sizeof(double) * (size_t) ctx->L * ctx->Lp, ctx->stream));     // rows of other ranks stay 0
  if (enqueue_statistic(ctx, 0, 1, stat, covclass, allow_mask(allowpair), ctx->stream, 1)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(cov_sums, ctx->d_covsum, sizeof(double) * (ctx->L + 4), cudaMemcpyDeviceToHost, ctx->stream));
  return check_flags(ctx, "corr_Probs");
}

int rsb_sharded_correct(rsb_ctx *ctx, const double *cov_sums, int actype, int mode, double w, double bmin, double *cov, double *minmax)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_covsum, cov_sums, sizeof(double) * (ctx->L + 4), cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_w, &w, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (enqueue_statistic(ctx, 0, 1, RSB_GT, RSB_C16, 0, ctx->stream, 2)) return 1;                          // covsum -> COVx, COVavg
  if (enqueue_correct(ctx, 0, 1, actype, mode, bmin, ctx->stream)) return 1;
  if ((mode & 2) && w > 0.0) {                                                                            // pairs owned by this rank
    unsigned long long mine = 0;
    for (int i = 0; i < ctx->L; i++) if ((i / RSB_ICOLS) % ctx->shard_world == ctx->shard_rank) mine += (unsigned long long) (ctx->L - 1 - i);
    if (ctx->d_m2p) { rsb_set_error(ctx, "a pair exclusion (rsb_set_pair_exclusion) is not offered on a sharded pair grid"); return 1; }
    ctx->hist_n += mine;
  }
  double mmx[2];
  if (cov && copy_matrix_out(ctx, ctx->d_cov, cov)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(mmx, ctx->d_minmax, sizeof(mmx), cudaMemcpyDeviceToHost, ctx->stream));
  if (check_flags(ctx, "corr_CalculateCOVCorrected")) return 1;
  if (minmax) { minmax[0] = mmx[0]; minmax[1] = mmx[1]; }
  return 0;
}

/* histograms of the scan left by rsb_scan / rsb_correct / rsb_statistic: ha (all pairs), and with pairmask also hb / ht */
int rsb_scan_hist(rsb_ctx *ctx, const uint8_t *pairmask, double w, double bmin, int nb, uint64_t *ha, uint64_t *hb, uint64_t *ht)
{
  RSB_RANGE("rsb_scan_hist");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (nb < 1 || !(w > 0.0)) { rsb_set_error(ctx, "bad histogram geometry"); return 1; }
  const size_t L = ctx->L;
  unsigned long long *d3 = nullptr; uint8_t *dmask = nullptr;
  int rc = 1;
#define SH_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rsb_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)
  SH_OK(cudaMalloc(&d3, sizeof(unsigned long long) * 3 * (size_t) nb));
  SH_OK(cudaMemsetAsync(d3, 0, sizeof(unsigned long long) * 3 * (size_t) nb, ctx->stream));
  if (pairmask) {
    SH_OK(cudaMalloc(&dmask, L * L));
    SH_OK(cudaMemcpyAsync(dmask, pairmask, L * L, cudaMemcpyHostToDevice, ctx->stream));
  }
  SH_OK(rsb_launch_hist3(ctx->d_cov, ctx->L, ctx->Lp, dmask, bmin, w, nb, d3, d3 + nb, d3 + 2 * (size_t) nb, ctx->d_flags, ctx->d_m2p, ctx->mind, ctx->stream));
  ctx->launches++;
  if (ha) SH_OK(cudaMemcpyAsync(ha, d3, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  if (hb) SH_OK(cudaMemcpyAsync(hb, d3 + nb, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  if (ht) SH_OK(cudaMemcpyAsync(ht, d3 + 2 * (size_t) nb, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  rc = check_flags(ctx, "cov_SignificantPairs_Ranking");
done:
#undef SH_OK
  cudaFree(d3); cudaFree(dmask);
  return rc;
}

/* pairs with both columns in the PDB sequence and fewer than `mind` positions apart stay out of every histogram, src/covariation.c:421-427 */
int rsb_set_pair_exclusion(rsb_ctx *ctx, const int *msa2pdb, int mind)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->N == 0) { rsb_set_error(ctx, "rsb_configure first"); return 1; }
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const size_t L = ctx->L;
  ctx->pairs_in_hist = (unsigned long long) L * (L - 1) / 2;
  if (!msa2pdb) { dfree(ctx->d_m2p); ctx->mind = 1; return 0; }
  if (!ctx->d_m2p) RSB_CUDA_OK(cudaMalloc(&ctx->d_m2p, sizeof(int) * L));
  RSB_CUDA_OK(cudaMemcpy(ctx->d_m2p, msa2pdb, sizeof(int) * L, cudaMemcpyHostToDevice));
  ctx->mind = mind;
  unsigned long long excl = 0;
  for (size_t i = 0; i + 1 < L; i++)
    for (size_t j = i + 1; j < L; j++) excl += (msa2pdb[i] >= 0 && msa2pdb[j] >= 0 && msa2pdb[j] - msa2pdb[i] < mind);
  ctx->pairs_in_hist -= excl;
  return 0;
}

/* scores written or changed on the host (mi->COV) -> the device matrix the histogram / E-value stages read */
int rsb_load_scores(rsb_ctx *ctx, const double *cov)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!cov || !ctx->d_cov) { rsb_set_error(ctx, "rsb_load_scores: no matrix / context not configured"); return 1; }
  RSB_CUDA_OK(cudaMemcpy2DAsync(ctx->d_cov, sizeof(double) * ctx->Lp, cov, sizeof(double) * ctx->L, sizeof(double) * ctx->L, ctx->L,
                                cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));                                        // the caller may reuse cov at once
  return 0;
}

/* E-values and significant pairs of the scan left in d_cov: the per-pair loop of cov_CreateHitList, src/covariation.c:828-910 */
int rsb_scan_hits(rsb_ctx *ctx, const rsb_nullfit *null, const uint8_t *pairmask, uint64_t Nb, uint64_t Nt, int expBP, double thresh,
                  double *eval, int64_t cap, int64_t *hit_i, int64_t *hit_j, double *hit_sc, double *hit_eval, double *hit_pval, int64_t *nhit)
{
  RSB_RANGE("rsb_scan_hits");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!null || !null->obs || null->nb < 1 || !(null->w > 0.0) || null->imax < null->imin || null->imin < 0 || null->imax >= null->nb || null->Nc == 0) {
    rsb_set_error(ctx, "rsb_scan_hits: empty or inconsistent null histogram"); return 1;
  }
  if (cap < 0 || (cap > 0 && !(hit_i && hit_j))) { rsb_set_error(ctx, "rsb_scan_hits: bad hit list arguments"); return 1; }
  if (expBP > 0 && ctx->shard_world > 1) {
    rsb_set_error(ctx, "rsb_scan_hits: the expBP rule (covariation.c:852) follows the order of the whole pair list and is not offered on a sharded pair grid");
    return 1;
  }
  const size_t L = ctx->L, nb = (size_t) null->nb;
  const long long P = (long long) L * (long long) (L - 1) / 2;
  const int report_all = thresh > 1000.0;                                                   // MAX_EVAL, src/correlators.h:24
  const long long dcap = std::max<long long>(1, std::min<long long>(cap, P));

  // suffix sums of the bins (exact): what the reference's loop at :2390 adds up for every pair
  std::vector<unsigned long long> csum(nb, 0ull);
  unsigned long long run = 0;
  for (long long b = null->imax; b >= 0; b--) { run += null->obs[b]; csum[(size_t) b] = run; }

  unsigned long long *d_csum = nullptr, *d_n = nullptr;
  double *d_surv = nullptr, *d_eval = nullptr, *d_hd = nullptr;
  long long *d_ij = nullptr;
  uint8_t *d_mask = nullptr;
  int rc = 1;
  std::vector<long long> ij;
  std::vector<double> hd;
  unsigned long long n_dev = 0;
  rsb_nullview nv;
  long long switch_n = (expBP > 0) ? P : -1;                                               // first pass: every pair outside the structure uses expBP
#define HITS_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rsb_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)
  HITS_OK(cudaMalloc(&d_csum, sizeof(unsigned long long) * nb));
  HITS_OK(cudaMalloc(&d_n, sizeof(unsigned long long)));
  HITS_OK(cudaMalloc(&d_ij, sizeof(long long) * (size_t) dcap));
  HITS_OK(cudaMalloc(&d_hd, sizeof(double) * 3 * (size_t) dcap));
  HITS_OK(cudaMemcpyAsync(d_csum, csum.data(), sizeof(unsigned long long) * nb, cudaMemcpyHostToDevice, ctx->stream));
  if (null->survfit) {
    HITS_OK(cudaMalloc(&d_surv, sizeof(double) * 2 * nb));
    HITS_OK(cudaMemcpyAsync(d_surv, null->survfit, sizeof(double) * 2 * nb, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (pairmask) {
    HITS_OK(cudaMalloc(&d_mask, L * L));
    HITS_OK(cudaMemcpyAsync(d_mask, pairmask, L * L, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (eval) {
    HITS_OK(cudaMalloc(&d_eval, sizeof(double) * L * L));
    HITS_OK(cudaMemsetAsync(d_eval, 0, sizeof(double) * L * L, ctx->stream));               // pairs of other ranks' rows stay 0 on a sharded grid
  }
  nv.bmin = null->bmin; nv.w = null->w; nv.xmax = null->xmax; nv.phi = null->phi; nv.Nc = (double) null->Nc;
  nv.nb = null->nb; nv.imin = null->imin; nv.imax = null->imax; nv.csum = d_csum; nv.survfit = d_surv;

  for (int pass = 0; pass < 2; pass++) {
    HITS_OK(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), ctx->stream));
    HITS_OK(rsb_launch_evalue_hits(ctx->d_cov, ctx->L, ctx->Lp, nv, d_mask, (double) Nb, (double) Nt, (double) expBP, switch_n, thresh, report_all,
                                   ctx->shard_rank, ctx->shard_world, d_eval, cap > 0 ? dcap : 0, d_ij, d_hd, d_hd + dcap, d_hd + 2 * dcap,
                                   d_n, ctx->d_flags, ctx->stream));
    ctx->launches++;
    HITS_OK(cudaMemcpyAsync(&n_dev, d_n, sizeof(n_dev), cudaMemcpyDeviceToHost, ctx->stream));
    if (check_flags(ctx, "cov_CreateHitList")) goto done;                                   // synchronises the stream
    const size_t kept = (size_t) std::min<unsigned long long>(n_dev, cap > 0 ? (unsigned long long) dcap : 0ull);
    if (kept) {                                                                             // host staging: [kept] keys + 3 x [kept] doubles
      ij.resize(kept); hd.resize(3 * kept);
      HITS_OK(cudaMemcpyAsync(ij.data(), d_ij, sizeof(long long) * kept, cudaMemcpyDeviceToHost, ctx->stream));
      for (int f = 0; f < 3; f++)
        HITS_OK(cudaMemcpyAsync(hd.data() + f * kept, d_hd + f * (size_t) dcap, sizeof(double) * kept, cudaMemcpyDeviceToHost, ctx->stream));
      HITS_OK(cudaStreamSynchronize(ctx->stream));
    }
    if (pass == 1 || expBP <= 0) break;
    // the expBP-th hit in row-major order closes the expBP regime; with fewer hits it never closes and this pass stands
    if (n_dev < (unsigned long long) expBP) break;
    if (n_dev > kept) { rsb_set_error(ctx, "rsb_scan_hits: the expBP rule needs the whole first-pass hit list (%llu hits, capacity %lld)", n_dev, (long long) cap); goto done; }
    std::vector<long long> order(ij.begin(), ij.begin() + kept);
    std::nth_element(order.begin(), order.begin() + (expBP - 1), order.end());
    const long long key = order[(size_t) expBP - 1], ki = key >> 32, kj = key & 0xffffffffll;
    switch_n = ki * (long long) L - ki * (ki + 1) / 2 + (kj - ki - 1);
  }
  {
    const size_t kept = (size_t) std::min<unsigned long long>(n_dev, cap > 0 ? (unsigned long long) dcap : 0ull);
    std::vector<size_t> idx(kept);
    for (size_t k = 0; k < kept; k++) idx[k] = k;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return ij[a] < ij[b]; });   // (i << 32 | j): the reference's row-major order
    for (size_t k = 0; k < kept; k++) {
      const size_t s = idx[k];
      hit_i[k] = ij[s] >> 32; hit_j[k] = ij[s] & 0xffffffffll;
      if (hit_sc)   hit_sc[k]   = hd[s];
      if (hit_eval) hit_eval[k] = hd[kept + s];
      if (hit_pval) hit_pval[k] = hd[2 * kept + s];
    }
    if (nhit) *nhit = (int64_t) n_dev;
    if (eval) {
      HITS_OK(cudaMemcpyAsync(eval, d_eval, sizeof(double) * L * L, cudaMemcpyDeviceToHost, ctx->stream));
      HITS_OK(cudaStreamSynchronize(ctx->stream));
    }
  }
  rc = 0;
done:
#undef HITS_OK
  cudaFree(d_csum); cudaFree(d_n); cudaFree(d_ij); cudaFree(d_hd); cudaFree(d_surv); cudaFree(d_mask); cudaFree(d_eval);
  return rc;
}

/* Tree_Substitutions after its Fitch pass, src/msatree.c:1455-1540: one row per branch, then the unweighted pair contraction */
int rsb_tree_substitutions(rsb_ctx *ctx, int ntaxa, const int *left, const int *right, const uint8_t *leaves, int64_t leaf_stride,
                           const uint8_t *internal, int64_t internal_stride, int includegaps, int *nsubs, int *ndouble, int *njoin)
{
  RSB_RANGE("rsb_tree_substitutions");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const int nrows = 2 * (ntaxa - 1);
  if (ntaxa < 2 || !left || !right || !leaves || !internal) { rsb_set_error(ctx, "rsb_tree_substitutions: bad arguments"); return 1; }
  if (ctx->N != nrows || ctx->Rcap < 1) {
    rsb_set_error(ctx, "rsb_tree_substitutions: the context must be configured for one row per branch: nseq = 2 (ntaxa - 1) = %d (it is %d)", nrows, ctx->N);
    return 1;
  }
  if (nrows > 65535) { rsb_set_error(ctx, "rsb_tree_substitutions: more than 65535 branches"); return 1; }
  if (ctx->shard_world > 1) { rsb_set_error(ctx, "rsb_tree_substitutions is not offered on a sharded pair grid"); return 1; }
  for (int v = 0; v < ntaxa - 1; v++) {
    const int kids[2] = { left[v], right[v] };
    for (int k : kids)
      if (k >= ntaxa - 1 || -k >= ntaxa || (k > 0 && k <= v)) {            // k <= 0 is leaf -k; internal children come after their parent
        rsb_set_error(ctx, "rsb_tree_substitutions: node %d has child %d outside the tree", v, k); return 1;
      }
  }
  const size_t L = ctx->L;
  uint8_t *d_leaves = nullptr, *d_internal = nullptr;
  int *d_lr = nullptr, *d_ns = nullptr, *d_tab = nullptr;
  int rc = 1;
  const bool pairs = (ndouble || njoin);
#define TS_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rsb_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)
  TS_OK(cudaMalloc(&d_leaves, (size_t) ntaxa * L));
  TS_OK(cudaMalloc(&d_internal, (size_t) (ntaxa - 1) * L));
  TS_OK(cudaMalloc(&d_lr, sizeof(int) * 2 * (size_t) (ntaxa - 1)));
  TS_OK(cudaMalloc(&d_ns, sizeof(int) * L));
  if (pairs) TS_OK(cudaMalloc(&d_tab, sizeof(int) * 2 * L * L));
  TS_OK(cudaMemcpy2DAsync(d_leaves, L, leaves, (size_t) leaf_stride, L, (size_t) ntaxa, cudaMemcpyHostToDevice, ctx->stream));
  TS_OK(cudaMemcpy2DAsync(d_internal, L, internal, (size_t) internal_stride, L, (size_t) (ntaxa - 1), cudaMemcpyHostToDevice, ctx->stream));
  TS_OK(cudaMemcpyAsync(d_lr, left, sizeof(int) * (size_t) (ntaxa - 1), cudaMemcpyHostToDevice, ctx->stream));
  TS_OK(cudaMemcpyAsync(d_lr + (ntaxa - 1), right, sizeof(int) * (size_t) (ntaxa - 1), cudaMemcpyHostToDevice, ctx->stream));
  TS_OK(rsb_launch_branch_rows(d_leaves, d_internal, d_lr, d_lr + (ntaxa - 1), ntaxa, ctx->L, includegaps, ctx->d_res, d_ns, ctx->stream));   // replicate slot 0
  ctx->launches++;
  if (nsubs) TS_OK(cudaMemcpyAsync(nsubs, d_ns, sizeof(int) * L, cudaMemcpyDeviceToHost, ctx->stream));
  if (pairs) {
    if (enqueue_counts(ctx, 1, 0, 1, ctx->d_res, ctx->stream)) goto done;                  // unit weights: the counts are plain integers
    TS_OK(rsb_launch_subs_tables(ctx->d_cnt, ctx->L, ctx->Lp, ndouble ? d_tab : nullptr, njoin ? d_tab + L * L : nullptr, ctx->stream));
    ctx->launches++;
    if (ndouble) TS_OK(cudaMemcpyAsync(ndouble, d_tab, sizeof(int) * L * L, cudaMemcpyDeviceToHost, ctx->stream));
    if (njoin)   TS_OK(cudaMemcpyAsync(njoin, d_tab + L * L, sizeof(int) * L * L, cudaMemcpyDeviceToHost, ctx->stream));
  }
  TS_OK(cudaStreamSynchronize(ctx->stream));
  rc = 0;
done:
#undef TS_OK
  cudaFree(d_leaves); cudaFree(d_internal); cudaFree(d_lr); cudaFree(d_ns); cudaFree(d_tab);
  return rc;
}

int rsb_hist_reset(rsb_ctx *ctx)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  RSB_CUDA_OK(cudaMemsetAsync(ctx->d_hist, 0, sizeof(unsigned long long) * HIST_BINS, ctx->stream));
  ctx->hist_n = 0;
  return 0;
}

int rsb_hist_read(rsb_ctx *ctx, uint64_t *bins, int nb_cap, uint64_t *n_out, int *imax_out)
{
  RSB_RANGE("rsb_hist_read");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const int nb = std::min(nb_cap, HIST_BINS);
  RSB_CUDA_OK(cudaMemcpyAsync(bins, ctx->d_hist, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  int imax = -1;
  for (int b = 0; b < nb; b++) if (bins[b]) imax = b;
  if (n_out) *n_out = ctx->hist_n;
  if (imax_out) *imax_out = imax;
  return 0;
}

int rsb_hist_exchange(rsb_ctx *ctx, void *device_buf, int nb, int to_library)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->d_hist || !device_buf || nb < 1 || nb > HIST_BINS) { rsb_set_error(ctx, "rsb_hist_exchange: bad arguments"); return 1; }
  if (to_library) RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_hist, device_buf, sizeof(uint64_t) * nb, cudaMemcpyDeviceToDevice, ctx->stream));
  else            RSB_CUDA_OK(cudaMemcpyAsync(device_buf, ctx->d_hist, sizeof(uint64_t) * nb, cudaMemcpyDeviceToDevice, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));                  // the caller's collective runs on another stream
  return 0;
}

int rsb_last_nseff(rsb_ctx *ctx, double *nseff, double *ngap)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->cur_geo == 1) { rsb_set_error(ctx, "no weighted counts resident"); return 1; }
  Geo &g = ctx->geo[ctx->cur_geo];
  const size_t L = ctx->L;
  if (!ctx->d_pp_out) {
    RSB_CUDA_OK(cudaMalloc(&ctx->d_pp_out, L * L * 16 * sizeof(double)));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_nseff_out, L * L * sizeof(double)));
    RSB_CUDA_OK(cudaMalloc(&ctx->d_ngap_out, L * L * sizeof(double)));
  }
  if (ctx->cur_rec)
    RSB_CUDA_OK(rsb_launch_export_nseff_rec((const double *) (ctx->d_cnt + (size_t) ctx->last_slot * 16 * L * ctx->Lp), ctx->L, ctx->Lp,
                                            (double) g.wtot * g.scale, ctx->d_nseff_out, ctx->d_ngap_out, ctx->stream));
  else
    RSB_CUDA_OK(rsb_launch_export_probs(ctx->d_cnt + (size_t) ctx->last_slot * 16 * L * ctx->Lp, ctx->L, ctx->Lp, g.scale, g.wtot,
                                        ctx->d_pp_out, ctx->d_nseff_out, ctx->d_ngap_out, ctx->stream));
  ctx->launches++;
  if (nseff) RSB_CUDA_OK(cudaMemcpyAsync(nseff, ctx->d_nseff_out, L * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (ngap)  RSB_CUDA_OK(cudaMemcpyAsync(ngap, ctx->d_ngap_out, L * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------- generators
int rsb_set_tree(rsb_ctx *ctx, const int *left, const int *right, const int *parent, const double *ld, const double *rd)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_gen));        // a generator may still be walking the previous tree
  const int N = ctx->N, nn = N - 1;
  if (N < 2) { rsb_set_error(ctx, "a tree needs at least 2 leaves"); return 1; }
  for (int v = 0; v < nn; v++) {
    const int kids[2] = { left[v], right[v] };
    for (int c : kids) if (c > 0 && (c <= v || c >= nn)) { rsb_set_error(ctx, "tree node %d: child %d must have a larger index than its parent", v, c); return 1; }
      else if (c <= 0 && -c >= N) { rsb_set_error(ctx, "tree node %d: leaf %d out of range", v, -c); return 1; }
  }
  ctx->h_left.assign(left, left + nn); ctx->h_right.assign(right, right + nn); ctx->h_parent.assign(parent, parent + nn);
  ctx->h_ld.assign(ld, ld + nn); ctx->h_rd.assign(rd, rd + nn);
  // nodes grouped by depth (level order): every level is a set of independent branches
  std::vector<int> depth(nn, 0), order(nn), level_start;
  for (int v = 0; v < nn; v++) { if (left[v] > 0) depth[left[v]] = depth[v] + 1; if (right[v] > 0) depth[right[v]] = depth[v] + 1; }
  int maxd = 0; for (int v = 0; v < nn; v++) maxd = std::max(maxd, depth[v]);
  for (int v = 0; v < nn; v++) order[v] = v;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return depth[a] < depth[b]; });
  level_start.assign(maxd + 2, 0);
  for (int v = 0; v < nn; v++) level_start[depth[v] + 1]++;
  for (int d = 0; d <= maxd; d++) level_start[d + 1] += level_start[d];
  ctx->nlevels = maxd + 1;
  ctx->h_level_start = level_start;
  dfree(ctx->d_left); dfree(ctx->d_right); dfree(ctx->d_parent); dfree(ctx->d_order); dfree(ctx->d_level_start);
  RSB_CUDA_OK(cudaMalloc(&ctx->d_left, sizeof(int) * nn));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_right, sizeof(int) * nn));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_parent, sizeof(int) * nn));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_order, sizeof(int) * nn));
  RSB_CUDA_OK(cudaMalloc(&ctx->d_level_start, sizeof(int) * (maxd + 2)));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_left, left, sizeof(int) * nn, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_right, right, sizeof(int) * nn, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_parent, parent, sizeof(int) * nn, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_order, order.data(), sizeof(int) * nn, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_level_start, level_start.data(), sizeof(int) * (maxd + 2), cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->sim_valid = false;                     // branch matrices depend on the branch lengths
  ctx->have_tree = true;
  return 0;
}

// P(t) = exp(tQ) by scaling and squaring in double on the host (4x4, N-1 branches x 2: negligible), with the
// reference's float-rounded floored time, negative clip and row renormalisation (e1_model.c:64-75, ratematrix.c:199-222)
static int branch_matrix(const double *Q, double t, double *P)
{
  float rt = (float) t;
  rt = (rt >= 0.0 && rt < 1e-5) ? 1e-5 : rt;
  if (rt < 0.0) { if (rt > -1e-5) rt = 1e-5; else return 1; }
  const double time = (rt > 10000.) ? 10000.0 : (double) rt;
  double M[16], T[16], X[16], norm = 0.0;
  for (int k = 0; k < 16; k++) { M[k] = time * Q[k]; norm += M[k] * M[k]; }
  norm = std::sqrt(norm);
  int z = 0; while (norm > 0.1) { norm *= 0.5; z++; }
  for (int k = 0; k < 16; k++) M[k] = std::ldexp(M[k], -z);
  for (int k = 0; k < 16; k++) { P[k] = (k % 5 == 0) ? 1.0 : 0.0; T[k] = P[k]; }
  for (int n = 1; n < 100; n++) {
    double delta = 0.0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += T[i * 4 + k] * M[k * 4 + j]; X[i * 4 + j] = s / n; }
    for (int k = 0; k < 16; k++) { T[k] = X[k]; P[k] += T[k]; delta += std::fabs(T[k]); }
    if (delta < 1e-18) break;
  }
  while (z-- > 0) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += P[i * 4 + k] * P[k * 4 + j]; X[i * 4 + j] = s; }
    memcpy(P, X, sizeof(X));
  }
  for (int i = 0; i < 4; i++) {
    double sum = 0.0;
    for (int j = 0; j < 4; j++) { if (P[i * 4 + j] < 0.0) { if (std::fabs(P[i * 4 + j]) < 0.001) P[i * 4 + j] = 0.0; else return 2; } sum += P[i * 4 + j]; }
    for (int j = 0; j < 4; j++) P[i * 4 + j] = (sum != 0.0) ? P[i * 4 + j] / sum : 0.25;
  }
  return 0;
}

int rsb_pool_reserve(rsb_ctx *ctx, int nrep)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->N == 0) { rsb_set_error(ctx, "rsb_configure first"); return 1; }
  if (nrep < 1) { rsb_set_error(ctx, "bad pool size"); return 1; }
  if (nrep <= ctx->Rpool) return 0;
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_gen));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->pool_ready.assign(nrep, nullptr);
  dfree(ctx->d_pool); dfree(ctx->d_simscratch); dfree(ctx->d_anc); dfree(ctx->d_shanc); dfree(ctx->d_perm);
  RSB_CUDA_OK(cudaMalloc(&ctx->d_pool, (size_t) nrep * ctx->N * ctx->L));
  ctx->Rpool = nrep;
  return 0;
}

int rsb_null_simulate(rsb_ctx *ctx, const double *Q, const uint8_t *root, const uint8_t *gapmask, int64_t gap_stride,
                      uint64_t seed, uint64_t first_id, int first_rep, int nrep)
{
  RSB_RANGE("rsb_null_simulate");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->have_tree) { rsb_set_error(ctx, "rsb_set_tree first"); return 1; }
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const int N = ctx->N, L = ctx->L, nn = N - 1;
  cudaStream_t sg = ctx->stream_gen;                     // generation stream, as generator A: after everything queued by the caller
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));
  RSB_CUDA_OK(cudaStreamWaitEvent(sg, ctx->ev_entry, 0));
  // branch matrices P(t) = exp(tQ) (ratematrix_ConditionalsFromRate, src/ratematrix.c:185-233) as cumulative integer
  // thresholds [node][side][4][4]; rebuilt only when the rate matrix or the tree changed
  if (!ctx->sim_valid || memcmp(ctx->sim_Q, Q, sizeof(ctx->sim_Q)) != 0) {
    std::vector<unsigned long long> th((size_t) nn * 32);
    for (int v = 0; v < nn; v++)
      for (int side = 0; side < 2; side++) {
        double P[16];
        if (branch_matrix(Q, side ? ctx->h_rd[v] : ctx->h_ld[v], P)) { rsb_set_error(ctx, "failed to evolve node %d to time %f", v, side ? ctx->h_rd[v] : ctx->h_ld[v]); return 1; }
        for (int a = 0; a < 4; a++) {
          double cdf = 0.0;
          for (int b = 0; b < 4; b++) {
            cdf += P[a * 4 + b];
            const double y = std::ceil(std::ldexp(cdf, 32));                             // rnd / 2^32 < cdf  <=>  rnd < ceil(cdf 2^32)
            th[((size_t) v * 2 + side) * 16 + a * 4 + b] = (y <= 0.0) ? 0ULL : (y >= 4294967296.0) ? 4294967296ULL : (unsigned long long) y;
          }
        }
      }
    if (!ctx->d_pthr) RSB_CUDA_OK(cudaMalloc(&ctx->d_pthr, th.size() * sizeof(unsigned long long)));
    RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_pthr, th.data(), th.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, sg));
    RSB_CUDA_OK(cudaStreamSynchronize(sg));                // th is a host temporary
    memcpy(ctx->sim_Q, Q, sizeof(ctx->sim_Q));
    ctx->sim_valid = true;
  }
  if (!ctx->d_root) RSB_CUDA_OK(cudaMalloc(&ctx->d_root, L));
  if (!ctx->d_simscratch) RSB_CUDA_OK(cudaMalloc(&ctx->d_simscratch, (size_t) ctx->Rpool * nn * L));
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_root, root, L, cudaMemcpyHostToDevice, sg));
  if (gapmask) {
    if (!ctx->d_gapmask) RSB_CUDA_OK(cudaMalloc(&ctx->d_gapmask, (size_t) N * L));
    RSB_CUDA_OK(cudaMemcpy2DAsync(ctx->d_gapmask, L, gapmask, (size_t) gap_stride, L, N, cudaMemcpyHostToDevice, sg));
  }
  // chunks of growing size with a readiness event each, as generator A
  int off = 0, chunks = 0, next = 4;
  while (off < nrep) {
    int n = (off == 0) ? 1 : next;
    if (off > 0) next *= 2;
    if (nrep <= 16 || chunks >= 5 || nrep - off - n < 4) n = nrep - off;
    RSB_CUDA_OK(rsb_launch_null_simulate(ctx->d_left, ctx->d_right, ctx->d_order, ctx->h_level_start.data(), ctx->nlevels, ctx->d_pthr, N, L, ctx->d_root,
                                         gapmask ? ctx->d_gapmask : nullptr, seed, first_id + (uint64_t) off, first_rep + off, n, ctx->d_pool,
                                         ctx->d_simscratch, sg));
    cudaEvent_t e = ctx->gen_ring[ctx->gen_next++ % ctx->gen_ring.size()];
    RSB_CUDA_OK(cudaEventRecord(e, sg));
    for (int r = first_rep + off; r < first_rep + off + n; r++) ctx->pool_ready[r] = e;
    off += n; chunks++;
  }
  ctx->launches += chunks * ctx->nlevels;
  return 0;
}

static int fitch_shuffle_impl(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, uint64_t seed, uint64_t first_id, const uint64_t *ids,
                              int first_rep, int nrep)
{
  RSB_RANGE("rsb:null_fitch_shuffle");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->have_tree) { rsb_set_error(ctx, "rsb_set_tree first"); return 1; }
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const int N = ctx->N, L = ctx->L, nn = N - 1;
  if (!ctx->d_msa0)  RSB_CUDA_OK(cudaMalloc(&ctx->d_msa0, (size_t) N * L));
  if (!ctx->d_anc)   RSB_CUDA_OK(cudaMalloc(&ctx->d_anc, (size_t) ctx->Rpool * nn * L));
  if (!ctx->d_shanc) RSB_CUDA_OK(cudaMalloc(&ctx->d_shanc, (size_t) ctx->Rpool * nn * L));
  if (!ctx->d_perm)  RSB_CUDA_OK(cudaMalloc(&ctx->d_perm, sizeof(int) * (size_t) ctx->Rpool * L));
  if (!ctx->d_sets)  RSB_CUDA_OK(cudaMalloc(&ctx->d_sets, (size_t) nn * L));
  if (!ctx->d_genflag) RSB_CUDA_OK(cudaMalloc(&ctx->d_genflag, sizeof(int)));
  // Runs on the generation stream, after everything already queued on the caller's stream (which may still read the
  // pool), in chunks of growing size: the first entry alone (the width pass wants it first), then 4, 8, 16 ...  Each chunk
  // records an event; the scans wait for the event of the entries they pack, so scanning starts while later replicates
  // are still being generated.
  cudaStream_t sg = ctx->stream_gen;
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));
  RSB_CUDA_OK(cudaStreamWaitEvent(sg, ctx->ev_entry, 0));
  RSB_CUDA_OK(cudaMemcpy2DAsync(ctx->d_msa0, L, msa, (size_t) row_stride, L, N, cudaMemcpyHostToDevice, sg));
  int unknown = 1;
  RSB_CUDA_OK(rsb_launch_unknown_check(ctx->d_msa0, (size_t) N * L, ctx->d_genflag, &unknown, sg));
  if (unknown & 2) {                                                 // the reference fails here too: "S not set up properly" (msatree.c:1740)
    rsb_set_error(ctx, "the alignment holds residue codes other than A C G U, gap and N (degenerate symbols must be converted to N first, "
                       "msamanip_ConvertDegen2N, src/R-scape.c:1855): the Fitch pass of the null generator cannot set them up");
    return 1;
  }
  if (getenv("RSCAPE_B200_FITCH_PER_REPLICATE")) unknown = 1;      // tests: force the general path
  uint8_t *sets = unknown ? nullptr : ctx->d_sets;
  unsigned long long *d_ids = nullptr;
  if (ids) {                                                         // explicit replicate ids (device copy; small, staged by the driver)
    if (ctx->ids_cap < (size_t) nrep) { dfree(ctx->d_ids); RSB_CUDA_OK(cudaMalloc(&ctx->d_ids, sizeof(unsigned long long) * (size_t) ctx->Rpool)); ctx->ids_cap = ctx->Rpool; }
    RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_ids, ids, sizeof(unsigned long long) * (size_t) nrep, cudaMemcpyHostToDevice, sg));
    d_ids = ctx->d_ids;
  }
  RSB_CUDA_OK(rsb_launch_permutations(L, seed, first_id, d_ids, first_rep, nrep, ctx->d_perm, sg));
  // chunk sizes 1, 4, 8, 16, 32, then the rest; a small batch (every level kernel is latency-bound then) goes in one piece
  static const int chunk0 = getenv("RSCAPE_B200_GEN_CHUNK") ? std::max(1, atoi(getenv("RSCAPE_B200_GEN_CHUNK"))) : 4;   // experiments
  int off = 0, chunks = 0, next = chunk0;
  while (off < nrep) {
    int n = (off == 0) ? 1 : next;
    if (off > 0) next *= 2;
    if (nrep <= 16 || chunks >= 5 || nrep - off - n < 4) n = nrep - off;
    RSB_CUDA_OK(rsb_launch_fitch_shuffle(ctx->d_left, ctx->d_right, ctx->d_parent, ctx->d_order, ctx->h_level_start.data(), ctx->nlevels, N, L,
                                         ctx->d_msa0, seed, first_id + (uint64_t) off, d_ids ? d_ids + off : nullptr, first_rep + off, n, ctx->d_pool, ctx->d_anc, ctx->d_shanc,
                                         ctx->d_perm, sets, off == 0, sg));
    cudaEvent_t e = ctx->gen_ring[ctx->gen_next++ % ctx->gen_ring.size()];
    RSB_CUDA_OK(cudaEventRecord(e, sg));
    for (int r = first_rep + off; r < first_rep + off + n; r++) ctx->pool_ready[r] = e;
    off += n; chunks++;
  }
  ctx->launches += 2 + chunks * (1 + 3 * ctx->nlevels);
  return 0;
}

int rsb_null_fitch_shuffle(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, uint64_t seed, uint64_t first_id, int first_rep, int nrep)
{
  return fitch_shuffle_impl(ctx, msa, row_stride, seed, first_id, nullptr, first_rep, nrep);
}

int rsb_null_fitch_shuffle_ids(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, uint64_t seed, const uint64_t *ids, int first_rep, int nrep)
{
  if (!ids) { rsb_set_error(ctx, "rsb_null_fitch_shuffle_ids: no ids"); return 1; }
  return fitch_shuffle_impl(ctx, msa, row_stride, seed, 0, ids, first_rep, nrep);
}

int rsb_pool_get(rsb_ctx *ctx, int first_rep, int nrep, uint8_t *out)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const size_t rb = (size_t) ctx->N * ctx->L;
  if (pool_wait(ctx, first_rep, nrep, ctx->stream)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(out, ctx->d_pool + first_rep * rb, rb * nrep, cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rsb_pool_get_internal(rsb_ctx *ctx, int which, int first_rep, int nrep, uint8_t *out)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const uint8_t *src = (which == 0) ? ctx->d_anc : (which == 1) ? ctx->d_shanc : nullptr;
  if (!src) { rsb_set_error(ctx, "rsb_pool_get_internal: generator A has not run (or bad selector %d)", which); return 1; }
  const size_t rb = (size_t) (ctx->N - 1) * ctx->L;
  if (pool_wait(ctx, first_rep, nrep, ctx->stream)) return 1;
  RSB_CUDA_OK(cudaMemcpyAsync(out, src + first_rep * rb, rb * nrep, cudaMemcpyDeviceToHost, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rsb_pool_put(rsb_ctx *ctx, int first_rep, int nrep, const uint8_t *in)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  const size_t rb = (size_t) ctx->N * ctx->L;
  if (pool_wait(ctx, first_rep, nrep, ctx->stream)) return 1;                      // a generator may still be writing these entries
  for (int r = first_rep; r < first_rep + nrep; r++) ctx->pool_ready[r] = nullptr;
  RSB_CUDA_OK(cudaMemcpyAsync(ctx->d_pool + first_rep * rb, in, rb * nrep, cudaMemcpyHostToDevice, ctx->stream));
  RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------- alignment preprocessing
namespace {
// the alignment on the device: the caller's buffer itself (on_device), else a temporary copy.  `owned` must be freed by the caller.
int stage_msa(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const uint8_t **dev, long long *dev_stride, uint8_t **owned)
{
  *owned = nullptr;
  if (nseq < 1 || alen < 1 || !msa || row_stride < alen) { rsb_set_error(ctx, "bad alignment arguments"); return 1; }
  if (on_device) { *dev = msa; *dev_stride = row_stride; return 0; }
  RSB_CUDA_OK(cudaMalloc(owned, (size_t) nseq * alen));
  cudaError_t e = cudaMemcpy2DAsync(*owned, alen, msa, (size_t) row_stride, alen, nseq, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { cudaFree(*owned); *owned = nullptr; rsb_set_error(ctx, "alignment upload: %s", cudaGetErrorString(e)); return 1; }
  *dev = *owned; *dev_stride = alen;
  return 0;
}
} // namespace

#define PREP_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rsb_set_error(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)

/* msamanip_RemoveGapColumns' column test, src/msamanip.c:486-500 */
int rsb_msa_gap_columns(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const double *wgt,
                        double gapthresh, uint8_t *useme)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const uint8_t *d = nullptr; long long ds = 0; uint8_t *owned = nullptr, *d_use = nullptr; double *d_w = nullptr;
  int rc = 1;
  if (stage_msa(ctx, msa, nseq, alen, row_stride, on_device, &d, &ds, &owned)) return 1;
  PREP_OK(cudaMalloc(&d_use, alen));
  if (wgt) { PREP_OK(cudaMalloc(&d_w, sizeof(double) * nseq)); PREP_OK(cudaMemcpyAsync(d_w, wgt, sizeof(double) * nseq, cudaMemcpyHostToDevice, ctx->stream)); }
  PREP_OK(rsb_launch_gap_columns(d, nseq, alen, ds, d_w, 1.0 - gapthresh, d_use, ctx->stream));
  ctx->launches++;
  PREP_OK(cudaMemcpyAsync(useme, d_use, alen, cudaMemcpyDeviceToHost, ctx->stream));
  PREP_OK(cudaStreamSynchronize(ctx->stream));
  rc = 0;
done:
  cudaFree(owned); cudaFree(d_use); cudaFree(d_w);
  return rc;
}

/* esl_msaweight_PB (src/R-scape.c:1556) */
int rsb_msa_pb_weights(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, double *wgt)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const uint8_t *d = nullptr; long long ds = 0; uint8_t *owned = nullptr; double *d_coef = nullptr, *d_w = nullptr;
  int rc = 1;
  if (stage_msa(ctx, msa, nseq, alen, row_stride, on_device, &d, &ds, &owned)) return 1;
  PREP_OK(cudaMalloc(&d_coef, sizeof(double) * 4 * (size_t) alen));
  PREP_OK(cudaMalloc(&d_w, sizeof(double) * nseq));
  PREP_OK(rsb_launch_pb_weights(d, nseq, alen, ds, d_coef, d_w, ctx->stream));
  ctx->launches += 3;
  PREP_OK(cudaMemcpyAsync(wgt, d_w, sizeof(double) * nseq, cudaMemcpyDeviceToHost, ctx->stream));
  PREP_OK(cudaStreamSynchronize(ctx->stream));
  rc = 0;
done:
  cudaFree(owned); cudaFree(d_coef); cudaFree(d_w);
  return rc;
}

/* esl_dst_XPairId for a list of sequence pairs (pairs: int [npairs][2]) -> pid[npairs]; pairs == NULL: every pair, and out is the
 * distance matrix double [nseq][nseq] = 1 - pid (diagonal 0), what esl_msaweight_GSC builds its UPGMA tree from */
int rsb_msa_pair_identity(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const int *pairs, int64_t npairs,
                          double *out)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const uint8_t *d = nullptr; long long ds = 0; uint8_t *owned = nullptr; int *d_pairs = nullptr; double *d_out = nullptr;
  int rc = 1;
  const long long np = pairs ? (long long) npairs : (long long) nseq * (nseq - 1) / 2;
  const size_t nout = pairs ? (size_t) npairs : (size_t) nseq * nseq;
  if (pairs) for (int64_t k = 0; k < npairs; k++)
    if (pairs[2 * k] < 0 || pairs[2 * k] >= nseq || pairs[2 * k + 1] < 0 || pairs[2 * k + 1] >= nseq) { rsb_set_error(ctx, "rsb_msa_pair_identity: pair %lld out of range", (long long) k); return 1; }
  if (stage_msa(ctx, msa, nseq, alen, row_stride, on_device, &d, &ds, &owned)) return 1;
  PREP_OK(cudaMalloc(&d_out, sizeof(double) * std::max<size_t>(nout, 1)));
  if (pairs && npairs > 0) { PREP_OK(cudaMalloc(&d_pairs, sizeof(int) * 2 * (size_t) npairs)); PREP_OK(cudaMemcpyAsync(d_pairs, pairs, sizeof(int) * 2 * (size_t) npairs, cudaMemcpyHostToDevice, ctx->stream)); }
  PREP_OK(rsb_launch_pair_identity(d, nseq, alen, ds, d_pairs, np, d_out, ctx->stream));
  ctx->launches += pairs ? 1 : 2;
  PREP_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
  PREP_OK(cudaStreamSynchronize(ctx->stream));
  rc = 0;
done:
  cudaFree(owned); cudaFree(d_pairs); cudaFree(d_out);
  return rc;
}

/* struct_ColumnSubset's residue part: out [nseq][nkeep] = the columns with useme != 0 (out may be a device pointer when out_on_device) */
int rsb_msa_column_subset(rsb_ctx *ctx, const uint8_t *msa, int nseq, int alen, int64_t row_stride, int on_device, const uint8_t *useme,
                          uint8_t *out, int out_on_device, int *nkeep_out)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  const uint8_t *d = nullptr; long long ds = 0; uint8_t *owned = nullptr, *d_out = nullptr; int *d_cols = nullptr;
  int rc = 1;
  std::vector<int> cols;
  for (int c = 0; c < alen; c++) if (useme[c]) cols.push_back(c);
  const int nkeep = (int) cols.size();
  if (nkeep_out) *nkeep_out = nkeep;
  if (nkeep == 0) return 0;
  if (stage_msa(ctx, msa, nseq, alen, row_stride, on_device, &d, &ds, &owned)) return 1;
  PREP_OK(cudaMalloc(&d_cols, sizeof(int) * nkeep));
  PREP_OK(cudaMemcpyAsync(d_cols, cols.data(), sizeof(int) * nkeep, cudaMemcpyHostToDevice, ctx->stream));
  if (!out_on_device) PREP_OK(cudaMalloc(&d_out, (size_t) nseq * nkeep));
  PREP_OK(rsb_launch_column_subset(d, nseq, ds, d_cols, nkeep, out_on_device ? out : d_out, ctx->stream));
  ctx->launches++;
  if (!out_on_device) PREP_OK(cudaMemcpyAsync(out, d_out, (size_t) nseq * nkeep, cudaMemcpyDeviceToHost, ctx->stream));
  PREP_OK(cudaStreamSynchronize(ctx->stream));
  rc = 0;
done:
  cudaFree(owned); cudaFree(d_cols); cudaFree(d_out);
  return rc;
}
#undef PREP_OK

// ---------------------------------------------------------------------------------------------- communicator (multi-GPU)
int rsb_comm_id(uint8_t *id128)
{
  rsb_nccl::Api &a = rsb_nccl::api();
  if (!a.ok) { snprintf(g_create_err, sizeof(g_create_err), "NCCL is not available: %s", a.why); return 1; }
  rsb_nccl::ncclUniqueId id;
  const int rc = a.GetUniqueId(&id);
  if (rc != rsb_nccl::ncclSuccess) { snprintf(g_create_err, sizeof(g_create_err), "ncclGetUniqueId: %s", a.GetErrorString(rc)); return 1; }
  memcpy(id128, id.internal, 128);
  return 0;
}

int rsb_comm_init(rsb_ctx *ctx, const uint8_t *id128, int nranks, int rank)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  rsb_nccl::Api &a = rsb_nccl::api();
  if (!a.ok) { rsb_set_error(ctx, "NCCL is not available: %s", a.why); return 1; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { rsb_set_error(ctx, "bad rank %d of %d", rank, nranks); return 1; }
  if (ctx->comm) { cudaDeviceSynchronize(); peer_teardown(ctx); a.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  rsb_nccl::ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  const int rc = a.CommInitRank(&ctx->comm, nranks, id, rank);
  if (rc != rsb_nccl::ncclSuccess) { ctx->comm = nullptr; rsb_set_error(ctx, "ncclCommInitRank: %s", a.GetErrorString(rc)); return 1; }
  ctx->comm_rank = rank; ctx->comm_size = nranks;
  return 0;
}

/* one process driving several devices: a communicator over the n contexts (one per device), rank k = ctxs[k] */
int rsb_comm_init_all(rsb_ctx **ctxs, int n)
{
  if (n < 1 || !ctxs || !ctxs[0]) return 1;
  rsb_nccl::Api &a = rsb_nccl::api();
  if (!a.ok) { rsb_set_error(ctxs[0], "NCCL is not available: %s", a.why); return 1; }
  std::vector<int> devs(n);
  std::vector<rsb_nccl::ncclComm_t> comms(n, nullptr);
  for (int k = 0; k < n; k++) {
    devs[k] = ctxs[k]->device;
    for (int q = 0; q < k; q++) if (devs[q] == devs[k]) { rsb_set_error(ctxs[0], "rsb_comm_init_all: device %d appears twice", devs[k]); return 1; }
  }
  const int rc = a.CommInitAll(comms.data(), n, devs.data());
  if (rc != rsb_nccl::ncclSuccess) { rsb_set_error(ctxs[0], "ncclCommInitAll: %s", a.GetErrorString(rc)); return 1; }
  for (int k = 0; k < n; k++) {
    if (ctxs[k]->comm) { cudaSetDevice(devs[k]); cudaDeviceSynchronize(); peer_teardown(ctxs[k]); a.CommDestroy(ctxs[k]->comm); }
    ctxs[k]->comm = comms[k]; ctxs[k]->comm_rank = k; ctxs[k]->comm_size = n;
  }
  // exchange blocks of the one-shot all-reduce: inside one process the devices map each other's memory directly
  bool peer = n >= 2 && n <= RSB_PEER_MAX && !(getenv("RSCAPE_B200_PEER_REDUCE") && atoi(getenv("RSCAPE_B200_PEER_REDUCE")) == 0);
  size_t cap = 0;
  for (int k = 0; k < n && peer; k++) {
    if (ctxs[k]->L <= 0) peer = false;
    cap = std::max(cap, (size_t) 4 * ctxs[k]->L * std::max(1, ctxs[k]->Rcap) + 64);
    for (int q = 0; q < n && peer; q++) {
      int can = 0;
      if (q != k && (cudaDeviceCanAccessPeer(&can, devs[k], devs[q]) != cudaSuccess || !can)) peer = false;
    }
  }
  void *blocks[RSB_PEER_MAX] = { nullptr };
  for (int k = 0; k < n && peer; k++) {
    cudaSetDevice(devs[k]);
    for (int q = 0; q < n; q++) if (q != k) { const cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false; cudaGetLastError(); }
    const size_t bytes = rsb_peer_block_bytes(n, cap);
    if (peer && (cudaMalloc(&ctxs[k]->peer_block, bytes) != cudaSuccess || cudaMemset(ctxs[k]->peer_block, 0, bytes) != cudaSuccess)) { cudaGetLastError(); peer = false; }
    blocks[k] = ctxs[k]->peer_block;
  }
  for (int k = 0; k < n; k++) {
    cudaSetDevice(devs[k]);
    if (peer) { cudaDeviceSynchronize(); peer_fill_view(ctxs[k], blocks, cap); ctxs[k]->peer_state = 1; }
    else { peer_teardown(ctxs[k]); ctxs[k]->peer_state = -1; }
  }
  return 0;
}

/* how the small per-scan vectors are reduced: *peer_path = 1 one-shot kernel over peer memory, 0 NCCL (or not decided yet);
 * *reductions = all-reduces done by that kernel so far */
int rsb_comm_info(rsb_ctx *ctx, int *nranks, int *rank, int *peer_path, int64_t *reductions)
{
  if (nranks) *nranks = ctx->comm ? ctx->comm_size : 1;
  if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
  if (peer_path) *peer_path = ctx->peer_state == 1 ? 1 : 0;
  if (reductions) *reductions = (int64_t) ctx->peer_reductions;
  return 0;
}

/* Self-test and latency of the small all-reduce (collective: every rank calls it with the same arguments).  A vector of `count`
 * doubles with rank-dependent contents is summed over the ranks once and checked against the closed form (*max_err = largest
 * deviation; 0 expected: the values are small integers), then all-reduced `iters` times (MAX, so that the values stay put)
 * between two CUDA events: *us_per_allreduce = mean device time of one all-reduce on the path rsb_comm_info reports. */
int rsb_comm_selftest(rsb_ctx *ctx, int count, int iters, double *us_per_allreduce, double *max_err)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->comm || ctx->comm_size < 2) { rsb_set_error(ctx, "rsb_comm_selftest: no communicator of at least two ranks"); return 1; }
  if (count < 1 || iters < 1) { rsb_set_error(ctx, "rsb_comm_selftest: bad arguments"); return 1; }
  std::vector<double> h((size_t) count);
  for (int i = 0; i < count; i++) h[i] = (double) ((ctx->comm_rank + 1) * (i % 7 + 1));
  // (no cudaMalloc / cudaFree here: with peer access enabled they synchronise the peer devices, and a peer whose all-reduce kernel
  // is waiting for THIS rank would never go idle.  The plan's scratch matrix serves.)
  if (ctx->L <= 0 || !ctx->d_tmp || (size_t) count > (size_t) ctx->L * ctx->Lp) { rsb_set_error(ctx, "rsb_comm_selftest: configure a plan with L * L >= count first"); return 1; }
  double *d = ctx->d_tmp;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = 1;
  float ms = 0.f;
  double err = 0.0;
  const double tot = 0.5 * ctx->comm_size * (ctx->comm_size + 1);
  if (cudaMemcpyAsync(d, h.data(), sizeof(double) * (size_t) count, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) goto done;
  if (comm_allreduce(ctx, d, (size_t) count, rsb_nccl::ncclFloat64, rsb_nccl::ncclSum, ctx->stream)) goto done;
  if (cudaMemcpyAsync(h.data(), d, sizeof(double) * (size_t) count, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) goto done;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) goto done;
  for (int i = 0; i < count; i++) err = std::max(err, std::fabs(h[i] - tot * (i % 7 + 1)));
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) goto done;
  for (int k = 0; k < 3; k++) if (comm_allreduce(ctx, d, (size_t) count, rsb_nccl::ncclFloat64, rsb_nccl::ncclMax, ctx->stream)) goto done;   // warm-up
  cudaEventRecord(e0, ctx->stream);
  for (int k = 0; k < iters; k++) if (comm_allreduce(ctx, d, (size_t) count, rsb_nccl::ncclFloat64, rsb_nccl::ncclMax, ctx->stream)) goto done;
  cudaEventRecord(e1, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) goto done;
  cudaEventElapsedTime(&ms, e0, e1);
  if (us_per_allreduce) *us_per_allreduce = 1e3 * (double) ms / iters;
  if (max_err) *max_err = err;
  rc = 0;
done:
  if (rc && !ctx->err[0]) rsb_set_error(ctx, "rsb_comm_selftest: %s", cudaGetErrorString(cudaGetLastError()));
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

int rsb_comm_destroy(rsb_ctx *ctx)
{
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    peer_teardown(ctx);
    rsb_nccl::api().CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_size = 1; ctx->comm_rank = 0;
  }
  return 0;
}

/* Null alignments generated by ONE rank made resident on all of them: pool entries [first_rep, first_rep + nrep) are broadcast from
 * rank `root` over the communicator (NVLink), on the generation stream -- behind the generator calls already queued there -- and
 * the entries' readiness events are renewed, so scans of those entries wait for the copy.  With a sharded pair grid every rank
 * scans every null; each rank then generates 1/W of them and the blocks are exchanged (1.4 GB in all at the LSU shape) instead of
 * every rank generating all of them.  Every rank calls this for every block, in the same order. */
int rsb_pool_broadcast(rsb_ctx *ctx, int first_rep, int nrep, int root)
{
  RSB_RANGE("rsb_pool_broadcast");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (pool_range_ok(ctx, first_rep, nrep)) return 1;
  rsb_nccl::Api &a = rsb_nccl::api();
  if (!ctx->comm || !a.Broadcast) { rsb_set_error(ctx, "rsb_pool_broadcast: no communicator (rsb_comm_init)"); return 1; }
  if (root < 0 || root >= ctx->comm_size) { rsb_set_error(ctx, "rsb_pool_broadcast: bad root %d", root); return 1; }
  cudaStream_t sg = ctx->stream_gen;
  RSB_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));          // the caller's stream may still read these entries
  RSB_CUDA_OK(cudaStreamWaitEvent(sg, ctx->ev_entry, 0));
  const size_t rb = (size_t) ctx->N * ctx->L;
  uint8_t *p = ctx->d_pool + (size_t) first_rep * rb;
  const int rc = a.Broadcast(p, p, rb * (size_t) nrep, 0 /* ncclInt8 */, root, ctx->comm, sg);
  if (rc != rsb_nccl::ncclSuccess) { rsb_set_error(ctx, "ncclBroadcast: %s", a.GetErrorString(rc)); return 1; }
  cudaEvent_t e = ctx->gen_ring[ctx->gen_next++ % ctx->gen_ring.size()];
  RSB_CUDA_OK(cudaEventRecord(e, sg));
  for (int r = first_rep; r < first_rep + nrep; r++) ctx->pool_ready[r] = e;
  ctx->launches++;
  return 0;
}

/* null_add2cumranklist across ranks (src/R-scape.c:1565-1612): sum the first nb bins of the device histograms of all ranks in place */
int rsb_hist_allreduce(rsb_ctx *ctx, int nb)
{
  RSB_RANGE("rsb_hist_allreduce");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (nb < 1 || nb > HIST_BINS) { rsb_set_error(ctx, "rsb_hist_allreduce: bad bin count"); return 1; }
  if (!ctx->comm) { rsb_set_error(ctx, "rsb_hist_allreduce: no communicator (rsb_comm_init)"); return 1; }
  if (comm_allreduce(ctx, ctx->d_hist, (size_t) nb, rsb_nccl::ncclUint64, rsb_nccl::ncclSum, ctx->stream)) return 1;
  unsigned long long *d_n = nullptr;                                  // the number of scores travels the same way
  RSB_CUDA_OK(cudaMalloc(&d_n, sizeof(unsigned long long)));
  RSB_CUDA_OK(cudaMemcpyAsync(d_n, &ctx->hist_n, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  int rc = comm_allreduce(ctx, d_n, 1, rsb_nccl::ncclUint64, rsb_nccl::ncclSum, ctx->stream);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(&ctx->hist_n, d_n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { rsb_set_error(ctx, "rsb_hist_allreduce: %s", cudaGetErrorString(e)); rc = 1; }
  }
  cudaFree(d_n);
  return rc;
}

/* global (min over ranks of lo, max over ranks of hi, min over ranks of aux): the score range of a rank's nulls and e.g. the
 * histogram width its replicate 0 asks for */
int rsb_comm_range(rsb_ctx *ctx, double *lo, double *hi, double *aux_min)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->comm || ctx->comm_size <= 1) return 0;
  double h[4] = { -*lo, *hi, aux_min ? -*aux_min : 0.0, 0.0 }, *d = nullptr;
  RSB_CUDA_OK(cudaMalloc(&d, sizeof(h)));
  int rc = 1;
  if (cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
      comm_allreduce(ctx, d, 4, rsb_nccl::ncclFloat64, rsb_nccl::ncclMax, ctx->stream) == 0 &&
      cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
      cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
    *lo = -h[0]; *hi = h[1]; if (aux_min) *aux_min = -h[2];
    rc = 0;
  } else if (!ctx->err[0]) rsb_set_error(ctx, "rsb_comm_range failed");
  cudaFree(d);
  return rc;
}

/* rsb_scan with the pair grid sharded over the communicator's ranks (BASELINE config 4): every rank contracts and scores the
 * rows it owns, the marginal sums / APC row sums / score range are all-reduced on the device, and the corrected matrix is
 * assembled on every rank by one SUM all-reduce (rows of other ranks are zero).  Probabilities are not exported here. */
int rsb_sharded_scan(rsb_ctx *ctx, const uint8_t *msa, int64_t row_stride, int on_device, int stat, int covclass, int actype,
                     const double *allowpair, double tol, double *cov, double *mincov, double *maxcov)
{
  RSB_RANGE("rsb_sharded_scan");
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (resolve_stat(ctx, stat, covclass)) return 1;
  if (!grid_comm(ctx)) { rsb_set_error(ctx, "rsb_sharded_scan needs rsb_set_shard and a communicator over the shards"); return 1; }
  if (stat == RSB_RAF || stat == RSB_RAFS || stat == RSB_CCF) { rsb_set_error(ctx, "statistic not available with a sharded pair grid"); return 1; }
  if (actype != RSB_APC && actype != RSB_ASC) { rsb_set_error(ctx, "rsb_sharded_scan: APC or ASC"); return 1; }
  if (ensure_geo(ctx, 0)) return 1;
  if (upload_msa(ctx, msa, row_stride, 0, 1, 0, on_device, ctx->stream)) return 1;
  if (run_pipeline(ctx, 1, stat, covclass, allow_mask(allowpair), tol)) return 1;          // all-reduces inside (grid_comm)
  if (enqueue_correct(ctx, 0, 1, actype, 1, 0.0, ctx->stream)) return 1;
  if (cov) {
    RSB_CUDA_OK(rsb_launch_zero_unowned_rows(ctx->d_cov, ctx->L, ctx->Lp, ctx->shard_rank, ctx->shard_world, ctx->stream));
    if (comm_allreduce(ctx, ctx->d_cov, (size_t) ctx->L * ctx->Lp, rsb_nccl::ncclFloat64, rsb_nccl::ncclSum, ctx->stream)) return 1;
    RSB_CUDA_OK(rsb_launch_symmetrize(ctx->d_cov, ctx->L, ctx->Lp, ctx->stream));
    ctx->launches += 2;
    if (copy_matrix_out(ctx, ctx->d_cov, cov)) return 1;
  }
  double mmx[2];
  RSB_CUDA_OK(cudaMemcpyAsync(mmx, ctx->d_minmax, sizeof(mmx), cudaMemcpyDeviceToHost, ctx->stream));
  if (check_flags(ctx, "corr_CalculateCOVCorrected")) return 1;
  if (mincov) *mincov = mmx[0];
  if (maxcov) *maxcov = mmx[1];
  return 0;
}

// ---------------------------------------------------------------------------------------------- pinned host buffers
/* Page-lock a host buffer the caller keeps handing to the library (mi->COV, mi->Eval, the pp slab, the residue rows ...): copies
 * to and from a registered buffer run at the PCIe rate and do not block the enqueuing thread; a pageable buffer costs the driver's
 * staging (about half the rate).  Failure is not an error of the path: the buffer simply stays pageable (return 1). */
int rsb_host_register(void *ptr, size_t bytes)
{
  if (!ptr || bytes == 0) return 1;
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return 1; }
  return 0;
}
int rsb_host_unregister(void *ptr)
{
  if (!ptr) return 1;
  if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return 1; }
  return 0;
}

// ---------------------------------------------------------------------------------------------- instrumentation
int rsb_profile_gram(rsb_ctx *ctx, int enable) { ctx->profile = enable != 0; return 0; }

/* contraction launches and their summed CUDA-event time per operand geometry, as of the last rsb_counters call (before its reset):
 * which = 0 input-alignment weights, 1 unit weights (RAF tables, substitution counts), 2 the nulls' own weights (rsb_set_null_slices) */
int rsb_counters_geometry(rsb_ctx *ctx, int which, double *gram_ms, int64_t *gram_launches, int *nslices)
{
  if (which < 0 || which > 2) { rsb_set_error(ctx, "geometry %d out of range", which); return 1; }
  if (gram_ms) *gram_ms = ctx->last_gram_ms_geo[which];
  if (gram_launches) *gram_launches = ctx->last_gram_launches_geo[which];
  if (nslices) *nslices = ctx->geo[which].S;
  return 0;
}

#ifdef RSB_BLOCKTRACE
extern "C" void rsb_trace_set_stats(unsigned long long *buf);
extern "C" void rsb_trace_set_gram(unsigned long long *buf);
static unsigned long long *g_trace = nullptr;
extern "C" int rsb_blocktrace(int on, const char *path)
{
  const size_t bytes = 8 * (1 + 3 * (1ULL << 20));
  if (on) {
    if (!g_trace) cudaMalloc(&g_trace, bytes);
    cudaMemset(g_trace, 0, bytes);
    rsb_trace_set_stats(g_trace); rsb_trace_set_gram(g_trace);
  } else {
    cudaDeviceSynchronize();
    rsb_trace_set_stats(nullptr); rsb_trace_set_gram(nullptr);
    std::vector<unsigned long long> h(bytes / 8);
    cudaMemcpy(h.data(), g_trace, bytes, cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "wb"); if (!f) return 1;
    const size_t n = std::min<unsigned long long>(h[0], 1ULL << 20);
    fwrite(h.data(), 8, 1 + 3 * n, f); fclose(f);
  }
  return 0;
}
#endif

int rsb_counters(rsb_ctx *ctx, int64_t *launches, double *gram_ms, int64_t *gram_launches, int reset)
{
  RSB_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->pending.empty()) {
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < ctx->pending.size(); k++) {
      auto &p = ctx->pending[k];
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) { ctx->gram_ms += ms; ctx->gram_ms_geo[ctx->pending_geo[k]] += ms; }
      cudaEventDestroy(p.first); cudaEventDestroy(p.second);
    }
    ctx->pending.clear(); ctx->pending_geo.clear();
  }
  if (!ctx->pending_aux.empty()) {
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_aux));
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_aux2));
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_aux3));
    RSB_CUDA_OK(cudaStreamSynchronize(ctx->stream_aux4));
    for (size_t k = 0; k < ctx->pending_aux.size(); k++) {
      auto &p = ctx->pending_aux[k]; auto &q = ctx->pending_stage[k];
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) ctx->aux_ms += ms;
      if (cudaEventElapsedTime(&ms, p.first, q.first) == cudaSuccess) ctx->stage_ms[0] += ms;
      if (cudaEventElapsedTime(&ms, q.first, q.second) == cudaSuccess) ctx->stage_ms[1] += ms;
      if (cudaEventElapsedTime(&ms, q.second, p.second) == cudaSuccess) ctx->stage_ms[2] += ms;
      cudaEventDestroy(p.first); cudaEventDestroy(p.second); cudaEventDestroy(q.first); cudaEventDestroy(q.second);
    }
    ctx->pending_aux.clear(); ctx->pending_stage.clear();
  }
  if (launches) *launches = ctx->launches;
  if (gram_ms) *gram_ms = ctx->gram_ms;
  if (gram_launches) *gram_launches = ctx->gram_launches;
  if (getenv("RSCAPE_B200_TRACE")) {
    const double na = ctx->aux_chains ? (double) ctx->aux_chains : 1.0;
    fprintf(stderr, "[rsb] gram %.3f ms x %lld, statistics chain %.3f ms x %lld (statistic %.3f, gap %.3f, reductions+correct+hist %.3f)\n",
            ctx->gram_launches ? ctx->gram_ms / ctx->gram_launches : 0.0, ctx->gram_launches, ctx->aux_ms / na, ctx->aux_chains,
            ctx->stage_ms[0] / na, ctx->stage_ms[1] / na, ctx->stage_ms[2] / na);
  }
  for (int k = 0; k < 3; k++) { ctx->last_gram_ms_geo[k] = ctx->gram_ms_geo[k]; ctx->last_gram_launches_geo[k] = ctx->gram_launches_geo[k]; }
  if (reset) { for (int k = 0; k < 3; k++) { ctx->gram_ms_geo[k] = 0.0; ctx->gram_launches_geo[k] = 0; }
               ctx->launches = 0; ctx->gram_ms = 0.0; ctx->gram_launches = 0; ctx->aux_ms = 0.0; ctx->aux_chains = 0; ctx->stage_ms[0] = ctx->stage_ms[1] = ctx->stage_ms[2] = 0.0; }
  return 0;
}

} // extern "C"
