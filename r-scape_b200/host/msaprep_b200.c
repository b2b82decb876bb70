/* msaprep_b200.c -- host-side mirror of the preprocessing that defines the scanned alignment and its weights:
 *
 *   msaweight_b200              msaweight's default branch, src/R-scape.c:1545-1562: GSC for nseq <= maxsq_gsc, else PB
 *   esl_msaweight_PB_b200       Henikoff position-based weights: column counts and per-sequence sums on the device
 *   esl_msaweight_GSC_b200      the N x N distance matrix 1 - pid on the device; UPGMA and the Gerstein/Sonnhammer/Chothia tree
 *                               weights (sequential O(N^3) / O(N) bookkeeping over nseq <= 1000 taxa) on the host
 *   msamanip_GapColumns_b200    the column test of msamanip_RemoveGapColumns, src/msamanip.c:486-500 (the removal of broken base
 *                               pairs and the column subset itself, :503-506, stay with the caller: they edit SS_cons)
 *   esl_dst_XAverageId_b200     esl_dst_XAverageId as called by msamanip_XStats, src/msamanip.c:1967: exhaustive below
 *                               max_comparisons pairs, else that many pairs drawn from a Mersenne Twister seeded with 42
 *
 * Easel is not part of the reference tree; the arithmetic follows SURVEY 9.7 (pinned through the tutorial transcript only).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "rscape_b200_host.h"

static rsb_ctx *prep_ctx = NULL;

/* one lazily created device context for the preprocessing calls (device: RSCAPE_B200_DEVICE, default 0) */
rsb_ctx *
rsb_host_prep_context(char *errbuf)
{
  if (!prep_ctx) {
    const char *env = getenv("RSCAPE_B200_DEVICE");
    if (rsb_create(env ? atoi(env) : 0, NULL, &prep_ctx) != 0) {
      if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "%s", rsb_create_error());
      prep_ctx = NULL;
    }
  }
  return prep_ctx;
}

void
rsb_host_prep_release(void)
{
  if (prep_ctx) { rsb_destroy(prep_ctx); prep_ctx = NULL; }
}

/* residues of a digital ESL_MSA as one block [nseq][alen] (ax rows are separate allocations with sentinels) */
static uint8_t *
flat_residues(const ESL_MSA *msa)
{
  size_t   N = (size_t) msa->nseq, L = (size_t) msa->alen, s;
  uint8_t *buf = malloc((N * L > 0) ? N * L : 1);
  if (!buf) return NULL;
  for (s = 0; s < N; s++) memcpy(buf + s * L, msa->ax[s] + 1, L);
  return buf;
}

int
esl_msaweight_PB_b200(ESL_MSA *msa)
{
  rsb_ctx *ctx = rsb_host_prep_context(NULL);
  uint8_t *res;
  int      status;
  if (!ctx) return eslFAIL;
  if (msa->nseq == 1) { msa->wgt[0] = 1.0; return eslOK; }
  if ((res = flat_residues(msa)) == NULL) return eslEMEM;
  status = rsb_msa_pb_weights(ctx, res, msa->nseq, (int) msa->alen, msa->alen, 0, msa->wgt) == 0 ? eslOK : eslFAIL;
  free(res);
  return status;
}

/* UPGMA on the distance matrix with Easel's bookkeeping (first strict minimum in row-major order, merged pair swapped to the end
 * of the active block), then the GSC weights on the tree.  D [N][N] is destroyed. */
static int
gsc_from_distances(double *D, int N, double *wgt)
{
  int     nn = N - 1, M, i, j, a, b, v, k, status = eslEMEM;
  int    *left = NULL, *right = NULL, *idx = NULL, *nin = NULL, *csize = NULL;
  double *ld = NULL, *rd = NULL, *height = NULL, *x = NULL, tot;

  left = malloc(sizeof(int) * nn); right = malloc(sizeof(int) * nn); idx = malloc(sizeof(int) * N); nin = malloc(sizeof(int) * N);
  csize = malloc(sizeof(int) * nn);
  ld = calloc(nn, sizeof(double)); rd = calloc(nn, sizeof(double)); height = calloc(nn, sizeof(double)); x = calloc(nn, sizeof(double));
  if (!left || !right || !idx || !nin || !csize || !ld || !rd || !height || !x) goto DONE;
  for (i = 0; i < N; i++) { idx[i] = -i; nin[i] = 1; }

  for (M = N; M >= 2; M--) {
    double minD = INFINITY;
    int    mi = 0, mj = 1;
    for (i = 0; i < M; i++)
      for (j = i + 1; j < M; j++)
        if (D[(size_t) i * N + j] < minD) { minD = D[(size_t) i * N + j]; mi = i; mj = j; }
    i = mi; j = mj;
    v = M - 2;
    left[v] = idx[i]; right[v] = idx[j];
    height[v] = minD / 2.0;
    ld[v] = height[v] - (idx[i] > 0 ? height[idx[i]] : 0.0);
    rd[v] = height[v] - (idx[j] > 0 ? height[idx[j]] : 0.0);
    /* swap j -> M-1, then i -> M-2 (rows, columns, idx, nin) */
    for (k = 0; k < 2; k++) {
      a = (k == 0) ? j : i; b = (k == 0) ? M - 1 : M - 2;
      if (a != b) {
        int    t, c;
        double td;
        for (c = 0; c < N; c++) { td = D[(size_t) a * N + c]; D[(size_t) a * N + c] = D[(size_t) b * N + c]; D[(size_t) b * N + c] = td; }
        for (c = 0; c < N; c++) { td = D[(size_t) c * N + a]; D[(size_t) c * N + a] = D[(size_t) c * N + b]; D[(size_t) c * N + b] = td; }
        t = idx[a]; idx[a] = idx[b]; idx[b] = t;
        t = nin[a]; nin[a] = nin[b]; nin[b] = t;
      }
      if (k == 0 && i == M - 1) i = j;                             /* i was moved by the first swap */
    }
    i = M - 2; j = M - 1;
    tot = (double) (nin[i] + nin[j]);
    for (k = 0; k < M; k++) {
      D[(size_t) i * N + k] = ((double) nin[i] * D[(size_t) i * N + k] + (double) nin[j] * D[(size_t) j * N + k]) / tot;
      D[(size_t) k * N + i] = D[(size_t) i * N + k];
    }
    D[(size_t) i * N + i] = 0.0;
    nin[i] += nin[j];
    idx[i] = v;
  }
  /* GSC */
  for (v = nn - 1; v >= 0; v--) {
    x[v] = ld[v] + rd[v];
    if (left[v]  > 0) x[v] += x[left[v]];
    if (right[v] > 0) x[v] += x[right[v]];
    csize[v] = (left[v] > 0 ? csize[left[v]] : 1) + (right[v] > 0 ? csize[right[v]] : 1);
  }
  for (i = 0; i < N; i++) wgt[i] = 0.0;
  x[0] = 0.0;
  for (v = 0; v < nn; v++) {
    double lw = ld[v] + (left[v]  > 0 ? x[left[v]]  : 0.0);
    double rw = rd[v] + (right[v] > 0 ? x[right[v]] : 0.0);
    double lx, rx;
    if (lw + rw == 0.0) {
      double ls = left[v] > 0 ? csize[left[v]] : 1, rs = right[v] > 0 ? csize[right[v]] : 1;
      lx = x[v] * ls / (ls + rs); rx = x[v] * rs / (ls + rs);
    } else { lx = x[v] * lw / (lw + rw); rx = x[v] * rw / (lw + rw); }
    if (left[v]  > 0) x[left[v]]  = lx + ld[v]; else wgt[-left[v]]  = lx + ld[v];
    if (right[v] > 0) x[right[v]] = rx + rd[v]; else wgt[-right[v]] = rx + rd[v];
  }
  tot = 0.0;
  for (i = 0; i < N; i++) tot += wgt[i];
  for (i = 0; i < N; i++) wgt[i] = (tot > 0.0) ? wgt[i] * ((double) N / tot) : 1.0;
  status = eslOK;
 DONE:
  free(left); free(right); free(idx); free(nin); free(csize); free(ld); free(rd); free(height); free(x);
  return status;
}

int
esl_msaweight_GSC_b200(ESL_MSA *msa)
{
  rsb_ctx *ctx = rsb_host_prep_context(NULL);
  uint8_t *res = NULL;
  double  *D = NULL;
  int      N = msa->nseq, status = eslFAIL;
  if (!ctx) return eslFAIL;
  if (N == 1) { msa->wgt[0] = 1.0; return eslOK; }
  res = flat_residues(msa);
  D   = malloc(sizeof(double) * (size_t) N * N);
  if (!res || !D) { status = eslEMEM; goto DONE; }
  if (rsb_msa_pair_identity(ctx, res, N, (int) msa->alen, msa->alen, 0, NULL, 0, D) != 0) goto DONE;     /* 1 - pid for every pair */
  status = gsc_from_distances(D, N, msa->wgt);
 DONE:
  free(res); free(D);
  return status;
}

/* src/R-scape.c:1545-1562, the default (non-gremlin) branch */
int
msaweight_b200(ESL_MSA *msa, int maxsq_gsc)
{
  return (msa->nseq <= maxsq_gsc) ? esl_msaweight_GSC_b200(msa) : esl_msaweight_PB_b200(msa);
}

/* useme[apos] = TRUE iff column apos passes the gap threshold (src/msamanip.c:486-500), with the alignment's current weights */
int
msamanip_GapColumns_b200(double gapthresh, ESL_MSA *msa, int *useme, char *errbuf)
{
  rsb_ctx *ctx = rsb_host_prep_context(errbuf);
  uint8_t *res = NULL, *keep = NULL;
  int      L = (int) msa->alen, c, s, unit = TRUE, status = eslFAIL;
  if (!ctx) return eslFAIL;
  res = flat_residues(msa); keep = malloc(L ? L : 1);
  if (!res || !keep) { status = eslEMEM; goto DONE; }
  for (s = 0; s < msa->nseq && unit; s++) unit = (msa->wgt[s] == 1.0);
  if (rsb_msa_gap_columns(ctx, res, msa->nseq, L, msa->alen, 0, unit ? NULL : msa->wgt, gapthresh, keep) != 0) {
    if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "%s", rsb_error(ctx));
    goto DONE;
  }
  for (c = 0; c < L; c++) useme[c] = keep[c] ? TRUE : FALSE;
  status = eslOK;
 DONE:
  free(res); free(keep);
  return status;
}

int
esl_dst_XAverageId_b200(ESL_MSA *msa, int max_comparisons, double *ret_id)
{
  rsb_ctx *ctx = rsb_host_prep_context(NULL);
  uint8_t *res = NULL;
  int     *pairs = NULL;
  double  *pid = NULL, sum = 0.0;
  int64_t  N = msa->nseq, np, k;
  int      i, j, status = eslFAIL;
  if (!ctx) return eslFAIL;
  if (N <= 1) { *ret_id = 1.0; return eslOK; }
  np = (N * (N - 1) / 2 <= max_comparisons) ? N * (N - 1) / 2 : max_comparisons;
  res = flat_residues(msa); pairs = malloc(sizeof(int) * 2 * (size_t) np); pid = malloc(sizeof(double) * (size_t) np);
  if (!res || !pairs || !pid) { status = eslEMEM; goto DONE; }
  if (N * (N - 1) / 2 <= max_comparisons) {
    for (k = 0, i = 0; i < N; i++) for (j = i + 1; j < N; j++, k++) { pairs[2 * k] = i; pairs[2 * k + 1] = j; }
  } else {                                                            /* a stochastic sample with a fixed seed */
    ESL_RANDOMNESS *r = esl_randomness_Create(42);
    if (!r) { status = eslEMEM; goto DONE; }
    for (k = 0; k < np; k++) {
      do { i = esl_rnd_Roll(r, N); j = esl_rnd_Roll(r, N); } while (j == i);
      pairs[2 * k] = i; pairs[2 * k + 1] = j;
    }
    esl_randomness_Destroy(r);
  }
  if (rsb_msa_pair_identity(ctx, res, (int) N, (int) msa->alen, msa->alen, 0, pairs, np, pid) != 0) goto DONE;
  for (k = 0; k < np; k++) sum += pid[k];                             /* the reference's summation order */
  *ret_id = sum / (double) np;
  status = eslOK;
 DONE:
  free(res); free(pairs); free(pid);
  return status;
}
