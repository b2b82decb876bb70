/* covariation_b200.c -- host-side callers of the covariation API, mirrored from the reference so that the
 * hot path can be driven end to end:
 *
 *   cov_CalculateCOV        the covariation-matrix part of cov_Calculate, src/covariation.c:78-258
 *   cov_CreateRankList ...  rank-list (score histogram) bookkeeping, src/covariation.c:641-736, :2334-2362
 *   cov_RankListFromCOV     the "ha" fill of cov_SignificantPairs_Ranking, src/covariation.c:415-432
 *   null_add2cumranklist    src/R-scape.c:1565-1612
 *   null_rscape_b200        batched form of null_rscape's loop body, src/R-scape.c:1650-1697: all nulls are scanned
 *                           on the device and only the cumulative histogram comes back.
 *   cov_CreateHitList_b200  the per-pair loop of cov_CreateHitList, src/covariation.c:828-910: E-values of every pair
 *                           (mi->Eval) and the list of significant pairs, computed on the device from mi->COV and
 *                           data->ranklist_null.
 *
 * The tail fit of the null histogram, the output files, power and CaCoFold (src/covariation.c:460-530, 911-1006) stay
 * R-scape host code; they consume the RANKLIST / HITLIST / mutual_s produced here unchanged.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include <pthread.h>

#include "rscape_b200_host.h"

/* ------------------------------------------------------------------ cov_Calculate dispatch */
int
cov_CalculateCOV(struct data_s *data, ESL_MSA *msa)
{
  struct mutual_s *mi = data->mi;
  COVCLASS covclass = mi->class;
  int      base, corr, status;

  if (data->covtype >= PTFp) ESL_FAIL(eslFAIL, data->errbuf, "wrong covariation type\n");       /* Potts: out of scope */
  base = ((int) data->covtype / 3) * 3;
  corr = (int) data->covtype % 3;                                                                /* 0 raw, 1 APC (p), 2 ASC (a) */

  if (base != RAF && base != RAFS) {                                                             /* :82-84 */
    status = corr_Probs(data->r, msa, data->T, data->ribosum, mi, data->covmethod, data->tol, data->verbose, data->errbuf);
    if (status != eslOK) return status;
  }
  switch (base) {
  case CHI:  status = corr_CalculateCHI (covclass, data);      break;
  case GT:   status = corr_CalculateGT  (covclass, data);      break;
  case MI:   status = corr_CalculateMI  (covclass, data);      break;
  case MIr:  status = corr_CalculateMIr (covclass, data);      break;
  case MIg:  status = corr_CalculateMIg (covclass, data);      break;
  case OMES: status = corr_CalculateOMES(covclass, data);      break;
  case RAF:  status = corr_CalculateRAF (covclass, data, msa); break;
  case RAFS: status = corr_CalculateRAFS(covclass, data, msa); break;
  case CCF:  status = corr_CalculateCCF (covclass, data);      break;
  default:   ESL_FAIL(eslFAIL, data->errbuf, "wrong covariation type\n");
  }
  if (status != eslOK) return status;
  if (corr == 1) status = corr_CalculateCOVCorrected(APC, data, FALSE);
  if (corr == 2) status = corr_CalculateCOVCorrected(ASC, data, FALSE);
  return status;
}

/* ------------------------------------------------------------------ rank lists */
RANKLIST *
cov_CreateRankList(double bmax, double bmin, double w)
{
  RANKLIST *rl = calloc(1, sizeof(RANKLIST));
  if (!rl) return NULL;
  rl->ha = esl_histogram_CreateFull(bmin, bmax, w);
  rl->ht = esl_histogram_CreateFull(bmin, bmax, w);
  rl->hb = esl_histogram_CreateFull(bmin, bmax, w);
  if (!rl->ha || !rl->ht || !rl->hb) { cov_FreeRankList(rl); return NULL; }
  return rl;
}

void
cov_FreeRankList(RANKLIST *rl)
{
  if (!rl) return;
  esl_histogram_Destroy(rl->ha);
  esl_histogram_Destroy(rl->ht);
  esl_histogram_Destroy(rl->hb);
  free(rl->survfit);
  free(rl);
}

int
cov_ranklist_Bin2Bin(int b, ESL_HISTOGRAM *h, ESL_HISTOGRAM *newh, int *ret_newb)
{
  double x = esl_histogram_Bin2LBound(h, b);
  *ret_newb = -1;
  if (!isfinite(x)) return eslERANGE;
  x = round((x - newh->bmin) / newh->w);
  if (x < (double) INT_MIN || x > (double) INT_MAX) return eslERANGE;
  if ((int) x > newh->nb) return eslERANGE;
  *ret_newb = (int) x;
  return eslOK;
}

static void
copy_hist_meta(const ESL_HISTOGRAM *from, ESL_HISTOGRAM *to)
{
  to->n = from->n; to->xmin = from->xmin; to->xmax = from->xmax; to->imin = from->imin; to->imax = from->imax;
  to->Nc = from->Nc; to->No = from->No;
}

int
cov_GrowRankList(RANKLIST **oranklist, double bmax, double bmin)
{
  RANKLIST *old = *oranklist, *grown;
  double    new_bmin = old->ha->bmin;
  int       b, nb2;

  if (bmin < old->ha->bmin) new_bmin -= fabs(bmin) * 2. * old->ha->w;             /* bmin stays a w-multiple (:694-696) */
  grown = cov_CreateRankList(ESL_MAX(bmax, old->ha->bmax), new_bmin, old->ha->w);
  if (!grown) return eslFAIL;
  copy_hist_meta(old->ha, grown->ha);
  copy_hist_meta(old->ht, grown->ht);
  copy_hist_meta(old->hb, grown->hb);
  for (b = old->ha->imin; b <= old->ha->imax; b++) {
    cov_ranklist_Bin2Bin(b, old->ha, grown->ha, &nb2);
    if (nb2 >= 0 && nb2 < grown->ha->nb) grown->ha->obs[nb2] = old->ha->obs[b];
  }
  cov_FreeRankList(old);
  *oranklist = grown;
  return eslOK;
}

int
cov_RankListFromCOV_b200(struct data_s *data, const uint8_t *pairmask, RANKLIST **ret_ranklist)
{
  struct mutual_s *mi = data->mi;
  RANKLIST *rl;
  double    bmax = mi->maxCOV + 5 * data->w, add;
  int64_t   L = mi->alen, i, j;
  int       mind = data->msa2pdb ? RSB_DATA_MIND(data) : 1;
  int       two_sets = (data->mode == GIVSS || data->mode == FOLDSS);                /* :436 */
  int       select;

  *ret_ranklist = NULL;
  while (fabs(bmax - data->bmin) < data->tol) bmax += data->w;
  if ((rl = cov_CreateRankList(bmax, data->bmin, data->w)) == NULL) ESL_FAIL(eslFAIL, data->errbuf, "rank list allocation failed");
  for (i = 0; i < L - 1; i++)
    for (j = i + 1; j < L; j++) {
      if (data->msa2pdb && data->msa2pdb[i] >= 0 && data->msa2pdb[j] >= 0 && data->msa2pdb[j] - data->msa2pdb[i] < mind) continue;   /* :421-427 */
      add = ESL_MAX(mi->COV->mx[i][j], data->bmin + data->w);
      esl_histogram_Add(rl->ha, add);
      if (two_sets) {
        select = (data->samplesize != SAMPLE_ALL && pairmask) ? (pairmask[(size_t) i * (size_t) L + (size_t) j] != 0) : FALSE;
        esl_histogram_Add(select ? rl->hb : rl->ht, add);
      }
    }
  *ret_ranklist = rl;
  return eslOK;
}

int
cov_RankListFromCOV(struct data_s *data, RANKLIST **ret_ranklist)
{
  return cov_RankListFromCOV_b200(data, NULL, ret_ranklist);
}

int
null_add2cumranklist(RANKLIST *ranklist, RANKLIST **ocumranklist, int verbose, char *errbuf)
{
  RANKLIST *cum;
  int       b, cumb;
  (void) verbose; (void) errbuf;

  if (ranklist == NULL) return eslOK;
  if (*ocumranklist == NULL) {
    *ocumranklist = cum = cov_CreateRankList(ranklist->ha->bmax, ranklist->ha->bmin, ranklist->ha->w);
    if (!cum) return eslFAIL;
    cum->ha->n = ranklist->ha->n; cum->ha->xmin = ranklist->ha->xmin; cum->ha->xmax = ranklist->ha->xmax;
    cum->ha->imin = ranklist->ha->imin; cum->ha->imax = ranklist->ha->imax;
  } else {
    if (cov_GrowRankList(ocumranklist, ranklist->ha->bmax, ranklist->ha->bmin) != eslOK) return eslFAIL;
    cum = *ocumranklist;
    cum->ha->n   += ranklist->ha->n;
    cum->ha->xmin = ESL_MIN(cum->ha->xmin, ranklist->ha->xmin);
    cum->ha->xmax = ESL_MAX(cum->ha->xmax, ranklist->ha->xmax);
    cum->ha->imin = ESL_MIN(cum->ha->imin, ranklist->ha->imin);
    cum->ha->imax = ESL_MAX(cum->ha->imax, ranklist->ha->imax);
  }
  for (b = ranklist->ha->imin; b <= ranklist->ha->imax; b++) {
    cov_ranklist_Bin2Bin(b, ranklist->ha, cum->ha, &cumb);
    if (cumb >= 0 && cumb < cum->ha->nb) {
      cum->ha->obs[cumb] += ranklist->ha->obs[b];
      cum->ha->Nc        += ranklist->ha->obs[b];
      cum->ha->No        += ranklist->ha->obs[b];
    }
  }
  return eslOK;
}

/* ------------------------------------------------------------------ the null loop, batched on the device */
static int
slots_for(int nseq, int alen, int nnull)
{
  /* enough replicates in flight to give every SM a few tiles, bounded by ~24 GB of operand + count planes */
  double tiles = ((double) alen / 32.0) * ((double) alen / 12.0) * 0.5 + 1.0;
  double bytes = 24.0 * alen * (double) nseq + 16.0 * 8.0 * alen * (double) alen + 3.0 * 8.0 * alen * (double) alen + (double) nseq * alen;
  int    r = (int) ceil(4.0 * 148.0 / tiles);
  int    rmem = (int) (24e9 / bytes);
  if (r < 4) r = 4;                     /* four slot groups: the statistics chain of a chunk has three contractions to finish in */
  if (r > rmem) r = rmem;
  if (r > nnull && nnull >= 2) r = nnull;
  if (r > 64) r = 64;
  if (r < 1) r = 1;
  return r;
}

/* One device's share of the null loop: replicates [r0, r1) scanned into that device's histogram with bin width w. */
struct null_worker {
  /* in */
  struct data_s *data; ESL_MSA **nulls; int r0, r1, device, slices, base, cls, corr; double w; const double *ap; int want_last;
  /* out */
  double *minmax;            /* [r1 - r0][2] */
  uint64_t *bins; int nb;    /* the device's histogram up to the bin of its largest score (+ margin) */
  uint64_t n_added;
  int status; char err[eslERRBUFSIZE];
};

static void *
null_worker_run(void *arg)
{
  struct null_worker *wk = arg;
  struct data_s   *data = wk->data;
  struct mutual_s *mi = data->mi;
  rsb_ctx  *ctx = NULL;
  uint8_t  *stage = NULL;
  size_t    N = (size_t) mi->nseq, L = (size_t) mi->alen;
  int       nmine = wk->r1 - wk->r0, slots, batch, r0, r, s, imax;
  double    hi = -eslINFINITY;

  wk->status = eslFAIL; wk->err[0] = 0; wk->bins = NULL; wk->nb = 0; wk->n_added = 0;
  slots = slots_for((int) N, (int) L, nmine);
  if (rsb_create(wk->device, NULL, &ctx) != 0) { snprintf(wk->err, eslERRBUFSIZE, "%s", rsb_create_error()); goto DONE; }
  /* RSCAPE_B200_NULL_SLICES: mixed precision -- the nulls contracted with fewer digit slices of the weights than the input alignment */
  if (getenv("RSCAPE_B200_NULL_SLICES") && rsb_set_null_slices(ctx, atoi(getenv("RSCAPE_B200_NULL_SLICES"))) != 0)
    { snprintf(wk->err, eslERRBUFSIZE, "%s", rsb_error(ctx)); goto DONE; }
  if (rsb_configure(ctx, (int) N, (int) L, slots, wk->slices) != 0 ||
      rsb_set_weights(ctx, wk->nulls[0]->wgt) != 0 ||                              /* nulls carry the input's weights (:1668) */
      (data->msa2pdb && rsb_set_pair_exclusion(ctx, data->msa2pdb, RSB_DATA_MIND(data)) != 0))   /* covariation.c:421-427 */
    { snprintf(wk->err, eslERRBUFSIZE, "%s", rsb_error(ctx)); goto DONE; }
  /* nulls handed to one rsb_null_hist call: the pipeline over the slot groups fills and drains once per call, so a call carries up
   * to 32 of them (at most 1 GB of staging), not just one per slot */
  batch = ESL_MIN(nmine, 32);
  while (batch > slots && (double) batch * (double) N * (double) L > 1e9) batch /= 2;
  if (batch < slots) batch = ESL_MIN(slots, ESL_MAX(nmine, 1));
  if ((stage = malloc((size_t) batch * N * L)) == NULL) { snprintf(wk->err, eslERRBUFSIZE, "allocation failed"); goto DONE; }
  for (r0 = wk->r0; r0 < wk->r1; r0 += batch) {
    int n = ESL_MIN(batch, wk->r1 - r0);
    for (r = 0; r < n; r++)
      for (s = 0; s < (int) N; s++) memcpy(stage + ((size_t) r * N + (size_t) s) * L, wk->nulls[r0 + r]->ax[s] + 1, L);
    if (rsb_null_hist(ctx, stage, n, (int64_t) L, (int64_t) (N * L), 0, wk->base, wk->cls, wk->corr, wk->ap,
                      data->tol, wk->w, data->bmin, wk->minmax + 2 * (r0 - wk->r0)) != 0)
      { snprintf(wk->err, eslERRBUFSIZE, "%s.\nFailed to run null R-scape", rsb_error(ctx)); goto DONE; }
  }
  for (r = 0; r < nmine; r++) hi = ESL_MAX(hi, wk->minmax[2 * r + 1]);
  if (wk->w > 0.) {
    double nbd = ceil((ESL_MAX(hi, data->bmin + wk->w) - data->bmin) / wk->w) + 8.;
    wk->nb = (nbd < 64.) ? 64 : (nbd > 4194304.) ? 4194304 : (int) nbd;
    if ((wk->bins = calloc((size_t) wk->nb, sizeof(uint64_t))) == NULL) { snprintf(wk->err, eslERRBUFSIZE, "allocation failed"); goto DONE; }
    if (rsb_hist_read(ctx, wk->bins, wk->nb, &wk->n_added, &imax) != 0) { snprintf(wk->err, eslERRBUFSIZE, "%s", rsb_error(ctx)); goto DONE; }
  }
  /* quirk Q3: mi keeps the last null's nseff / ngap (read by power_SPAIR_Create, src/power.c:94-95) */
  if (wk->want_last && wk->base != RAF && wk->base != RAFS && rsb_last_nseff(ctx, mi->nseff[0], mi->ngap[0]) != 0)
    { snprintf(wk->err, eslERRBUFSIZE, "%s", rsb_error(ctx)); goto DONE; }
  wk->status = eslOK;
 DONE:
  free(stage);
  if (ctx) rsb_destroy(ctx);
  return NULL;
}

/* The null loop of null_rscape (src/R-scape.c:1650-1697) on RSCAPE_B200_GPUS devices (default 1; "all" = every visible device):
 * contiguous blocks of replicates per device, one host thread and one context per device, the integer histograms summed on the
 * host (tens of KB).  The width pass (calculate_width_histo, :1681-1684) is fused with the scan of replicate 0: every device
 * histograms with the incoming data->w (cfg->w, 0.05) and the loop is repeated only if the width replicate 0 asks for is
 * different -- the scan draws no random numbers, so the result is the reference's either way. */
int
null_rscape_b200(struct data_s *data, ESL_MSA **nulls, int nnull, int hpts, RANKLIST **ret_cumranklist)
{
  struct mutual_s *mi = data->mi;
  struct null_worker *wk = NULL;
  pthread_t *th = NULL;
  RANKLIST  *cum = NULL;
  double    *minmax = NULL, ap[16], w, w_true, bmax = -eslINFINITY, xmin = eslINFINITY, xmax = -eslINFINITY;
  const char *env;
  uint64_t   n_added = 0;
  int        base, corr, cls, device = 0, slices = 0, ngpu = 1, ndev = 1, attempt, k, r, s, b, status = eslFAIL, x, y;

  *ret_cumranklist = NULL;
  if (nnull < 1) return eslOK;
  if (data->covtype >= PTFp) ESL_FAIL(eslFAIL, data->errbuf, "wrong covariation type\n");
  base = ((int) data->covtype / 3) * 3;
  corr = (int) data->covtype % 3;
  corr = (corr == 1) ? RSB_CORR_APC : (corr == 2) ? RSB_CORR_ASC : RSB_CORR_NONE;
  cls  = (int) mi->class;
  if (cls == CSELECT) cls = (mi->nseq <= mi->nseqthresh || mi->alen <= mi->alenthresh) ? C2 : C16;
  if (base == CCF) cls = C16;
  for (x = 0; x < 4; x++) for (y = 0; y < 4; y++) ap[x * 4 + y] = data->allowpair ? data->allowpair->mx[x][y] : 0.0;

  if ((env = getenv("RSCAPE_B200_DEVICE")) != NULL) device = atoi(env);
  if ((env = getenv("RSCAPE_B200_SLICES")) != NULL) slices = atoi(env);
  ndev = rsb_device_count();
  if ((env = getenv("RSCAPE_B200_GPUS")) != NULL) ngpu = (strcmp(env, "all") == 0) ? ndev : atoi(env);
  if (ngpu > ndev && !getenv("RSCAPE_B200_GPUS_OVERSUBSCRIBE")) ngpu = ndev;      /* (tests: several contexts on one device) */
  if (ngpu > nnull) ngpu = nnull;
  if (ngpu < 1) ngpu = 1;

  wk = calloc((size_t) ngpu, sizeof(*wk)); th = calloc((size_t) ngpu, sizeof(*th));
  minmax = malloc(sizeof(double) * 2 * (size_t) nnull);
  if (!wk || !th || !minmax) { snprintf(data->errbuf, eslERRBUFSIZE, "allocation failed"); goto DONE; }

  w = data->w;
  for (attempt = 0; attempt < 2; attempt++) {
    for (k = 0; k < ngpu; k++) {
      int q = nnull / ngpu, extra = nnull % ngpu;
      free(wk[k].bins); wk[k].bins = NULL;
      wk[k].data = data; wk[k].nulls = nulls; wk[k].r0 = k * q + ESL_MIN(k, extra); wk[k].r1 = wk[k].r0 + q + (k < extra ? 1 : 0);
      wk[k].device = (ngpu == 1) ? device : (ndev > 0 ? k % ndev : 0); wk[k].slices = slices; wk[k].base = base; wk[k].cls = cls; wk[k].corr = corr;
      wk[k].w = w; wk[k].ap = data->allowpair ? ap : NULL; wk[k].minmax = minmax + 2 * wk[k].r0; wk[k].want_last = (wk[k].r1 == nnull);
    }
    if (ngpu == 1) null_worker_run(&wk[0]);
    else {
      for (k = 0; k < ngpu; k++) if (pthread_create(&th[k], NULL, null_worker_run, &wk[k]) != 0) { wk[k].status = eslFAIL; snprintf(wk[k].err, eslERRBUFSIZE, "pthread_create failed"); th[k] = 0; }
      for (k = 0; k < ngpu; k++) if (th[k]) pthread_join(th[k], NULL);
    }
    for (k = 0; k < ngpu; k++) if (wk[k].status != eslOK) { snprintf(data->errbuf, eslERRBUFSIZE, "%s", wk[k].err); goto DONE; }
    /* calculate_width_histo on replicate 0's score range (src/R-scape.c:1355-1360) */
    if (!(minmax[1] > data->bmin)) { snprintf(data->errbuf, eslERRBUFSIZE, "bmin %f should be larger than maxCOV %f.\nFailed to calculate the width of the histogram", data->bmin, minmax[1]); goto DONE; }
    w_true = ESL_MIN(data->w, (minmax[1] - ESL_MAX(data->bmin, minmax[0])) / (double) hpts);
    if (w_true < data->tol) w_true = 0.0;
    if (w_true == w) break;
    w = w_true;                                                                    /* rare: the first null spans less than hpts * w */
  }
  data->w = w;

  if (w >= 1e-20) {                                                                /* else "covariation scores are almost constant" (covariation.c:349) */
    /* cumulative rank list in the reference's form: bmax = largest per-replicate maxCOV + 5w (covariation.c:415, :699) */
    for (r = 0; r < nnull; r++) {
      double lo = ESL_MAX(minmax[2 * r], data->bmin + w), hi = ESL_MAX(minmax[2 * r + 1], data->bmin + w), bm = minmax[2 * r + 1] + 5 * w;
      while (fabs(bm - data->bmin) < data->tol) bm += w;
      bmax = ESL_MAX(bmax, bm); xmin = ESL_MIN(xmin, lo); xmax = ESL_MAX(xmax, hi);
    }
    if ((cum = cov_CreateRankList(bmax, data->bmin, w)) == NULL) { snprintf(data->errbuf, eslERRBUFSIZE, "rank list allocation failed"); goto DONE; }
    for (k = 0; k < ngpu; k++) {                                                   /* null_add2cumranklist across devices: integer sums */
      for (b = 0; b < wk[k].nb; b++) {
        if (wk[k].bins[b] == 0) continue;
        if (b >= cum->ha->nb) { snprintf(data->errbuf, eslERRBUFSIZE, "internal: null score beyond the rank list"); goto DONE; }
        cum->ha->obs[b] += wk[k].bins[b]; cum->ha->Nc += wk[k].bins[b]; cum->ha->No += wk[k].bins[b];
      }
      n_added += wk[k].n_added;
    }
    cum->ha->n    = n_added;
    cum->ha->xmin = xmin;
    cum->ha->xmax = xmax;
    esl_histogram_Score2Bin(cum->ha, xmin, &cum->ha->imin);
    esl_histogram_Score2Bin(cum->ha, xmax, &cum->ha->imax);
    (void) s;
  }

  *ret_cumranklist = cum; cum = NULL;
  status = eslOK;

 DONE:
  if (cum) cov_FreeRankList(cum);
  if (wk) for (k = 0; k < ngpu; k++) free(wk[k].bins);
  free(wk); free(th); free(minmax);
  return status;
}

/* ------------------------------------------------------------------ E-values and the significant-pair list */
/* The structure's base pairs as the pair mask of rsb_scan_hist / cov_CreateHitList_b200.  ct is data->ct in Easel's convention
 * (esl_wuss2ct: indices 1..alen, ct[i] = j and ct[j] = i for a pair, 0 = unpaired; ct[0] unused).  Replaces, for a structure
 * given as SS_cons, the per-pair CMAP_IsBPLocal / CMAP_GetBPTYPE scans of src/covariation.c:437-456 and :832-840. */
int
cov_PairMaskFromCT(const int *ct, int64_t alen, uint8_t *pairmask)
{
  int64_t i, j;
  memset(pairmask, 0, (size_t) alen * (size_t) alen);
  for (i = 1; i <= alen; i++) {
    j = ct[i];
    if (j == 0) continue;
    if (j < 1 || j > alen || j == i || ct[j] != i) return eslFAIL;      /* not a consistent pairing */
    if (i < j) pairmask[(size_t) (i - 1) * (size_t) alen + (size_t) (j - 1)] = 1;
  }
  return eslOK;
}

void
cov_FreeHitList(HITLIST *hitlist)
{
  if (!hitlist) return;
  free(hitlist->srthit);
  free(hitlist->hit);
  free(hitlist);
}

/* The loop at src/covariation.c:828-910 on the device.  pairmask (uint8 [alen][alen], may be NULL) flags the pairs that the
 * reference finds with CMAP_GetBPTYPE + data->samplesize (:832-840): building that mask once from data->clist is the
 * caller's O(ncnt) job and replaces a linear scan of the contact list per pair.  Filled per hit: i, j, sc, Eval, pval;
 * nsubs / power (data->spair), bptype and is_compatible (:897-899) are O(nhit) lookups left to the caller, as are
 * data->spair[n].sc / Eval / Pval, which the caller can read from mi->COV / mi->Eval.  mi->Eval is written only when a
 * null rank list exists (:855), like the reference. */
int
cov_CreateHitList_b200(struct data_s *data, struct mutual_s *mi, RANKLIST *ranklist, const uint8_t *pairmask, HITLIST **ret_hitlist)
{
  rsb_ctx  *ctx = corr_b200_context(mi);
  RANKLIST *null = data->ranklist_null;
  HITLIST  *hitlist = NULL;
  int64_t  *hi = NULL, *hj = NULL, nhit = 0, h, P, cap;
  double   *sc = NULL, *ev = NULL, *pv = NULL;
  int64_t   L = mi->alen, i, j;
  int       status = eslFAIL;
  int       all = (data->thresh->val > MAX_EVAL);

  *ret_hitlist = NULL;
  if (!ctx) ESL_FAIL(eslFAIL, data->errbuf, "mutual_s was not created by corr_Create()");
  P = L * (L - 1) / 2;
  if ((hitlist = calloc(1, sizeof(HITLIST))) == NULL) goto ERROR;
  hitlist->Nt = (int64_t) ranklist->ht->Nc;                                         /* :811-812 */
  hitlist->Nb = (int64_t) ranklist->hb->Nc;

  if (null != NULL && ranklist->ht->Nc + ranklist->hb->Nc == 0)                    /* every E-value would be pval * 0 */
    ESL_XFAIL(eslFAIL, data->errbuf, "cov_CreateHitList_b200: the rank list's ht / hb histograms are empty (fill them with cov_RankListFromCOV_b200 in GIVSS or FOLDSS mode)");

  if (null == NULL) {                          /* naive method: E-values carry no significance, every pair has pval = eval = 0 (:844-847) */
    nhit = (0. < data->thresh->val || all) ? P : 0;
    if ((hitlist->hit = calloc((size_t) nhit + 1, sizeof(HIT))) == NULL) goto ERROR;
    if (nhit)
      for (h = 0, i = 0; i < L - 1; i++)
        for (j = i + 1; j < L; j++, h++) { hitlist->hit[h].i = i; hitlist->hit[h].j = j; hitlist->hit[h].sc = mi->COV->mx[i][j]; }
  }
  else {
    rsb_nullfit nf;
    nf.bmin = null->ha->bmin; nf.w = null->ha->w; nf.nb = null->ha->nb; nf.imin = null->ha->imin; nf.imax = null->ha->imax;
    nf.xmax = null->ha->xmax; nf.phi = null->ha->phi; nf.Nc = null->ha->Nc; nf.obs = null->ha->obs; nf.survfit = null->survfit;
    /* the hit list reads whatever mi->COV holds now (:845) */
    if (rsb_load_scores(ctx, mi->COV->mx[0]) != 0) { snprintf(data->errbuf, eslERRBUFSIZE, "%s", rsb_error(ctx)); goto ERROR; }
    /* lists are short unless every pair is reported: start with room for 4096 hits and repeat the call once if there are more
     * (the expBP rule needs its whole first-pass list, so it starts with room for every pair) */
    cap = (all || data->expBP > 0) ? P : 4096;
    for (;;) {
      free(hi); free(hj); free(sc); free(ev); free(pv);
      hi = malloc(sizeof(int64_t) * (size_t) (cap + 1)); hj = malloc(sizeof(int64_t) * (size_t) (cap + 1));
      sc = malloc(sizeof(double) * (size_t) (cap + 1));  ev = malloc(sizeof(double) * (size_t) (cap + 1)); pv = malloc(sizeof(double) * (size_t) (cap + 1));
      if (!hi || !hj || !sc || !ev || !pv) { snprintf(data->errbuf, eslERRBUFSIZE, "allocation failed"); goto ERROR; }
      if (rsb_scan_hits(ctx, &nf, pairmask, ranklist->hb->Nc, ranklist->ht->Nc, data->expBP, data->thresh->val,
                        mi->Eval->mx[0], cap, hi, hj, sc, ev, pv, &nhit) != 0) {
        snprintf(data->errbuf, eslERRBUFSIZE, "%s", rsb_error(ctx));
        goto ERROR;
      }
      if (nhit <= cap) break;
      cap = nhit;
    }
    if ((hitlist->hit = calloc((size_t) nhit + 1, sizeof(HIT))) == NULL) goto ERROR;
    for (h = 0; h < nhit; h++) {
      hitlist->hit[h].i = hi[h]; hitlist->hit[h].j = hj[h]; hitlist->hit[h].sc = sc[h]; hitlist->hit[h].Eval = ev[h]; hitlist->hit[h].pval = pv[h];
    }
  }
  if ((hitlist->srthit = calloc((size_t) nhit + 1, sizeof(HIT *))) == NULL) goto ERROR;
  for (h = 0; h < nhit; h++) hitlist->srthit[h] = hitlist->hit + h;
  hitlist->srthit[0] = hitlist->hit;                                                /* :817 */
  hitlist->nhit = (int) nhit;
  *ret_hitlist = hitlist; hitlist = NULL;
  status = eslOK;

 ERROR:
  free(hi); free(hj); free(sc); free(ev); free(pv);
  if (hitlist) cov_FreeHitList(hitlist);
  return status;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * Tail fit of the cumulative null histogram: the "Histogram and Fit" block of cov_SignificantPairs_Ranking
 * (src/covariation.c:459-487).  O(bins) host arithmetic between the device's histogram and the device's E-value pass. */
static double
histogram_pmass(ESL_HISTOGRAM *h, double target_pmass, double target_fracfit)       /* cov_histogram_pmass, :2484-2505 */
{
  int      i, tp = 0, nfit = 0;
  uint64_t c = 0;
  double   pmass = NAN;                                                            /* stays undefined when no bin holds a score */
  for (i = h->imax; i >= h->imin; i--) if (h->obs[i] > 0) tp++;
  for (i = h->imax; i >= h->imin; i--) {
    c += h->obs[i];
    if (h->obs[i] > 0) {
      nfit++;
      pmass = (double) c / (double) h->Nc;
      if ((double) nfit / (double) tp >= target_fracfit || pmass >= target_pmass) break;
    }
  }
  return pmass;
}

int
cov_NullFit_b200(ESL_HISTOGRAM *h, double pmass_target, double fracfit, int doexpfit, double **ret_survfit, double *ret_newmass,
                 double *ret_mu, double *ret_lambda, double *ret_tau, char *errbuf)
{
  double  ep[3] = { 0.0, 0.0, 0.0 }, newmass = 0.0, pmass, *survfit = NULL;
  int     b, status;

  *ret_survfit = NULL;
  pmass = histogram_pmass(h, pmass_target, fracfit);
  if (isnan(pmass)) ESL_FAIL(eslFAIL, errbuf, "bad Null histogram fit, pmass is nan.");                          /* :470 */
  if (esl_histogram_SetTailByMass(h, pmass, &newmass) != eslOK) ESL_FAIL(eslFAIL, errbuf, "could not set TailByMass");
  if (doexpfit) status = esl_exp_FitCompleteBinned(h, &ep[0], &ep[1]);
  else          status = esl_gam_FitCompleteBinned(h, &ep[0], &ep[1], &ep[2]);
  if (status != eslOK) ESL_FAIL(eslFAIL, errbuf, doexpfit ? "could not do exponential fit" : "could not do a Gamma fit");
  if (!isinf(ep[1])) {                                                                                          /* :1930, :1961 */
    if ((survfit = calloc((size_t) 2 * (size_t) h->nb, sizeof(double))) == NULL) ESL_FAIL(eslEMEM, errbuf, "allocation failed");
    for (b = h->cmin; b < 2 * h->nb; b++) {                                                                     /* cov_histogram_SetSurvFitTail */
      const double bi = esl_histogram_Bin2UBound(h, b);
      survfit[b] = newmass * (doexpfit ? esl_exp_generic_surv(bi, ep) : esl_gam_generic_surv(bi, ep));
    }
  }
  *ret_survfit = survfit;
  if (ret_newmass) *ret_newmass = newmass;
  if (ret_mu)      *ret_mu      = ep[0];
  if (ret_lambda)  *ret_lambda  = ep[1];
  if (ret_tau)     *ret_tau     = ep[2];
  return eslOK;
}
