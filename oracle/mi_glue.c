/* mi_glue.c -- tiny C helpers so that Python tests can drive an implementation of the
 * reference's covariation API (corr_Create / corr_Probs / corr_Calculate* /
 * corr_CalculateCOVCorrected over struct mutual_s and struct data_s) through ctypes without
 * knowing struct layouts.  TEST INFRASTRUCTURE.  Compiled twice:
 *   -DGLUE_REFERENCE : against the reference's own src/correlators.h -> oracle/_ref/librscape_ref.so
 *   (default)        : against include/rscape_compat.h               -> r-scape_b200/librscape_b200_host.so
 * so each library is handled with the header it was built with.
 */
#ifdef GLUE_REFERENCE
#include "rscape_config.h"
#include "easel.h"
#include "correlators.h"
#define MI_CLASS(mi) ((mi)->class)
#else
#include "rscape_compat.h"
#include "rscape_b200_host.h"
#define MI_CLASS(mi) ((mi)->class)
#endif

ESL_ALPHABET *glue_abc_rna(void) { static ESL_ALPHABET *abc = NULL; if (!abc) abc = esl_alphabet_Create(eslRNA); return abc; }

/* residues [nseq][L] (no sentinels) -> digital ESL_MSA with ax[s][1..L] */
ESL_MSA *
glue_msa_create(int nseq, int L, const uint8_t *res, const double *wgt)
{
  ESL_MSA *msa = esl_msa_CreateDigital(glue_abc_rna(), nseq, L);
  int      s;
  if (!msa) return NULL;
  for (s = 0; s < nseq; s++) {
    memcpy(msa->ax[s] + 1, res + (size_t) s * L, (size_t) L);
    msa->wgt[s] = wgt ? wgt[s] : 1.0;
  }
  return msa;
}
void glue_msa_destroy(ESL_MSA *msa) { esl_msa_Destroy(msa); }

/* WC + GU, src/R-scape.c:883-887 */
ESL_DMATRIX *
glue_allowpair_default(void)
{
  ESL_DMATRIX *ap = esl_dmatrix_Create(4, 4);
  esl_dmatrix_Set(ap, 0.0);
  ap->mx[0][3] = ap->mx[3][0] = 1.0;
  ap->mx[1][2] = ap->mx[2][1] = 1.0;
  ap->mx[2][3] = ap->mx[3][2] = 1.0;
  return ap;
}
ESL_DMATRIX *
glue_allowpair_from(const double *v16)
{
  ESL_DMATRIX *ap = esl_dmatrix_Create(4, 4);
  int i;
  for (i = 0; i < 16; i++) ap->mx[0][i] = v16[i];
  return ap;
}

struct data_s *
glue_data_create(struct mutual_s *mi, ESL_DMATRIX *allowpair, int covtype, double tol)
{
  struct data_s *d = calloc(1, sizeof(struct data_s));
  if (!d) return NULL;
  d->mi        = mi;
  d->allowpair = allowpair;
  d->covtype   = (COVTYPE) covtype;
  d->covmethod = NONPARAM;
  d->mode      = RANSS;
  d->tol       = tol;
  d->verbose   = 0;
  d->errbuf    = calloc(1, eslERRBUFSIZE);
  d->bmin      = -10.0;
  d->w         = 0.05;
  return d;
}
void        glue_data_destroy(struct data_s *d) { if (d) { free(d->errbuf); free(d); } }
const char *glue_data_errbuf(struct data_s *d)  { return d->errbuf; }
size_t      glue_sizeof_data(void)              { return sizeof(struct data_s); }
size_t      glue_sizeof_mutual(void)            { return sizeof(struct mutual_s); }

/* flatten the state block; any pointer may be NULL */
void
glue_mi_export(struct mutual_s *mi, double *pp, double *pm, double *ps, double *nseff, double *ngap,
               double *cov, double *minmax, int *type_class)
{
  int L = (int) mi->alen, i, j;
  for (i = 0; i < L; i++) {
    if (pm) memcpy(pm + (size_t) i * 4, mi->pm[i], sizeof(double) * 4);
    if (ps) memcpy(ps + (size_t) i * 5, mi->ps[i], sizeof(double) * 5);
    if (nseff) memcpy(nseff + (size_t) i * L, mi->nseff[i], sizeof(double) * (size_t) L);
    if (ngap)  memcpy(ngap  + (size_t) i * L, mi->ngap[i],  sizeof(double) * (size_t) L);
    if (cov)   memcpy(cov   + (size_t) i * L, mi->COV->mx[i], sizeof(double) * (size_t) L);
    if (pp) for (j = 0; j < L; j++) memcpy(pp + ((size_t) i * L + j) * 16, mi->pp[i][j], sizeof(double) * 16);
  }
  if (minmax) { minmax[0] = mi->minCOV; minmax[1] = mi->maxCOV; }
  if (type_class) { type_class[0] = (int) mi->type; type_class[1] = (int) MI_CLASS(mi); }
}

#ifndef GLUE_REFERENCE
/* drive null_rscape_b200 (batched null loop of the host mirror) from flat arrays; returns the cumulative
 * rank list's geometry and bins.  meta: { bmin, bmax, w, xmin, xmax }, imeta: { nb, imin, imax }, counts: { n, Nc, No } */
int
glue_null_rscape(struct data_s *data, int nnull, int nseq, int L, const uint8_t *nulls, const double *wgt, int hpts,
                 double *meta, int *imeta, uint64_t *counts, uint64_t *bins, int nb_cap)
{
  ESL_MSA **arr = malloc(sizeof(ESL_MSA *) * (size_t) nnull);
  RANKLIST *cum = NULL;
  int r, b, status;
  for (r = 0; r < nnull; r++) arr[r] = glue_msa_create(nseq, L, nulls + (size_t) r * nseq * L, wgt);
  status = null_rscape_b200(data, arr, nnull, hpts, &cum);
  if (status == eslOK && cum) {
    meta[0] = cum->ha->bmin; meta[1] = cum->ha->bmax; meta[2] = cum->ha->w; meta[3] = cum->ha->xmin; meta[4] = cum->ha->xmax;
    imeta[0] = cum->ha->nb; imeta[1] = cum->ha->imin; imeta[2] = cum->ha->imax;
    counts[0] = cum->ha->n; counts[1] = cum->ha->Nc; counts[2] = cum->ha->No;
    for (b = 0; b < cum->ha->nb && b < nb_cap; b++) bins[b] = cum->ha->obs[b];
    cov_FreeRankList(cum);
  } else if (status == eslOK) { imeta[0] = 0; }
  for (r = 0; r < nnull; r++) esl_msa_Destroy(arr[r]);
  free(arr);
  return status;
}
/* drive cov_CreateHitList_b200 from flat arrays.  geom: { bmin, w, xmax, phi }, ig: { nb, imin, imax }; survfit may be NULL,
 * obs == NULL means "no null rank list" (the naive method).  Returns the Easel status; *nhit = length of the list, of which
 * the first min(*nhit, cap) entries are copied out; eval_out gets mi->Eval. */
int
glue_hitlist(struct data_s *data, struct mutual_s *mi, const double *geom, const int *ig, uint64_t *obs, double *survfit,
             uint64_t Nb, uint64_t Nt, int expBP, double thresh, const uint8_t *pairmask, int64_t cap,
             int64_t *hi, int64_t *hj, double *sc, double *ev, double *pv, double *eval_out, int64_t *nhit)
{
  ESL_HISTOGRAM ha, hb, ht;
  RANKLIST      null, rl;
  THRESH        th;
  HITLIST      *hl = NULL;
  int64_t       h, L = mi->alen;
  int           status, i;

  memset(&ha, 0, sizeof(ha)); memset(&hb, 0, sizeof(hb)); memset(&ht, 0, sizeof(ht));
  memset(&null, 0, sizeof(null)); memset(&rl, 0, sizeof(rl));
  if (obs) {
    ha.bmin = geom[0]; ha.w = geom[1]; ha.xmax = geom[2]; ha.phi = geom[3];
    ha.nb = ig[0]; ha.imin = ig[1]; ha.imax = ig[2]; ha.bmax = ha.bmin + ha.w * ha.nb; ha.obs = obs;
    for (i = 0; i < ha.nb; i++) { ha.Nc += obs[i]; ha.No += obs[i]; ha.n += obs[i]; }
    null.ha = &ha; null.survfit = survfit;
  }
  hb.Nc = Nb; ht.Nc = Nt;
  rl.hb = &hb; rl.ht = &ht;
  th.type = Eval; th.val = thresh; th.sc_bp = th.sc_nbp = 0.;
  data->ranklist_null = obs ? &null : NULL;
  data->thresh = &th;
  data->expBP  = expBP;
  status = cov_CreateHitList_b200(data, mi, &rl, pairmask, &hl);
  data->ranklist_null = NULL; data->thresh = NULL;
  if (status != eslOK) return status;
  *nhit = hl->nhit;
  for (h = 0; h < hl->nhit && h < cap; h++) {
    if (hl->srthit[h] != hl->hit + h) { cov_FreeHitList(hl); return eslFAIL; }
    hi[h] = hl->hit[h].i; hj[h] = hl->hit[h].j; sc[h] = hl->hit[h].sc; ev[h] = hl->hit[h].Eval; pv[h] = hl->hit[h].pval;
  }
  if (eval_out) for (i = 0; i < L; i++) memcpy(eval_out + (size_t) i * L, mi->Eval->mx[i], sizeof(double) * (size_t) L);
  cov_FreeHitList(hl);
  return eslOK;
}
/* drive Tree_Substitutions_b200 from flat arrays: all = [2N-1][L] rows of the Fitch reconstruction */
int
glue_tree_substitutions(int N, const int *left, const int *right, int L, const uint8_t *all, int includegaps, int want_pairs,
                        int *nsubs, int *ndouble, int *njoin, char *errbuf)
{
  ESL_TREE *T = esl_tree_Create(N);
  ESL_MSA  *msa = glue_msa_create(N, L, all, NULL), *allmsa = glue_msa_create(2 * N - 1, L, all, NULL);
  int      *a = NULL, *b = NULL, *c = NULL, v, status;
  for (v = 0; v < N - 1; v++) { T->left[v] = left[v]; T->right[v] = right[v]; }
  status = Tree_Substitutions_b200(msa, allmsa, T, &a, want_pairs ? &b : NULL, want_pairs ? &c : NULL, includegaps, errbuf, 0);
  if (status == eslOK) {
    memcpy(nsubs, a, sizeof(int) * (size_t) L);
    if (want_pairs) { memcpy(ndouble, b, sizeof(int) * (size_t) L * L); memcpy(njoin, c, sizeof(int) * (size_t) L * L); }
  }
  free(a); free(b); free(c);
  esl_msa_Destroy(msa); esl_msa_Destroy(allmsa); esl_tree_Destroy(T);
  return status;
}
double glue_data_w(struct data_s *d) { return d->w; }
int    glue_cov_calculate(struct data_s *d, ESL_MSA *msa) { return cov_CalculateCOV(d, msa); }
#endif
#ifndef GLUE_REFERENCE
/* drive the preprocessing mirrors (host/msaprep_b200.c) from flat arrays: what = 0 msaweight_b200 (GSC | PB by maxsq_gsc),
 * 1 esl_msaweight_PB_b200, 2 esl_msaweight_GSC_b200 -> wgt[nseq]; useme (int[L], may be NULL) <- msamanip_GapColumns_b200 with
 * the INPUT weights; avgid (may be NULL) <- esl_dst_XAverageId_b200(max_comparisons) */
int
glue_msaprep(int nseq, int L, const uint8_t *res, const double *wgt_in, int what, int maxsq_gsc, double gapthresh, int max_comparisons,
             double *wgt_out, int *useme, double *avgid)
{
  ESL_MSA *msa = glue_msa_create(nseq, L, res, wgt_in);
  char     errbuf[eslERRBUFSIZE];
  int      status = eslOK, s;
  if (useme && status == eslOK) status = msamanip_GapColumns_b200(gapthresh, msa, useme, errbuf);
  if (avgid && status == eslOK) status = esl_dst_XAverageId_b200(msa, max_comparisons, avgid);
  if (wgt_out && status == eslOK) {
    status = (what == 0) ? msaweight_b200(msa, maxsq_gsc) : (what == 1) ? esl_msaweight_PB_b200(msa) : esl_msaweight_GSC_b200(msa);
    for (s = 0; s < nseq; s++) wgt_out[s] = msa->wgt[s];
  }
  esl_msa_Destroy(msa);
  return status;
}
#endif

/* Tail fit of a null histogram given as flat arrays: geom = { bmin, w, xmax }, ig = { nb, imin, imax }; n = scores held.
 * out = { newmass, mu, lambda, tau, phi, cmin }; survfit: double [2 nb] (zeros when the fit has no finite rate).
 *   host library : cov_NullFit_b200 (the mirror of src/covariation.c:459-487)
 *   reference    : its own cov_histogram_pmass + cov_NullFitGamma / cov_NullFitExponential (ref_glue_evalue.c) */
#ifndef GLUE_REFERENCE
int
glue_nullfit(const double *geom, const int *ig, uint64_t n, uint64_t *obs, double pmass, double fracfit, int doexpfit, double *survfit, double *out)
{
  ESL_HISTOGRAM h;
  double       *sf = NULL;
  char          errbuf[eslERRBUFSIZE];
  int           status, b;
  memset(&h, 0, sizeof(h));
  h.bmin = geom[0]; h.w = geom[1]; h.xmax = geom[2]; h.nb = ig[0]; h.imin = ig[1]; h.imax = ig[2]; h.cmin = h.imin;
  h.bmax = h.bmin + h.w * h.nb; h.n = h.Nc = h.No = n; h.obs = obs;
  status = cov_NullFit_b200(&h, pmass, fracfit, doexpfit, &sf, &out[0], &out[1], &out[2], &out[3], errbuf);
  if (status != eslOK) return status;
  out[4] = h.phi; out[5] = (double) h.cmin;
  for (b = 0; b < 2 * h.nb; b++) survfit[b] = sf ? sf[b] : 0.0;
  free(sf);
  return eslOK;
}
#endif
