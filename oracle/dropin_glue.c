/* dropin_glue.c -- the drop-in proof.  TEST INFRASTRUCTURE, built only where the reference tree is present
 * (oracle/Makefile, target _ref/libdropin_b200.so; the prebuilt .so travels to the GPU box).
 *
 * The reference's UNMODIFIED src/covariation.c is compiled where it lies and linked against
 * r-scape_b200/librscape_b200_host.so INSTEAD OF correlators.o: every corr_* call of the real cov_Calculate
 * (src/covariation.c:64-306) and the histogram fill of the real cov_SignificantPairs_Ranking (:326-530) then run on
 * the device-backed implementation, over the reference's own struct data_s / struct mutual_s (real correlators.h).
 * Whatever else covariation.c references (plots, power, CaCoFold, Potts, Easel's fits) becomes an abort-stub
 * (ref_stubs_gen.sh) and is unreachable from the entry below (mode RANSS, no figures).
 *
 * dropin_cov_calculate: flat arrays in, cov_Calculate(data, msa, &ranklist, NULL, NULL, analyze) in between,
 * mi->COV / min / max / mi->type and the rank list's ha histogram out.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rscape_config.h"
#include "easel.h"
#include "correlators.h"
#include "covariation.h"

static char dropin_err[eslERRBUFSIZE];
const char *dropin_errbuf(void) { return dropin_err; }

/* meta: { minCOV, maxCOV, ha->bmin, ha->bmax, ha->w, ha->xmin, ha->xmax }; imeta: { mi->type, mi->class, ha->nb, ha->imin, ha->imax }; n: ha->n */
int
dropin_cov_calculate(int nseq, int L, const uint8_t *res, const double *wgt, int covtype, int covclass, int analyze, double w, double bmin,
                     double *cov_out, double *meta, int *imeta, uint64_t *n_out, uint64_t *bins, int nb_cap)
{
  ESL_ALPHABET     *abc = esl_alphabet_Create(eslRNA);
  ESL_MSA          *msa = esl_msa_CreateDigital(abc, nseq, L);
  ESL_DMATRIX      *allowpair = esl_dmatrix_Create(4, 4);
  struct mutual_s  *mi = NULL;
  struct data_s     data;
  struct outfiles_s ofile;
  THRESH            thresh;
  CLIST             clist;
  RANKLIST         *ranklist = NULL;
  int              *msa2pdb = malloc(sizeof(int) * (size_t) L), *msamap = malloc(sizeof(int) * (size_t) L);
  int               s, i, b, status = eslFAIL;

  dropin_err[0] = 0;
  for (s = 0; s < nseq; s++) { memcpy(msa->ax[s] + 1, res + (size_t) s * L, (size_t) L); msa->wgt[s] = wgt ? wgt[s] : 1.0; }
  esl_dmatrix_Set(allowpair, 0.0);                                          /* WC + GU, src/R-scape.c:883-887 */
  allowpair->mx[0][3] = allowpair->mx[3][0] = allowpair->mx[1][2] = allowpair->mx[2][1] = allowpair->mx[2][3] = allowpair->mx[3][2] = 1.0;
  for (i = 0; i < L; i++) { msa2pdb[i] = i; msamap[i] = i; }                /* R-scape's defaults without a PDB file */
  memset(&data, 0, sizeof(data)); memset(&ofile, 0, sizeof(ofile)); memset(&thresh, 0, sizeof(thresh)); memset(&clist, 0, sizeof(clist));
  clist.mind = 1;
  thresh.type = Eval; thresh.val = 0.05;

  mi = corr_Create(L, nseq, FALSE, 8, 50, abc, (COVCLASS) covclass);       /* nseqthresh, alenthresh: src/R-scape.c:2432 */
  if (!mi) { snprintf(dropin_err, sizeof(dropin_err), "corr_Create failed"); goto DONE; }
  data.ofile = &ofile; data.thresh = &thresh; data.clist = &clist; data.msa2pdb = msa2pdb; data.msamap = msamap;
  data.mi = mi; data.covtype = (COVTYPE) covtype; data.covmethod = NONPARAM; data.statsmethod = NULLPHYLO; data.mode = RANSS;
  data.samplesize = SAMPLE_WC; data.nseq = nseq; data.expBP = -1; data.allowpair = allowpair; data.abcisRNA = TRUE;
  data.tol = 1e-6; data.bmin = bmin; data.w = w; data.nofigures = TRUE; data.verbose = FALSE; data.errbuf = dropin_err;

  status = cov_Calculate(&data, msa, &ranklist, NULL, NULL, analyze);      /* the reference's own function */
  if (status != eslOK) goto DONE;

  for (i = 0; i < L; i++) memcpy(cov_out + (size_t) i * L, mi->COV->mx[i], sizeof(double) * (size_t) L);
  meta[0] = mi->minCOV; meta[1] = mi->maxCOV;
  imeta[0] = (int) mi->type; imeta[1] = (int) mi->class; imeta[2] = 0;
  if (ranklist) {
    meta[2] = ranklist->ha->bmin; meta[3] = ranklist->ha->bmax; meta[4] = ranklist->ha->w; meta[5] = ranklist->ha->xmin; meta[6] = ranklist->ha->xmax;
    imeta[2] = ranklist->ha->nb; imeta[3] = ranklist->ha->imin; imeta[4] = ranklist->ha->imax;
    *n_out = ranklist->ha->n;
    for (b = 0; b < ranklist->ha->nb && b < nb_cap; b++) bins[b] = ranklist->ha->obs[b];
  }

 DONE:
  if (ranklist) cov_FreeRankList(ranklist);
  if (mi) corr_Destroy(mi);
  esl_dmatrix_Destroy(allowpair);
  esl_msa_Destroy(msa);
  esl_alphabet_Destroy(abc);
  free(msa2pdb); free(msamap);
  return status;
}


/* --savenull / --givennull (src/R-scape.c:2466, 2480): the reference's own cov_WriteNullHistogram and cov_ReadNullHistogram
 * (src/covariation.c:1845-1866, 1718-1843), unmodified, on a cumulative null histogram as the B200 null loop leaves it
 * (bmin, w, integer bins).  The list is built by the reference's cov_CreateRankList; the file is then read back by the reference's
 * reader.  out_meta: { bmin, bmax, w, xmin, xmax } of the list read back; out_imeta: { nb, imin, imax }; its bins go to out_bins. */
int
dropin_null_histogram_roundtrip(const char *path, double bmin, double w, int nb, const uint64_t *bins,
                                double *out_meta, int *out_imeta, uint64_t *out_n, uint64_t *out_bins, int out_cap)
{
  RANKLIST *rl = cov_CreateRankList(bmin + nb * w, bmin, w), *back = NULL;
  int       b, status = eslFAIL;
  dropin_err[0] = 0;
  if (rl == NULL) return eslEMEM;
  for (b = 0; b < nb && b < rl->ha->nb; b++) {
    rl->ha->obs[b] = bins[b];
    rl->ha->n += bins[b]; rl->ha->Nc += bins[b]; rl->ha->No += bins[b];
    if (bins[b]) { if (b < rl->ha->imin) rl->ha->imin = b; if (b > rl->ha->imax) rl->ha->imax = b; }
  }
  if ((status = cov_WriteNullHistogram((char *) path, rl, dropin_err, FALSE)) != eslOK) goto DONE;
  if ((status = cov_ReadNullHistogram((char *) path, &back, dropin_err, FALSE)) != eslOK) goto DONE;
  out_meta[0] = back->ha->bmin; out_meta[1] = back->ha->bmax; out_meta[2] = back->ha->w; out_meta[3] = back->ha->xmin; out_meta[4] = back->ha->xmax;
  out_imeta[0] = back->ha->nb; out_imeta[1] = back->ha->imin; out_imeta[2] = back->ha->imax;
  *out_n = back->ha->n;
  for (b = 0; b < back->ha->nb && b < out_cap; b++) out_bins[b] = back->ha->obs[b];
 DONE:
  if (back) cov_FreeRankList(back);
  cov_FreeRankList(rl);
  return status;
}
