/* rscape_b200_host.h -- the reference's covariation API, served by the B200 library.
 *
 * Same names, argument meaning and error behaviour as src/correlators.h:445-482 of the reference, so
 * that src/covariation.c (cov_Calculate), src/R-scape.c, the power and CaCoFold code link against this
 * layer instead of correlators.o and run unchanged on the results.  Implemented in
 * r-scape_b200/host/correlators_b200.c on top of the plain C-ABI of rscape_b200.h.
 *
 * Error convention: int Easel status (eslOK / eslFAIL) with the message written into the caller's
 * errbuf (char[eslERRBUFSIZE]); corr_Create returns NULL.  CUDA failures map to eslFAIL + text.
 *
 * The second half declares the host-side callers of that API that this repo mirrors so that the path can
 * be driven end to end without the rest of R-scape: cov_Calculate's dispatch (src/covariation.c:64-306),
 * the rank-list histograms (:415-457, :641-736, :2334-2362), the E-value / hit-list loop (:828-910) and the null loop
 * (src/R-scape.c:1565-1724), the latter in a batched form (rsb_null_*) that a 20-line edit of null_rscape would call
 * (INTEGRATION.md).
 */
#ifndef RSCAPE_B200_HOST_INCLUDED
#define RSCAPE_B200_HOST_INCLUDED

#include "rscape_compat.h"
#include "rscape_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/correlators.h:445-482 ---- */
extern int              corr_CalculateCHI     (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateCHI_C16 (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateCHI_C2  (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateOMES    (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateOMES_C16(struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateOMES_C2 (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateGT      (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateGT_C16  (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateGT_C2   (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateGT_CWC  (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateMI      (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateMI_C16  (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateMI_C2   (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateMIr     (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateMIr_C16 (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateMIr_C2  (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateMIg     (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateMIg_C16 (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateMIg_C2  (struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf);
extern int              corr_CalculateRAF     (COVCLASS covclass, struct data_s *data, ESL_MSA *msa);
extern int              corr_CalculateRAFS    (COVCLASS covclass, struct data_s *data, ESL_MSA *msa);
extern int              corr_CalculateCCF     (COVCLASS covclass, struct data_s *data);
extern int              corr_CalculateCCF_C16 (struct mutual_s *mi,                         int verbose, char *errbuf);
extern int              corr_CalculateCOVCorrected(ACTYPE actype, struct data_s *data, int shiftnonneg);
extern struct mutual_s *corr_Create(int64_t alen, int64_t nseq, int isshuffled, int nseqthresh, int thresh, ESL_ALPHABET *abc, COVCLASS covclass);
extern int              corr_Reuse(struct mutual_s *mi, int ishuffled, COVTYPE mitype, COVCLASS miclass);
extern int              corr_ReuseCOV(struct mutual_s *mi, COVTYPE mitype, COVCLASS covclass);
extern void             corr_Destroy(struct mutual_s *mi);
extern int              corr_NaivePP(ESL_RANDOMNESS *r, ESL_MSA *msa, struct mutual_s *mi, double tol, int verbose, char *errbuf);
extern int              corr_NaivePS(ESL_RANDOMNESS *r, ESL_MSA *msa, struct mutual_s *mi, double tol, int verbose, char *errbuf);
extern int              corr_Marginals(struct mutual_s *mi, double tol, int verbose, char *errbuf);
extern int              corr_PostOrderPP(ESL_MSA *msa, ESL_TREE *T, struct ribomatrix_s *ribosum, struct mutual_s *mi,
                                         double tol, int verbose, char *errbuf);
extern int              corr_Probs(ESL_RANDOMNESS *r, ESL_MSA *msa, ESL_TREE *T, struct ribomatrix_s *ribosum, struct mutual_s *mi,
                                   METHOD method, double tol, int verbose, char *errbuf);
extern int              corr_ValidateProbs(struct mutual_s *mi, double tol, int verbose, char *errbuf);
extern int              corr_COVTYPEString(char **ret_covtype, COVTYPE type, char *errbuf);
extern int              corr_String2COVTYPE(char *covtype, COVTYPE *ret_type, char *errbuf);
extern int              corr_THRESHTYPEString(char **ret_threshtype, THRESHTYPE type, char *errbuf);

/* the device context behind a mutual_s (created by corr_Create, freed by corr_Destroy) */
extern rsb_ctx         *corr_b200_context(struct mutual_s *mi);

/* ---- callers mirrored from src/covariation.c and src/R-scape.c ---- */
/* covariation matrix part of cov_Calculate, src/covariation.c:78-258 (no ranking, plots or power) */
extern int        cov_CalculateCOV(struct data_s *data, ESL_MSA *msa);
/* src/covariation.c:641-665, :683-736, :2334-2362 */
extern RANKLIST  *cov_CreateRankList(double bmax, double bmin, double w);
extern int        cov_GrowRankList(RANKLIST **oranklist, double bmax, double bmin);
extern void       cov_FreeRankList(RANKLIST *ranklist);
extern int        cov_ranklist_Bin2Bin(int b, ESL_HISTOGRAM *h, ESL_HISTOGRAM *newh, int *ret_newb);
/* The histogram fill of cov_SignificantPairs_Ranking, src/covariation.c:415-457, from mi->COV on the host: every pair that passes
 * the PDB-distance rule (:421-427, data->msa2pdb and the contact list's mind) goes to ha; in GIVSS / FOLDSS mode also to hb when
 * pairmask flags it as a member of the structure set chosen by data->samplesize (never for SAMPLE_ALL), else to ht.
 * pairmask: uint8 [alen][alen], entries i<j, or NULL for an empty structure (what CMAP_Is*Local return on an empty contact list). */
extern int        cov_RankListFromCOV_b200(struct data_s *data, const uint8_t *pairmask, RANKLIST **ret_ranklist);
/* the same with no structure mask */
extern int        cov_RankListFromCOV(struct data_s *data, RANKLIST **ret_ranklist);
/* src/R-scape.c:1565-1612 */
extern int        null_add2cumranklist(RANKLIST *ranklist, RANKLIST **ocumranklist, int verbose, char *errbuf);

/* Batched replacement for the body of null_rscape's loop (src/R-scape.c:1650-1697): nulls[r] are alignments with
 * msa->nseq rows and msa->alen columns (weights are taken from data->mi's last corr_Probs call / msa->wgt, quirk Q1).
 * The first null fixes the histogram width (calculate_width_histo), every null is scanned and added to the cumulative
 * rank list, which is returned in the reference's own RANKLIST form.  data->w is updated as cfg->w is at :1359. */
extern int        null_rscape_b200(struct data_s *data, ESL_MSA **nulls, int nnull, int hpts, RANKLIST **ret_cumranklist);

/* The per-pair loop of cov_CreateHitList (src/covariation.c:828-910) on the device: mi->Eval and the significant pairs
 * (i, j, sc, Eval, pval) from mi->COV, data->ranklist_null (cumulative null histogram + fitted tail), ranklist->hb/ht->Nc,
 * data->expBP and data->thresh->val.  pairmask: uint8 [alen][alen], nonzero = the pair belongs to the structure set selected by
 * data->samplesize (built once by the caller from data->clist), or NULL.  The file output, power and CaCoFold parts of the
 * reference function (:911-1006) stay host code on the returned list. */
extern int        cov_CreateHitList_b200(struct data_s *data, struct mutual_s *mi, RANKLIST *ranklist, const uint8_t *pairmask,
                                         HITLIST **ret_hitlist);
extern void       cov_FreeHitList(HITLIST *hitlist);
/* "Histogram and Fit" of cov_SignificantPairs_Ranking, src/covariation.c:459-487: choose the censored tail mass
 * (cov_histogram_pmass, :2484-2505, a `static` of the reference), fit an exponential (data->doexpfit) or a gamma to the tail of
 * the cumulative null histogram h (cov_NullFitExponential / cov_NullFitGamma, :1915-1973) and tabulate the fitted survival
 * (cov_histogram_SetSurvFitTail, :1677-1699).  *ret_survfit: double[2 h->nb], malloc'd, NULL when the fit has no finite rate.
 * Sets h->phi / cmin / z as esl_histogram_SetTailByMass does.  Host arithmetic on O(bins) data; the device never sees it. */
extern int        cov_NullFit_b200(ESL_HISTOGRAM *h, double pmass, double fracfit, int doexpfit, double **ret_survfit, double *ret_newmass,
                                   double *ret_mu, double *ret_lambda, double *ret_tau, char *errbuf);
/* pair mask (uint8 [alen][alen], entries i<j) of the base pairs of a ct array in Easel's convention (1-based, 0 = unpaired) */
extern int        cov_PairMaskFromCT(const int *ct, int64_t alen, uint8_t *pairmask);

/* Tree_Substitutions (src/msatree.c:1423-1554) from its Fitch reconstruction on: allmsa holds the 2N-1 rows written by
 * Tree_FitchAlgorithmAncenstral (:1451; leaves first, internal node v at row N+v).  Same outputs and allocation as the
 * reference (nsubs int[alen]; ndouble, njoin int[alen*alen], entries i<j), computed on the device. */
extern int        Tree_Substitutions_b200(ESL_MSA *msa, ESL_MSA *allmsa, ESL_TREE *T, int **ret_nsubs, int **ret_ndouble, int **ret_njoin,
                                          int includegaps, char *errbuf, int verbose);

/* ---- preprocessing that defines the scanned alignment and its weights (r-scape_b200/host/msaprep_b200.c) ---- */
/* msaweight's default branch, src/R-scape.c:1545-1562: esl_msaweight_GSC for nseq <= maxsq_gsc, else esl_msaweight_PB; fills msa->wgt */
extern int        msaweight_b200(ESL_MSA *msa, int maxsq_gsc);
extern int        esl_msaweight_PB_b200(ESL_MSA *msa);
extern int        esl_msaweight_GSC_b200(ESL_MSA *msa);
/* the column test of msamanip_RemoveGapColumns, src/msamanip.c:486-500: useme int[alen] */
extern int        msamanip_GapColumns_b200(double gapthresh, ESL_MSA *msa, int *useme, char *errbuf);
/* esl_dst_XAverageId as msamanip_XStats calls it (max_comparisons = 10000), src/msamanip.c:1967: fraction, not percent */
extern int        esl_dst_XAverageId_b200(ESL_MSA *msa, int max_comparisons, double *ret_id);
/* the device context these calls share (created on first use, RSCAPE_B200_DEVICE), and its release */
extern rsb_ctx   *rsb_host_prep_context(char *errbuf);
extern void       rsb_host_prep_release(void);

#ifdef __cplusplus
}
#endif
#endif
