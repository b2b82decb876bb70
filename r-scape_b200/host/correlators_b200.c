/* correlators_b200.c -- the reference's covariation API (src/correlators.h:445-482) served by the B200 library.
 *
 * Drop-in for the reference's src/correlators.c: same exported names, argument meaning, state carried in
 * struct mutual_s, and error behaviour (Easel status + errbuf).  No arithmetic happens here: every function
 * gathers its inputs, calls the C-ABI of include/rscape_b200.h (sm_100a kernels) and lets it write the results
 * straight into the host fields that R-scape's downstream code reads (mi->pp/pm/ps/nseff/ngap, mi->COV,
 * mi->minCOV/maxCOV, mi->type/class).  There is no CPU fallback: without a B200 corr_Create returns NULL.
 *
 * Memory: the reference mallocs pp[i][j] separately for each of the L^2 pairs (src/correlators.c:1192-1194);
 * here pp, nseff and ngap are single slabs with pointer tables on top, so mi->pp[i][j][k] keeps working for
 * cacofold.c / power.c while a scan lands with one device-to-host copy per field.
 *
 * Environment: RSCAPE_B200_DEVICE (CUDA ordinal, default 0), RSCAPE_B200_SLICES (weight slices 1..6, default auto).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "rscape_b200_host.h"

#ifdef RSB_USE_RSCAPE_HEADERS
#define MI_CLASS(mi) ((mi)->class)
#else
#define MI_CLASS(mi) ((mi)->class)
#endif

/* ------------------------------------------------------------------ side table: mutual_s -> device context */
typedef struct side_s {
  struct mutual_s *mi;
  rsb_ctx         *ctx;
  uint8_t         *stage;       /* contiguous copy of the alignment rows (ESL_MSA keeps one malloc per row) */
  double          *pp_slab, *nseff_slab, *ngap_slab, *pm_slab, *ps_slab;
  int              pinned;       /* the slabs, mi->COV and mi->Eval are page-locked (rsb_host_register) */
  struct side_s   *next;
} SIDE;

static SIDE *side_head = NULL;

static SIDE *
side_of(struct mutual_s *mi)
{
  SIDE *s;
  for (s = side_head; s; s = s->next) if (s->mi == mi) return s;
  return NULL;
}

rsb_ctx *
corr_b200_context(struct mutual_s *mi)
{
  SIDE *s = side_of(mi);
  return s ? s->ctx : NULL;
}

static int
fail(char *errbuf, rsb_ctx *ctx, const char *what)
{
  if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "%s: %s", what, ctx ? rsb_error(ctx) : rsb_create_error());
  return eslFAIL;
}

/* ------------------------------------------------------------------ lifecycle: src/correlators.c:1161-1297 */
struct mutual_s *
corr_Create(int64_t alen, int64_t nseq, int ishuffled, int nseqthresh, int alenthresh, ESL_ALPHABET *abc, COVCLASS covclass)
{
  struct mutual_s *mi = NULL;
  SIDE            *sd = NULL;
  const char      *env;
  size_t           L = (size_t) alen;
  int64_t          i, j;
  int              device = 0, slices = 0;

  if (alen < 1 || nseq < 1 || abc == NULL || abc->K != 4) return NULL;          /* GAPASCHAR 0, K = 4 (rscape_config.h:48) */
  if ((env = getenv("RSCAPE_B200_DEVICE")) != NULL) device = atoi(env);
  if ((env = getenv("RSCAPE_B200_SLICES")) != NULL) slices = atoi(env);

  if ((mi = calloc(1, sizeof(struct mutual_s))) == NULL) return NULL;
  if ((sd = calloc(1, sizeof(SIDE))) == NULL) { free(mi); return NULL; }
  mi->alen = alen; mi->nseq = nseq; mi->nseqthresh = nseqthresh; mi->alenthresh = alenthresh;
  mi->ishuffled = ishuffled; mi->abc = abc;

  sd->pp_slab    = calloc(L * L * 16, sizeof(double));
  sd->nseff_slab = calloc(L * L, sizeof(double));
  sd->ngap_slab  = calloc(L * L, sizeof(double));
  sd->pm_slab    = calloc(L * 4, sizeof(double));
  sd->ps_slab    = calloc(L * 5, sizeof(double));
  sd->stage      = malloc((size_t) nseq * L);
  mi->pp    = malloc(sizeof(double **) * L);
  mi->nseff = malloc(sizeof(double *) * L);
  mi->ngap  = malloc(sizeof(double *) * L);
  mi->pm    = malloc(sizeof(double *) * L);
  mi->ps    = malloc(sizeof(double *) * L);
  if (!sd->pp_slab || !sd->nseff_slab || !sd->ngap_slab || !sd->pm_slab || !sd->ps_slab || !sd->stage ||
      !mi->pp || !mi->nseff || !mi->ngap || !mi->pm || !mi->ps) goto ERROR;
  for (i = 0; i < alen; i++) {
    if ((mi->pp[i] = malloc(sizeof(double *) * L)) == NULL) goto ERROR;
    for (j = 0; j < alen; j++) mi->pp[i][j] = sd->pp_slab + ((size_t) i * L + (size_t) j) * 16;
    mi->nseff[i] = sd->nseff_slab + (size_t) i * L;
    mi->ngap[i]  = sd->ngap_slab  + (size_t) i * L;
    mi->pm[i]    = sd->pm_slab    + (size_t) i * 4;
    mi->ps[i]    = sd->ps_slab    + (size_t) i * 5;
  }
  mi->COV  = esl_dmatrix_Create((int) alen, (int) alen);
  mi->Eval = esl_dmatrix_Create((int) alen, (int) alen);
  if (!mi->COV || !mi->Eval) goto ERROR;

  if (rsb_create(device, NULL, &sd->ctx) != 0) { fprintf(stderr, "corr_Create(): %s\n", rsb_create_error()); goto ERROR; }
  if (rsb_configure(sd->ctx, (int) nseq, (int) alen, 1, slices) != 0) { fprintf(stderr, "corr_Create(): %s\n", rsb_error(sd->ctx)); goto ERROR; }

  /* page-lock what the device writes into on every scan (a pageable target costs about half the copy rate and blocks the
   * enqueuing thread); RSCAPE_B200_PIN=0 leaves everything pageable */
  sd->pinned = !(getenv("RSCAPE_B200_PIN") && atoi(getenv("RSCAPE_B200_PIN")) == 0);
  if (sd->pinned && mi->COV && mi->Eval && sd->pp_slab && sd->nseff_slab && sd->ngap_slab && sd->stage) {
    rsb_host_register(mi->COV->mx[0],  sizeof(double) * L * L);
    rsb_host_register(mi->Eval->mx[0], sizeof(double) * L * L);
    rsb_host_register(sd->pp_slab,     sizeof(double) * L * L * 16);
    rsb_host_register(sd->nseff_slab,  sizeof(double) * L * L);
    rsb_host_register(sd->ngap_slab,   sizeof(double) * L * L);
    rsb_host_register(sd->stage,       (size_t) nseq * L);
  }
  sd->mi = mi;
  sd->next = side_head;
  side_head = sd;
  corr_ReuseCOV(mi, COVNONE, covclass);
  return mi;

 ERROR:
  if (sd) {
    if (sd->ctx) rsb_destroy(sd->ctx);
    free(sd->pp_slab); free(sd->nseff_slab); free(sd->ngap_slab); free(sd->pm_slab); free(sd->ps_slab); free(sd->stage);
    free(sd);
  }
  if (mi) {
    if (mi->pp) for (i = 0; i < alen; i++) free(mi->pp[i]);
    free(mi->pp); free(mi->nseff); free(mi->ngap); free(mi->pm); free(mi->ps);
    if (mi->COV) esl_dmatrix_Destroy(mi->COV);
    if (mi->Eval) esl_dmatrix_Destroy(mi->Eval);
    free(mi);
  }
  return NULL;
}

int
corr_Reuse(struct mutual_s *mi, int ishuffled, COVTYPE mitype, COVCLASS miclass)
{
  SIDE  *sd = side_of(mi);
  size_t L = (size_t) mi->alen;

  mi->ishuffled = ishuffled;
  if (sd) {
    memset(sd->pp_slab,    0, sizeof(double) * L * L * 16);
    memset(sd->nseff_slab, 0, sizeof(double) * L * L);
    memset(sd->ngap_slab,  0, sizeof(double) * L * L);
    memset(sd->pm_slab,    0, sizeof(double) * L * 4);
    memset(sd->ps_slab,    0, sizeof(double) * L * 5);
  }
  return corr_ReuseCOV(mi, mitype, miclass);
}

int
corr_ReuseCOV(struct mutual_s *mi, COVTYPE mitype, COVCLASS miclass)
{
  mi->type     = mitype;
  MI_CLASS(mi) = miclass;
  if (mi->COV)  esl_dmatrix_Set(mi->COV,  -eslINFINITY);
  if (mi->Eval) esl_dmatrix_Set(mi->Eval,  eslINFINITY);
  mi->besthreshCOV = -eslINFINITY;
  mi->minCOV       =  eslINFINITY;
  mi->maxCOV       = -eslINFINITY;
  return eslOK;
}

static void
unpin_buffers(SIDE *sd, struct mutual_s *mi)
{
  if (!sd || !sd->pinned) return;
  if (mi && mi->COV)  rsb_host_unregister(mi->COV->mx[0]);
  if (mi && mi->Eval) rsb_host_unregister(mi->Eval->mx[0]);
  rsb_host_unregister(sd->pp_slab); rsb_host_unregister(sd->nseff_slab); rsb_host_unregister(sd->ngap_slab); rsb_host_unregister(sd->stage);
  sd->pinned = 0;
}

void
corr_Destroy(struct mutual_s *mi)
{
  SIDE **pp, *sd;
  int64_t i;

  if (mi == NULL) return;
  for (pp = &side_head; (sd = *pp) != NULL; pp = &sd->next)
    if (sd->mi == mi) {
      *pp = sd->next;
      unpin_buffers(sd, mi);
      rsb_destroy(sd->ctx);
      free(sd->pp_slab); free(sd->nseff_slab); free(sd->ngap_slab); free(sd->pm_slab); free(sd->ps_slab); free(sd->stage);
      free(sd);
      break;
    }
  for (i = 0; i < mi->alen; i++) free(mi->pp[i]);
  free(mi->pp); free(mi->nseff); free(mi->ngap); free(mi->pm); free(mi->ps);
  if (mi->COV)  esl_dmatrix_Destroy(mi->COV);
  if (mi->Eval) esl_dmatrix_Destroy(mi->Eval);
  free(mi);
}

/* ------------------------------------------------------------------ probabilities: src/correlators.c:1301-1545 */
static int
stage_alignment(SIDE *sd, ESL_MSA *msa, char *errbuf)
{
  struct mutual_s *mi = sd->mi;
  int s;
  if (msa->nseq != mi->nseq || msa->alen != mi->alen) {
    if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "alignment is %d x %d but mutual_s was created for %d x %d",
                         msa->nseq, (int) msa->alen, (int) mi->nseq, (int) mi->alen);
    return eslFAIL;
  }
  for (s = 0; s < msa->nseq; s++) memcpy(sd->stage + (size_t) s * mi->alen, msa->ax[s] + 1, (size_t) mi->alen);
  if (rsb_set_weights(sd->ctx, msa->wgt) != 0) return fail(errbuf, sd->ctx, "sequence weights");
  return eslOK;
}

int
corr_NaivePP(ESL_RANDOMNESS *r, ESL_MSA *msa, struct mutual_s *mi, double tol, int verbose, char *errbuf)
{
  SIDE *sd = side_of(mi);
  int   status;
  (void) r; (void) verbose;
  if (!sd) ESL_FAIL(eslFAIL, errbuf, "mutual_s was not created by corr_Create()");
  if ((status = stage_alignment(sd, msa, errbuf)) != eslOK) return status;
  if (rsb_probs(sd->ctx, sd->stage, mi->alen, 0, tol, sd->pp_slab, NULL, NULL, sd->nseff_slab, sd->ngap_slab) != 0)
    return fail(errbuf, sd->ctx, "corr_NaivePP()");
  return eslOK;
}

int
corr_NaivePS(ESL_RANDOMNESS *r, ESL_MSA *msa, struct mutual_s *mi, double tol, int verbose, char *errbuf)
{
  SIDE *sd = side_of(mi);
  (void) r; (void) msa; (void) tol; (void) verbose;
  if (!sd) ESL_FAIL(eslFAIL, errbuf, "mutual_s was not created by corr_Create()");
  if (rsb_fetch_probs(sd->ctx, NULL, NULL, sd->ps_slab, NULL, NULL) != 0) return fail(errbuf, sd->ctx, "corr_NaivePS()");
  return eslOK;
}

int
corr_Marginals(struct mutual_s *mi, double tol, int verbose, char *errbuf)
{
  SIDE *sd = side_of(mi);
  int   i, status;
  (void) verbose;
  if (!sd) ESL_FAIL(eslFAIL, errbuf, "mutual_s was not created by corr_Create()");
  if (rsb_fetch_probs(sd->ctx, NULL, sd->pm_slab, NULL, NULL, NULL) != 0) return fail(errbuf, sd->ctx, "corr_Marginals()");
  for (i = 0; i < mi->alen; i++)
    if ((status = esl_vec_DValidate(mi->pm[i], 4, tol, errbuf)) != eslOK) return status;
  return eslOK;
}

int
corr_PostOrderPP(ESL_MSA *msa, ESL_TREE *T, struct ribomatrix_s *ribosum, struct mutual_s *mi, double tol, int verbose, char *errbuf)
{
  (void) msa; (void) T; (void) ribosum; (void) mi; (void) tol; (void) verbose;
  ESL_FAIL(eslFAIL, errbuf, "corr_PostOrderPP(): the AKMAEV method is not supported by the B200 path");
}

int
corr_ValidateProbs(struct mutual_s *mi, double tol, int verbose, char *errbuf)
{
  int i, j;
  (void) verbose;
  for (i = 0; i < mi->alen - 1; i++)
    for (j = i + 1; j < mi->alen; j++)
      if (esl_vec_DValidate(mi->pp[i][j], 16, tol, errbuf) != eslOK) ESL_FAIL(eslFAIL, errbuf, "pp validation failed");
  for (i = 0; i < mi->alen; i++)
    if (esl_vec_DValidate(mi->pm[i], 4, tol, errbuf) != eslOK) ESL_FAIL(eslFAIL, errbuf, "pm validation failed");
  for (i = 0; i < mi->alen; i++)
    if (esl_vec_DValidate(mi->ps[i], 5, tol, errbuf) != eslOK) ESL_FAIL(eslFAIL, errbuf, "ps validation failed");
  return eslOK;
}

/* corr_Probs = NaivePP + NaivePS + Marginals + ValidateProbs (src/correlators.c:1424-1456) in one device pass */
int
corr_Probs(ESL_RANDOMNESS *r, ESL_MSA *msa, ESL_TREE *T, struct ribomatrix_s *ribosum, struct mutual_s *mi,
           METHOD method, double tol, int verbose, char *errbuf)
{
  SIDE *sd = side_of(mi);
  int   status;
  (void) r; (void) T; (void) ribosum; (void) verbose;
  if (!sd) ESL_FAIL(eslFAIL, errbuf, "mutual_s was not created by corr_Create()");
  if (method == AKMAEV) return corr_PostOrderPP(msa, T, ribosum, mi, tol, verbose, errbuf);
  if (method != NONPARAM && method != POTTS) ESL_FAIL(eslFAIL, errbuf, "bad method option");
  if ((status = stage_alignment(sd, msa, errbuf)) != eslOK) return status;
  if (rsb_probs(sd->ctx, sd->stage, mi->alen, 0, tol, sd->pp_slab, sd->pm_slab, sd->ps_slab, sd->nseff_slab, sd->ngap_slab) != 0)
    return fail(errbuf, sd->ctx, "corr_Probs()");
  return eslOK;
}

/* ------------------------------------------------------------------ statistics: src/correlators.c:50-1061 */
static void
allowpair_flat(ESL_DMATRIX *allowpair, double *ap16)
{
  int x, y;
  for (x = 0; x < 4; x++)
    for (y = 0; y < 4; y++) ap16[x * 4 + y] = allowpair ? allowpair->mx[x][y] : 0.0;
}

/* one statistic of one class; labels mi->type / mi->class the way the reference's _C16/_C2/_CWC functions do */
static int
run_statistic(struct mutual_s *mi, int stat, COVCLASS cls, ESL_DMATRIX *allowpair, ESL_MSA *msa, char *errbuf, const char *who)
{
  SIDE   *sd = side_of(mi);
  double  ap[16];
  int     status;
  if (!sd) ESL_FAIL(eslFAIL, errbuf, "mutual_s was not created by corr_Create()");
  corr_ReuseCOV(mi, (COVTYPE) stat, (cls == CWC) ? C16 : cls);                   /* quirk Q9: GT_CWC labels itself C16 */
  allowpair_flat(allowpair, ap);
  if (msa) {
    if ((status = stage_alignment(sd, msa, errbuf)) != eslOK) return status;
  }
  if (rsb_statistic(sd->ctx, stat, (int) cls, (allowpair ? ap : NULL), msa ? sd->stage : NULL, mi->alen, 0,
                    mi->COV->mx[0], &mi->minCOV, &mi->maxCOV) != 0)
    return fail(errbuf, sd->ctx, who);
  return eslOK;
}

static COVCLASS
select_class(COVCLASS covclass, struct mutual_s *mi)     /* CSELECT rule, src/correlators.c:336 */
{
  if (covclass != CSELECT) return covclass;
  return (mi->nseq <= mi->nseqthresh || mi->alen <= mi->alenthresh) ? C2 : C16;
}

#define RSB_DEFINE_STAT(NAME, STATCODE, ALLOW_CWC)                                                                   \
  int corr_Calculate##NAME##_C16(struct mutual_s *mi, int verbose, char *errbuf)                                     \
  { (void) verbose; return run_statistic(mi, STATCODE, C16, NULL, NULL, errbuf, "corr_Calculate" #NAME "_C16()"); }  \
  int corr_Calculate##NAME##_C2(struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf)              \
  { (void) verbose; return run_statistic(mi, STATCODE, C2, allowpair, NULL, errbuf, "corr_Calculate" #NAME "_C2()"); } \
  int corr_Calculate##NAME(COVCLASS covclass, struct data_s *data)                                                   \
  {                                                                                                                  \
    struct mutual_s *mi = data->mi;                                                                                  \
    COVCLASS cls = select_class(covclass, mi);                                                                       \
    int status;                                                                                                      \
    if (cls == CWC && !(ALLOW_CWC)) ESL_FAIL(eslFAIL, data->errbuf, "corr_Calculate" #NAME "() CWC not implemented\n"); \
    status = run_statistic(mi, STATCODE, cls, data->allowpair, NULL, data->errbuf, "corr_Calculate" #NAME "()");     \
    if (status != eslOK) return eslFAIL;                                                                             \
    return eslOK;                                                                                                    \
  }

RSB_DEFINE_STAT(CHI,  RSB_STAT_CHI,  0)
RSB_DEFINE_STAT(OMES, RSB_STAT_OMES, 0)
RSB_DEFINE_STAT(GT,   RSB_STAT_GT,   1)
RSB_DEFINE_STAT(MI,   RSB_STAT_MI,   0)
RSB_DEFINE_STAT(MIr,  RSB_STAT_MIr,  0)
RSB_DEFINE_STAT(MIg,  RSB_STAT_MIg,  0)

int
corr_CalculateGT_CWC(struct mutual_s *mi, ESL_DMATRIX *allowpair, int verbose, char *errbuf)
{
  (void) verbose;
  return run_statistic(mi, RSB_STAT_GT, CWC, allowpair, NULL, errbuf, "corr_CalculateGT_CWC()");
}

int
corr_CalculateCCF_C16(struct mutual_s *mi, int verbose, char *errbuf)
{
  (void) verbose;
  return run_statistic(mi, RSB_STAT_CCF, C16, NULL, NULL, errbuf, "corr_CalculateCCF_C16()");
}

int
corr_CalculateCCF(COVCLASS covclass, struct data_s *data)
{
  (void) covclass;
  if (corr_CalculateCCF_C16(data->mi, data->verbose, data->errbuf) != eslOK) return eslFAIL;
  return eslOK;
}

/* RAF / RAFS read the alignment, ignore the weights and label themselves C2 (src/correlators.c:889,948) */
int
corr_CalculateRAF(COVCLASS covclass, struct data_s *data, ESL_MSA *msa)
{
  (void) covclass;
  return run_statistic(data->mi, RSB_STAT_RAF, C2, data->allowpair, msa, data->errbuf, "corr_CalculateRAF()");
}

int
corr_CalculateRAFS(COVCLASS covclass, struct data_s *data, ESL_MSA *msa)
{
  (void) covclass;
  return run_statistic(data->mi, RSB_STAT_RAFS, C2, data->allowpair, msa, data->errbuf, "corr_CalculateRAFS()");
}

/* ------------------------------------------------------------------ background correction: src/correlators.c:1064-1157 */
static const struct { COVTYPE type; const char *name; } covtype_names[] = {
  { CHI, "CHI" }, { CHIp, "CHIp" }, { CHIa, "CHIa" }, { GT, "GT" }, { GTp, "GTp" }, { GTa, "GTa" },
  { MI, "MI" }, { MIp, "MIp" }, { MIa, "MIa" }, { MIr, "MIr" }, { MIrp, "MIrp" }, { MIra, "MIra" },
  { MIg, "MIg" }, { MIgp, "MIgp" }, { MIga, "MIga" }, { OMES, "OMES" }, { OMESp, "OMESp" }, { OMESa, "OMESa" },
  { RAF, "RAF" }, { RAFp, "RAFp" }, { RAFa, "RAFa" }, { RAFS, "RAFS" }, { RAFSp, "RAFSp" }, { RAFSa, "RAFSa" },
  { CCF, "CCF" }, { CCFp, "CCFp" }, { CCFa, "CCFa" }, { PTFp, "PTFp" }, { PTAp, "PTAp" }, { PTDp, "PTDp" },
};
#define N_COVTYPE_NAMES ((int) (sizeof(covtype_names) / sizeof(covtype_names[0])))

int
corr_COVTYPEString(char **ret_covtype, COVTYPE type, char *errbuf)
{
  int k;
  for (k = 0; k < N_COVTYPE_NAMES; k++)
    if (covtype_names[k].type == type) return esl_sprintf(ret_covtype, "%s", covtype_names[k].name);
  ESL_FAIL(eslFAIL, errbuf, "wrong COVTYPE");
}

int
corr_String2COVTYPE(char *covtype, COVTYPE *ret_type, char *errbuf)
{
  int k;
  for (k = 0; k < N_COVTYPE_NAMES; k++)
    if (esl_strcmp(covtype, covtype_names[k].name) == 0) { *ret_type = covtype_names[k].type; return eslOK; }
  ESL_FAIL(eslFAIL, errbuf, "wrong COVTYPE %s", covtype);
}

int
corr_THRESHTYPEString(char **ret_threshtype, THRESHTYPE type, char *errbuf)
{
  if (type == Eval) return esl_sprintf(ret_threshtype, "Eval");
  ESL_FAIL(eslFAIL, errbuf, "wrong THRESHTYPE");
}

int
corr_CalculateCOVCorrected(ACTYPE actype, struct data_s *data, int shiftnonneg)
{
  struct mutual_s *mi = data->mi;
  SIDE            *sd = side_of(mi);
  char            *type = NULL, *covtype = NULL;
  double          *raw;
  double           mn, mx;
  size_t           L = (size_t) mi->alen, k;
  int              i, j;

  if (!sd) ESL_FAIL(eslFAIL, data->errbuf, "mutual_s was not created by corr_Create()");
  if (actype != APC && actype != ASC) ESL_FAIL(eslFAIL, data->errbuf, "wrong correction type\n");

  /* the type is renamed through its string, GT -> GTp / GTa (:1076-1090); an unknown name leaves it unchanged */
  corr_COVTYPEString(&type, mi->type, data->errbuf);
  esl_sprintf(&covtype, "%s%s", type ? type : "", actype == APC ? "p" : "a");

  /* the correction applies to whatever mi->COV holds on the host (it may have been written by host code) */
  if ((raw = malloc(sizeof(double) * L * L)) == NULL) { free(type); free(covtype); ESL_FAIL(eslFAIL, data->errbuf, "allocation failed"); }
  memcpy(raw, mi->COV->mx[0], sizeof(double) * L * L);
  corr_String2COVTYPE(covtype, &mi->type, NULL);
  corr_ReuseCOV(mi, mi->type, MI_CLASS(mi));
  free(type); free(covtype);

  for (k = 0; k < L * L; k++) if (isinf(raw[k])) raw[k] = 0.0;                   /* the -inf diagonal is never read (:1105) */
  if (rsb_correct_host(sd->ctx, (int) actype, raw, &mn, &mx) != 0) {
    free(raw);
    if (data->errbuf) snprintf(data->errbuf, eslERRBUFSIZE, "%s", rsb_error(sd->ctx));
    return eslFAIL;
  }
  memcpy(mi->COV->mx[0], raw, sizeof(double) * L * L);
  free(raw);
  mi->minCOV = mn;
  mi->maxCOV = mx;

  if (shiftnonneg)                                                               /* Potts only (:1130-1134) */
    for (i = 0; i < mi->alen; i++)
      for (j = 0; j < mi->alen; j++)
        if (i != j) mi->COV->mx[i][j] -= mi->minCOV;
  return eslOK;
}
