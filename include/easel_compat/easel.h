/* easel.h -- minimal Easel-compatible shim (NOT Easel; written for this repo).
 *
 * R-scape links against the Easel C library, which is an un-vendored submodule of
 * the reference tree (configure.ac:128-129) and is absent from this box.  This
 * header declares only the handful of Easel types, macros and functions that
 * R-scape's covariation hot path touches (the list is `nm src/correlators.o`,
 * SURVEY.md section 8c), with the semantics of upstream Easel as restated in
 * SURVEY.md section 9.7.  Three consumers share it so that all three agree on
 * struct layouts:
 *   1. oracle/_ref : the reference's own src/correlators.c compiled unchanged,
 *   2. oracle/     : the CPU restatement (test infrastructure only),
 *   3. r-scape_b200/host : the host-side mirror of the reference API that calls
 *      the sm_100a kernels through the C-ABI in include/rscape_b200.h.
 * Inside a real R-scape tree this directory is simply left off the include path
 * and the real Easel headers are used instead.
 *
 * Every other esl_*.h in this directory just includes this file.
 */
#ifndef RSB_EASEL_COMPAT_INCLUDED
#define RSB_EASEL_COMPAT_INCLUDED

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <inttypes.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include <float.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (upstream easel.h values) ---- */
#define eslOK              0
#define eslFAIL            1
#define eslEOL             2
#define eslEOF             3
#define eslEOD             4
#define eslEMEM            5
#define eslENOTFOUND       6
#define eslEFORMAT         7
#define eslEAMBIGUOUS      8
#define eslEDIVZERO        9
#define eslEINCOMPAT      10
#define eslEINVAL         11
#define eslESYS           12
#define eslECORRUPT       13
#define eslEINCONCEIVABLE 14
#define eslESYNTAX        15
#define eslERANGE         16
#define eslEDUP           17
#define eslENOHALT        18
#define eslENORESULT      19
#define eslENODATA        20
#define eslETYPE          21
#define eslEOVERWRITE     22
#define eslENOSPACE       23
#define eslEUNIMPLEMENTED 24
#define eslENOFORMAT      25
#define eslENOALPHABET    26
#define eslEWRITE         27
#define eslEINACCURATE    28

#define eslERRBUFSIZE 128
#define eslINFINITY   INFINITY
#define eslNaN        NAN
#define eslCONST_LOG2 0.69314718055994529
#define eslSMALLX1    5e-9

#ifndef TRUE
#define TRUE  1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define ESL_DASSERT1(x)
#define ESL_DASSERT2(x)
#define ESL_RALLOC(p, tmp, newsize) do {                                 \
    if ((p) == NULL) (tmp) = malloc(newsize);                            \
    else             (tmp) = realloc((p), (newsize));                    \
    if ((tmp) != NULL) (p) = (tmp);                                      \
    else { status = eslEMEM;                                             \
      esl_exception(status, FALSE, __FILE__, __LINE__, "realloc for size %d failed", (int)(newsize)); goto ERROR; } \
  } while (0)
#define ESL_SWAP(x, y, type) do { type esl_swap_tmp_ = (x); (x) = (y); (y) = esl_swap_tmp_; } while (0)
#define ESL_MIN(a,b) (((a)<(b))?(a):(b))
#define ESL_MAX(a,b) (((a)>(b))?(a):(b))

extern void esl_exception(int errcode, int use_errno, char *sourcefile, int sourceline, char *format, ...);
extern void esl_fail(char *errbuf, const char *format, ...);
extern void esl_fatal(const char *format, ...);
extern int  esl_sprintf(char **ret_s, const char *format, ...);
extern int  esl_strcmp(const char *s1, const char *s2);
extern int  esl_strdup(const char *s, int64_t n, char **ret_dup);

#define ESL_FAIL(code, errbuf, ...) do {                                 \
    esl_fail(errbuf, __VA_ARGS__);                                       \
    return code; } while (0)
#define ESL_XFAIL(code, errbuf, ...) do {                                \
    status = code;                                                       \
    esl_fail(errbuf, __VA_ARGS__);                                       \
    goto ERROR; } while (0)
#define ESL_EXCEPTION(code, ...) do {                                    \
    esl_exception(code, FALSE, __FILE__, __LINE__, __VA_ARGS__);         \
    return code; } while (0)
#define ESL_XEXCEPTION(code, ...) do {                                   \
    status = code;                                                       \
    esl_exception(code, FALSE, __FILE__, __LINE__, __VA_ARGS__);         \
    goto ERROR; } while (0)
#define ESL_ALLOC(p, size) do {                                          \
    size_t esl_alloc_size_ = (size);                                     \
    if (esl_alloc_size_ == 0) { (p) = NULL; status = eslEMEM;            \
      esl_exception(status, FALSE, __FILE__, __LINE__, "zero malloc disallowed"); goto ERROR; } \
    if (((p) = malloc(esl_alloc_size_)) == NULL) { status = eslEMEM;     \
      esl_exception(status, FALSE, __FILE__, __LINE__, "malloc of size %d failed", (int) esl_alloc_size_); goto ERROR; } \
  } while (0)
#define ESL_REALLOC(p, newsize) do {                                     \
    void *esl_tmp_;                                                      \
    if ((p) == NULL) esl_tmp_ = malloc(newsize);                         \
    else             esl_tmp_ = realloc((p), (newsize));                 \
    if (esl_tmp_ != NULL) (p) = esl_tmp_;                                \
    else { status = eslEMEM;                                             \
      esl_exception(status, FALSE, __FILE__, __LINE__, "realloc for size %d failed", (int)(newsize)); goto ERROR; } \
  } while (0)

/* ---- alphabet ---- */
typedef uint8_t ESL_DSQ;
#define eslDSQ_SENTINEL 255
#define eslDSQ_ILLEGAL  254
#define eslDSQ_IGNORED  253
#define eslDSQ_EOL      252
#define eslDSQ_EOD      251

#define eslUNKNOWN     0
#define eslRNA         1
#define eslDNA         2
#define eslAMINO       3
#define eslCOINS       4
#define eslDICE        5
#define eslNONSTANDARD 6

/* digital RNA: A0 C1 G2 U3 -4 R5 Y6 M7 K8 S9 W10 H11 B12 V13 D14 N15 *16 ~17; K=4, Kp=18 */
typedef struct {
  int      type;
  int      K;
  int      Kp;
  char    *sym;
  ESL_DSQ  inmap[128];
  char   **degen;
  int     *ndegen;
  ESL_DSQ *complement;
} ESL_ALPHABET;

extern ESL_ALPHABET *esl_alphabet_Create(int type);
extern int64_t       esl_dsq_GetLen(const ESL_DSQ *dsq);
extern int64_t       esl_dsq_GetRawLen(const ESL_ALPHABET *abc, const ESL_DSQ *dsq);
extern void          esl_alphabet_Destroy(ESL_ALPHABET *a);

#define esl_abc_XIsValid(a, x)       ((x) < (a)->Kp)
#define esl_abc_XIsResidue(a, x)     ((x) < (a)->K || ((x) > (a)->K && (x) < (a)->Kp-2))
#define esl_abc_XIsCanonical(a, x)   ((x) < (a)->K)
#define esl_abc_XIsGap(a, x)         ((x) == (a)->K)
#define esl_abc_XIsDegenerate(a, x)  ((x) >  (a)->K && (x) < (a)->Kp-2)
#define esl_abc_XIsUnknown(a, x)     ((x) == (a)->Kp-3)
#define esl_abc_XIsNonresidue(a, x)  ((x) == (a)->Kp-2)
#define esl_abc_XIsMissing(a, x)     ((x) == (a)->Kp-1)
#define esl_abc_XGetGap(a)           ((a)->K)
#define esl_abc_XGetUnknown(a)       ((a)->Kp-3)

/* ---- dense matrix ---- */
typedef struct {
  double **mx;
  int      n;
  int      m;
  enum { eslGENERAL, eslUPPER } type;
  int      ncells;
} ESL_DMATRIX;

extern ESL_DMATRIX *esl_dmatrix_Create(int n, int m);
extern ESL_DMATRIX *esl_dmatrix_Clone(const ESL_DMATRIX *old);
extern int          esl_dmatrix_Copy(const ESL_DMATRIX *src, ESL_DMATRIX *dest);
extern void         esl_dmatrix_Destroy(ESL_DMATRIX *A);
extern int          esl_dmatrix_Set(ESL_DMATRIX *A, double x);
extern int          esl_dmatrix_SetZero(ESL_DMATRIX *A);
extern int          esl_dmatrix_SetIdentity(ESL_DMATRIX *A);
extern int          esl_dmatrix_Dump(FILE *ofp, const ESL_DMATRIX *A, const char *rowlabel, const char *collabel);
extern int          esl_dmx_Exp(const ESL_DMATRIX *Q, double t, ESL_DMATRIX *P);
extern int          esl_dmx_Multiply(const ESL_DMATRIX *A, const ESL_DMATRIX *B, ESL_DMATRIX *C);
extern int          esl_dmx_Scale(ESL_DMATRIX *A, double k);
extern int          esl_rmx_ValidateP(ESL_DMATRIX *P, double tol, char *errbuf);

/* ---- vector ops ---- */
extern void   esl_vec_DSet(double *vec, int n, double value);
extern void   esl_vec_ISet(int *vec, int n, int value);
extern void   esl_vec_FSet(float *vec, int n, float value);
extern void   esl_vec_DCopy(const double *src, int n, double *dest);
extern void   esl_vec_ICopy(const int *src, int n, int *dest);
extern double esl_vec_DSum(const double *vec, int n);
extern float  esl_vec_FSum(const float *vec, int n);
extern void   esl_vec_FScale(float *vec, int n, float scale);
extern void   esl_vec_DNorm(double *vec, int n);
extern int    esl_vec_DValidate(const double *vec, int n, double tol, char *errbuf);
extern int    esl_vec_DDump(FILE *ofp, const double *v, int n, const char *label);

/* ---- integer pushdown stack ---- */
typedef struct {
  int   *idata;
  int    n;
  int    nalloc;
} ESL_STACK;
extern ESL_STACK *esl_stack_ICreate(void);
extern int        esl_stack_IPush(ESL_STACK *s, int x);
extern int        esl_stack_IPop(ESL_STACK *s, int *ret_x);
extern void       esl_stack_Destroy(ESL_STACK *s);
extern int        esl_stack_ObjectCount(ESL_STACK *s);

/* ---- random numbers: Mersenne Twister MT19937 as in upstream esl_random ---- */
typedef struct {
  int      type;
  int      mti;
  uint32_t mt[624];
  uint32_t x;
  uint32_t seed;
} ESL_RANDOMNESS;
extern ESL_RANDOMNESS *esl_randomness_Create(uint32_t seed);
extern void            esl_randomness_Destroy(ESL_RANDOMNESS *r);
extern double          esl_random(ESL_RANDOMNESS *r);
extern int             esl_rnd_FChoose(ESL_RANDOMNESS *r, const float *p, int N);
extern int             esl_rnd_DChoose(ESL_RANDOMNESS *r, const double *p, int N);
extern int             esl_vec_IShuffle(ESL_RANDOMNESS *r, int *v, int n);
#define esl_rnd_Roll(r, n) ((int)(esl_random(r) * (n)))

/* ---- multiple alignment (digital mode subset; leading fields in upstream order) ---- */
#define eslMSA_HASWGTS (1 << 0)
#define eslMSA_DIGITAL (1 << 1)
typedef struct {
  char        **aseq;
  char        **sqname;
  double       *wgt;
  int64_t       alen;
  int           nseq;
  int           flags;
  ESL_ALPHABET *abc;
  ESL_DSQ     **ax;
  char         *name;
  char         *desc;
  char         *acc;
  char         *au;
  char         *ss_cons;
  char         *sa_cons;
  char         *pp_cons;
  char         *rf;
  char         *mm;
  char        **sqacc;
  char        **sqdesc;
  char        **ss;
  char        **sa;
  char        **pp;
  float         cutoff[6];
  int           cutset[6];
  int           sqalloc;
  int64_t      *sqlen;
  int64_t      *sslen;
  int64_t      *salen;
  int64_t      *pplen;
  int           lastidx;
  /* unparsed Stockholm markup (never touched on the hot path; present so that
   * reference sources that mention the fields compile) */
  char        **comment;
  int           ncomment;
  int           alloc_ncomment;
  char        **gf_tag;
  char        **gf;
  int           ngf;
  int           alloc_ngf;
  char        **gs_tag;
  char       ***gs;
  int           ngs;
  char        **gc_tag;
  char        **gc;
  int           ngc;
  char        **gr_tag;
  char       ***gr;
  int           ngr;
  void         *index;
  void         *gs_idx;
  void         *gc_idx;
  void         *gr_idx;
  int64_t       offset;
} ESL_MSA;

extern ESL_MSA *esl_msa_CreateDigital(const ESL_ALPHABET *abc, int nseq, int64_t alen);
extern ESL_MSA *esl_msa_Clone(const ESL_MSA *msa);
extern void     esl_msa_Destroy(ESL_MSA *msa);
extern int      esl_msa_SequenceSubset(const ESL_MSA *msa, const int *useme, ESL_MSA **ret_new);

/* ---- tree ---- */
typedef struct {
  int      N;
  int     *parent;
  int     *left;
  int     *right;
  double  *ld;
  double  *rd;
  int     *taxaparent;
  int     *cladesize;
  char   **taxonlabel;
  char   **nodelabel;
  int      is_linkage_tree;
  int      show_unrooted;
  int      show_node_labels;
  int      show_root_branchlength;
  int      show_branchlengths;
  int      show_quoted_labels;
  int      show_numeric_taxonlabels;
  int      nalloc;
} ESL_TREE;
extern ESL_TREE *esl_tree_Create(int ntaxa);
extern void      esl_tree_Destroy(ESL_TREE *T);
/* easel_shim_fit.c: what Tree_CalculateExtFromMSA / Tree_RootAtMidPoint (src/msatree.c:49-105, 524-790) call around FastTree's output */
extern int       esl_tree_ReadNewick(FILE *fp, char *errbuf, ESL_TREE **ret_T);
extern int       esl_tree_RenumberNodes(ESL_TREE *T);
extern int       esl_tree_SetTaxaParents(ESL_TREE *T);
extern int       esl_tree_SetCladesizes(ESL_TREE *T);
extern int       esl_tree_Validate(ESL_TREE *T, char *errbuf);
extern int       esl_tree_Grow(ESL_TREE *T);

/* ---- histogram ("full" histogram subset used by src/covariation.c) ---- */
typedef struct {
  uint64_t *obs;
  int       nb;
  double    w;
  double    bmin, bmax;
  int       imin, imax;
  double    xmin, xmax;
  uint64_t  n;
  double   *x;
  uint64_t  nalloc;
  double    phi;
  int       cmin;
  uint64_t  z;
  uint64_t  Nc;
  uint64_t  No;
  double   *expect;
  int       emin;
  double    tailbase;
  double    tailmass;
  int       is_full;
  int       is_done;
  int       is_sorted;
  int       is_tailfit;
  int       is_rounded;
  enum { COMPLETE, VIRTUAL_CENSORED, TRUE_CENSORED } dataset_is;
} ESL_HISTOGRAM;

extern ESL_HISTOGRAM *esl_histogram_Create    (double bmin, double bmax, double w);
extern ESL_HISTOGRAM *esl_histogram_CreateFull(double bmin, double bmax, double w);
extern void           esl_histogram_Destroy(ESL_HISTOGRAM *h);
extern int            esl_histogram_Score2Bin(ESL_HISTOGRAM *h, double x, int *ret_b);
extern int            esl_histogram_Add(ESL_HISTOGRAM *h, double x);
#define esl_histogram_Bin2LBound(h,b)  ((h)->w*(b) + (h)->bmin)
#define esl_histogram_Bin2UBound(h,b)  ((h)->w*((b)+1) + (h)->bmin)

/* MSA file format codes (upstream esl_msafile.h values) */
#define eslMSAFILE_UNKNOWN     0
#define eslMSAFILE_STOCKHOLM 101
#define eslMSAFILE_PFAM      102
#define eslMSAFILE_A2M       103
#define eslMSAFILE_PSIBLAST  104
#define eslMSAFILE_SELEX     105
#define eslMSAFILE_AFA       106
#define eslMSAFILE_CLUSTAL   107

/* ---- opaque types only named in prototypes of R-scape headers ---- */
typedef int64_t esl_pos_t;
typedef struct esl_sq_s {
  char    *name;
  char    *acc;
  char    *desc;
  int32_t  tax_id;
  char    *seq;
  ESL_DSQ *dsq;
  char    *ss;
  int64_t  n;
  int64_t  start, end, C, W, L;
  char    *source;
} ESL_SQ;
typedef struct esl_getopts_s ESL_GETOPTS;
typedef struct esl_sqfile_s  ESL_SQFILE;
typedef struct esl_msafile_s ESL_MSAFILE;
typedef struct esl_fileparser_s ESL_FILEPARSER;
/* whitespace-delimited token files (esl_fileparser): what cov_ReadNullHistogram (--givennull, src/covariation.c:1718-1843) reads with */
extern int  esl_fileparser_Open(const char *filename, const char *envvar, ESL_FILEPARSER **ret_efp);
extern int  esl_fileparser_SetCommentChar(ESL_FILEPARSER *efp, char c);
extern int  esl_fileparser_NextLine(ESL_FILEPARSER *efp);
extern int  esl_fileparser_GetTokenOnLine(ESL_FILEPARSER *efp, char **opt_tok, int *opt_toklen);
extern void esl_fileparser_Close(ESL_FILEPARSER *efp);
/* tail fits of the null histogram (src/covariation.c:1915-1973), restated in easel_shim_fit.c */
extern double esl_exp_generic_surv(double x, void *params);
extern double esl_gam_generic_surv(double x, void *params);
extern double esl_gam_cdf (double x, double mu, double lambda, double tau);
extern double esl_gam_surv(double x, double mu, double lambda, double tau);
extern int    esl_stats_IncompleteGamma(double a, double x, double *ret_pax, double *ret_qax);
extern int    esl_histogram_SetTailByMass(ESL_HISTOGRAM *h, double pmass, double *ret_newmass);
extern int    esl_exp_FitCompleteBinned(ESL_HISTOGRAM *h, double *ret_mu, double *ret_lambda);
extern int    esl_gam_FitCompleteBinned(ESL_HISTOGRAM *h, double *ret_mu, double *ret_lambda, double *ret_tau);

#ifdef __cplusplus
}
#endif
#endif /* RSB_EASEL_COMPAT_INCLUDED */
