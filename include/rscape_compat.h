/* rscape_compat.h -- the R-scape types that cross the covariation drop-in boundary.
 *
 * Field-compatible restatement of the parts of the reference's src/correlators.h that the
 * hot path touches: the enums (src/correlators.h:29-105), struct mutual_s (:107-130),
 * RANKLIST (:133-138), THRESH (:170-176) and struct data_s (:374-442).  Members of data_s that
 * the hot path never dereferences are declared as opaque pointers of the same size, so the
 * layout is identical on LP64 (checked against the real header by tests/test_layout.py when the
 * reference tree is present).  Inside a real R-scape tree, define RSB_USE_RSCAPE_HEADERS and the
 * reference's own correlators.h is used instead.
 */
#ifndef RSB_RSCAPE_COMPAT_INCLUDED
#define RSB_RSCAPE_COMPAT_INCLUDED

#ifdef RSB_USE_RSCAPE_HEADERS
#include "correlators.h"
#include "covariation.h"
/* minimum backbone distance of the contact list (lib/R-view/src/rview_contacts.h:102), read by the histogram fill */
#define RSB_DATA_MIND(data) ((data)->clist ? (data)->clist->mind : 1)
#else
/* Without R-view's headers CLIST is opaque; the one field the hot path reads (mind, rview_contacts.h:102) is reached through
 * this stand-in: in the compat build data->clist, when set, points to a struct rsb_clist_compat. */
struct rsb_clist_compat { int mind; };
#define RSB_DATA_MIND(data) ((data)->clist ? ((const struct rsb_clist_compat *) (data)->clist)->mind : 1)

#include "easel.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef IDX
#define IDX(i,j,L)  ( (i) * (L) + (j) )
#endif

typedef enum { SAMPLE_CONTACTS = 0, SAMPLE_BP = 1, SAMPLE_WC = 2, SAMPLE_ALL = 3 } SAMPLESIZE;
typedef enum { C16 = 0, C2 = 1, CWC = 2, CSELECT = 3 } COVCLASS;
typedef enum {
  CHI  = 0,  CHIp  = 1,  CHIa  = 2,
  GT   = 3,  GTp   = 4,  GTa   = 5,
  MI   = 6,  MIp   = 7,  MIa   = 8,
  MIr  = 9,  MIrp  = 10, MIra  = 11,
  MIg  = 12, MIgp  = 13, MIga  = 14,
  OMES = 15, OMESp = 16, OMESa = 17,
  RAF  = 18, RAFp  = 19, RAFa  = 20,
  RAFS = 21, RAFSp = 22, RAFSa = 23,
  CCF  = 24, CCFp  = 25, CCFa  = 26,
  PTFp = 27, PTAp  = 28, PTDp  = 29,
  COVNONE = 30
} COVTYPE;
typedef enum { APC = 0, ASC = 1 } ACTYPE;
typedef enum { NONPARAM = 0, POTTS = 1, AKMAEV = 2 } METHOD;
typedef enum { NAIVE = 0, NULLPHYLO = 1, GIVENNULL = 2 } STATSMETHOD;
typedef enum { GIVSS = 0, FOLDSS = 1, RANSS = 2 } MODE;
typedef enum { Eval = 0 } THRESHTYPE;

struct mutual_s {
  int64_t         alen;
  int64_t         nseq;
  double       ***pp;          /* [alen][alen][K*K] joint probabilities               */
  double        **pm;          /* [alen][K]        partner-averaged marginals         */
  double        **nseff;       /* [alen][alen]     effective number of sequences      */
  double        **ps;          /* [alen][K+1]      single-column probabilities        */
  double        **ngap;        /* [alen][alen]     weighted gaps, i<j only            */

  COVTYPE         type;
#ifdef __cplusplus
  COVCLASS        class_;
#else
  COVCLASS        class;
#endif
  ESL_DMATRIX    *COV;
  ESL_DMATRIX    *Eval;

  double          besthreshCOV;
  double          minCOV;
  double          maxCOV;

  int             ishuffled;
  int             nseqthresh;
  int             alenthresh;

  ESL_ALPHABET   *abc;
};

typedef struct ranklist_s {
  ESL_HISTOGRAM *ha;
  ESL_HISTOGRAM *ht;
  ESL_HISTOGRAM *hb;
  double        *survfit;
} RANKLIST;

/* src/correlators.h:141-165; bptype is R-view's BPTYPE enum (lib/R-view/src/rview_contacts.h:22-62), an int here */
#ifndef MAX_EVAL
#define MAX_EVAL 1000               /* forcing -E > MAX_EVAL reports all pairs, src/correlators.h:24 */
#endif
typedef struct hit_s {
  int64_t i;
  int64_t j;
  double  sc;
  double  Eval;
  double  pval;
  int64_t nsubs;
  double  power;
  int     bptype;
  int     is_compatible;
} HIT;

typedef struct hitlist_s {
  int       nhit;
  HIT     **srthit;
  HIT      *hit;
  int64_t   Nt;
  int64_t   Nb;
} HITLIST;

typedef struct thresh_s {
  THRESHTYPE type;
  double     val;
  double     sc_bp;
  double     sc_nbp;
} THRESH;

struct ribomatrix_s;   /* opaque here; src/ribosum_matrix.h:26-70 */

struct data_s {
  void                *ofile;            /* struct outfiles_s * */
  char                *gnuplot;
  int                  R2Rall;
  int                  R2Rmsa;
  ESL_RANDOMNESS      *r;

  SAMPLESIZE           samplesize;
  RANKLIST            *ranklist_null;
  RANKLIST            *ranklist_aux;
  struct mutual_s     *mi;
  void                *pt;               /* PT * */
  THRESH              *thresh;
  STATSMETHOD          statsmethod;
  METHOD               covmethod;
  MODE                 mode;
  int                  abcisRNA;
  int                  hasss;
  COVTYPE              covtype;

  int                  OL;
  int                  nseq;
  void                *ctlist;           /* CTLIST * */
  int                  expBP;
  int                  onbpairs;
  int                  nbpairs;
  int                  nbpairs_fold;
  int                 *nsubs;
  int                 *ndouble;
  int                 *njoin;
  void                *spair;            /* SPAIR * */
  void                *power;            /* POWER * */

  void                *r3d;              /* R3D * */
  int                  helix_unpaired;
  int                  nagg;
  double               agg_Eval;
  int                 *agg_method;       /* enum agg_e * */

  int                  pc_codon_thresh;

  ESL_TREE            *T;
  struct ribomatrix_s *ribosum;

  int                  gapthresh;
  int                 *ct;
  void                *clist;            /* CLIST * */
  int                 *msa2pdb;
  int                 *msamap;
  int                  firstpos;
  double               bmin;
  double               w;
  double               fracfit;
  double               pmass;
  int                  doexpfit;
  double               tau;
  double               mu;
  double               lambda;
  ESL_DMATRIX         *allowpair;
  double               tol;
  int                  nofigures;
  int                  verbose;
  char                *errbuf;
  int                  doR2R;
  int                  doDotPlot;
  int                  ignorebps;
  int                  prep_onehot;
  int                  prep_RF;
};

#ifdef __cplusplus
}
#endif
#endif /* RSB_USE_RSCAPE_HEADERS */
#endif
