/* easel_shim.c -- implementations behind include/easel_compat/easel.h.
 *
 * HOST INFRASTRUCTURE shared with the test oracle.  This is NOT Easel: it is a from-scratch restatement of the
 * few Easel routines that R-scape's covariation path calls (`nm src/correlators.o` in the
 * reference; SURVEY.md section 8c lists them), following the upstream semantics summarised in
 * SURVEY.md section 9.7.  Easel is an un-vendored submodule of the reference
 * (configure.ac:128-129, no version pin), so parity of these routines against real Easel is
 * UNPINNED except where the tutorial transcript pins the composition (tests/test_golden_tutorial.py).
 *
 * Linked into: r-scape_b200/librscape_b200_host.so (the host-side mirror of the reference API; inside a
 * real R-scape tree the real libeasel is linked instead), the CPU checker built under oracle/ and
 * oracle/_ref/librscape_ref.so (reference sources compiled unchanged).
 */
#include <stdarg.h>
#include "easel.h"

/* ------------------------------------------------------------------ errors / strings */
void
esl_exception(int errcode, int use_errno, char *sourcefile, int sourceline, char *format, ...)
{
  va_list ap;
  (void) use_errno;
  fprintf(stderr, "Fatal exception (source file %s, line %d): ", sourcefile, sourceline);
  va_start(ap, format);
  vfprintf(stderr, format, ap);
  va_end(ap);
  fprintf(stderr, " [status %d]\n", errcode);
  fflush(stderr);
  abort();
}

void
esl_fail(char *errbuf, const char *format, ...)
{
  va_list ap;
  if (errbuf == NULL) return;
  va_start(ap, format);
  vsnprintf(errbuf, eslERRBUFSIZE, format, ap);
  va_end(ap);
}

void
esl_fatal(const char *format, ...)
{
  va_list ap;
  va_start(ap, format);
  vfprintf(stderr, format, ap);
  va_end(ap);
  fprintf(stderr, "\n");
  exit(1);
}

int
esl_sprintf(char **ret_s, const char *format, ...)
{
  va_list ap, ap2;
  int     n;
  char   *s;

  if (format == NULL) { *ret_s = NULL; return eslOK; }
  va_start(ap, format);
  va_copy(ap2, ap);
  n = vsnprintf(NULL, 0, format, ap);
  va_end(ap);
  if (n < 0 || (s = malloc((size_t) n + 1)) == NULL) { va_end(ap2); *ret_s = NULL; return eslEMEM; }
  vsnprintf(s, (size_t) n + 1, format, ap2);
  va_end(ap2);
  *ret_s = s;
  return eslOK;
}

/* NULL-tolerant strcmp: two NULLs compare equal, one NULL compares unequal. */
int
esl_strcmp(const char *s1, const char *s2)
{
  if (s1 && s2) return strcmp(s1, s2);
  if (s1)       return  1;
  if (s2)       return -1;
  return 0;
}

int
esl_strdup(const char *s, int64_t n, char **ret_dup)
{
  char *d;
  if (s == NULL) { *ret_dup = NULL; return eslOK; }
  if (n < 0) n = (int64_t) strlen(s);
  if ((d = malloc((size_t) n + 1)) == NULL) { *ret_dup = NULL; return eslEMEM; }
  memcpy(d, s, (size_t) n);
  d[n] = '\0';
  *ret_dup = d;
  return eslOK;
}

/* ------------------------------------------------------------------ alphabet */
ESL_ALPHABET *
esl_alphabet_Create(int type)
{
  static const char rna[] = "ACGU-RYMKSWHBVDN*~";
  static const char dna[] = "ACGT-RYMKSWHBVDN*~";
  ESL_ALPHABET *a;
  const char   *sym;
  int           x;

  if      (type == eslRNA) sym = rna;
  else if (type == eslDNA) sym = dna;
  else return NULL;

  if ((a = calloc(1, sizeof(ESL_ALPHABET))) == NULL) return NULL;
  a->type = type;
  a->K    = 4;
  a->Kp   = 18;
  a->sym  = malloc(a->Kp + 1);
  memcpy(a->sym, sym, a->Kp + 1);
  for (x = 0; x < 128; x++) a->inmap[x] = eslDSQ_ILLEGAL;
  for (x = 0; x < a->Kp; x++) {
    a->inmap[(int) sym[x]] = (ESL_DSQ) x;
    if (sym[x] >= 'A' && sym[x] <= 'Z') a->inmap[(int) sym[x] + 32] = (ESL_DSQ) x;
  }
  /* synonyms upstream defines for nucleic alphabets */
  a->inmap['T'] = a->inmap['t'] = 3;
  a->inmap['U'] = a->inmap['u'] = 3;
  a->inmap['X'] = a->inmap['x'] = 15;
  a->inmap['I'] = a->inmap['i'] = 0;
  a->inmap['_'] = a->inmap['.'] = 4;
  return a;
}

void
esl_alphabet_Destroy(ESL_ALPHABET *a)
{
  if (a == NULL) return;
  free(a->sym);
  free(a);
}

/* digital sequences are sentinel-delimited: dsq[0] = dsq[L+1] = eslDSQ_SENTINEL */
int64_t
esl_dsq_GetLen(const ESL_DSQ *dsq)
{
  int64_t n = 0;
  while (dsq[n + 1] != eslDSQ_SENTINEL) n++;
  return n;
}

int64_t
esl_dsq_GetRawLen(const ESL_ALPHABET *abc, const ESL_DSQ *dsq)
{
  int64_t n = 0, i;
  for (i = 1; dsq[i] != eslDSQ_SENTINEL; i++)
    if (esl_abc_XIsResidue(abc, dsq[i])) n++;
  return n;
}

/* ------------------------------------------------------------------ dense matrices */
ESL_DMATRIX *
esl_dmatrix_Create(int n, int m)
{
  ESL_DMATRIX *A;
  int          r;

  if ((A = malloc(sizeof(ESL_DMATRIX))) == NULL) return NULL;
  A->mx = malloc(sizeof(double *) * (size_t) n);
  if (A->mx == NULL) { free(A); return NULL; }
  A->mx[0] = malloc(sizeof(double) * (size_t) n * (size_t) m);
  if (A->mx[0] == NULL) { free(A->mx); free(A); return NULL; }
  for (r = 1; r < n; r++) A->mx[r] = A->mx[0] + (size_t) r * (size_t) m;
  A->n = n;
  A->m = m;
  A->type = eslGENERAL;
  A->ncells = n * m;
  return A;
}

int
esl_dmatrix_Copy(const ESL_DMATRIX *src, ESL_DMATRIX *dest)
{
  if (src->n != dest->n || src->m != dest->m) return eslEINCOMPAT;
  memcpy(dest->mx[0], src->mx[0], sizeof(double) * (size_t) src->n * (size_t) src->m);
  return eslOK;
}

ESL_DMATRIX *
esl_dmatrix_Clone(const ESL_DMATRIX *old)
{
  ESL_DMATRIX *A = esl_dmatrix_Create(old->n, old->m);
  if (A == NULL) return NULL;
  esl_dmatrix_Copy(old, A);
  return A;
}

void
esl_dmatrix_Destroy(ESL_DMATRIX *A)
{
  if (A == NULL) return;
  if (A->mx) { free(A->mx[0]); free(A->mx); }
  free(A);
}

int
esl_dmatrix_Set(ESL_DMATRIX *A, double x)
{
  size_t i, tot = (size_t) A->n * (size_t) A->m;
  for (i = 0; i < tot; i++) A->mx[0][i] = x;
  return eslOK;
}

int esl_dmatrix_SetZero(ESL_DMATRIX *A) { return esl_dmatrix_Set(A, 0.0); }

int
esl_dmatrix_SetIdentity(ESL_DMATRIX *A)
{
  int i;
  if (A->n != A->m) return eslEINVAL;
  esl_dmatrix_Set(A, 0.0);
  for (i = 0; i < A->n; i++) A->mx[i][i] = 1.0;
  return eslOK;
}

int
esl_dmatrix_Dump(FILE *ofp, const ESL_DMATRIX *A, const char *rowlabel, const char *collabel)
{
  int i, j;
  (void) rowlabel; (void) collabel;
  for (i = 0; i < A->n; i++) {
    for (j = 0; j < A->m; j++) fprintf(ofp, "%12.6g ", A->mx[i][j]);
    fprintf(ofp, "\n");
  }
  return eslOK;
}

int
esl_dmx_Multiply(const ESL_DMATRIX *A, const ESL_DMATRIX *B, ESL_DMATRIX *C)
{
  int i, j, k;
  if (A->m != B->n || A->n != C->n || B->m != C->m) return eslEINVAL;
  for (i = 0; i < A->n; i++)
    for (j = 0; j < B->m; j++) {
      double s = 0.0;
      for (k = 0; k < A->m; k++) s += A->mx[i][k] * B->mx[k][j];
      C->mx[i][j] = s;
    }
  return eslOK;
}

int
esl_dmx_Scale(ESL_DMATRIX *A, double k)
{
  size_t i, tot = (size_t) A->n * (size_t) A->m;
  for (i = 0; i < tot; i++) A->mx[0][i] *= k;
  return eslOK;
}

/* every row of a conditional matrix is a probability vector */
int
esl_rmx_ValidateP(ESL_DMATRIX *P, double tol, char *errbuf)
{
  int i, j;
  if (P->n != P->m) ESL_FAIL(eslFAIL, errbuf, "a conditional matrix P must be square");
  for (i = 0; i < P->n; i++) {
    double sum = 0.0;
    for (j = 0; j < P->m; j++) {
      if (P->mx[i][j] < 0.0 || P->mx[i][j] > 1.0) ESL_FAIL(eslFAIL, errbuf, "element %d,%d is not a probability (%f)", i, j, P->mx[i][j]);
      sum += P->mx[i][j];
    }
    if (fabs(sum - 1.0) > tol) ESL_FAIL(eslFAIL, errbuf, "row %d does not sum to 1.0", i);
  }
  return eslOK;
}

/* P = exp(tQ) by scaling-and-squaring around a Taylor series.
 * Upstream esl_dmx_Exp: scale tQ down by 2^z until its (Frobenius) norm is small, sum the
 * Taylor series to convergence, square z times.  Same scheme, own constants; results agree
 * with any convergent expm to ~1e-15 which is all the callers (ratematrix.c:208) rely on. */
int
esl_dmx_Exp(const ESL_DMATRIX *Q, double t, ESL_DMATRIX *P)
{
  int          n = Q->n;
  ESL_DMATRIX *M, *term, *tmp;
  double       norm = 0.0, fac;
  int          z = 0, i, j, k;

  if (Q->n != Q->m || P->n != n || P->m != n) return eslEINVAL;
  M    = esl_dmatrix_Create(n, n);
  term = esl_dmatrix_Create(n, n);
  tmp  = esl_dmatrix_Create(n, n);
  if (!M || !term || !tmp) return eslEMEM;

  for (i = 0; i < n; i++)
    for (j = 0; j < n; j++) { M->mx[i][j] = t * Q->mx[i][j]; norm += M->mx[i][j] * M->mx[i][j]; }
  norm = sqrt(norm);
  while (norm > 0.1) { norm *= 0.5; z++; }
  fac = ldexp(1.0, -z);
  esl_dmx_Scale(M, fac);

  esl_dmatrix_SetIdentity(P);
  esl_dmatrix_SetIdentity(term);
  for (k = 1; k < 100; k++) {
    double delta = 0.0;
    esl_dmx_Multiply(term, M, tmp);
    esl_dmx_Scale(tmp, 1.0 / (double) k);
    esl_dmatrix_Copy(tmp, term);
    for (i = 0; i < n; i++)
      for (j = 0; j < n; j++) { P->mx[i][j] += term->mx[i][j]; delta += fabs(term->mx[i][j]); }
    if (delta < 1e-18) break;
  }
  while (z-- > 0) { esl_dmx_Multiply(P, P, tmp); esl_dmatrix_Copy(tmp, P); }

  esl_dmatrix_Destroy(M);
  esl_dmatrix_Destroy(term);
  esl_dmatrix_Destroy(tmp);
  return eslOK;
}

/* ------------------------------------------------------------------ vector ops */
void esl_vec_DSet(double *vec, int n, double value) { int i; for (i = 0; i < n; i++) vec[i] = value; }
void esl_vec_ISet(int    *vec, int n, int    value) { int i; for (i = 0; i < n; i++) vec[i] = value; }
void esl_vec_FSet(float  *vec, int n, float  value) { int i; for (i = 0; i < n; i++) vec[i] = value; }
void esl_vec_DCopy(const double *src, int n, double *dest) { memcpy(dest, src, sizeof(double) * (size_t) n); }
void esl_vec_ICopy(const int    *src, int n, int    *dest) { memcpy(dest, src, sizeof(int)    * (size_t) n); }
void esl_vec_FScale(float *vec, int n, float scale) { int i; for (i = 0; i < n; i++) vec[i] *= scale; }

/* Sums are Kahan-compensated, as upstream's esl_vec_{D,F}Sum. */
double
esl_vec_DSum(const double *vec, int n)
{
  double sum = 0.0, c = 0.0, y, t;
  int    i;
  for (i = 0; i < n; i++) { y = vec[i] - c; t = sum + y; c = (t - sum) - y; sum = t; }
  return sum;
}

float
esl_vec_FSum(const float *vec, int n)
{
  float sum = 0.0f, c = 0.0f, y, t;
  int   i;
  for (i = 0; i < n; i++) { y = vec[i] - c; t = sum + y; c = (t - sum) - y; sum = t; }
  return sum;
}

/* divide by the sum; an exactly-zero sum yields the uniform vector */
void
esl_vec_DNorm(double *vec, int n)
{
  double sum = esl_vec_DSum(vec, n);
  int    i;
  if (sum != 0.0) for (i = 0; i < n; i++) vec[i] /= sum;
  else            for (i = 0; i < n; i++) vec[i] = 1.0 / (double) n;
}

int
esl_vec_DValidate(const double *vec, int n, double tol, char *errbuf)
{
  double sum = 0.0;
  int    i;
  if (errbuf) errbuf[0] = '\0';
  if (n == 0) return eslOK;
  for (i = 0; i < n; i++) {
    if (!isfinite(vec[i]) || vec[i] < 0.0 || vec[i] > 1.0) ESL_FAIL(eslFAIL, errbuf, "value %d is not a probability between 0..1", i);
    sum += vec[i];
  }
  if (fabs(sum - 1.0) > tol) ESL_FAIL(eslFAIL, errbuf, "vector does not sum to 1.0");
  return eslOK;
}

int
esl_vec_DDump(FILE *ofp, const double *v, int n, const char *label)
{
  int i;
  fprintf(ofp, "     ");
  for (i = 0; i < n; i++) { if (label) fprintf(ofp, "         %c ", label[i]); else fprintf(ofp, "%10d ", i + 1); }
  fprintf(ofp, "\n      ");
  for (i = 0; i < n; i++) fprintf(ofp, "%10.6f ", v[i]);
  fprintf(ofp, "\n");
  return eslOK;
}

/* ------------------------------------------------------------------ integer stack */
ESL_STACK *
esl_stack_ICreate(void)
{
  ESL_STACK *s = malloc(sizeof(ESL_STACK));
  if (s == NULL) return NULL;
  s->nalloc = 128;
  s->n      = 0;
  s->idata  = malloc(sizeof(int) * (size_t) s->nalloc);
  if (s->idata == NULL) { free(s); return NULL; }
  return s;
}

int
esl_stack_IPush(ESL_STACK *s, int x)
{
  if (s->n == s->nalloc) {
    int *p = realloc(s->idata, sizeof(int) * (size_t) s->nalloc * 2);
    if (p == NULL) return eslEMEM;
    s->idata = p;
    s->nalloc *= 2;
  }
  s->idata[s->n++] = x;
  return eslOK;
}

int
esl_stack_IPop(ESL_STACK *s, int *ret_x)
{
  if (s->n == 0) { *ret_x = 0; return eslEOD; }
  *ret_x = s->idata[--s->n];
  return eslOK;
}

int  esl_stack_ObjectCount(ESL_STACK *s) { return s->n; }
void esl_stack_Destroy(ESL_STACK *s) { if (s) { free(s->idata); free(s); } }

/* ------------------------------------------------------------------ Mersenne Twister (MT19937) */
static void
mt_refill(ESL_RANDOMNESS *r)
{
  static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
  uint32_t y;
  int      z;
  for (z = 0; z < 227; z++) { y = (r->mt[z] & 0x80000000u) | (r->mt[z+1] & 0x7fffffffu); r->mt[z] = r->mt[z+397] ^ (y >> 1) ^ mag01[y & 1u]; }
  for (     ; z < 623; z++) { y = (r->mt[z] & 0x80000000u) | (r->mt[z+1] & 0x7fffffffu); r->mt[z] = r->mt[z-227] ^ (y >> 1) ^ mag01[y & 1u]; }
  y = (r->mt[623] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
  r->mt[623] = r->mt[396] ^ (y >> 1) ^ mag01[y & 1u];
  r->mti = 0;
}

ESL_RANDOMNESS *
esl_randomness_Create(uint32_t seed)
{
  ESL_RANDOMNESS *r = calloc(1, sizeof(ESL_RANDOMNESS));
  int             z;
  if (r == NULL) return NULL;
  if (seed == 0) seed = 42;
  r->seed  = seed;
  r->mt[0] = seed;
  for (z = 1; z < 624; z++) r->mt[z] = 69069u * r->mt[z-1];
  mt_refill(r);
  return r;
}

void esl_randomness_Destroy(ESL_RANDOMNESS *r) { free(r); }

/* uniform on [0,1) */
double
esl_random(ESL_RANDOMNESS *r)
{
  uint32_t x;
  if (r->mti >= 624) mt_refill(r);
  x  = r->mt[r->mti++];
  x ^= (x >> 11);
  x ^= (x <<  7) & 0x9d2c5680u;
  x ^= (x << 15) & 0xefc60000u;
  x ^= (x >> 18);
  return (double) x / 4294967296.0;
}

int
esl_rnd_DChoose(ESL_RANDOMNESS *r, const double *p, int N)
{
  double roll = esl_random(r), sum = 0.0;
  int    i;
  for (i = 0; i < N; i++) { sum += p[i]; if (roll < sum) return i; }
  do { i = (int) (esl_random(r) * N); } while (p[i] == 0.0);
  return i;
}

int
esl_rnd_FChoose(ESL_RANDOMNESS *r, const float *p, int N)
{
  float roll = (float) esl_random(r), sum = 0.0f;
  int   i;
  for (i = 0; i < N; i++) { sum += p[i]; if (roll < sum) return i; }
  do { i = (int) (esl_random(r) * N); } while (p[i] == 0.0f);
  return i;
}

/* Fisher-Yates */
int
esl_vec_IShuffle(ESL_RANDOMNESS *r, int *v, int n)
{
  int w, t;
  for (; n > 1; n--) { w = (int) (esl_random(r) * n); t = v[w]; v[w] = v[n-1]; v[n-1] = t; }
  return eslOK;
}

/* ------------------------------------------------------------------ MSA (digital subset) */
ESL_MSA *
esl_msa_CreateDigital(const ESL_ALPHABET *abc, int nseq, int64_t alen)
{
  ESL_MSA *msa = calloc(1, sizeof(ESL_MSA));
  int      i;
  if (msa == NULL) return NULL;
  msa->alen    = alen;
  msa->nseq    = nseq;
  msa->sqalloc = nseq;
  msa->flags   = eslMSA_DIGITAL;
  msa->abc     = (ESL_ALPHABET *) abc;
  msa->lastidx = 0;
  msa->sqname  = calloc((size_t) nseq, sizeof(char *));
  msa->wgt     = malloc(sizeof(double) * (size_t) nseq);
  msa->sqlen   = calloc((size_t) nseq, sizeof(int64_t));
  msa->ax      = calloc((size_t) nseq, sizeof(ESL_DSQ *));
  for (i = 0; i < nseq; i++) {
    msa->wgt[i] = 1.0;
    if (alen >= 0) {
      msa->ax[i] = malloc((size_t) alen + 2);
      memset(msa->ax[i], abc ? abc->K : 4, (size_t) alen + 2);
      msa->ax[i][0] = msa->ax[i][alen+1] = eslDSQ_SENTINEL;
    }
  }
  return msa;
}

ESL_MSA *
esl_msa_Clone(const ESL_MSA *msa)
{
  ESL_MSA *nw = esl_msa_CreateDigital(msa->abc, msa->nseq, msa->alen);
  int      i;
  if (nw == NULL) return NULL;
  for (i = 0; i < msa->nseq; i++) {
    memcpy(nw->ax[i], msa->ax[i], (size_t) msa->alen + 2);
    nw->wgt[i] = msa->wgt[i];
    if (msa->sqname && msa->sqname[i]) esl_strdup(msa->sqname[i], -1, &nw->sqname[i]);
    if (msa->sqlen) nw->sqlen[i] = msa->sqlen[i];
  }
  nw->flags = msa->flags;
  if (msa->ss_cons) esl_strdup(msa->ss_cons, -1, &nw->ss_cons);
  if (msa->name)    esl_strdup(msa->name,    -1, &nw->name);
  return nw;
}

int
esl_msa_SequenceSubset(const ESL_MSA *msa, const int *useme, ESL_MSA **ret_new)
{
  ESL_MSA *nw;
  int      i, n = 0, k = 0;
  for (i = 0; i < msa->nseq; i++) if (useme[i]) n++;
  if (n == 0) { *ret_new = NULL; return eslFAIL; }
  nw = esl_msa_CreateDigital(msa->abc, n, msa->alen);
  if (nw == NULL) { *ret_new = NULL; return eslEMEM; }
  for (i = 0; i < msa->nseq; i++) {
    if (!useme[i]) continue;
    memcpy(nw->ax[k], msa->ax[i], (size_t) msa->alen + 2);
    nw->wgt[k] = msa->wgt[i];
    if (msa->sqname && msa->sqname[i]) esl_strdup(msa->sqname[i], -1, &nw->sqname[k]);
    k++;
  }
  if (msa->ss_cons) esl_strdup(msa->ss_cons, -1, &nw->ss_cons);
  *ret_new = nw;
  return eslOK;
}

void
esl_msa_Destroy(ESL_MSA *msa)
{
  int i;
  if (msa == NULL) return;
  for (i = 0; i < msa->nseq; i++) {
    if (msa->ax)     free(msa->ax[i]);
    if (msa->sqname) free(msa->sqname[i]);
  }
  free(msa->ax); free(msa->sqname); free(msa->wgt); free(msa->sqlen);
  free(msa->ss_cons); free(msa->name); free(msa->acc); free(msa->desc);
  free(msa);
}

/* ------------------------------------------------------------------ tree */
ESL_TREE *
esl_tree_Create(int ntaxa)
{
  ESL_TREE *T = calloc(1, sizeof(ESL_TREE));
  int       nn = (ntaxa > 1) ? ntaxa - 1 : 1, i;
  if (T == NULL) return NULL;
  T->N      = ntaxa;
  T->nalloc = ntaxa;
  T->parent = malloc(sizeof(int)    * (size_t) nn);
  T->left   = malloc(sizeof(int)    * (size_t) nn);
  T->right  = malloc(sizeof(int)    * (size_t) nn);
  T->ld     = malloc(sizeof(double) * (size_t) nn);
  T->rd     = malloc(sizeof(double) * (size_t) nn);
  for (i = 0; i < nn; i++) { T->parent[i] = T->left[i] = T->right[i] = 0; T->ld[i] = T->rd[i] = 0.0; }
  T->show_branchlengths = TRUE;
  return T;
}

void
esl_tree_Destroy(ESL_TREE *T)
{
  if (T == NULL) return;
  free(T->parent); free(T->left); free(T->right); free(T->ld); free(T->rd);
  free(T->taxaparent); free(T->cladesize);
  if (T->taxonlabel) { int i; for (i = 0; i < T->nalloc; i++) free(T->taxonlabel[i]); free(T->taxonlabel); }
  if (T->nodelabel)  { int i; for (i = 0; i < T->nalloc - 1; i++) free(T->nodelabel[i]); free(T->nodelabel); }
  free(T);
}

/* ------------------------------------------------------------------ histogram */
ESL_HISTOGRAM *
esl_histogram_Create(double bmin, double bmax, double w)
{
  ESL_HISTOGRAM *h = calloc(1, sizeof(ESL_HISTOGRAM));
  if (h == NULL) return NULL;
  h->xmin = DBL_MAX;
  h->xmax = -DBL_MAX;
  h->bmin = bmin;
  h->bmax = bmax;
  h->w    = w;
  h->nb   = (int) ((bmax - bmin) / w);
  h->imin = h->nb;
  h->imax = -1;
  h->cmin = h->imin;
  h->emin = -1;
  h->tailmass = 1.0;
  h->dataset_is = COMPLETE;
  h->obs  = calloc((size_t) (h->nb > 0 ? h->nb : 1), sizeof(uint64_t));
  if (h->obs == NULL) { free(h); return NULL; }
  return h;
}

ESL_HISTOGRAM *
esl_histogram_CreateFull(double bmin, double bmax, double w)
{
  ESL_HISTOGRAM *h = esl_histogram_Create(bmin, bmax, w);
  if (h == NULL) return NULL;
  h->is_full = TRUE;
  h->nalloc  = 128;
  h->x       = malloc(sizeof(double) * h->nalloc);
  if (h->x == NULL) { esl_histogram_Destroy(h); return NULL; }
  return h;
}

void
esl_histogram_Destroy(ESL_HISTOGRAM *h)
{
  if (h == NULL) return;
  free(h->x); free(h->obs); free(h->expect);
  free(h);
}

/* bin b covers (bmin + b*w, bmin + (b+1)*w] */
int
esl_histogram_Score2Bin(ESL_HISTOGRAM *h, double x, int *ret_b)
{
  if (!isfinite(x)) { *ret_b = 0; ESL_EXCEPTION(eslERANGE, "value added to histogram is not finite"); }
  x = ceil(((x - h->bmin) / h->w) - 1.0);
  if (x < (double) INT_MIN || x > (double) INT_MAX) { *ret_b = 0; ESL_EXCEPTION(eslERANGE, "value %f isn't going to fit in histogram", x); }
  *ret_b = (int) x;
  return eslOK;
}

int
esl_histogram_Add(ESL_HISTOGRAM *h, double x)
{
  int b, i, nnew, status;

  if ((status = esl_histogram_Score2Bin(h, x, &b)) != eslOK) return status;
  h->is_sorted = FALSE;

  if (h->is_full) {
    if (h->n == h->nalloc) {
      double *p = realloc(h->x, sizeof(double) * h->nalloc * 2);
      if (p == NULL) return eslEMEM;
      h->x = p;
      h->nalloc *= 2;
    }
    h->x[h->n] = x;
  }

  if (b < 0) {                      /* grow below: shift everything up by nnew bins */
    uint64_t *p;
    nnew = -b * 2;
    p = realloc(h->obs, sizeof(uint64_t) * (size_t) (nnew + h->nb));
    if (p == NULL) return eslEMEM;
    h->obs = p;
    memmove(h->obs + nnew, h->obs, sizeof(uint64_t) * (size_t) h->nb);
    h->nb   += nnew;
    b       += nnew;
    h->bmin -= nnew * h->w;
    h->imin += nnew;
    h->cmin += nnew;
    if (h->imax > -1) h->imax += nnew;
    for (i = 0; i < nnew; i++) h->obs[i] = 0;
  } else if (b >= h->nb) {          /* grow above */
    uint64_t *p;
    nnew = (b - h->nb + 1) * 2;
    p = realloc(h->obs, sizeof(uint64_t) * (size_t) (nnew + h->nb));
    if (p == NULL) return eslEMEM;
    h->obs = p;
    for (i = h->nb; i < h->nb + nnew; i++) h->obs[i] = 0;
    if (h->imin == h->nb) { h->imin += nnew; h->cmin += nnew; }
    h->bmax += nnew * h->w;
    h->nb   += nnew;
  }

  h->obs[b]++;
  h->n++;
  h->Nc++;
  h->No++;
  if (b > h->imax) h->imax = b;
  if (b < h->imin) { h->imin = b; h->cmin = b; }
  if (x > h->xmax) h->xmax = x;
  if (x < h->xmin) h->xmin = x;
  return eslOK;
}

/* ------------------------------------------------------------------ esl_fileparser (the subset R-scape's readers use)
 * Lines of whitespace-delimited tokens; everything from the comment character to the end of a line is skipped, and so are
 * lines without a token.  esl_fileparser_NextLine moves to the next line that holds a token (eslEOF at the end of the file),
 * esl_fileparser_GetTokenOnLine hands out its tokens one by one (eslEOL when the line is used up).  The token points into the
 * parser's line buffer and stays valid until the next NextLine. */
struct esl_fileparser_s { FILE *fp; char *buf; size_t cap; char *pos; char comment; };

int
esl_fileparser_Open(const char *filename, const char *envvar, ESL_FILEPARSER **ret_efp)
{
  ESL_FILEPARSER *efp;
  (void) envvar;
  *ret_efp = NULL;
  if ((efp = calloc(1, sizeof(*efp))) == NULL) return eslEMEM;
  if ((efp->fp = fopen(filename, "r")) == NULL) { free(efp); return eslENOTFOUND; }
  *ret_efp = efp;
  return eslOK;
}

int
esl_fileparser_SetCommentChar(ESL_FILEPARSER *efp, char c)
{
  efp->comment = c;
  return eslOK;
}

static int
fileparser_readline(ESL_FILEPARSER *efp)
{
  size_t n = 0;
  int    c;
  while ((c = fgetc(efp->fp)) != EOF) {
    if (n + 2 > efp->cap) {
      size_t cap = efp->cap ? 2 * efp->cap : 256;
      char  *p = realloc(efp->buf, cap);
      if (p == NULL) return eslEMEM;
      efp->buf = p; efp->cap = cap;
    }
    if (c == '\n') break;
    efp->buf[n++] = (char) c;
  }
  if (c == EOF && n == 0) return eslEOF;
  if (efp->buf == NULL) { if ((efp->buf = malloc(256)) == NULL) return eslEMEM; efp->cap = 256; }
  efp->buf[n] = 0;
  if (efp->comment) { char *h = strchr(efp->buf, efp->comment); if (h) *h = 0; }
  efp->pos = efp->buf;
  return eslOK;
}

int
esl_fileparser_NextLine(ESL_FILEPARSER *efp)
{
  int status;
  while ((status = fileparser_readline(efp)) == eslOK) {
    char *p = efp->pos;
    while (*p == ' ' || *p == '\t' || *p == '\r') p++;
    if (*p) { efp->pos = p; return eslOK; }          /* a line with at least one token */
  }
  return status;
}

int
esl_fileparser_GetTokenOnLine(ESL_FILEPARSER *efp, char **opt_tok, int *opt_toklen)
{
  char *p = efp->pos, *t;
  if (p == NULL) return eslEOL;
  while (*p == ' ' || *p == '\t' || *p == '\r') p++;
  if (*p == 0) { efp->pos = p; return eslEOL; }
  t = p;
  while (*p && *p != ' ' && *p != '\t' && *p != '\r') p++;
  if (opt_toklen) *opt_toklen = (int) (p - t);
  if (*p) { *p = 0; p++; }
  efp->pos = p;
  if (opt_tok) *opt_tok = t;
  return eslOK;
}

void
esl_fileparser_Close(ESL_FILEPARSER *efp)
{
  if (efp == NULL) return;
  if (efp->fp) fclose(efp->fp);
  free(efp->buf);
  free(efp);
}
