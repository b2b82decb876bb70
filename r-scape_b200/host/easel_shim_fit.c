/* easel_shim_fit.c -- second part of the Easel-compatible shim: what R-scape's host code needs around the covariation path to go
 * from FastTree's output to E-values without an Easel checkout:
 *
 *   esl_tree_ReadNewick / RenumberNodes / SetTaxaParents / SetCladesizes / Validate / Grow
 *        callers: Tree_CalculateExtFromMSA, Tree_ReorderTaxaAccordingMSA, Tree_RootAtMidPoint  (src/msatree.c:49-105, 524-790, 823-912)
 *   esl_histogram_SetTailByMass, esl_gam_FitCompleteBinned, esl_gam_generic_surv, esl_exp_FitCompleteBinned, esl_exp_generic_surv,
 *   esl_stats_IncompleteGamma
 *        callers: cov_NullFitGamma / cov_NullFitExponential                                   (src/covariation.c:1915-1973)
 *
 * Easel is an un-vendored submodule of the reference (no version pin in the tree), so these are restatements of its published
 * behaviour, NOT copies: PARITY UNPINNED except through the tutorial transcript's 11 significant pairs
 * (tests/test_gpu_config1.py).  Known freedoms: (i) a Newick polytomy (FastTree emits them for identical sequences, and a
 * trifurcation at the top of its unrooted trees) is resolved into binary nodes joined by zero-length branches; (ii) the binned
 * maximum-likelihood gamma fit is found by a simplex search on (log lambda, log tau) to a tighter tolerance than a conjugate-
 * gradient run would reach -- the optimum itself is the same.  Inside a real R-scape tree this file is left off the link line. */
#include <ctype.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "easel.h"

/* ------------------------------------------------------------------ tree helpers */
int
esl_tree_SetTaxaParents(ESL_TREE *T)
{
  int v;
  if (T->taxaparent == NULL) {
    T->taxaparent = malloc(sizeof(int) * (size_t) (T->nalloc > T->N ? T->nalloc : T->N));
    if (T->taxaparent == NULL) return eslEMEM;
  }
  for (v = 0; v < T->N; v++) T->taxaparent[v] = 0;
  for (v = 0; v < T->N - 1; v++) {
    if (T->left[v]  <= 0) T->taxaparent[-T->left[v]]  = v;
    if (T->right[v] <= 0) T->taxaparent[-T->right[v]] = v;
  }
  return eslOK;
}

/* number of taxa below every internal node; relies on preorder numbering (children have larger indices than their parent) */
int
esl_tree_SetCladesizes(ESL_TREE *T)
{
  int v;
  if (T->cladesize == NULL) {
    T->cladesize = malloc(sizeof(int) * (size_t) (T->nalloc > T->N ? T->nalloc : T->N));
    if (T->cladesize == NULL) return eslEMEM;
  }
  for (v = 0; v < T->N - 1; v++) T->cladesize[v] = 0;
  for (v = T->N - 2; v >= 0; v--) {
    T->cladesize[v] += (T->left[v]  > 0) ? T->cladesize[T->left[v]]  : 1;
    T->cladesize[v] += (T->right[v] > 0) ? T->cladesize[T->right[v]] : 1;
  }
  return eslOK;
}

/* internal nodes renumbered in preorder (root 0, a node's left subtree before its right one); taxon indices unchanged */
int
esl_tree_RenumberNodes(ESL_TREE *T)
{
  int  nn = T->N - 1, top = 0, next = 0, v, changed = FALSE;
  int *map, *stack, *parent, *left, *right, *tp = NULL, *cs = NULL;
  double *ld, *rd;
  char **nl = NULL;

  if (nn < 1) return eslOK;
  map = malloc(sizeof(int) * (size_t) nn); stack = malloc(sizeof(int) * (size_t) (nn + 1));
  if (!map || !stack) { free(map); free(stack); return eslEMEM; }
  stack[top++] = 0;
  while (top > 0) {
    v = stack[--top];
    if (v != next) changed = TRUE;
    map[v] = next++;
    if (T->right[v] > 0) stack[top++] = T->right[v];
    if (T->left[v]  > 0) stack[top++] = T->left[v];
  }
  free(stack);
  if (next != nn) { free(map); return eslEINCONCEIVABLE; }
  if (!changed) { free(map); return eslOK; }

  {
    size_t na = (size_t) (T->nalloc > T->N ? T->nalloc : T->N);
    parent = malloc(sizeof(int) * na); left = malloc(sizeof(int) * na); right = malloc(sizeof(int) * na);
    ld = malloc(sizeof(double) * na); rd = malloc(sizeof(double) * na);
    if (T->taxaparent) tp = malloc(sizeof(int) * na);
    if (T->cladesize)  cs = malloc(sizeof(int) * na);
    if (T->nodelabel)  nl = calloc(na, sizeof(char *));
  }
  for (v = 0; v < nn; v++) {
    const int m = map[v];
    parent[m] = map[T->parent[v]];
    left[m]   = (T->left[v]  > 0) ? map[T->left[v]]  : T->left[v];
    right[m]  = (T->right[v] > 0) ? map[T->right[v]] : T->right[v];
    ld[m] = T->ld[v]; rd[m] = T->rd[v];
    if (tp) { if (T->left[v] <= 0) tp[-T->left[v]] = m; if (T->right[v] <= 0) tp[-T->right[v]] = m; }
    if (cs) cs[m] = T->cladesize[v];
    if (nl) nl[m] = T->nodelabel[v];
  }
  parent[0] = 0;
  free(T->parent); free(T->left); free(T->right); free(T->ld); free(T->rd);
  T->parent = parent; T->left = left; T->right = right; T->ld = ld; T->rd = rd;
  if (tp) { free(T->taxaparent); T->taxaparent = tp; }
  if (cs) { free(T->cladesize);  T->cladesize  = cs; }
  if (nl) { free(T->nodelabel);  T->nodelabel  = nl; }
  free(map);
  return eslOK;
}

int
esl_tree_Validate(ESL_TREE *T, char *errbuf)
{
  int v, *seen;
  if (T == NULL || T->N < 1) ESL_FAIL(eslFAIL, errbuf, "number of taxa is < 1");
  if (T->N == 1) return eslOK;
  if (T->parent[0] != 0) ESL_FAIL(eslFAIL, errbuf, "parent of root 0 should be set to 0");
  seen = calloc((size_t) T->N, sizeof(int));
  for (v = 0; v < T->N - 1; v++) {
    const int c[2] = { T->left[v], T->right[v] };
    int k;
    for (k = 0; k < 2; k++) {
      if (c[k] > 0) {
        if (c[k] >= T->N - 1 || c[k] <= v)  { free(seen); ESL_FAIL(eslFAIL, errbuf, "node %d: child %d is not numbered in preorder", v, c[k]); }
        if (T->parent[c[k]] != v)          { free(seen); ESL_FAIL(eslFAIL, errbuf, "node %d: child %d has parent %d", v, c[k], T->parent[c[k]]); }
      } else {
        if (-c[k] >= T->N || seen[-c[k]]++) { free(seen); ESL_FAIL(eslFAIL, errbuf, "node %d: taxon %d out of range or seen twice", v, -c[k]); }
        if (T->taxaparent && T->taxaparent[-c[k]] != v) { free(seen); ESL_FAIL(eslFAIL, errbuf, "taxon %d: taxaparent is %d, not %d", -c[k], T->taxaparent[-c[k]], v); }
      }
    }
    if (!(T->ld[v] >= 0.0) || !(T->rd[v] >= 0.0)) { free(seen); ESL_FAIL(eslFAIL, errbuf, "node %d: negative or NaN branch length", v); }
  }
  for (v = 0; v < T->N; v++) if (!seen[v]) { free(seen); ESL_FAIL(eslFAIL, errbuf, "taxon %d is not in the tree", v); }
  free(seen);
  return eslOK;
}

int
esl_tree_Grow(ESL_TREE *T)
{
  size_t na;
  int    i;
  if (T->N < T->nalloc) return eslOK;
  na = (size_t) T->nalloc * 2;
#define RSB_GROW(p, type) do { type *q_ = realloc(T->p, sizeof(type) * na); if (!q_) return eslEMEM; T->p = q_; } while (0)
  RSB_GROW(parent, int); RSB_GROW(left, int); RSB_GROW(right, int); RSB_GROW(ld, double); RSB_GROW(rd, double);
  if (T->taxaparent) RSB_GROW(taxaparent, int);
  if (T->cladesize)  RSB_GROW(cladesize, int);
  if (T->taxonlabel) { RSB_GROW(taxonlabel, char *); for (i = T->nalloc; i < (int) na; i++) T->taxonlabel[i] = NULL; }
  if (T->nodelabel)  { RSB_GROW(nodelabel, char *);  for (i = T->nalloc; i < (int) na; i++) T->nodelabel[i]  = NULL; }
#undef RSB_GROW
  for (i = T->nalloc - 1; i < (int) na - 1; i++) { T->parent[i] = T->left[i] = T->right[i] = 0; T->ld[i] = T->rd[i] = 0.0; }
  T->nalloc = (int) na;
  return eslOK;
}

/* ------------------------------------------------------------------ Newick */
typedef struct nwk_node_s {
  struct nwk_node_s *kid[2];
  double             len;          /* length of the branch above this node */
  char              *label;        /* taxon name (leaves) */
} NWK_NODE;

typedef struct { const char *s; size_t pos, n; int ntaxa; char err[128]; } NWK_PARSE;

static void nwk_skip(NWK_PARSE *p)
{
  for (;;) {
    while (p->pos < p->n && isspace((unsigned char) p->s[p->pos])) p->pos++;
    if (p->pos < p->n && p->s[p->pos] == '[') {                      /* comment */
      while (p->pos < p->n && p->s[p->pos] != ']') p->pos++;
      if (p->pos < p->n) p->pos++;
    } else break;
  }
}

static void nwk_free(NWK_NODE *x)
{
  if (!x) return;
  nwk_free(x->kid[0]); nwk_free(x->kid[1]);
  free(x->label); free(x);
}

static char *nwk_label(NWK_PARSE *p)
{
  size_t a, len = 0, cap = 32;
  char  *out = malloc(cap);
  nwk_skip(p);
  if (p->pos < p->n && p->s[p->pos] == '\'') {                       /* quoted: '' is a literal quote */
    p->pos++;
    while (p->pos < p->n) {
      if (p->s[p->pos] == '\'') { if (p->pos + 1 < p->n && p->s[p->pos + 1] == '\'') p->pos++; else { p->pos++; break; } }
      if (len + 2 > cap) out = realloc(out, cap *= 2);
      out[len++] = p->s[p->pos++];
    }
  } else {
    a = p->pos;
    while (p->pos < p->n && !strchr("(),:;[", p->s[p->pos]) && !isspace((unsigned char) p->s[p->pos])) p->pos++;
    len = p->pos - a;
    if (len + 1 > cap) out = realloc(out, cap = len + 1);
    memcpy(out, p->s + a, len);
  }
  out[len] = '\0';
  return out;
}

static NWK_NODE *nwk_subtree(NWK_PARSE *p)
{
  NWK_NODE *x = calloc(1, sizeof(NWK_NODE));
  char     *lab;
  nwk_skip(p);
  if (p->pos < p->n && p->s[p->pos] == '(') {
    NWK_NODE *acc = NULL;
    int       nk = 0;
    p->pos++;
    for (;;) {
      NWK_NODE *k = nwk_subtree(p);
      if (!k) { nwk_free(acc); free(x); return NULL; }
      if (nk == 0) acc = k;
      else if (nk == 1) { x->kid[0] = acc; x->kid[1] = k; acc = NULL; }
      else {                                                         /* polytomy: ((a,b):0,c) */
        NWK_NODE *j = calloc(1, sizeof(NWK_NODE));
        j->kid[0] = x->kid[0]; j->kid[1] = x->kid[1]; j->len = 0.0;
        x->kid[0] = j; x->kid[1] = k;
      }
      nk++;
      nwk_skip(p);
      if (p->pos < p->n && p->s[p->pos] == ',') { p->pos++; continue; }
      if (p->pos < p->n && p->s[p->pos] == ')') { p->pos++; break; }
      snprintf(p->err, sizeof(p->err), "expected , or ) at position %zu", p->pos);
      nwk_free(acc); nwk_free(x); return NULL;
    }
    if (nk == 1) {                                                   /* a single child in parentheses: splice it out */
      NWK_NODE *only = acc;
      free(x);
      x = only;
      lab = nwk_label(p); free(lab);
      nwk_skip(p);
      if (p->pos < p->n && p->s[p->pos] == ':') { p->pos++; nwk_skip(p); x->len += strtod(p->s + p->pos, (char **) &lab); p->pos = (size_t) (lab - p->s); }
      return x;
    }
    lab = nwk_label(p);                                              /* internal label (FastTree: a support value) -> T->nodelabel */
    if (lab[0]) x->label = lab; else free(lab);
  } else {
    x->label = nwk_label(p);
    if (x->label[0] == '\0') { snprintf(p->err, sizeof(p->err), "empty taxon name at position %zu", p->pos); nwk_free(x); return NULL; }
    p->ntaxa++;
  }
  nwk_skip(p);
  if (p->pos < p->n && p->s[p->pos] == ':') {
    char *end;
    p->pos++; nwk_skip(p);
    x->len = strtod(p->s + p->pos, &end);
    if (end == p->s + p->pos) { snprintf(p->err, sizeof(p->err), "bad branch length at position %zu", p->pos); nwk_free(x); return NULL; }
    p->pos = (size_t) (end - p->s);
    if (x->len < 0.0) x->len = 0.0;                                  /* FastTree can print tiny negative lengths */
  }
  return x;
}

/* preorder numbering of the internal nodes; taxa numbered in order of appearance */
static int nwk_fill(ESL_TREE *T, NWK_NODE *x, int parent, int *next_node, int *next_taxon)
{
  int v, k;
  if (!x->kid[0]) {
    const int t = (*next_taxon)++;
    T->taxonlabel[t] = x->label; x->label = NULL;
    T->taxaparent[t] = parent;
    return -t;
  }
  v = (*next_node)++;
  T->parent[v] = (parent < 0) ? 0 : parent;
  T->nodelabel[v] = x->label; x->label = NULL;
  for (k = 0; k < 2; k++) {
    const int c = nwk_fill(T, x->kid[k], v, next_node, next_taxon);
    if (k == 0) { T->left[v]  = c; T->ld[v] = x->kid[k]->len; }
    else        { T->right[v] = c; T->rd[v] = x->kid[k]->len; }
  }
  return v;
}

int
esl_tree_ReadNewick(FILE *fp, char *errbuf, ESL_TREE **ret_T)
{
  NWK_PARSE p;
  NWK_NODE *root;
  ESL_TREE *T;
  char     *buf = NULL;
  size_t    len = 0, cap = 0;
  int       c, nn = 0, nt = 0;

  *ret_T = NULL;
  while ((c = fgetc(fp)) != EOF) {
    if (len + 2 > cap) { cap = cap ? cap * 2 : 4096; buf = realloc(buf, cap); }
    buf[len++] = (char) c;
    if (c == ';') break;
  }
  if (!buf) ESL_FAIL(eslEOF, errbuf, "no tree in the file");
  buf[len] = '\0';
  memset(&p, 0, sizeof(p));
  p.s = buf; p.n = len;
  root = nwk_subtree(&p);
  if (!root) { if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "Newick: %.100s", p.err); free(buf); return eslEFORMAT; }
  if (p.ntaxa < 2 || !root->kid[0]) { nwk_free(root); free(buf); ESL_FAIL(eslEFORMAT, errbuf, "Newick: fewer than two taxa"); }
  T = esl_tree_Create(p.ntaxa);
  T->taxonlabel = calloc((size_t) p.ntaxa, sizeof(char *));
  T->taxaparent = calloc((size_t) p.ntaxa, sizeof(int));
  T->nodelabel  = calloc((size_t) p.ntaxa, sizeof(char *));          /* always present: Tree_RootAtMidPoint reads it unconditionally (src/msatree.c:654) */
  nwk_fill(T, root, -1, &nn, &nt);
  nwk_free(root); free(buf);
  if (nn != p.ntaxa - 1 || nt != p.ntaxa) { esl_tree_Destroy(T); ESL_FAIL(eslEFORMAT, errbuf, "Newick: tree is not binary after resolving polytomies"); }
  *ret_T = T;
  return eslOK;
}

/* ------------------------------------------------------------------ incomplete gamma, gamma / exponential tails */
/* regularised P(a,x) and Q(a,x) = 1 - P: series for x < a + 1, Lentz continued fraction otherwise */
int
esl_stats_IncompleteGamma(double a, double x, double *ret_pax, double *ret_qax)
{
  double pax, qax;
  if (!(a > 0.0) || !(x >= 0.0)) return eslERANGE;
  if (x == 0.0) { pax = 0.0; qax = 1.0; }
  else if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    int    it;
    for (it = 0; it < 100000; it++) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
    pax = sum * exp(-x + a * log(x) - lgamma(a));
    qax = 1.0 - pax;
  } else {
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    int    it;
    for (it = 1; it < 100000; it++) {
      const double an = -(double) it * ((double) it - a);
      double del;
      b += 2.0;
      d = an * d + b; if (fabs(d) < tiny) d = tiny;
      c = b + an / c; if (fabs(c) < tiny) c = tiny;
      d = 1.0 / d;
      del = d * c;
      h *= del;
      if (fabs(del - 1.0) < 1e-16) break;
    }
    qax = exp(-x + a * log(x) - lgamma(a)) * h;
    pax = 1.0 - qax;
  }
  if (ret_pax) *ret_pax = pax;
  if (ret_qax) *ret_qax = qax;
  return eslOK;
}

double
esl_gam_cdf(double x, double mu, double lambda, double tau)
{
  double p;
  if (x <= mu) return 0.0;
  esl_stats_IncompleteGamma(tau, lambda * (x - mu), &p, NULL);
  return p;
}

double
esl_gam_surv(double x, double mu, double lambda, double tau)
{
  double q;
  if (x <= mu) return 1.0;
  esl_stats_IncompleteGamma(tau, lambda * (x - mu), NULL, &q);
  return q;
}

double esl_gam_generic_surv(double x, void *params) { double *p = (double *) params; return esl_gam_surv(x, p[0], p[1], p[2]); }
double esl_exp_generic_surv(double x, void *params) { double *p = (double *) params; return (x < p[0]) ? 1.0 : exp(-p[1] * (x - p[0])); }

/* phi, cmin, z for a tail holding (at least) the fraction pmass of the scores, counted from the top bin down */
int
esl_histogram_SetTailByMass(ESL_HISTOGRAM *h, double pmass, double *ret_newmass)
{
  uint64_t sum = 0;
  int      b;
  for (b = h->imax; b >= h->imin; b--) {
    sum += h->obs[b];
    if ((double) sum >= pmass * (double) h->n) break;
  }
  if (b < h->imin) b = h->imin;
  h->phi        = esl_histogram_Bin2LBound(h, b);
  h->z          = h->n - sum;
  h->cmin       = b;
  h->Nc         = h->n;
  h->No         = h->n - h->z;
  h->dataset_is = VIRTUAL_CENSORED;
  h->is_tailfit = TRUE;
  if (ret_newmass) *ret_newmass = (double) sum / (double) h->n;
  return eslOK;
}

static double tail_mu(ESL_HISTOGRAM *h)
{
  if (h->dataset_is == VIRTUAL_CENSORED) return h->phi;
  return h->is_rounded ? esl_histogram_Bin2LBound(h, h->imin) : h->xmin;
}

/* complete exponential, binned data: closed-form maximum-likelihood rate */
int
esl_exp_FitCompleteBinned(ESL_HISTOGRAM *h, double *ret_mu, double *ret_lambda)
{
  const double mu = tail_mu(h), delta = h->w;
  double sa = 0.0, sb = 0.0;
  int    i;
  if (h->dataset_is == TRUE_CENSORED) return eslEINVAL;
  for (i = h->cmin; i <= h->imax; i++) {
    if (h->obs[i] == 0) continue;
    sa += (double) h->obs[i] * (esl_histogram_Bin2LBound(h, i) - mu);
    sb += (double) h->obs[i];
  }
  *ret_mu     = mu;
  *ret_lambda = (sa > 0.0) ? log(sb * delta / sa + 1.0) / delta : eslINFINITY;
  return eslOK;
}

/* -log likelihood of the binned tail under a gamma(mu; lambda, tau) */
static double
gam_binned_nll(const ESL_HISTOGRAM *h, double mu, double loglambda, double logtau)
{
  const double lambda = exp(loglambda), tau = exp(logtau);
  double nll = 0.0;
  int    i;
  for (i = h->cmin; i <= h->imax; i++) {
    double ai, bi, d;
    if (h->obs[i] == 0) continue;
    ai = esl_histogram_Bin2LBound(h, i); bi = esl_histogram_Bin2UBound(h, i);
    if (ai < mu) ai = mu;
    d = esl_gam_cdf(bi, mu, lambda, tau) - esl_gam_cdf(ai, mu, lambda, tau);
    if (!(d > 0.0)) d = esl_gam_surv(ai, mu, lambda, tau) - esl_gam_surv(bi, mu, lambda, tau);   /* far tail: the survival side keeps digits */
    if (!(d > 0.0)) return eslINFINITY;
    nll -= (double) h->obs[i] * log(d);
  }
  return nll;
}

int
esl_gam_FitCompleteBinned(ESL_HISTOGRAM *h, double *ret_mu, double *ret_lambda, double *ret_tau)
{
  const double mu = tail_mu(h);
  double s[3][2], f[3], mean = 0.0, var = 0.0, n = 0.0;
  int    i, it;

  if (h->dataset_is == TRUE_CENSORED) return eslEINVAL;
  /* starting point: method of moments on the bin centres */
  for (i = h->cmin; i <= h->imax; i++) { const double x = esl_histogram_Bin2LBound(h, i) + 0.5 * h->w - mu; mean += (double) h->obs[i] * x; n += (double) h->obs[i]; }
  if (!(n > 0.0)) { *ret_mu = mu; *ret_lambda = eslINFINITY; *ret_tau = 1.0; return eslOK; }
  mean /= n;
  for (i = h->cmin; i <= h->imax; i++) { const double x = esl_histogram_Bin2LBound(h, i) + 0.5 * h->w - mu - mean; var += (double) h->obs[i] * x * x; }
  var = (n > 1.0) ? var / (n - 1.0) : mean * mean;
  if (!(mean > 0.0)) { *ret_mu = mu; *ret_lambda = eslINFINITY; *ret_tau = 1.0; return eslOK; }
  if (!(var > 0.0)) var = mean * mean;
  s[0][0] = log(mean / var); s[0][1] = log(mean * mean / var);
  s[1][0] = s[0][0] + 0.5;   s[1][1] = s[0][1];
  s[2][0] = s[0][0];         s[2][1] = s[0][1] + 0.5;
  for (i = 0; i < 3; i++) f[i] = gam_binned_nll(h, mu, s[i][0], s[i][1]);

  /* Nelder-Mead on (log lambda, log tau) */
  for (it = 0; it < 2000; it++) {
    int lo = 0, hi = 0, mid, k;
    double c[2], r[2], fr;
    for (k = 1; k < 3; k++) { if (f[k] < f[lo]) lo = k; if (f[k] > f[hi]) hi = k; }
    mid = 3 - lo - hi; if (lo == hi) mid = 1;
    if (fabs(f[hi] - f[lo]) <= 1e-13 * (fabs(f[lo]) + 1e-300) &&
        fabs(s[hi][0] - s[lo][0]) + fabs(s[hi][1] - s[lo][1]) < 1e-9) break;
    c[0] = 0.5 * (s[lo][0] + s[mid][0]); c[1] = 0.5 * (s[lo][1] + s[mid][1]);
    r[0] = 2.0 * c[0] - s[hi][0];        r[1] = 2.0 * c[1] - s[hi][1];
    fr = gam_binned_nll(h, mu, r[0], r[1]);
    if (fr < f[lo]) {
      double e[2] = { 3.0 * c[0] - 2.0 * s[hi][0], 3.0 * c[1] - 2.0 * s[hi][1] };
      const double fe = gam_binned_nll(h, mu, e[0], e[1]);
      if (fe < fr) { s[hi][0] = e[0]; s[hi][1] = e[1]; f[hi] = fe; } else { s[hi][0] = r[0]; s[hi][1] = r[1]; f[hi] = fr; }
    } else if (fr < f[mid]) { s[hi][0] = r[0]; s[hi][1] = r[1]; f[hi] = fr; }
    else {
      double q[2]; double fq;
      if (fr < f[hi]) { q[0] = 0.5 * (c[0] + r[0]); q[1] = 0.5 * (c[1] + r[1]); } else { q[0] = 0.5 * (c[0] + s[hi][0]); q[1] = 0.5 * (c[1] + s[hi][1]); }
      fq = gam_binned_nll(h, mu, q[0], q[1]);
      if (fq < ((fr < f[hi]) ? fr : f[hi])) { s[hi][0] = q[0]; s[hi][1] = q[1]; f[hi] = fq; }
      else for (k = 0; k < 3; k++) if (k != lo) {                    /* shrink towards the best vertex */
        s[k][0] = 0.5 * (s[k][0] + s[lo][0]); s[k][1] = 0.5 * (s[k][1] + s[lo][1]);
        f[k] = gam_binned_nll(h, mu, s[k][0], s[k][1]);
      }
    }
  }
  { int lo = 0, k; for (k = 1; k < 3; k++) if (f[k] < f[lo]) lo = k;
    *ret_mu = mu; *ret_lambda = exp(s[lo][0]); *ret_tau = exp(s[lo][1]); }
  return eslOK;
}
