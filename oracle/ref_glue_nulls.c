/* ref_glue_nulls.c -- drive the REFERENCE's two null-alignment generators from flat arrays.
 * TEST INFRASTRUCTURE, compiled only into oracle/_ref/librscape_ref.so (needs the reference tree).
 *
 *   glue_ref_fitch_shuffle : Tree_FitchAlgorithmAncenstral (src/msatree.c:173) +
 *                            msamanip_ShuffleTreeSubstitutions (src/msamanip.c:1449), i.e. the body of
 *                            null_rscape's loop at src/R-scape.c:1653-1661
 *   glue_ref_simulate      : cov_GenerateAlignment (src/cov_simulate.c:61) with noss + noindels, as
 *                            R-scape-sim's SAMPLE_NAIVE path (src/R-scape-sim.c:967)
 *   glue_ref_ptime         : ratematrix_ConditionalsFromRate (src/ratematrix.c:185)
 *   glue_ref_tree_substitutions : Tree_Substitutions (src/msatree.c:1423), the substitution counts behind R-scape's power
 * Randomness comes from the shim's MT19937 (easel_shim.c), the same stream oracle.c consumes,
 * so the restatement in oracle.c can be compared with the reference residue for residue.
 */
#ifdef GLUE_REFERENCE
#include "rscape_config.h"
#include "easel.h"
#include "msatree.h"
#include "msamanip.h"
#include "cov_simulate.h"
#include "e1_rate.h"
#include "ratematrix.h"

extern ESL_ALPHABET *glue_abc_rna(void);
extern ESL_MSA      *glue_msa_create(int nseq, int L, const uint8_t *res, const double *wgt);

ESL_TREE *
glue_tree_create(int N, const int *left, const int *right, const int *parent, const double *ld, const double *rd)
{
  ESL_TREE *T = esl_tree_Create(N);
  int v;
  for (v = 0; v < N - 1; v++) { T->left[v] = left[v]; T->right[v] = right[v]; T->parent[v] = parent[v]; T->ld[v] = ld[v]; T->rd[v] = rd[v]; }
  return T;
}
void glue_tree_destroy(ESL_TREE *T) { esl_tree_Destroy(T); }

static void
flatten(ESL_MSA *m, int nrows, int L, uint8_t *out)
{
  int s;
  for (s = 0; s < nrows; s++) memcpy(out + (size_t) s * L, m->ax[s] + 1, (size_t) L);
}

int
glue_ref_fitch_shuffle(ESL_RANDOMNESS *r, ESL_TREE *T, int L, const uint8_t *msa_flat, uint8_t *shmsa_flat, uint8_t *allmsa_flat, int *sc)
{
  ESL_MSA *msa = glue_msa_create(T->N, L, msa_flat, NULL), *allmsa = NULL, *shmsa = NULL;
  char     errbuf[eslERRBUFSIZE];
  int     *usecol = malloc(sizeof(int) * (size_t) (L + 1)), status, c;

  for (c = 0; c <= L; c++) usecol[c] = TRUE;
  status = Tree_FitchAlgorithmAncenstral(r, T, msa, &allmsa, sc, FALSE, errbuf, FALSE);
  if (status == eslOK) status = msamanip_ShuffleTreeSubstitutions(r, T, msa, allmsa, usecol, &shmsa, errbuf, FALSE);
  if (status == eslOK) {
    flatten(shmsa, T->N, L, shmsa_flat);
    if (allmsa_flat) flatten(allmsa, 2 * T->N - 1, L, allmsa_flat);
  }
  esl_msa_Destroy(msa); esl_msa_Destroy(allmsa); esl_msa_Destroy(shmsa); free(usecol);
  return status;
}

/* Tree_Substitutions (src/msatree.c:1423), Fitch pass included: nsubs int [L], ndouble / njoin int [L][L]; any may be NULL */
int
glue_ref_tree_substitutions(ESL_RANDOMNESS *r, ESL_TREE *T, int L, const uint8_t *msa_flat, int includegaps, int *nsubs, int *ndouble, int *njoin)
{
  ESL_MSA *msa = glue_msa_create(T->N, L, msa_flat, NULL);
  char     errbuf[eslERRBUFSIZE];
  int     *a = NULL, *b = NULL, *c = NULL, status;

  status = Tree_Substitutions(r, msa, T, nsubs ? &a : NULL, ndouble ? &b : NULL, njoin ? &c : NULL, includegaps, errbuf, FALSE);
  if (status == eslOK) {
    if (nsubs)   memcpy(nsubs,   a, sizeof(int) * (size_t) L);
    if (ndouble) memcpy(ndouble, b, sizeof(int) * (size_t) L * (size_t) L);
    if (njoin)   memcpy(njoin,   c, sizeof(int) * (size_t) L * (size_t) L);
  }
  free(a); free(b); free(c);
  esl_msa_Destroy(msa);
  return status;
}

/* e1_rate.c itself needs Easel's file parser; e1_model_Transitions only asks it for a label (src/e1_model.c:123) */
char *e1_rate_EvomodelType(EVOM evomodel) { (void) evomodel; return "GG"; }

int
glue_ref_ptime(const double *Q16, double t, double *P16)
{
  ESL_DMATRIX *Q = esl_dmatrix_Create(4, 4), *P;
  int i;
  for (i = 0; i < 16; i++) Q->mx[0][i] = Q16[i];
  P = ratematrix_ConditionalsFromRate(t, Q, 1e-6, NULL, FALSE);
  esl_dmatrix_Destroy(Q);
  if (!P) return eslFAIL;
  for (i = 0; i < 16; i++) P16[i] = P->mx[0][i];
  esl_dmatrix_Destroy(P);
  return eslOK;
}

int
glue_ref_simulate(ESL_RANDOMNESS *r, ESL_TREE *T, const double *Q16, const uint8_t *root_flat, int L, uint8_t *leaves_flat)
{
  static double f[4] = { 0.25, 0.25, 0.25, 0.25 };
  ESL_MSA *root = glue_msa_create(1, L, root_flat, NULL), *full = NULL;
  E1_RATE  R;
  EMRATE   em;
  char     errbuf[eslERRBUFSIZE];
  int      i, k = 0, status;

  memset(&R, 0, sizeof(R)); memset(&em, 0, sizeof(em));
  em.Qstar = esl_dmatrix_Create(4, 4);
  for (i = 0; i < 16; i++) em.Qstar->mx[0][i] = Q16[i];
  em.f     = f;
  em.abc_r = glue_abc_rna();
  R.evomodel = GG;          /* no indel transitions are built (src/e1_model.c:154); RenormNoIndels overwrites them anyway */
  R.em       = &em;
  esl_strdup("root", -1, &root->name);

  status = cov_GenerateAlignment(r, GIVEN, T->N, 0.0, T, root, &R, NULL, NULL, &full, TRUE, TRUE, FALSE, "sim", 1e-6, errbuf, FALSE);
  if (status == eslOK) {
    for (i = 0; i < full->nseq; i++)                 /* leaves are the rows not named v<k> (src/R-scape-sim.c:990-993) */
      if (full->sqname[i][0] != 'v') { memcpy(leaves_flat + (size_t) k * L, full->ax[i] + 1, (size_t) L); k++; }
    if (k != T->N) status = eslFAIL;
  }
  esl_dmatrix_Destroy(em.Qstar);
  esl_msa_Destroy(root); esl_msa_Destroy(full);
  return status;
}

/* FastTree's Newick file -> the rooted tree R-scape hands to its null generator, as Tree_CalculateExtFromMSA does after the
 * FastTree run (src/msatree.c:84-95): esl_tree_ReadNewick (the shim's restatement) + esl_tree_Validate, then the REFERENCE's own
 * Tree_ReorderTaxaAccordingMSA (:823-912) and Tree_RootAtMidPoint (:524-790).  names: the alignment's sequence names in row order.
 * Outputs: int [N-1] left / right / parent, double [N-1] ld / rd in Easel's convention (children <= 0 are taxa = alignment rows). */
int
glue_ref_tree_from_newick(const char *path, int nseq, const char **names, int rootatmid, int *left, int *right, int *parent, double *ld, double *rd)
{
  ESL_MSA  *msa = esl_msa_CreateDigital(glue_abc_rna(), nseq, 1);
  ESL_TREE *T = NULL;
  FILE     *fp = fopen(path, "r");
  char      errbuf[eslERRBUFSIZE];
  int       status = eslFAIL, v;

  if (!fp || !msa) goto done;
  for (v = 0; v < nseq; v++) { free(msa->sqname[v]); msa->sqname[v] = NULL; esl_strdup(names[v], -1, &msa->sqname[v]); }
  if ((status = esl_tree_ReadNewick(fp, errbuf, &T)) != eslOK) { fprintf(stderr, "%s\n", errbuf); goto done; }
  if (T->N != nseq) { status = eslFAIL; goto done; }
  if ((status = esl_tree_Validate(T, errbuf)) != eslOK) { fprintf(stderr, "%s\n", errbuf); goto done; }
  if ((status = Tree_ReorderTaxaAccordingMSA(msa, T, errbuf, FALSE)) != eslOK) { fprintf(stderr, "%s\n", errbuf); goto done; }
  if (rootatmid && (status = Tree_RootAtMidPoint(&T, NULL, errbuf, FALSE)) != eslOK) { fprintf(stderr, "%s\n", errbuf); goto done; }
  for (v = 0; v < nseq - 1; v++) { left[v] = T->left[v]; right[v] = T->right[v]; parent[v] = T->parent[v]; ld[v] = T->ld[v]; rd[v] = T->rd[v]; }
  status = eslOK;
done:
  if (fp) fclose(fp);
  esl_tree_Destroy(T);
  esl_msa_Destroy(msa);
  return status;
}
#endif
