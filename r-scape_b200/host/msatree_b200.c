/* msatree_b200.c -- host-side mirror of Tree_Substitutions (src/msatree.c:1423-1554), the substitution counts behind
 * R-scape's power calculation (callers src/R-scape.c:2784, :2809, :2840).
 *
 * The reference function first reconstructs the ancestral sequences (Tree_FitchAlgorithmAncenstral, :1451) and then
 * loops over columns / column pairs and branches (:1455-1540).  The loops -- O(L^2 N) for the pair tables -- run on the
 * device (rsb_tree_substitutions); the Fitch pass, which draws from the caller's RNG stream, stays where it is: inside an
 * R-scape tree the body of Tree_Substitutions after :1452 becomes one call of Tree_Substitutions_b200.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rscape_b200_host.h"

/* msa: the alignment (alen is taken from it, as the reference does when T == NULL); allmsa: the 2N-1 rows written by
 * Tree_FitchAlgorithmAncenstral (leaves first, internal node v at row N+v); outputs as the reference allocates them:
 * nsubs int[alen], ndouble / njoin int[alen*alen] with the entries i<j filled.  Any ret_ pointer may be NULL. */
int
Tree_Substitutions_b200(ESL_MSA *msa, ESL_MSA *allmsa, ESL_TREE *T, int **ret_nsubs, int **ret_ndouble, int **ret_njoin,
                        int includegaps, char *errbuf, int verbose)
{
  rsb_ctx    *ctx = NULL;
  uint8_t    *leaves = NULL, *internal = NULL;
  int        *nsubs = NULL, *ndouble = NULL, *njoin = NULL;
  const char *env;
  size_t      L = (size_t) msa->alen;
  int         N, s, device = 0, status = eslFAIL;
  (void) verbose;

  if (T == NULL) {                                           /* :1437-1449: zero counts (njoin is not touched there) */
    if (ret_nsubs   && (nsubs   = calloc(L ? L : 1, sizeof(int))) == NULL)     goto ERROR;
    if (ret_ndouble && (ndouble = calloc(L ? L * L : 1, sizeof(int))) == NULL) goto ERROR;
    if (ret_nsubs)   *ret_nsubs   = nsubs;
    if (ret_ndouble) *ret_ndouble = ndouble;
    return eslOK;
  }
  N = T->N;
  if (!allmsa || allmsa->nseq < 2 * N - 1 || allmsa->alen != msa->alen) {
    if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "Tree_Substitutions: allmsa must hold the %d rows of the Fitch reconstruction", 2 * N - 1);
    return eslFAIL;
  }
  if ((env = getenv("RSCAPE_B200_DEVICE")) != NULL) device = atoi(env);
  leaves   = malloc((size_t) N * L + 1);
  internal = malloc((size_t) (N - 1) * L + 1);
  if (ret_nsubs)   nsubs   = malloc(sizeof(int) * (L ? L : 1));
  if (ret_ndouble) ndouble = malloc(sizeof(int) * (L ? L * L : 1));
  if (ret_njoin)   njoin   = malloc(sizeof(int) * (L ? L * L : 1));
  if (!leaves || !internal || (ret_nsubs && !nsubs) || (ret_ndouble && !ndouble) || (ret_njoin && !njoin)) {
    if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "allocation failed");
    goto ERROR;
  }
  for (s = 0; s < N; s++)     memcpy(leaves   + (size_t) s * L, allmsa->ax[s] + 1, L);
  for (s = 0; s < N - 1; s++) memcpy(internal + (size_t) s * L, allmsa->ax[N + s] + 1, L);

  if (rsb_create(device, NULL, &ctx) != 0) { if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "%s", rsb_create_error()); goto ERROR; }
  if (rsb_configure(ctx, 2 * (N - 1), (int) L, 1, 1) != 0 ||                               /* one row per branch, unit weights */
      rsb_tree_substitutions(ctx, N, T->left, T->right, leaves, (int64_t) L, internal, (int64_t) L, includegaps, nsubs, ndouble, njoin) != 0) {
    if (errbuf) snprintf(errbuf, eslERRBUFSIZE, "%s", rsb_error(ctx));
    goto ERROR;
  }
  if (ret_nsubs)   { *ret_nsubs   = nsubs;   nsubs   = NULL; }
  if (ret_ndouble) { *ret_ndouble = ndouble; ndouble = NULL; }
  if (ret_njoin)   { *ret_njoin   = njoin;   njoin   = NULL; }
  status = eslOK;

 ERROR:
  free(leaves); free(internal); free(nsubs); free(ndouble); free(njoin);
  if (ctx) rsb_destroy(ctx);
  return status;
}
