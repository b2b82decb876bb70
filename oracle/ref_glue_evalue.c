/* ref_glue_evalue.c -- reach the REFERENCE's E-value arithmetic from flat arrays.
 * TEST INFRASTRUCTURE, compiled only into oracle/_ref/librscape_ref.so (needs the reference tree).
 *
 * cov2evalue and evalue2cov are `static` in src/covariation.c (:2370, :2404), so this translation unit includes that
 * source file where it lies (unchanged, nothing copied) and wraps the two functions.  The rest of covariation.c is
 * compiled along with them; whatever it leaves undefined (plots, power, CaCoFold, Easel's fits) becomes an
 * abort-stub (ref_stubs_gen.sh) and is unreachable from the entry points below.
 *
 *   glue_ref_cov2evalue  : cov2evalue(cov, Nc, h, survfit)          src/covariation.c:2370-2400
 *   glue_ref_evalue2cov  : evalue2cov(eval_thresh, Nc, h, survfit)  src/covariation.c:2404-2435
 */
#ifdef GLUE_REFERENCE
#include "covariation.c"

static void
hist_view(ESL_HISTOGRAM *h, const double *geom, const int *ig, uint64_t Nc, uint64_t No, uint64_t *obs)
{
  memset(h, 0, sizeof(*h));
  h->bmin = geom[0]; h->w = geom[1]; h->xmax = geom[2]; h->phi = geom[3];
  h->nb = ig[0]; h->imin = ig[1]; h->imax = ig[2]; h->cmin = ig[3];
  h->bmax = h->bmin + h->w * h->nb;
  h->Nc = Nc; h->No = No; h->obs = obs;
}

/* geom = { bmin, w, xmax, phi }, ig = { nb, imin, imax, cmin } */
double
glue_ref_cov2evalue(double cov, int Nc, const double *geom, const int *ig, uint64_t hNc, uint64_t hNo, uint64_t *obs, double *survfit)
{
  ESL_HISTOGRAM h;
  hist_view(&h, geom, ig, hNc, hNo, obs);
  return cov2evalue(cov, Nc, &h, survfit);
}

double
glue_ref_evalue2cov(double eval_thresh, int Nc, const double *geom, const int *ig, uint64_t hNc, uint64_t hNo, uint64_t *obs, double *survfit)
{
  ESL_HISTOGRAM h;
  hist_view(&h, geom, ig, hNc, hNo, obs);
  return evalue2cov(eval_thresh, Nc, &h, survfit);
}
#endif
