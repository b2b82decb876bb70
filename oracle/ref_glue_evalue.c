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

/* the reference's own "Histogram and Fit" block (src/covariation.c:459-487): static cov_histogram_pmass (:2484) + cov_NullFitGamma /
 * cov_NullFitExponential (:1915-1973) over the shim's restatement of Easel's fits.  Same flat interface as glue_nullfit (mi_glue.c). */
int
glue_ref_nullfit(const double *geom, const int *ig, uint64_t n, uint64_t *obs, double pmass_target, double fracfit, int doexpfit, double *survfit, double *out)
{
  ESL_HISTOGRAM h;
  double       *sf = NULL, pmass;
  char          errbuf[eslERRBUFSIZE];
  int           status, b;
  memset(&h, 0, sizeof(h));
  h.bmin = geom[0]; h.w = geom[1]; h.xmax = geom[2]; h.nb = ig[0]; h.imin = ig[1]; h.imax = ig[2]; h.cmin = h.imin;
  h.bmax = h.bmin + h.w * h.nb; h.n = h.Nc = h.No = n; h.obs = obs;
  pmass = cov_histogram_pmass(&h, pmass_target, fracfit);
  if (isnan(pmass)) return eslFAIL;
  out[3] = 0.0;
  if (doexpfit) status = cov_NullFitExponential(&h, &sf, pmass, &out[0], &out[1], &out[2], FALSE, errbuf);
  else          status = cov_NullFitGamma(&h, &sf, pmass, &out[0], &out[1], &out[2], &out[3], FALSE, errbuf);
  if (status != eslOK) return status;
  out[4] = h.phi; out[5] = (double) h.cmin;
  for (b = 0; b < 2 * h.nb; b++) survfit[b] = sf ? sf[b] : 0.0;
  free(sf);
  return eslOK;
}
#endif
