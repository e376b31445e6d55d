/* Host build of longsom_b200/csrc/ls_cephes.h + the pairwise summation used by the K2 kernel,
 * so that the CPU test-suite can compare the restated special functions with scipy
 * (test infrastructure; the product runs the same header on the GPU). */
#include <stdint.h>
#include "../../longsom_b200/csrc/ls_cephes.h"
#include "../../longsom_b200/csrc/ls_pairwise.h"

double h_lbeta(double a, double b) { return ls_lbeta(a, b); }
double h_lgam(double x) { return ls_lgam(x); }
double h_Gamma(double x) { return ls_Gamma(x); }

void h_betabinom_sf(const int32_t *k, const int32_t *n, double a, double b, double *p, int64_t m, double *scratch) {
  const double lab = ls_lbeta(a, b);
  for (int64_t i = 0; i < m; ++i) {
    p[i] = ls_sf_from_terms_host(k[i], n[i], a, b, lab, scratch);
  }
}
