// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE),
// restated so that cdf = np.sum(pmf(arange(0, k))) (scipy rv_discrete._cdf_single,
// _distn_infrastructure.py:3463-3466) is accumulated in the same order as the reference.
#pragma once
#include "ls_cephes.h"

#ifdef __CUDACC__
#define LS_HDN __host__ __device__
#else
#define LS_HDN static
#endif

LS_HDN double ls_pairwise_sum(const double *a, long n) {
  if (n < 8) {
    double res = 0.;  /* numpy starts from -0.0 for floats; identical for non-negative terms */
    for (long i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    long i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return ls_pairwise_sum(a, n2) + ls_pairwise_sum(a + n2, n - n2);
  }
}

/* scipy betabinom.sf(k - eps, n, a, b), 0 < eps < 1, integer k, with the argument handling of
 * rv_discrete.sf (_distn_infrastructure.py: x < 0 -> 1, x >= n -> 0, bad n -> NaN). */
LS_HDN double ls_sf_finish(int k, int n, double cdf) {
  if (n < 0) return NAN;
  if (k <= 0) return 1.0;   /* k - eps < 0 */
  if (k > n) return 0.0;    /* k - eps >= n  <=>  k >= n + 1 */
  double s = 1.0 - cdf;
  return s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
}

#ifndef __CUDACC__
static double ls_sf_from_terms_host(int k, int n, double a, double b, double lab, double *scratch) {
  if (n < 0 || k <= 0 || k > n) return ls_sf_finish(k, n, 0.0);
  const double lnp1 = log((double)n + 1.0);
  for (int i = 0; i < k; ++i) scratch[i] = exp(ls_betabinom_logpmf((double)i, (double)n, a, b, lnp1, lab));
  return ls_sf_finish(k, n, ls_pairwise_sum(scratch, k));
}
#endif
