// Restatement of the special functions behind scipy.stats.betabinom (un-vendored third-party
// dependency of the reference: scipy -> xsf/cephes lbeta / lgam / Gamma; call sites
// BaseCellCalling.step1.py:196,201,329-330,427-428, SingleCellGenotype.py:204).
//
// scipy's betaln is cephes `lbeta`: for a+b <= 171.62 it forms Gamma(a)*Gamma(b)/Gamma(a+b)
// directly and takes one log; above that it falls back to lgam differences.  Reproducing the
// reference's p-values to ~1e-15 therefore needs the same branches and the same polynomial
// coefficients (the published cephes 2.x tables), not just "some lgamma".  Compiles for the
// device (nvcc) and for the host (gcc, used by the CPU-side validation in tests/).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LS_HD __host__ __device__ __forceinline__
#else
#define LS_HD static inline
#endif

LS_HD double ls_polevl(double x, const double *c, int n) {
  double a = c[0];
  for (int i = 1; i <= n; ++i) a = a * x + c[i];
  return a;
}
LS_HD double ls_p1evl(double x, const double *c, int n) {
  double a = x + c[0];
  for (int i = 1; i < n; ++i) a = a * x + c[i];
  return a;
}

#define LS_MAXGAM 171.624376956302725
#define LS_MAXSTIR 143.01608
#define LS_SQTPI 2.50662827463100050242E0
#define LS_LS2PI 0.91893853320467274178  /* log(sqrt(2*pi)) */
#define LS_MAXLGM 2.556348e305
#define LS_LOGPI 1.14472988584940017414

/* Stirling's formula for Gamma(x), 33 <= x <= 172 */
LS_HD double ls_stirf(double x) {
  const double STIR[5] = {7.87311395793093628397E-4, -2.29549961613378126380E-4, -2.68132617805781232825E-3,
                          3.47222221605458667310E-3, 8.33333333333482257126E-2};
  if (x >= LS_MAXGAM) return INFINITY;
  double w = 1.0 / x;
  w = 1.0 + w * ls_polevl(w, STIR, 4);
  double y = exp(x);
  if (x > LS_MAXSTIR) { /* avoid overflow in pow() */
    double v = pow(x, 0.5 * x - 0.25);
    y = v * (v / y);
  } else {
    y = pow(x, x - 0.5) / y;
  }
  y = LS_SQTPI * y * w;
  return y;
}

/* cephes Gamma(x) for x > 0 (the only domain the beta-binomial needs) */
LS_HD double ls_Gamma(double x) {
  const double P[7] = {1.60119522476751861407E-4, 1.19135147006586384913E-3, 1.04213797561761569935E-2,
                       4.76367800457137231464E-2, 2.07448227648435975150E-1, 4.94214826801497100753E-1,
                       9.99999999999999996796E-1};
  const double Q[8] = {-2.31581873324120129819E-5, 5.39605580493303397842E-4, -4.45641913851797240494E-3,
                       1.18139785222060435552E-2,  3.58236398605498653373E-2, -2.34591795718243348568E-1,
                       7.14304917030273074085E-2,  1.00000000000000000320E0};
  double p, q, z;
  if (!isfinite(x)) return x;
  q = fabs(x);
  if (q > 33.0) {
    if (x < 0.0) return NAN; /* not needed here */
    return ls_stirf(x);
  }
  z = 1.0;
  while (x >= 3.0) {
    x -= 1.0;
    z *= x;
  }
  while (x < 0.0) {
    if (x > -1.E-9) goto small;
    z /= x;
    x += 1.0;
  }
  while (x < 2.0) {
    if (x < 1.e-9) goto small;
    z /= x;
    x += 1.0;
  }
  if (x == 2.0) return z;
  x -= 2.0;
  p = ls_polevl(x, P, 6);
  q = ls_polevl(x, Q, 7);
  return z * p / q;
small:
  if (x == 0.0) return INFINITY;
  return z / ((1.0 + 0.5772156649015329 * x) * x);
}

/* cephes lgam(x) for x > 0 */
LS_HD double ls_lgam(double x) {
  const double A[5] = {8.11614167470508450300E-4, -5.95061904284301438324E-4, 7.93650340457716943945E-4,
                       -2.77777777730099687205E-3, 8.33333333333331927722E-2};
  const double B[6] = {-1.37825152569120859100E3, -3.88016315134637840924E4, -3.31612992738871184744E5,
                       -1.16237097492762307383E6, -1.72173700820839662146E6, -8.53555664245765465627E5};
  const double C[6] = {-3.51815701436523470549E2, -1.70642106651881159223E4, -2.20528590553854454839E5,
                       -1.13933444367982507207E6, -2.53252307177582951285E6, -2.01889141433532773231E6};
  double p, q, u, z;
  if (!isfinite(x)) return x;
  if (x < -34.0) return NAN; /* negative arguments never occur on this path */
  if (x < 13.0) {
    z = 1.0;
    p = 0.0;
    u = x;
    while (u >= 3.0) {
      p -= 1.0;
      u = x + p;
      z *= u;
    }
    while (u < 2.0) {
      if (u == 0.0) return INFINITY;
      z /= u;
      p += 1.0;
      u = x + p;
    }
    if (z < 0.0) z = -z;
    if (u == 2.0) return log(z);
    p -= 2.0;
    x = x + p;
    p = x * ls_polevl(x, B, 5) / ls_p1evl(x, C, 6);
    return log(z) + p;
  }
  if (x > LS_MAXLGM) return INFINITY;
  q = (x - 0.5) * log(x) - x + LS_LS2PI;
  if (x > 1.0e8) return q;
  p = 1.0 / (x * x);
  if (x >= 1000.0)
    q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
  else
    q += ls_polevl(p, A, 4) / x;
  return q;
}

/* cephes lbeta_asymp: a >> b, a > 1e6 */
LS_HD double ls_lbeta_asymp(double a, double b) {
  double r = ls_lgam(b);
  r -= b * log(a);
  r += b * (1 - b) / (2 * a);
  r += b * (1 - b) * (1 - 2 * b) / (12 * a * a);
  r += -b * b * (1 - b) * (1 - b) / (12 * a * a * a);
  return r;
}

/* cephes lbeta(a, b) for a, b > 0  == scipy.special.betaln */
LS_HD double ls_lbeta(double a, double b) {
  double y;
  if (fabs(a) < fabs(b)) {
    y = a;
    a = b;
    b = y;
  }
  if (fabs(a) > 1e6 * fabs(b) && a > 1e6) return ls_lbeta_asymp(a, b);
  y = a + b;
  if (fabs(y) > LS_MAXGAM || fabs(a) > LS_MAXGAM || fabs(b) > LS_MAXGAM) {
    y = ls_lgam(y);
    y = ls_lgam(b) - y;
    y = ls_lgam(a) + y;
    return y;
  }
  y = ls_Gamma(y);
  a = ls_Gamma(a);
  b = ls_Gamma(b);
  if (y == 0.0) return INFINITY;
  if (fabs(fabs(a) - fabs(y)) > fabs(fabs(b) - fabs(y))) {
    y = b / y;
    y *= a;
  } else {
    y = a / y;
    y *= b;
  }
  if (y < 0) y = -y;
  return log(y);
}

/* scipy betabinom._logpmf (scipy/stats/_discrete_distns.py:247-249), same association */
LS_HD double ls_betabinom_logpmf(double k, double n, double a, double b, double log_np1, double lbeta_ab) {
  double combiln = -log_np1 - ls_lbeta(n - k + 1.0, k + 1.0);
  return combiln + ls_lbeta(k + a, n - k + b) - lbeta_ab;
}
