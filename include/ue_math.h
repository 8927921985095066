/* include/ue_math.h — portable IEEE-754 elementary functions (exp, log, log10, pow, cos).
 *
 * Why: the finite-difference Jacobian keeps an entry iff the perturbed residual differs from the
 * unperturbed one by at least one bit (jaccliplim = 0, bbb/oderhs.m:8706), and entries are
 * differences of O(1e-8) increments.  glibc's and CUDA's libm differ by ~1 ulp, which is enough to
 * move 7 % of the Jacobian entries by more than 1e-8 relative.  These routines use only + - * /,
 * comparisons and exponent-field manipulation, in a fixed order, so (with FMA contraction disabled
 * on both sides) the CPU checker and the CUDA kernels produce BIT-IDENTICAL results.
 *
 * Algorithms: the classical argument reductions and minimax polynomials published with FreeBSD
 * msun / fdlibm (exp: r = x - k ln2, Remez rational on |r| <= ln2/2; log: x = 2^k (1+f),
 * s = f/(2+f), even/odd polynomial in s^2; sin/cos kernels on |x| <= pi/4).  Accuracy ~1 ulp
 * (pow(x,y) = exp(y log x): ~|y ln x| ulp), far below the 1e-12 residual tolerance.
 *
 * Domain: finite arguments of the sizes met in the hot path.  log of a non-positive number
 * returns NaN / -inf like libm; cos reduces by pi/2 in plain double (absolute error 2 ulp(1); its only
 * uses have the argument 0).
 */
#ifndef UE_MATH_H
#define UE_MATH_H
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define UE_HD __host__ __device__ __forceinline__
/* The long functions (log, exp, pow, cos) are NOT inlined on the device: the kernels are single cold passes whose time is
 * dominated by instruction fetch, and a physics kernel calls them at dozens of sites. */
#ifdef UE_MATH_FORCEINLINE
#define UE_HD_BIG UE_HD
#else
#define UE_HD_BIG static __host__ __device__ __noinline__
#endif
#else
#define UE_HD static inline
#define UE_HD_BIG static inline
#endif

UE_HD int64_t ue_d2bits(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t b; memcpy(&b, &x, 8); return b;
#endif
}
UE_HD double ue_bits2d(int64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x; memcpy(&x, &b, 8); return x;
#endif
}
UE_HD double ue_two_pow(int k) { /* 2^k, -1022 <= k <= 1023 */
  return ue_bits2d((int64_t)(k + 1023) << 52);
}

UE_HD_BIG double ue_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  if (!(x > 0.0)) {
    if (x == 0.0) return -1.0 / 0.0;
    return 0.0 / 0.0; /* negative or NaN */
  }
  int k = 0;
  if (x < 2.2250738585072014e-308) { x = x * 18014398509481984.0; k = -54; } /* subnormal: scale by 2^54 */
  int64_t b = ue_d2bits(x);
  if (b >= 0x7ff0000000000000LL) return x; /* +inf */
  k += (int)(b >> 52) - 1023;
  b = (b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL; /* mantissa in [1,2) */
  double m = ue_bits2d(b);
  if (m > 1.4142135623730951) { m = m * 0.5; k += 1; } /* (sqrt(2)/2, sqrt(2)] */
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

UE_HD double ue_log10(double x) { return ue_log(x) * 4.34294481903251816668e-01; }

UE_HD_BIG double ue_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return 1.0 / 0.0;
  if (x < -7.45133219101941108420e+02) return 0.0;
  const int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
  const double dk = (double)k;
  const double hi = x - dk * ln2HI;
  const double lo = dk * ln2LO;
  const double r = hi - lo;
  const double t = r * r;
  const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  if (k >= -1021 && k <= 1023) return y * ue_two_pow(k);
  if (k > 1023) return y * ue_two_pow(1023) * ue_two_pow(k - 1023);
  return y * ue_two_pow(k + 1000) * ue_two_pow(-1000);
}

UE_HD double ue_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}

/* x**y for x >= 0 (Fortran real power).  Exponents 0, 1, 2, 0.5 are exact shortcuts. */
UE_HD_BIG double ue_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (y == 1.0) return x;
  if (y == 2.0) return x * x;
  if (y == 0.5) return ue_sqrt(x);
  if (x == 0.0) return (y > 0.0) ? 0.0 : 1.0 / 0.0;
  return ue_exp(y * ue_log(x));
}

UE_HD double ue_ksin(double x) { /* |x| <= pi/4 */
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x, v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x + v * (S1 + z * r);
}
UE_HD double ue_kcos(double x) { /* |x| <= pi/4 */
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return 1.0 - (0.5 * z - z * r);
}
UE_HD_BIG double ue_cos(double x) { /* intended for |x| <= pi; larger |x| is folded by 2*pi steps */
  const double pi = 3.14159265358979311600e+00, pio2 = 1.57079632679489655800e+00, pio4 = 7.85398163397448278999e-01;
  if (x < 0.0) x = -x;
  if (x != x || x > 1.0e15) return 0.0 / 0.0;
  if (x > pi) { const double n = (double)(int64_t)(x / (2.0 * pi) + 0.5); x = x - n * (2.0 * pi); if (x < 0.0) x = -x; }
  if (x <= pio4) return ue_kcos(x);
  if (x <= 3.0 * pio4) return -ue_ksin(x - pio2);
  return -ue_kcos(pi - x);
}
#endif /* UE_MATH_H */
