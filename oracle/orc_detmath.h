/* orc_detmath.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * Deterministic double-precision log / sin / cos built only from IEEE-754
 * correctly-rounded +,-,*,/ (no FMA contraction: compile with
 * -ffp-contract=off).  The CUDA product has its own copy of the same
 * operation sequence (abeille_b200/csrc/detmath.cuh, compiled -fmad=false),
 * so that the oracle in "det" math mode and the kernels agree BIT FOR BIT.
 *
 * Why: the reference calls glibc's log/sin/cos (include/utils/rng.hpp:74-79,
 * include/utils/direction.hpp:130-151, src/isotropic.cpp:28-36).  CUDA's libm
 * differs from glibc by <=1 ulp on a fraction of arguments, which would make
 * every position differ in the last bit.  Both sides therefore evaluate the
 * same published algorithm (Sun fdlibm e_log.c / k_sin.c / k_cos.c /
 * e_rem_pio2.c medium-size path, restated here), whose error is < 1 ulp --
 * the same accuracy class as glibc.  The oracle's "libm" math mode keeps
 * std::log/sin/cos and is used to show that integer outcomes do not depend on
 * which <1ulp libm is used (tests/test_oracle_math.py).
 */
#ifndef ORC_DETMATH_H
#define ORC_DETMATH_H
#include <stdint.h>
#include <math.h>
#include <string.h>

static inline uint64_t orc_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double orc_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline int32_t orc_hi(double x) { return (int32_t)(orc_d2u(x) >> 32); }
static inline uint32_t orc_lo(double x) { return (uint32_t)(orc_d2u(x) & 0xffffffffu); }
static inline double orc_set_hi(double x, int32_t hi) {
  return orc_u2d(((uint64_t)(uint32_t)hi << 32) | (orc_d2u(x) & 0xffffffffu));
}

static inline double orc_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
               Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
               Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
  double hfsq, f, s, z, R, w, t1, t2, dk;
  int32_t k, hx, i, j;
  uint32_t lx;
  hx = orc_hi(x);
  lx = orc_lo(x);
  k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / 0.0;
    if (hx < 0) return (x - x) / 0.0;
    k -= 54;
    x *= two54;
    hx = orc_hi(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  i = (hx + 0x95f64) & 0x100000;
  x = orc_set_hi(x, hx | (i ^ 0x3ff00000));
  k += (i >> 20);
  f = x - 1.0;
  if ((0x000fffff & (2 + hx)) < 3) {
    if (f == 0.0) {
      if (k == 0) return 0.0;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  s = f / (2.0 + f);
  dk = (double)k;
  z = s * s;
  i = hx - 0x6147a;
  w = z * z;
  j = 0x6b851 - hx;
  t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  R = t2 + t1;
  if (i > 0) {
    hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// fdlibm e_exp.c (same operation sequence as det_exp in abeille_b200/csrc/detmath.cuh)
static inline double orc_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
               P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08,
               o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
               twom1000 = 9.33263618503218878990e-302;
  int32_t hx = orc_hi(x);
  const int xsb = (hx >> 31) & 1;
  hx &= 0x7fffffff;
  double hi = 0., lo = 0.;
  int32_t k = 0;
  if (hx >= 0x40862E42) {  // |x| >= 709.78...
    if (hx >= 0x7ff00000) {
      if (((hx & 0xfffff) | orc_lo(x)) != 0) return x + x;  // NaN
      return xsb == 0 ? x : 0.0;                        // exp(+-inf)
    }
    if (x > o_threshold) return 1.0e+300 * 1.0e+300;
    if (x < u_threshold) return 0.0;
  }
  if (hx > 0x3fd62e42) {     // |x| > 0.5 ln2
    if (hx < 0x3FF0A2B2) {   // |x| < 1.5 ln2
      hi = xsb ? x + ln2HI : x - ln2HI;
      lo = xsb ? -ln2LO : ln2LO;
      k = 1 - xsb - xsb;
    } else {
      k = (int32_t)(invln2 * x + (xsb ? -0.5 : 0.5));
      const double t = (double)k;
      hi = x - t * ln2HI;
      lo = t * ln2LO;
    }
    x = hi - lo;
  } else if (hx < 0x3e300000) {  // |x| < 2^-28
    return 1.0 + x;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k >= -1021) return orc_set_hi(y, orc_hi(y) + k * 1048576);
  y = orc_set_hi(y, orc_hi(y) + (k + 1000) * 1048576);
  return y * twom1000;
}

static inline double orc_ksin(double x, double y, int iy) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  int32_t ix = orc_hi(x) & 0x7fffffff;
  if (ix < 0x3e400000) {
    if ((int)x == 0) return x;
  }
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (iy == 0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static inline double orc_kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  int32_t ix = orc_hi(x) & 0x7fffffff;
  if (ix < 0x3e400000) {
    if (((int)x) == 0) return 1.0;
  }
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  if (ix < 0x3FD33333) return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ix > 0x3fe90000)
    qx = 0.28125;
  else
    qx = orc_u2d((uint64_t)(uint32_t)(ix - 0x00200000) << 32);
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

/* reduce |x| (< 2^19*pi/2) to y0+y1 in [-pi/4,pi/4]; returns quadrant n */
static inline int orc_rem_pio2(double x, double* y0, double* y1) {
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
               pio2_1t = 6.07710050650619224932e-11, pio2_2 = 6.07710050630396597660e-11,
               pio2_2t = 2.02226624879595063154e-21, pio2_3 = 2.02226624871116645580e-21,
               pio2_3t = 8.47842766036889956997e-32;
  int32_t hx = orc_hi(x);
  int32_t ix = hx & 0x7fffffff;
  double t = orc_u2d(orc_d2u(x) & 0x7fffffffffffffffULL);
  int n = (int)(t * invpio2 + 0.5);
  double fn = (double)n;
  double r = t - fn * pio2_1;
  double w = fn * pio2_1t;
  int32_t j = ix >> 20;
  double a0 = r - w;
  int32_t i = j - ((orc_hi(a0) >> 20) & 0x7ff);
  if (i > 16) {
    t = r;
    w = fn * pio2_2;
    r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    a0 = r - w;
    i = j - ((orc_hi(a0) >> 20) & 0x7ff);
    if (i > 49) {
      t = r;
      w = fn * pio2_3;
      r = t - w;
      w = fn * pio2_3t - ((t - r) - w);
      a0 = r - w;
    }
  }
  double a1 = (r - a0) - w;
  if (hx < 0) {
    *y0 = -a0;
    *y1 = -a1;
    return -n;
  }
  *y0 = a0;
  *y1 = a1;
  return n;
}

static inline void orc_sincos(double x, double* sn, double* cs) {
  int32_t ix = orc_hi(x) & 0x7fffffff;
  if (ix <= 0x3fe921fb) {
    *sn = orc_ksin(x, 0.0, 0);
    *cs = orc_kcos(x, 0.0);
    return;
  }
  if (ix >= 0x7ff00000) {
    *sn = *cs = x - x;
    return;
  }
  double y0, y1;
  int n = orc_rem_pio2(x, &y0, &y1);
  double s = orc_ksin(y0, y1, 1), c = orc_kcos(y0, y1);
  switch (n & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}
// acos, for the higher harmonics of the flat-vibration noise source (same operation sequence as det_acos in detmath.cuh)
static inline double orc_acos(double x) {
  // fdlibm e_acos.c: a rational approximation of (asin(x) - x) / x^3 on [0, 0.5], the half-angle identity outside
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00,
               pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
               pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
               qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
               qS4 = 7.70381505559019352791e-02;
  const int32_t hx = orc_hi(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix >= 0x3ff00000) {  // |x| >= 1
    if (((uint32_t)(ix - 0x3ff00000) | orc_lo(x)) == 0) return hx > 0 ? 0.0 : pi + 2.0 * pio2_lo;
    return (x - x) / (x - x);
  }
  if (ix < 0x3fe00000) {  // |x| < 0.5
    if (ix <= 0x3c600000) return pio2_hi + pio2_lo;
    const double z = x * x;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double r = p / q;
    return pio2_hi - (x - (pio2_lo - x * r));
  } else if (hx < 0) {  // x < -0.5
    const double z = (1.0 + x) * 0.5;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double s = sqrt(z);
    const double r = p / q;
    const double w = r * s - pio2_lo;
    return pi - 2.0 * (s + w);
  } else {  // x > 0.5
    const double z = (1.0 - x) * 0.5;
    const double s = sqrt(z);
    const double df = orc_u2d(orc_d2u(s) & 0xffffffff00000000ULL);
    const double c = (z - df * df) / (s + df);
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double r = p / q;
    const double w = r * s + c;
    return 2.0 * (df + w);
  }
}

static inline double orc_sin(double x) { double s, c; orc_sincos(x, &s, &c); return s; }
static inline double orc_cos(double x) { double s, c; orc_sincos(x, &s, &c); return c; }
#endif
