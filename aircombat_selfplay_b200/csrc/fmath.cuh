// fmath.cuh -- fp64 division, square root and the few elementary functions of the per-frame FDM path, written for the
// way the substep kernels run: ONE in-order dependent chain per thread at two warps per scheduler, so every instruction
// on the chain costs its full latency.
//
// Why not `/`, sqrt() and libdevice (ncu + SASS of k_env_substeps, profiles/r2_k_env_substeps_65536envs.txt):
//  * every IEEE fp64 division / reciprocal / sqrt is MUFU seed + Newton steps + a RANGE GUARD (FSETP, FFMA, FSETP, BRA,
//    BSSY / BSYNC around a CALL to a slow path for zero / subnormal / huge operands): 17 / 12 / 16 instructions where the
//    arithmetic is 7 / 6 / 10, ~40 such sites per frame, and each BSSY / BSYNC pair fences the instruction scheduler;
//  * libdevice atan2 / sincos / exp / log materialise their polynomial coefficients with UMOV pairs (two per double, 8.6 %
//    of all executed instructions) and carry quadrant / special-case code the FDM's operand ranges never reach.
// Here the Newton sequences run without the guard (the operands are physical quantities of moderate magnitude; call
// sites whose operand can be exactly zero use the *0 variants or keep their own test), and the polynomial coefficients
// sit in __constant__ arrays indexed by literals: ptxas fetches them with uniform constant loads (LDCU.128, two
// coefficients per instruction, off the dependent chain) instead of two UMOV per coefficient, and the polynomials are
// evaluated as two interleaved half-length Horner chains (even / odd powers).
//
// Accuracy: division and sqrt within 1 ulp (correctly rounded except in rare half-way cases); atan / sin / cos / log /
// exp within a few ulp on the stated intervals -- against the 1e-9 relative per-step tolerance of the parity tests.
// Outside its interval every function falls back to libdevice (a branch the flight envelope rarely takes).
// tests/test_fmath_host.py compiles this header for the HOST (seeds emulated with a 2^-12 relative error, i.e. far
// worse than the MUFU seeds) and checks every function against libm.
//
// -DACS_IEEE_MATH restores `/`, sqrt() and libdevice everywhere (tuning builds, tools/exp_variants.py).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define FM_DEV __device__ __forceinline__
#define FM_CONST __constant__
#else
#define FM_DEV static inline
#define FM_CONST static const
#endif

// atan(t) = t * Q(t^2) on |t| <= 0.41422 (tan(pi/8)): tools/make_atan_poly.py (Chebyshev interpolation in exact
// rationals, max |error| 5.6e-17)
FM_CONST double FM_ATAN_Q[13] = {
  1.00000000000000000e+00, -3.33333333333333259e-01, 1.99999999999970118e-01, -1.42857142853295982e-01,
  1.11111110853156225e-01, -9.09090805880346103e-02, 7.69228109488473560e-02, -6.66620518638054438e-02,
  5.87683329138398852e-02, -5.21730801921760079e-02, 4.49948232104990381e-02, -3.33873979550728395e-02,
  1.51133334306833046e-02};
// sin(x) = x + x^3 S(x^2), cos(x) = 1 - x^2/2 + x^4 C(x^2) on |x| <= pi/4: the classic minimax kernels (Sun fdlibm
// k_sin.c / k_cos.c constants, public domain), error < 2^-57
FM_CONST double FM_SIN_S[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
FM_CONST double FM_COS_C[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                               -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
// log(x) = 2 s L(s^2), s = (x - 1) / (x + 1): L(v) = sum v^k / (2k + 1), k = 0..11 (|s| <= 0.19: truncation 2e-19)
FM_CONST double FM_LOG_L[12] = {1.0, 1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0,
                                1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0};
// exp(r) = sum r^k / k!, k = 0..13 on |r| <= ln2 / 2 (truncation 4e-18)
FM_CONST double FM_EXP_E[14] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                                1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0,
                                1.0 / 6227020800.0};

// sin(r) = r + r^3 T(r^2) on |r| <= pi/2: T(z) = sum (-1)^(k+1) z^k / (2k + 3)!, k = 0..9
FM_CONST double FM_SIN_T[10] = {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0,
                                -1.0 / 1307674368000.0, 1.0 / 355687428096000.0, -1.0 / 121645100408832000.0,
                                1.0 / 51090942171709440000.0};

// argument-reduction constants (in constant memory like the coefficients: a literal costs two UMOV per use):
// 2/pi, -(pi/2) in three parts | 1/pi, -pi in three parts | log2(e), -ln2 in two parts
FM_CONST double FM_RED[11] = {0.6366197723675814, -1.5707963267948966, -6.123233995736766e-17, 1.4973849048591698e-33,
                              0.3183098861837907, -3.141592653589793, -1.2246467991473532e-16, 2.9947698097183397e-33,
                              1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10};

#ifdef ACS_IEEE_MATH
FM_DEV double fm_div(double a, double b) { return a / b; }
FM_DEV double fm_rcp(double b) { return 1.0 / b; }
FM_DEV double fm_sqrt(double x) { return sqrt(x); }
FM_DEV double fm_sqrt0(double x) { return sqrt(x); }
FM_DEV double fm_rsqrt(double x) { return 1.0 / sqrt(x); }
FM_DEV void fm_sincos_small(double x, double* s, double* c) { sincos(x, s, c); }
FM_DEV double fm_angle_sc(double s, double c, double y, double x) { (void)s; (void)c; return atan2(y, x); }
FM_DEV double fm_atan2(double y, double x) { return atan2(y, x); }
FM_DEV double fm_powpos(double x, double y) { return exp(y * log(x)); }
FM_DEV double fm_pow_ratio(double num, double den, double y) { return exp(y * log(num / den)); }
FM_DEV double fm_exp(double z) { return exp(z); }
FM_DEV void fm_sincos(double x, double* s, double* c) { sincos(x, s, c); }
FM_DEV double fm_sin(double x) { return sin(x); }
FM_DEV double fm_cos(double x) { return cos(x); }
FM_DEV double fm_acos(double x) { return acos(x); }
FM_DEV double fm_log(double x) { return log(x); }
FM_DEV double fm_tanh(double x) { return tanh(x); }
FM_DEV double fm_atanh(double x) { return atanh(x); }
#else

#ifdef __CUDACC__
FM_DEV double fm_rcp_seed(double b) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r; }     // MUFU.RCP64H
FM_DEV double fm_rsq_seed(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }   // MUFU.RSQ64H
FM_DEV double fm_pow2i(int n) { return __hiloint2double((n + 1023) << 20, 0); }
FM_DEV double fm_rint(double x) { return rint(x); }
// x = m 2^e with m in [1, 2) for normal positive x
FM_DEV double fm_frexp1(double x, int* e) {
  const int hi = __double2hiint(x);
  *e = (hi >> 20) - 1023;
  return __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}
#else
// host build (tests): seeds with a deliberately poor relative error of 2^-12 -- a 24-bit significand like a float's, but
// over the whole double exponent range, as the hardware seeds have (the FDM takes 1 / sqrt of numbers near 1e102)
FM_DEV double fm_host_seed24(double v) {
  unsigned long long u;
  __builtin_memcpy(&u, &v, 8);
  u &= ~((1ull << 29) - 1ull);
  __builtin_memcpy(&v, &u, 8);
  return v;
}
FM_DEV double fm_rcp_seed(double b) { return fm_host_seed24(1.0 / b) * (1.0 + 1.0 / 4096.0); }
FM_DEV double fm_rsq_seed(double x) { return fm_host_seed24(1.0 / sqrt(x)) * (1.0 - 1.0 / 4096.0); }
FM_DEV double fm_pow2i(int n) { return ldexp(1.0, n); }
FM_DEV double fm_rint(double x) { return rint(x); }
FM_DEV double fm_frexp1(double x, int* e) { const double m = frexp(x, e); *e -= 1; return m + m; }
#endif

// 1 / b after one cubic Newton step: relative error (seed error)^3 + 2^-53
FM_DEV double fm_rcp1(double b) {
  const double r = fm_rcp_seed(b), e = fma(-b, r, 1.0), t = fma(e, e, e);
  return fma(r, t, r);
}
// a / b: quotient with the cubic reciprocal, then one exact-residual correction (error of the quotient squared)
FM_DEV double fm_div(double a, double b) {
  const double r = fm_rcp1(b), q = a * r;
  return fma(fma(-b, q, a), r, q);
}
FM_DEV double fm_rcp(double b) {
  const double r = fm_rcp1(b);
  return fma(r, fma(-b, r, 1.0), r);
}
// 1 / sqrt(x) after one cubic step: y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2
FM_DEV double fm_rsqrt1(double x) {
  const double y = fm_rsq_seed(x), e = fma(-x, y * y, 1.0), p = fma(e, 0.375, 0.5);
  return fma(p, y * e, y);
}
// sqrt(x), x > 0 and normal: g = x y, one Heron correction with y / 2 ~ 1 / (2 sqrt x)
FM_DEV double fm_sqrt(double x) {
  const double y = fm_rsqrt1(x), g = x * y;
  return fma(fma(-g, g, x), 0.5 * y, g);
}
// the same for operands that can be zero or subnormal (the seed flushes subnormals: +inf, then NaN): 0 below the
// smallest normal number -- also for negative operands, where sqrt() gives NaN; NaN stays NaN
FM_DEV double fm_sqrt0(double x) {
  const double g = fm_sqrt(x);
  return x < 2.2250738585072014e-308 ? 0.0 : g;
}
FM_DEV double fm_rsqrt(double x) {
  const double y = fm_rsqrt1(x), g = x * y;
  return fma(0.5 * y, fma(-g, y, 1.0), y);
}

// even / odd split of a polynomial in u: two half-length dependent chains instead of one
#define FM_POLY_EO(C, N, u, w, out)                                                          \
  {                                                                                          \
    double ev_ = C[((N) - 1) & ~1], od_ = C[(((N) - 2) & ~1) + 1];                             \
    _Pragma("unroll") for (int k_ = (((N) - 1) & ~1) - 2; k_ >= 0; k_ -= 2) ev_ = fma(ev_, w, C[k_]); \
    _Pragma("unroll") for (int k_ = ((((N) - 2) & ~1) + 1) - 2; k_ >= 1; k_ -= 2) od_ = fma(od_, w, C[k_]); \
    out = fma(od_, u, ev_);                                                                  \
  }

// atan2(y, x) for finite operands: the smaller magnitude over the larger, folded once more at tan(pi/8) through
// atan(a) = pi/4 + atan((a - 1) / (a + 1)) -- operands selected BEFORE the one division -- then the octant fix-ups.
// A few ulp; atan2(0, 0) = 0 and the sign of y as libm (the sign of a zero x is not distinguished).
FM_DEV double fm_atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  const bool big = mn > 0.41421356237309503 * mx;
  const double num = big ? mn - mx : mn, den = big ? mn + mx : mx;
  const double t = fm_div(num, den), u = t * t, w = u * u;
  double q;
  FM_POLY_EO(FM_ATAN_Q, 13, u, w, q)
  double r = fma(t, q, big ? 0.78539816339744831 : 0.0);
  if (ay > ax) r = 1.5707963267948966 - r;
  if (x < 0.0) r = 3.141592653589793 - r;
  r = mx == 0.0 ? 0.0 : r;
  return copysign(r, y);
}
// The angle whose sine s and cosine c are known (s^2 + c^2 = 1 to rounding): 2 atan(s / (1 + c)).  |angle| <= 45 deg
// takes the short form (angle of attack and sideslip of an aircraft that flies forwards); anything else fm_atan2(y, x).
FM_DEV double fm_angle_sc(double s, double c, double y, double x) {
  if (c >= 0.7072) {
    const double t = fm_div(s, 1.0 + c), u = t * t, w = u * u;
    double q;
    FM_POLY_EO(FM_ATAN_Q, 13, u, w, q)
    return (t + t) * q;
  }
  return fm_atan2(y, x);
}
// exp(z + zl), |z| <= 700, zl a rounding-error term: z = n ln2 + r, |r| <= ln2 / 2, Taylor through r^13, scale by 2^n
FM_DEV double fm_exp_core(double z, double zl) {
  const double n = fm_rint(z * FM_RED[8]);
  double r = fma(n, FM_RED[9], z);
  r = fma(n, FM_RED[10], r) + zl;
  const double r2 = r * r;
  double p;
  FM_POLY_EO(FM_EXP_E, 14, r, r2, p)
  return p * fm_pow2i((int)n);
}
// (num / den)^y for a ratio near one (the temperature ratio inside one ISA layer: 0.75 .. 1.33) and |y ln| < 700:
// ln through atanh, s = (num - den) / (num + den) -- ONE division for the ratio and the logarithm -- then exp by
// argument reduction to |r| <= ln2 / 2.  Other ratios: exp(y log(num / den)) as before.
FM_DEV double fm_pow_ratio(double num, double den, double y) {
  const double dn = num - den, sm = num + den;
  if (fabs(dn) <= 0.19 * sm && sm > 0.0) {
    const double s = fm_div(dn, sm), v = s * s, w = v * v;
    double l;
    FM_POLY_EO(FM_LOG_L, 12, v, w, l)
    // z = y * 2 s l with the product's rounding error carried into the reduction (z_hi + z_lo)
    const double ls = (s + s) * l, z = y * ls;
    return fm_exp_core(z, fma(y, ls, -z));
  }
  return exp(y * log(num / den));
}
FM_DEV double fm_powpos(double x, double y) { return fm_pow_ratio(x, 1.0, y); }
// exp(z) with z clamped to [-700, 700] (1e-304 .. 1e304: no overflow / underflow handling, no libdevice fallback)
FM_DEV double fm_exp(double z) {
  const double zc = z < -700.0 ? -700.0 : (z > 700.0 ? 700.0 : z);
  return fm_exp_core(zc, 0.0);
}
// sin and cos for |x| < 3e9 (the quadrant is an int): x = n pi/2 + r by a three-term Cody-Waite reduction -- fma keeps
// x - n P1 exact, so the reduction error is |n| 2^-160 + one rounding of r -- then the kernels above on |r| <= pi/4 and
// the quadrant swap.  No libdevice fallback: the angles of this simulator are bounded by construction (attitudes; a
// missile heading integrates <= 2 rad/s for <= 60 s); NaN and infinity still give NaN.
FM_DEV void fm_sincos(double x, double* s, double* c) {
  const double n = fm_rint(x * FM_RED[0]);
  double r = fma(n, FM_RED[1], x);
  r = fma(n, FM_RED[2], r);
  r = fma(n, FM_RED[3], r);
  const int q = (int)n;
  const double z = r * r;
  double ps = FM_SIN_S[5], pc = FM_COS_C[5];
#pragma unroll
  for (int k = 4; k >= 0; k--) { ps = fma(ps, z, FM_SIN_S[k]); pc = fma(pc, z, FM_COS_C[k]); }
  const double sr = fma(r * z, ps, r), cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
  *s = (q & 2) ? -s0 : s0;
  *c = ((q + 1) & 2) ? -c0 : c0;
}
// sin alone: x = n pi + r, one odd polynomial on |r| <= pi/2 (Taylor through r^21: truncation 1.3e-18); same domain
FM_DEV double fm_sin(double x) {
  const double n = fm_rint(x * FM_RED[4]);
  double r = fma(n, FM_RED[5], x);
  r = fma(n, FM_RED[6], r);
  r = fma(n, FM_RED[7], r);
  const double z = r * r, w = z * z;
  double p;
  FM_POLY_EO(FM_SIN_T, 10, z, w, p)
  const double sr = fma(r * z, p, r);
  return ((int)n & 1) ? -sr : sr;
}
// sin and cos of a small angle (the earth rotation angle: 7.3e-5 rad/s of simulated time)
FM_DEV void fm_sincos_small(double x, double* s, double* c) {
  if (fabs(x) <= 0.78539816339744831) {
    const double z = x * x;
    double ps = FM_SIN_S[5], pc = FM_COS_C[5];
#pragma unroll
    for (int k = 4; k >= 0; k--) { ps = fma(ps, z, FM_SIN_S[k]); pc = fma(pc, z, FM_COS_C[k]); }
    *s = fma(x * z, ps, x);
    *c = fma(z * z, pc, fma(-0.5, z, 1.0));
  } else fm_sincos(x, s, c);
}
FM_DEV double fm_cos(double x) { double sn, cs; fm_sincos(x, &sn, &cs); return cs; }
// acos(x), |x| <= 1: 2 atan2(sqrt(1 - x), sqrt(1 + x)) -- 1 - x is exact near 1, so small angles keep their relative accuracy
FM_DEV double fm_acos(double x) {
  const double r = fm_atan2(fm_sqrt0(1.0 - x), fm_sqrt0(1.0 + x));
  return r + r;
}
// log(x), x > 0 and normal: x = m 2^e with m in [sqrt(1/2), sqrt(2)), ln m through atanh as above; log(0) = -inf
FM_DEV double fm_log(double x) {
  int e;
  double m = fm_frexp1(x, &e);
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double s = fm_div(m - 1.0, m + 1.0), v = s * s, w = v * v;
  double l;
  FM_POLY_EO(FM_LOG_L, 12, v, w, l)
  const double ed = (double)e;
  const double r = fma(ed, -FM_RED[9], fma(s + s, l, ed * -FM_RED[10]));
  return x > 0.0 ? r : (x == 0.0 ? -INFINITY : NAN);
}
// tanh and atanh to an ABSOLUTE accuracy of a few 1e-16 (reward shaping, compared at 1e-6): (1 - t) / (1 + t) with
// t = exp(-2 |x|), and log((1 + y) / (1 - y)) / 2 (atanh(-1) = -inf, atanh(1) = +inf as libm)
FM_DEV double fm_tanh(double x) {
  const double t = fm_exp(-2.0 * fabs(x));
  return copysign(fm_div(1.0 - t, 1.0 + t), x);
}
FM_DEV double fm_atanh(double y) {
  if (y >= 1.0) return INFINITY;
  return 0.5 * fm_log(fm_div(1.0 + y, 1.0 - y));
}
#endif
