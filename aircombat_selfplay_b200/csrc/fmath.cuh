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

#ifdef ACS_IEEE_MATH
FM_DEV double fm_div(double a, double b) { return a / b; }
FM_DEV double fm_rcp(double b) { return 1.0 / b; }
FM_DEV double fm_sqrt(double x) { return sqrt(x); }
FM_DEV double fm_sqrt0(double x) { return sqrt(x); }
FM_DEV double fm_rsqrt(double x) { return 1.0 / sqrt(x); }
FM_DEV void fm_sincos_small(double x, double* s, double* c) { sincos(x, s, c); }
FM_DEV double fm_angle_sc(double s, double c, double y, double x) { (void)s; (void)c; return atan2(y, x); }
FM_DEV double fm_powpos(double x, double y) { return exp(y * log(x)); }
FM_DEV double fm_pow_ratio(double num, double den, double y) { return exp(y * log(num / den)); }
FM_DEV double fm_exp(double z) { return exp(z); }
FM_DEV void fm_sincos(double x, double* s, double* c) { sincos(x, s, c); }
FM_DEV double fm_sin(double x) { return sin(x); }
#else

#ifdef __CUDACC__
FM_DEV double fm_rcp_seed(double b) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r; }     // MUFU.RCP64H
FM_DEV double fm_rsq_seed(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }   // MUFU.RSQ64H
FM_DEV double fm_pow2i(int n) { return __hiloint2double((n + 1023) << 20, 0); }
FM_DEV double fm_rint(double x) { return rint(x); }
#else
// host build (tests): seeds with a deliberately poor relative error of 2^-12
FM_DEV double fm_rcp_seed(double b) { return (double)(float)(1.0 / b) * (1.0 + 1.0 / 4096.0); }
FM_DEV double fm_rsq_seed(double x) { return (double)(float)(1.0 / sqrt(x)) * (1.0 - 1.0 / 4096.0); }
FM_DEV double fm_pow2i(int n) { return ldexp(1.0, n); }
FM_DEV double fm_rint(double x) { return rint(x); }
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
// the same for operands that can be exactly zero (a seed of +inf would give NaN)
FM_DEV double fm_sqrt0(double x) {
  const double g = fm_sqrt(x);
  return x == 0.0 ? 0.0 : g;
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

// sin and cos of a small angle (the earth rotation angle: 7.3e-5 rad/s of simulated time)
FM_DEV void fm_sincos_small(double x, double* s, double* c) {
  if (fabs(x) <= 0.78539816339744831) {
    const double z = x * x;
    double ps = FM_SIN_S[5], pc = FM_COS_C[5];
#pragma unroll
    for (int k = 4; k >= 0; k--) { ps = fma(ps, z, FM_SIN_S[k]); pc = fma(pc, z, FM_COS_C[k]); }
    *s = fma(x * z, ps, x);
    *c = fma(z * z, pc, fma(-0.5, z, 1.0));
  } else sincos(x, s, c);
}
// The angle whose sine s and cosine c are known (s^2 + c^2 = 1 to rounding): 2 atan(s / (1 + c)).  |angle| <= 45 deg
// takes the polynomial (angle of attack and sideslip of an aircraft that flies forwards); anything else is atan2(y, x).
FM_DEV double fm_angle_sc(double s, double c, double y, double x) {
  if (c >= 0.7072) {
    const double t = fm_div(s, 1.0 + c), u = t * t, w = u * u;
    double q;
    FM_POLY_EO(FM_ATAN_Q, 13, u, w, q)
    return (t + t) * q;
  }
  return atan2(y, x);
}
// exp(z + zl), |z| <= 700, zl a rounding-error term: z = n ln2 + r, |r| <= ln2 / 2, Taylor through r^13, scale by 2^n
FM_DEV double fm_exp_core(double z, double zl) {
  const double n = fm_rint(z * 1.4426950408889634);
  double r = fma(n, -6.93147180369123816490e-01, z);
  r = fma(n, -1.90821492927058770002e-10, r) + zl;
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
// exp(z) for |z| <= 700 (no overflow / underflow handling below that bound); libdevice beyond
FM_DEV double fm_exp(double z) {
  if (fabs(z) <= 700.0) return fm_exp_core(z, 0.0);
  return exp(z);
}
// sin and cos for |x| <= 1e5: x = n pi/2 + r by a three-term Cody-Waite reduction (fma keeps every product exact), then
// the kernels above on |r| <= pi/4 and the quadrant swap; libdevice beyond (its own slow path starts at 1.05e5)
FM_DEV void fm_sincos(double x, double* s, double* c) {
  if (fabs(x) <= 1.0e5) {
    const double n = fm_rint(x * 0.6366197723675814);
    double r = fma(n, -1.5707963267948966, x);
    r = fma(n, -6.123233995736766e-17, r);
    r = fma(n, 1.4973849048591698e-33, r);
    const int q = (int)n;
    const double z = r * r;
    double ps = FM_SIN_S[5], pc = FM_COS_C[5];
#pragma unroll
    for (int k = 4; k >= 0; k--) { ps = fma(ps, z, FM_SIN_S[k]); pc = fma(pc, z, FM_COS_C[k]); }
    const double sr = fma(r * z, ps, r), cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
    *s = (q & 2) ? -s0 : s0;
    *c = ((q + 1) & 2) ? -c0 : c0;
  } else sincos(x, s, c);
}
// sin alone: x = n pi + r, one odd polynomial on |r| <= pi/2 (Taylor through r^21: truncation 1.3e-18)
FM_DEV double fm_sin(double x) {
  if (fabs(x) <= 1.0e5) {
    const double n = fm_rint(x * 0.3183098861837907);
    double r = fma(n, -3.141592653589793, x);
    r = fma(n, -1.2246467991473532e-16, r);
    r = fma(n, 2.9947698097183397e-33, r);
    const double z = r * r, w = z * z;
    double p;
    FM_POLY_EO(FM_SIN_T, 10, z, w, p)
    const double sr = fma(r * z, p, r);
    return ((int)n & 1) ? -sr : sr;
  }
  return sin(x);
}
FM_DEV double fm_powpos(double x, double y) { return fm_pow_ratio(x, 1.0, y); }
#endif
