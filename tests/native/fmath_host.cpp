#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "fmath.cuh"
static double ulp_err(double got, long double want) {
  if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
  int e; frexpl(want, &e);
  long double ulp = ldexpl(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}
int main() {
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double w_div = 0, w_rcp = 0, w_sqrt = 0, w_rsqrt = 0, w_sin = 0, w_cos = 0, w_ang = 0, w_pow = 0, w_pow_rel = 0;
  const int N = 2000000;
  for (int i = 0; i < N; i++) {
    const double a = ldexp(U(g) + 0.5, (int)(U(g) * 160) - 80) * (U(g) < 0.5 ? -1 : 1);
    const double b = ldexp(U(g) + 0.5, (int)(U(g) * 160) - 80) * (U(g) < 0.5 ? -1 : 1);
    w_div = fmax(w_div, ulp_err(fm_div(a, b), (long double)a / b));
    w_rcp = fmax(w_rcp, ulp_err(fm_rcp(b), 1.0L / b));
    const double x = fabs(a);
    w_sqrt = fmax(w_sqrt, ulp_err(fm_sqrt(x), sqrtl(x)));
    w_rsqrt = fmax(w_rsqrt, ulp_err(fm_rsqrt(x), 1.0L / sqrtl(x)));
    const double th = (U(g) * 2 - 1) * 0.785398163397448;
    double s, c; fm_sincos_small(th, &s, &c);
    w_sin = fmax(w_sin, ulp_err(s, sinl(th))); w_cos = fmax(w_cos, ulp_err(c, cosl(th)));
    // angle: an arbitrary (y, x) pair scaled like body velocities; s, c formed as the FDM forms them
    const double ang = (U(g) * 2 - 1) * (U(g) < 0.9 ? 0.78 : 3.14159);
    const double R = 100 + 900 * U(g), yy = R * sin(ang), xx = R * cos(ang);
    const double ir = fm_rcp(fm_sqrt(xx * xx + yy * yy));
    const double got = fm_angle_sc(yy * ir, xx * ir, yy, xx);
    w_ang = fmax(w_ang, ulp_err(got, atan2l(yy, xx)));
    // ISA-like power: ratio 0.7 .. 1.45, exponents +-(5.26 .. 34)
    const double den = 200 + 400 * U(g), num = den * (0.7 + 0.75 * U(g)), y = (U(g) < 0.5 ? -1 : 1) * (0.1 + 34 * U(g)) ;
    if (fabs(y * log(num / den)) < 5) {
      const double gp = fm_pow_ratio(num, den, y);
      const long double wp = powl((long double)num / den, y);
      w_pow = fmax(w_pow, ulp_err(gp, wp));
      w_pow_rel = fmax(w_pow_rel, (double)fabsl((gp - wp) / wp));
    }
  }
  double w_exp=0,w_gs=0,w_gc=0,w_s1=0, a_gs=0, a_gc=0, a_s1=0;
  for (int i = 0; i < N; i++) {
    const double z = (U(g) * 2 - 1) * (U(g) < 0.5 ? 3.0 : 600.0);
    w_exp = fmax(w_exp, ulp_err(fm_exp(z), expl(z)));
    const double x = (U(g) * 2 - 1) * (U(g) < 0.5 ? 7.0 : 9.0e4);
    double s, c; fm_sincos(x, &s, &c);
    w_gs = fmax(w_gs, ulp_err(s, sinl(x))); w_gc = fmax(w_gc, ulp_err(c, cosl(x)));
    a_gs = fmax(a_gs, (double)fabsl(s - sinl(x))); a_gc = fmax(a_gc, (double)fabsl(c - cosl(x)));
    const double s1 = fm_sin(x);
    w_s1 = fmax(w_s1, ulp_err(s1, sinl(x))); a_s1 = fmax(a_s1, (double)fabsl(s1 - sinl(x)));
  }
  printf("{\"exp\": %.3f, \"sincos_s\": %.3f, \"sincos_c\": %.3f, \"sin\": %.3f, \"abs_s\": %.3e, \"abs_c\": %.3e, \"abs_sin\": %.3e}\n", w_exp, w_gs, w_gc, w_s1, a_gs, a_gc, a_s1);
  // special values
  int bad = 0;
  if (fm_sqrt0(0.0) != 0.0) bad |= 1;
  if (fm_div(0.0, 3.0) != 0.0) bad |= 2;
  if (fm_angle_sc(0.0, 1.0, 0.0, 5.0) != 0.0) bad |= 4;
  if (fm_pow_ratio(300.0, 300.0, -5.2) != 1.0) bad |= 8;
  if (fabs(fm_pow_ratio(10.0, 300.0, 2.0) - (10.0 / 300.0) * (10.0 / 300.0)) > 1e-17) bad |= 16;   // fallback branch
  { double s, c; fm_sincos_small(2.5, &s, &c); if (fabs(s - sin(2.5)) > 1e-16 || fabs(c - cos(2.5)) > 1e-16) bad |= 32; }
  if (fabs(fm_angle_sc(sin(2.0), cos(2.0), sin(2.0), cos(2.0)) - 2.0) > 1e-15) bad |= 64;
  printf("{\"div\": %.3f, \"rcp\": %.3f, \"sqrt\": %.3f, \"rsqrt\": %.3f, \"sin\": %.3f, \"cos\": %.3f, \"angle\": %.3f, \"pow\": %.3f, \"pow_rel\": %.3e, \"bad\": %d}\n",
         w_div, w_rcp, w_sqrt, w_rsqrt, w_sin, w_cos, w_ang, w_pow, w_pow_rel, bad);
  return 0;
}
