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
    const double x = (U(g) * 2 - 1) * (U(g) < 0.5 ? 7.0 : (U(g) < 0.5 ? 9.0e4 : 2.0e8));
    double s, c; fm_sincos(x, &s, &c);
    w_gs = fmax(w_gs, ulp_err(s, sinl(x))); w_gc = fmax(w_gc, ulp_err(c, cosl(x)));
    a_gs = fmax(a_gs, (double)fabsl(s - sinl(x))); a_gc = fmax(a_gc, (double)fabsl(c - cosl(x)));
    const double s1 = fm_sin(x);
    w_s1 = fmax(w_s1, ulp_err(s1, sinl(x))); a_s1 = fmax(a_s1, (double)fabsl(s1 - sinl(x)));
  }
  printf("{\"exp\": %.3f, \"sincos_s\": %.3f, \"sincos_c\": %.3f, \"sin\": %.3f, \"abs_s\": %.3e, \"abs_c\": %.3e, \"abs_sin\": %.3e}\n", w_exp, w_gs, w_gc, w_s1, a_gs, a_gc, a_s1);
  double w_at2 = 0, a_at2 = 0;
  for (int i = 0; i < N; i++) {
    const double ang = (U(g) * 2 - 1) * 3.14159265, R = ldexp(1.0 + U(g), (int)(U(g) * 60) - 30);
    const double yy = R * sin(ang) * (U(g) < 0.02 ? 0.0 : 1.0), xx = R * cos(ang) * (U(g) < 0.02 ? 0.0 : 1.0);
    if (xx == 0.0 && yy == 0.0) continue;
    const double got = fm_atan2(yy, xx);
    const long double want = atan2l(yy, xx);
    if (want != 0.0L) w_at2 = fmax(w_at2, ulp_err(got, want));
    a_at2 = fmax(a_at2, (double)fabsl(got - want));
  }
  printf("{\"atan2\": %.3f, \"abs_atan2\": %.3e}\n", w_at2, a_at2);
  if (fm_atan2(0.0, 0.0) != 0.0 || fm_atan2(0.0, -2.0) != 3.141592653589793 || fm_atan2(-3.0, 0.0) != -1.5707963267948966) { printf("{\"atan2_special\": 1}\n"); }
  double a_acos = 0, r_acos = 0, r_log = 0, a_tanh = 0, a_atanh = 0, r_cos = 0;
  for (int i = 0; i < N; i++) {
    const double x = U(g) < 0.3 ? 1.0 - ldexp(U(g), -(int)(U(g) * 50)) : (U(g) * 2 - 1);
    const double xs = U(g) < 0.5 ? x : -x;
    const long double wa = acosl(xs);
    a_acos = fmax(a_acos, (double)fabsl(fm_acos(xs) - wa));
    if (wa > 0) r_acos = fmax(r_acos, ulp_err(fm_acos(xs), wa));
    const double lx = ldexp(0.5 + U(g), (int)(U(g) * 200) - 100);
    if (lx != 1.0) r_log = fmax(r_log, ulp_err(fm_log(lx), logl(lx)));
    const double tx = (U(g) * 2 - 1) * 12;
    a_tanh = fmax(a_tanh, (double)fabsl(fm_tanh(tx) - tanhl(tx)));
    const double ay = (U(g) * 2 - 1) * 0.9999;
    a_atanh = fmax(a_atanh, (double)fabsl(fm_atanh(ay) - atanhl(ay)));
    const double cx = (U(g) * 2 - 1) * 7.0;
    r_cos = fmax(r_cos, (double)fabsl(fm_cos(cx) - cosl(cx)));
  }
  int bad2 = 0;
  if (fm_acos(1.0) != 0.0 || fabs(fm_acos(-1.0) - 3.141592653589793) > 5e-16 || fabs(fm_acos(0.0) - 1.5707963267948966) > 3e-16) bad2 |= 1;
  if (!(fm_log(0.0) < -1e308) || fm_log(1.0) != 0.0 || !(fm_atanh(-1.0) < -1e308) || !(fm_atanh(1.0) > 1e308) || fm_tanh(0.0) != 0.0) bad2 |= 2;
  if (fm_exp(-1e4) > 1e-300 || fm_exp(-1e4) < 0.0 || fm_exp(0.0) != 1.0) bad2 |= 4;
  printf("{\"abs_acos\": %.3e, \"acos\": %.3f, \"log\": %.3f, \"abs_tanh\": %.3e, \"abs_atanh\": %.3e, \"abs_cos\": %.3e, \"bad2\": %d}\n",
         a_acos, r_acos, r_log, a_tanh, a_atanh, r_cos, bad2);
  // special values
  int bad = 0;
  if (fm_sqrt0(0.0) != 0.0 || fm_sqrt0(1e-310) != 0.0 || fm_sqrt0(4.0) != 2.0 || !(fm_sqrt0(NAN) != fm_sqrt0(NAN))) bad |= 1;
  if (fm_div(0.0, 3.0) != 0.0) bad |= 2;
  if (fm_angle_sc(0.0, 1.0, 0.0, 5.0) != 0.0) bad |= 4;
  if (fm_pow_ratio(300.0, 300.0, -5.2) != 1.0) bad |= 8;
  if (fabs(fm_pow_ratio(10.0, 300.0, 2.0) - (10.0 / 300.0) * (10.0 / 300.0)) > 1e-17) bad |= 16;   // fallback branch
  { double s, c; fm_sincos_small(2.5, &s, &c); if (fabs(s - sin(2.5)) > 3e-16 || fabs(c - cos(2.5)) > 3e-16) bad |= 32; }
  if (fabs(fm_angle_sc(sin(2.0), cos(2.0), sin(2.0), cos(2.0)) - 2.0) > 1e-15) bad |= 64;
  printf("{\"div\": %.3f, \"rcp\": %.3f, \"sqrt\": %.3f, \"rsqrt\": %.3f, \"sin\": %.3f, \"cos\": %.3f, \"angle\": %.3f, \"pow\": %.3f, \"pow_rel\": %.3e, \"bad\": %d}\n",
         w_div, w_rcp, w_sqrt, w_rsqrt, w_sin, w_cos, w_ang, w_pow, w_pow_rel, bad);
  return 0;
}
