// Batched F-16 6-DoF flight-dynamics core for sm_100a: one thread integrates one aircraft.
//
// Replaces, for the F-16 model the reference loads, one jsbsim.FGFDMExec.run() call per aircraft per
// substep (reference envs/JSBSim/core/simulatior.py:223).  The per-frame model order and every
// one-frame lag follow JSBSim's FGFDMExec::Run (reference data/src/FGFDMExec.cpp:407-431, model order
// :222-236): Propagate -> Inertial -> Atmosphere -> FCS -> MassBalance -> Auxiliary -> Propulsion ->
// Aerodynamics -> Aircraft -> Accelerations.  "J/" = reference envs/JSBSim/data/src/.
//
// All arithmetic is fp64 (the path's arithmetic type); state lives in registers for the whole K-substep
// loop and touches HBM once per interaction step.  Flight-control and aerodynamic code is generated
// (gen/f16_gen.cuh); the 43 lookup tables are read from shared memory (pointer T).
#pragma once
#include <cmath>
#include "fmath.cuh"

#define FDM_DEV __device__ __forceinline__
// helpers instantiated many times per frame; ACS_NOINLINE_HELPERS trades call overhead for instruction-cache footprint
#ifdef ACS_NOINLINE_HELPERS
#define FDM_HELPER __device__ __noinline__
#else
#define FDM_HELPER __device__ __forceinline__
#endif
// compile-time string equality (field-ownership predicates of the generated header fold to constants)
__host__ __device__ constexpr bool f16_streq(const char* a, const char* b) {
  while (*a && *a == *b) { ++a; ++b; }
  return *a == *b;
}

static constexpr double RADTODEG = 180.0 / 3.14159265358979323846;
static constexpr double DEGTORAD = 3.14159265358979323846 / 180.0;
static constexpr double FTTOM = 0.3048;
static constexpr double INCHTOFT = 1.0 / 12.0;
static constexpr double SLUGTOLB = 32.174049;
static constexpr double LBTOSLUG = 1.0 / SLUGTOLB;
static constexpr double KGTOSLUG = 0.06852168;
static constexpr double FPSTOKTS = 1.0 / 1.68781;
// J/models/FGInertial.cpp:55-61
static constexpr double EARTH_GM = 14.0764417572E15;
static constexpr double EARTH_J2 = 1.08262982E-03;
static constexpr double EARTH_A = 20925646.32546;
static constexpr double EARTH_B = 20855486.5951;
static constexpr double EARTH_OMEGA = 0.00007292115;
static constexpr double G_ACCEL_REF = 9.80665 / FTTOM;

struct V3 { double x, y, z; };
FDM_DEV V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
FDM_DEV V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
FDM_DEV V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
FDM_DEV V3 operator*(double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }
FDM_DEV V3 cross(const V3& a, const V3& b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
FDM_DEV double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FDM_DEV double mag(const V3& a) { return fm_sqrt0(a.x * a.x + a.y * a.y + a.z * a.z); }

struct M33 { double m[3][3]; };
FDM_DEV V3 mul(const M33& a, const V3& v) {
  return v3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
FDM_DEV V3 mulT(const M33& a, const V3& v) {  // a^T * v
  return v3(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z, a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
            a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z);
}
FDM_DEV M33 mul(const M33& a, const M33& b) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
FDM_DEV M33 mulABt(const M33& a, const M33& b) {  // a * b^T
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[j][0] + a.m[i][1] * b.m[j][1] + a.m[i][2] * b.m[j][2];
  return r;
}

// Clip (J/models/flight_control/FGFCSComponent.cpp:266-290).  Two compare+select pairs: sm_100a has no fp64 min/max
// instruction, fmin/fmax expand to more (NaN handling).
FDM_DEV double f16_constrain(double lo, double v, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------ table lookups (J/math/FGTable.cpp:443-517)
// Stateless bracket rule shared with the oracle: first upper row r>=1 (0-based) whose key >= lookup key.  The
// breakpoints ascend, so that row is 1 + #{1 <= i <= n-2 : k[i] < key}: a branch-free count (n is a literal at every
// call site, the loop unrolls into independent shared-memory loads).  The model compiler stores the reciprocal
// breakpoint spacings behind the keys (k[n + r] = 1 / (k[r] - k[r-1])), so the interpolation factor is a multiply.
struct Bracket { int r; double f; };
// The whole table array once more in constant memory.  Every breakpoint the search compares against sits at an address
// known at compile time (the table offset and the loop index are literals), so it is a constant-bank operand of the
// comparison itself -- no load instruction, no shared-memory latency on the search (ncu: the counting loop was the
// hottest source line of the frame, LDS + DSETP + add per breakpoint, ~58 breakpoints per frame).  What is indexed by the
// search RESULT (per-lane addresses) stays in shared memory: the constant cache serialises divergent addresses.
constexpr int F16_KC_MAX = 2048;
__constant__ double g_f16_kc[F16_KC_MAX];
FDM_HELPER Bracket f16_bracket(const double* __restrict__ T, const int off, const int n, const double key) {
  const double* __restrict__ k = T + off;
  Bracket b;
  int r = 1;
#pragma unroll
#ifdef ACS_BRACKET_SMEM      // tuning builds: the search reads shared memory as it did before
  for (int i = 1; i < n - 1; i++) r += (k[i] < key) ? 1 : 0;
#else
  for (int i = 1; i < n - 1; i++) r += (g_f16_kc[off + i] < key) ? 1 : 0;
#endif
  double f = (key - k[r - 1]) * k[n + r];
  b.r = r; b.f = f > 1.0 ? 1.0 : (f < 0.0 ? 0.0 : f);
  return b;
}
// Uniformly spaced breakpoints (the alpha, beta and elevator grids): the row comes from one multiply, then one exact
// fix-up step against the stored breakpoints reproduces the search rule bit for bit (the candidate is off by <= 1).
FDM_DEV Bracket f16_bracket_u(const double* __restrict__ T, const int off, const int n, const double key, const double k0, const double inv_step) {
  const double* __restrict__ k = T + off;
  Bracket b;
  const double t0 = (key - k0) * inv_step, t = t0 < 0.0 ? 0.0 : (t0 > (double)(n - 2) ? (double)(n - 2) : t0);
  int r = 1 + (int)t;
  if (r > 1 && k[r - 1] >= key) r--;
  else if (r < n - 1 && k[r] < key) r++;
  const double f = (key - k[r - 1]) * k[n + r];
  b.r = r; b.f = f > 1.0 ? 1.0 : (f < 0.0 ? 0.0 : f);
  return b;
}
// 1-D: clamp, no extrapolation.  Below the first key r = 1 and f = 0, which already yields v[0] exactly.
// The convex form (1 - f) lo + f hi returns v[r - 1] at f = 0 and v[r] at f = 1 EXACTLY, so the clamped factor alone
// gives the end values beyond the first and last key -- no "key >= last key" test and select per lookup (the reference's
// lo + f (hi - lo) does not reproduce hi at f = 1; in between the two forms differ by <= 1 ulp).
FDM_DEV double f16_tab1(const double* __restrict__ v, const int n, const Bracket& b) {
  (void)n;
  return fma(b.f, v[b.r], (1.0 - b.f) * v[b.r - 1]);
}
FDM_DEV double f16_tab2(const double* __restrict__ v, const int nc, const Bracket& rb, const Bracket& cb) {
  const double* r0 = v + (rb.r - 1) * nc;
  const double* r1 = v + rb.r * nc;
  const double col1 = rb.f * (r1[cb.r - 1] - r0[cb.r - 1]) + r0[cb.r - 1];
  const double col2 = rb.f * (r1[cb.r] - r0[cb.r]) + r0[cb.r];
  return col1 + cb.f * (col2 - col1);
}

// ------------------------------------------------------------------ FCS primitives
// J/models/flight_control/FGPID.cpp:154-214 (non-"standard" form, no pvdot)
FDM_DEV double f16_pid(double Input, double test, double kp, double ki, double kd, int int_type, double dt,
                       double& prev, double& prev2, double& itot) {
  double I_out_delta = 0.0;
  const double Dval = (Input - prev) * fm_rcp(dt);   // dt is kernel-uniform: the reciprocal is loop-invariant (<= 1 ulp from the quotient)
  if (fabs(test) < 0.000001) {
    switch (int_type) {
      case 1: I_out_delta = Input; break;
      case 2: I_out_delta = 0.5 * (Input + prev); break;
      case 3: I_out_delta = 1.5 * Input - 0.5 * prev; break;
      case 4: I_out_delta = (23.0 * Input - 16.0 * prev + 5.0 * prev2) / 12.0; break;
      default: I_out_delta = 0.0;
    }
  }
  if (test < 0.0) itot = 0.0;
  itot += ki * dt * I_out_delta;
  const double Output = kp * Input + itot + kd * Dval;
  prev2 = test < 0.0 ? 0.0 : prev;
  prev = Input;
  return Output;
}
FDM_DEV bool f16_equal_to_roundoff(double a, double b) {
  // JSBSim: |a - b| <= 2 eps max(|a|, |b|).  sm_100a has no fp64 max: the compiler expands it (and re-forms it from the
  // equivalent "d <= eps |a| or d <= eps |b|") into DSETP.MAX + selects + NaN fix-up, ~12 instructions on every actuator
  // of every frame.  When the test can hold at all, |a| and |b| agree to 2 eps, so scaling by |a| alone decides
  // differently only for d within a relative 4e-16 of the threshold -- and either decision leaves the actuator within
  // 2 eps of the other (it keeps its position instead of snapping to an input that is already that close).
  return fabs(a - b) <= (2.0 * 2.220446049250313e-16) * fabs(a);
}
// J/models/flight_control/FGKinemat.cpp:99-157.  Input already scaled by the last detent.
FDM_DEV double f16_kinemat(const double* __restrict__ det, const double* __restrict__ tim, int n, double Input, double Output, double dt) {
  double dt0 = dt;
  Input = f16_constrain(det[0], Input, det[n - 1]);
  for (int it = 0; it < 8 && dt0 > 0.0 && !f16_equal_to_roundoff(Input, Output); ++it) {
    int ind = 1;
    while (ind < n - 1 && ((Input < Output) ? det[ind] < Output : det[ind] <= Output)) ++ind;
    if (tim[ind] <= 0.0) { Output = Input; break; }
    const double Rate = fm_div(det[ind] - det[ind - 1], tim[ind]);      // tim[ind] > 0 here, detents ascend: Rate > 0
    const double ThisInput = f16_constrain(det[ind - 1], Input, det[ind]);
    double ThisDt = fabs(fm_div(ThisInput - Output, Rate));
    if (dt0 < ThisDt) { ThisDt = dt0; if (Output < Input) Output += ThisDt * Rate; else Output -= ThisDt * Rate; }
    else Output = ThisInput;
    dt0 -= ThisDt;
  }
  return Output;
}

// Two-detent kinematic (every F-16 actuator except the flaps): the traverse loop above runs exactly once, so the
// component collapses to one rate-limited move.  d0, d1, t1 are literals of the generated call site.
FDM_HELPER double f16_kinemat2(const double d0, const double d1, const double t1, double Input, double Output, const double dt) {
  Input = f16_constrain(d0, Input, d1);
  if (dt > 0.0 && !f16_equal_to_roundoff(Input, Output)) {
    if (t1 <= 0.0) return Input;
    // d0, d1, t1 are literals: Rate and its reciprocal fold at compile time (no fp64 division on the actuator chain)
    const double Rate = (d1 - d0) / t1, invRate = t1 / (d1 - d0);
    const double ThisDt = fabs((Input - Output) * invRate);
    if (dt < ThisDt) { if (Output < Input) Output += dt * Rate; else Output -= dt * Rate; }
    else Output = Input;
  }
  return Output;
}
// x^y for x > 0 as exp(y ln x): 2-3 ulp instead of pow()'s < 1 ulp at a quarter of its instruction count
FDM_HELPER double f16_powpos(const double x, const double y) { return fm_powpos(x, y); }

#include "gen/f16_gen.cuh"
static_assert(F16_NTAB <= F16_KC_MAX, "g_f16_kc holds the whole table array");

// ------------------------------------------------------------------ ISA-1976 (J/models/atmosphere/FGStandardAtmosphere.cpp)
// Layer constants are computed on the host exactly as the constructor does (:130-150, :405-462) and passed in.
struct AtmoConst {
  double H[9], Tt[9], Lapse[8], PB[9], DB[9], Tmb[8];  // Tmb[b] = GetTemperature(GeometricAltitude(H[b]))
  double Reng, g0, EarthRadius, SLdensity, StdDaySLsoundspeed, StdDaySLpressure;
  // per-layer constants of the expressions below, so that no division by a constant is left on the device:
  // invdH[r] = 1 / (H[r] - H[r-1]), Pexp[b] = g0 / (Reng * Lapse[b]), Piso[b] = -g0 / (Reng * Tmb[b]), invSLdensity
  double invdH[9], Pexp[8], Piso[8], invSLdensity;
};
struct Atmo { double T, P, rho, a, density_altitude; };
FDM_DEV double atmo_geopot(const AtmoConst& c, double h) { return fm_div(h * c.EarthRadius, c.EarthRadius + h); }
FDM_DEV double atmo_geomet(const AtmoConst& c, double H) { return (H * c.EarthRadius) / (c.EarthRadius - H); }
FDM_DEV void atmosphere_calculate(const AtmoConst& c, double altitude, Atmo& o) {
  const double G = atmo_geopot(c, altitude);
  // GetTemperature :244-271
  double Tm;
  if (G >= 0.0) {
    if (G >= c.H[8]) Tm = c.Tt[8];
    else if (G <= c.H[0]) Tm = c.Tt[0];
    else {
      int r = 1;
      while (r < 8 && c.H[r] < G) r++;
      double f = (G - c.H[r - 1]) * c.invdH[r];
      if (f > 1.0) f = 1.0;
      Tm = f * (c.Tt[r] - c.Tt[r - 1]) + c.Tt[r - 1];
    }
  } else Tm = c.Tt[0] + G * c.Lapse[0];
  // GetPressure :191-227
  int b = 0;
  double BaseAlt = c.H[0];
  for (; b < 7; ++b) { const double testAlt = c.H[b + 1]; if (G < testAlt) break; BaseAlt = testAlt; }
  const double Tmb = c.Tmb[b], deltaH = G - BaseAlt, Lmb = c.Lapse[b];
  double Pm;
  if (Lmb != 0.0) Pm = c.PB[b] * fm_pow_ratio(Tmb, Tmb + Lmb * deltaH, c.Pexp[b]);   // (T_base / T)^(g0 / (R L))
  else Pm = c.PB[b] * exp(c.Piso[b] * deltaH);
  o.T = Tm; o.P = Pm; o.rho = fm_div(Pm, c.Reng * Tm);
  o.a = fm_sqrt(1.4 * c.Reng * Tm);
  // CalculateDensityAltitude :464-492.  On a standard day (the reference never biases temperature or pressure) the
  // density altitude IS the geometric altitude: the reference's power-law inversion returns it to 6e-13 relative
  // (tests/test_oracle_fdm.py::test_density_altitude_is_identity_on_a_standard_day), so the inversion is skipped.
#ifndef ACS_EXACT_DENSITY_ALTITUDE
  o.density_altitude = altitude;
  return;
#endif
  int d = 0;
  for (; d < 7; d++) if (o.rho >= c.DB[d + 1]) break;
  const double Ld = c.Lapse[d];
  double da;
  if (Ld != 0.0) da = c.H[d] + (c.Tt[d] / Ld) * (f16_powpos(o.rho / c.DB[d], -1.0 / (1.0 + c.g0 / (c.Reng * Ld))) - 1);
  else da = c.H[d] + (-c.Reng * c.Tt[d] / c.g0) * log(o.rho / c.DB[d]);
  o.density_altitude = atmo_geomet(c, da);
}

// Host side: the layer constants, computed exactly as the reference's constructor does and copied to constant memory by
// acs_create (tests/native/fdm_host.cpp fills the same struct when it compiles this header for the host).
static inline void host_atmo(AtmoConst& c) {
  // FGAtmosphere.h / FGStandardAtmosphere.cpp ctor (reference data/src/models/atmosphere/FGStandardAtmosphere.cpp:60-150)
  const double Rstar = 8.31432 * KGTOSLUG / (1.8 * (FTTOM * FTTOM));
  const double Mair = 28.9645 * KGTOSLUG / 1000.0;
  c.g0 = 9.80665 / FTTOM;
  c.Reng = Rstar / Mair;
  c.EarthRadius = 6356766.0 / FTTOM;
  const double h[9] = {0.0000, 36089.2388, 65616.7979, 104986.8766, 154199.4751, 167322.8346, 232939.6325, 278385.8268, 298556.4304};
  const double t[9] = {518.67, 389.97, 389.97, 411.57, 487.17, 487.17, 386.37, 336.5028, 336.5028};
  for (int i = 0; i < 9; i++) { c.H[i] = h[i]; c.Tt[i] = t[i]; }
  for (int b = 0; b < 8; b++) c.Lapse[b] = (t[b + 1] - t[b]) / (h[b + 1] - h[b]) - 0.0;
  c.StdDaySLpressure = 2116.228;
  c.PB[0] = c.StdDaySLpressure;
  for (int b = 0; b < 8; b++) {
    const double deltaH = h[b + 1] - h[b], Tmb = t[b];
    if (c.Lapse[b] != 0.0) { const double L = c.Lapse[b]; c.PB[b + 1] = c.PB[b] * std::pow(Tmb / (Tmb + L * deltaH), c.g0 / (c.Reng * L)); }
    else c.PB[b + 1] = c.PB[b] * std::exp(-c.g0 * deltaH / (c.Reng * Tmb));
  }
  for (int i = 0; i < 9; i++) c.DB[i] = c.PB[i] / (c.Reng * t[i]);
  c.SLdensity = c.StdDaySLpressure / (c.Reng * t[0]);
  c.StdDaySLsoundspeed = std::sqrt(1.4 * c.Reng * 518.67);
  // Tmb[b] = GetTemperature(GeometricAltitude(H[b])): the geometric/geopotential round trip of the base altitude
  for (int b = 0; b < 8; b++) {
    const double geomet = (h[b] * c.EarthRadius) / (c.EarthRadius - h[b]);
    const double G = (geomet * c.EarthRadius) / (c.EarthRadius + geomet);
    double Tm;
    if (G >= 0.0) {
      if (G <= h[0]) Tm = t[0];
      else if (G >= h[8]) Tm = t[8];
      else { int r = 1; while (r < 8 && h[r] < G) r++; double f = (G - h[r - 1]) / (h[r] - h[r - 1]); if (f > 1.0) f = 1.0; Tm = f * (t[r] - t[r - 1]) + t[r - 1]; }
    } else Tm = t[0] + G * c.Lapse[0];
    c.Tmb[b] = Tm;
  }
  // reciprocals / exponents the device code multiplies by (fdm_core.cuh, AtmoConst)
  c.invdH[0] = 0.0;
  for (int r = 1; r < 9; r++) c.invdH[r] = 1.0 / (h[r] - h[r - 1]);
  for (int b = 0; b < 8; b++) {
    c.Pexp[b] = c.Lapse[b] != 0.0 ? c.g0 / (c.Reng * c.Lapse[b]) : 0.0;
    c.Piso[b] = -c.g0 / (c.Reng * c.Tmb[b]);
  }
  c.invSLdensity = 1.0 / c.SLdensity;
}

// J/FGJSBBase.cpp:245-296
FDM_DEV double pitot_total_pressure(double mach, double p) {
  if (mach < 0) return p;
  if (mach < 1) { const double t = 1 + 0.2 * mach * mach; return p * ((t * t) * t * fm_sqrt(t)); }    // t^3.5
  const double m2 = mach * mach, m7 = (m2 * m2) * (m2 * mach), x = 7 * m2 - 1;
  return fm_div(p * 166.92158009316827 * m7, (x * x) * fm_sqrt(x));                                  // M^7 / x^2.5
}
FDM_DEV double mach_from_impact_pressure(double qc, double p) {
  const double A = fm_div(qc, p) + 1;
  // A^(1/3.5) = A^(2/7): fp32 seed, two Newton steps on y^7 = A^2 (relative error 1e-6 -> 3e-12 -> < 1 ulp)
  double y = (double)exp2f(0.2857142857142857f * log2f((float)A));
#pragma unroll
  for (int i = 0; i < 2; i++) { const double y2 = y * y, y3 = y2 * y, y6 = y3 * y3; y -= fm_div(y6 * y - A * A, 7.0 * y6); }
  double M = fm_sqrt0(5.0 * (y - 1));
  if (M > 1.0)
    for (int i = 0; i < 10; i++) { const double y = 1 - fm_rcp(7.0 * M * M); M = 0.8812848543473311 * fm_sqrt(A * ((y * y) * fm_sqrt(y))); }
  return M;
}

// ------------------------------------------------------------------ persisted per-aircraft state
struct AcCore {
  double q0, q1, q2, q3;          // qAttitudeECI
  V3 wi, ri, vi;                  // vPQRi, vInertialPosition, vInertialVelocity
  double epa;
  V3 dqv0, dqv1, dqa0;            // Adams-Bashforth history: dqInertialVelocity[0..1], dqUVWidot[0]
  V3 pqridot, uvwidot, bodyaccel; // Accelerations outputs of the previous frame
  double N1, N2, N2norm, FF;      // turbine
  double engflags;                // bit0 Starved, bit1 Augmentation
  double tank0, tank1, tank2, tank3;
  V3 cg;                          // vXYZcg of the previous frame (tank inertia uses it, J/FGFDMExec.cpp:569)
  double sim_time;
};
#define FDM_CORE_FIELDS(X) \
  X("q0", a.q0) X("q1", a.q1) X("q2", a.q2) X("q3", a.q3) \
  X("wi_x", a.wi.x) X("wi_y", a.wi.y) X("wi_z", a.wi.z) X("ri_x", a.ri.x) X("ri_y", a.ri.y) X("ri_z", a.ri.z) \
  X("vi_x", a.vi.x) X("vi_y", a.vi.y) X("vi_z", a.vi.z) X("epa", a.epa) \
  X("dqv0_x", a.dqv0.x) X("dqv0_y", a.dqv0.y) X("dqv0_z", a.dqv0.z) X("dqv1_x", a.dqv1.x) X("dqv1_y", a.dqv1.y) X("dqv1_z", a.dqv1.z) \
  X("dqa0_x", a.dqa0.x) X("dqa0_y", a.dqa0.y) X("dqa0_z", a.dqa0.z) \
  X("pqridot_x", a.pqridot.x) X("pqridot_y", a.pqridot.y) X("pqridot_z", a.pqridot.z) \
  X("uvwidot_x", a.uvwidot.x) X("uvwidot_y", a.uvwidot.y) X("uvwidot_z", a.uvwidot.z) \
  X("bodyaccel_x", a.bodyaccel.x) X("bodyaccel_y", a.bodyaccel.y) X("bodyaccel_z", a.bodyaccel.z) \
  X("N1", a.N1) X("N2", a.N2) X("N2norm", a.N2norm) X("FuelFlow_pph", a.FF) X("engflags", a.engflags) \
  X("tank0", a.tank0) X("tank1", a.tank1) X("tank2", a.tank2) X("tank3", a.tank3) \
  X("cg_x", a.cg.x) X("cg_y", a.cg.y) X("cg_z", a.cg.z) X("sim_time", a.sim_time)
#define FDM_N_CORE 45

// derived outputs of the last frame, consumed by the per-step logic and by get_property-style reads
// (reference envs/JSBSim/core/catalog.py JsbsimCatalog names in comments)
struct AcOut {
  double lon_deg, lat_geod_deg, h_sl_ft;        // position/long-gc-deg, lat-geod-deg, h-sl-ft
  double roll, pitch, heading;                  // attitude/roll-rad, pitch-rad, heading-true-rad
  double vn, ve, vd, u, v, w;                   // velocities/v-north-fps ... w-fps
  double vc_fps;                                // velocities/vc-fps
  double npx, npy, npz;                         // accelerations/n-pilot-{x,y,z}-norm
  double p, q, r;                               // velocities/{p,q,r}-rad_sec
  double eci_vmag;                              // velocities/eci-velocity-mag-fps
  double geod_alt_ft, alpha, beta, mach, thrust;  // diagnostics
};
#define FDM_OUT_FIELDS(X) \
  X("lon_deg", o.lon_deg) X("lat_geod_deg", o.lat_geod_deg) X("h_sl_ft", o.h_sl_ft) X("roll_rad", o.roll) X("pitch_rad", o.pitch) \
  X("heading_rad", o.heading) X("v_north_fps", o.vn) X("v_east_fps", o.ve) X("v_down_fps", o.vd) X("u_fps", o.u) X("v_fps", o.v) \
  X("w_fps", o.w) X("vc_fps", o.vc_fps) X("n_pilot_x", o.npx) X("n_pilot_y", o.npy) X("n_pilot_z", o.npz) \
  X("p_rad_sec", o.p) X("q_rad_sec", o.q) X("r_rad_sec", o.r) X("eci_velocity_mag_fps", o.eci_vmag) \
  X("geod_alt_ft", o.geod_alt_ft) X("alpha_rad", o.alpha) X("beta_rad", o.beta) X("mach", o.mach) X("thrust_lbs", o.thrust)
#define FDM_N_OUT 25

// ------------------------------------------------------------------ per-frame scratch shared between the model stages
struct Frame {
  M33 Ti2b, Tl2b;          // body matrices of this frame
  M33 Ti2l;                // inertial -> local (NED) of this frame's position
  double sin_epa, cos_epa;
  V3 ecef;                 // vLocation
  double radius, rxy, geodAlt, sinLatGd, cosLatGd, sinLon, cosLon, cosLatGc, h_asl, gd_s1, gd_cc;
  V3 uvw, pqr, vel;        // body velocity, body rates (wrt ECEF), NED velocity
  V3 grav;                 // ECEF gravity
  Atmo atm;
  double alpha, beta, Vt, Vt2, qbar, mach, vcas;
  V3 pilotN;
  double Mass; V3 cg; M33 J, Jinv;
  double thrust;
};

FDM_DEV M33 quat_T(double q0, double q1, double q2, double q3) {  // J/math/FGQuaternion.cpp:187-216
  const double q0q0 = q0 * q0, q1q1 = q1 * q1, q2q2 = q2 * q2, q3q3 = q3 * q3;
  const double q0q1 = q0 * q1, q0q2 = q0 * q2, q0q3 = q0 * q3, q1q2 = q1 * q2, q1q3 = q1 * q3, q2q3 = q2 * q3;
  M33 t;
  t.m[0][0] = q0q0 + q1q1 - q2q2 - q3q3; t.m[0][1] = 2.0 * (q1q2 + q0q3); t.m[0][2] = 2.0 * (q1q3 - q0q2);
  t.m[1][0] = 2.0 * (q1q2 - q0q3); t.m[1][1] = q0q0 - q1q1 + q2q2 - q3q3; t.m[1][2] = 2.0 * (q2q3 + q0q1);
  t.m[2][0] = 2.0 * (q1q3 + q0q2); t.m[2][1] = 2.0 * (q2q3 - q0q1); t.m[2][2] = q0q0 - q1q1 - q2q2 + q3q3;
  return t;
}
// J/math/FGMatrix33.cpp:106-154
FDM_DEV void mat_quat(const M33& a, double q[4]) {
  const double m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2], m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2],
               m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
  const double t0 = 1.0 + m11 + m22 + m33, t1 = 1.0 + m11 - m22 - m33, t2 = 1.0 - m11 + m22 - m33, t3 = 1.0 - m11 - m22 + m33;
  int idx = 0; double best = t0;
  if (t1 > best) { idx = 1; best = t1; }
  if (t2 > best) { idx = 2; best = t2; }
  if (t3 > best) { idx = 3; best = t3; }
  // q_main = sqrt(t) / 2 and the other three 0.25 (..) / q_main = (..) / (2 sqrt(t)) from ONE inverse square root (the
  // reference takes a square root and three quotients: <= 2 ulp apart)
  const double iq = fm_rsqrt(best), h = 0.5 * iq, qm = 0.5 * (best * iq);
  if (idx == 0) { q[0] = qm; q[1] = h * (m23 - m32); q[2] = h * (m31 - m13); q[3] = h * (m12 - m21); }
  else if (idx == 1) { q[1] = qm; q[0] = h * (m23 - m32); q[2] = h * (m12 + m21); q[3] = h * (m31 + m13); }
  else if (idx == 2) { q[2] = qm; q[0] = h * (m31 - m13); q[1] = h * (m12 + m21); q[3] = h * (m23 + m32); }
  else { q[3] = qm; q[0] = h * (m12 - m21); q[1] = h * (m13 + m31); q[2] = h * (m23 + m32); }
}
// J/math/FGMatrix33.cpp:159-192
FDM_DEV void mat_euler(const M33& a, double& phi, double& tht, double& psi) {
  bool lock = false;
  const double m13 = a.m[0][2];
  if (m13 <= -1.0) { tht = 0.5 * M_PI; lock = true; }
  else if (1.0 <= m13) { tht = -0.5 * M_PI; lock = true; }
  else tht = fm_atan2(-m13, fm_sqrt0((1.0 - m13) * (1.0 + m13)));    // asin(-m13), |m13| < 1
  if (lock) { phi = fm_atan2(-a.m[2][1], a.m[1][1]); psi = 0.0; }
  else {
    phi = fm_atan2(a.m[1][2], a.m[2][2]);
    psi = fm_atan2(a.m[0][1], a.m[0][0]);
    if (psi < 0.0) psi += 2 * M_PI;
  }
}

FDM_DEV V3 structural_to_body(const V3& cg, double x, double y, double z) {  // J/models/FGMassBalance.cpp:347-374
  return v3(INCHTOFT * (cg.x - x), INCHTOFT * (y - cg.y), INCHTOFT * (cg.z - z));
}
FDM_DEV void add_pointmass_inertia(M33& J, const V3& cg, double mass_sl, double x, double y, double z) {
  const V3 v = structural_to_body(cg, x, y, z);
  const V3 sv = mass_sl * v;
  const double xx = sv.x * v.x, yy = sv.y * v.y, zz = sv.z * v.z, xy = -sv.x * v.y, xz = -sv.x * v.z, yz = -sv.y * v.z;
  J.m[0][0] += yy + zz; J.m[0][1] += xy; J.m[0][2] += xz;
  J.m[1][0] += xy; J.m[1][1] += xx + zz; J.m[1][2] += yz;
  J.m[2][0] += xz; J.m[2][1] += yz; J.m[2][2] += xx + yy;
}

// ---- location-dependent part of Propagate: ECEF -> geodetic quantities and Tec2l (J/math/FGLocation.cpp:283-370)
FDM_DEV void location_derived(Frame& f, M33& Tec2l) {
  const double x = f.ecef.x, y = f.ecef.y, z = f.ecef.z;
  const double rxy2 = x * x + y * y;
  f.radius = fm_sqrt(rxy2 + z * z);
  const double rxy = fm_sqrt0(rxy2);
  f.rxy = rxy;
  double sinLon, cosLon;
  if (rxy == 0.0) { sinLon = 0.0; cosLon = 1.0; } else { const double ir = fm_rcp(rxy); sinLon = y * ir; cosLon = x * ir; }
  f.sinLon = sinLon; f.cosLon = cosLon;
  // geocentric latitude only enters through its sine and cosine (J2 gravity, sea-level radius): z/r and rxy/r
  f.cosLatGc = fm_div(rxy, f.radius);
  const double ec = EARTH_B / EARTH_A, ec2 = ec * ec, e2 = 1.0 - ec2, c = EARTH_A * e2;
  const double s0 = fabs(z), zc = ec * s0, c0 = ec * rxy, c02 = c0 * c0, s02 = s0 * s0, a02 = c02 + s02;
  const double a0 = fm_sqrt(a02), a03 = a02 * a0;
  double s1 = zc * a03 + c * s02 * s0;
  const double c1 = rxy * a03 - c * c02 * c0, cs0c0 = c * c0 * s0;
  const double b0 = 1.5 * cs0c0 * ((rxy * s0 - zc * c0) * a0 - cs0c0);
  s1 = s1 * a03 - b0 * s0;
  const double cc = ec * (c1 * a03 - b0 * c0);
  const double s12 = s1 * s1, cc2 = cc * cc, inorm = fm_rsqrt(s12 + cc2);
  const double sgn = z < 0.0 ? -1.0 : 1.0;
  const double cosLat = cc * inorm, sinLat = sgn * s1 * inorm;
  f.cosLatGd = cosLat; f.sinLatGd = sinLat; f.gd_s1 = s1; f.gd_cc = cc;
  f.geodAlt = (rxy * cc + s0 * s1 - EARTH_A * fm_sqrt(ec2 * s12 + cc2)) * inorm;
  Tec2l.m[0][0] = -cosLon * sinLat; Tec2l.m[0][1] = -sinLon * sinLat; Tec2l.m[0][2] = cosLat;
  Tec2l.m[1][0] = -sinLon; Tec2l.m[1][1] = cosLon; Tec2l.m[1][2] = 0.0;
  Tec2l.m[2][0] = -cosLon * cosLat; Tec2l.m[2][1] = -sinLon * cosLat; Tec2l.m[2][2] = -sinLat;
}

// ============================================================== one FDM frame
// ============================================================== one FDM frame, stage by stage
// The stages are separate functions so that the single-thread frame (fdm_frame) and the two-warp role split
// (fdm_split.cuh) run the same arithmetic.  dt = 0 reproduces the suspended-integration passes of FGFDMExec::RunIC.
struct WindAxes { double sa, ca, sb, cb; };

// Propagate (J/models/FGPropagate.cpp:218-297) in three parts.  The position chain -- inertial position (Adams-Bashforth 3
// on the PAST velocities), earth rotation angle, ECEF location, geodetic quantities, local frame -- does not depend on
// this frame's velocity or attitude update, only on the velocity the previous frame ended with, so a multi-warp frame
// computes it (and the gravity and atmosphere that hang off it) one frame ahead on another warp.  Called in the order
// rot, pos, combine on one thread they are FGPropagate::Run.
FDM_DEV void fdm_propagate_rot(AcCore& a, const double dt, V3& vi_before) {   // attitude, rates, inertial velocity
  vi_before = a.vi;   // the position integrator uses the velocity from before this frame's velocity update (:228-231)
  if (dt != 0.0) {
    a.sim_time += dt;  // FGFDMExec::IncrTime
    // quaternion: rectangular Euler on qdot of the previous frame's state, then normalise (:371-470)
    {
      const double qd0 = -0.5 * (a.q1 * a.wi.x + a.q2 * a.wi.y + a.q3 * a.wi.z);
      const double qd1 = 0.5 * (a.q0 * a.wi.x - a.q3 * a.wi.y + a.q2 * a.wi.z);
      const double qd2 = 0.5 * (a.q3 * a.wi.x + a.q0 * a.wi.y - a.q1 * a.wi.z);
      const double qd3 = 0.5 * (-a.q2 * a.wi.x + a.q1 * a.wi.y + a.q0 * a.wi.z);
      a.q0 += dt * qd0; a.q1 += dt * qd1; a.q2 += dt * qd2; a.q3 += dt * qd3;
      // |q| and 1 / |q| from one inverse-square-root chain (the reference divides by sqrt: <= 1 ulp apart)
      const double qq = a.q0 * a.q0 + a.q1 * a.q1 + a.q2 * a.q2 + a.q3 * a.q3, iq = fm_rsqrt(qq), n = qq * iq;
      // Select, not branch: about half of the aircraft renormalise in a given frame, and with a branch here ptxas has put
      // the reconvergence point behind the rest of Propagate (ncu: ~700 instructions per frame executed 2.25 times with 14
      // active lanes, +23 % warp instructions in the multi-warp frames).  x * 1.0 is exact, so the bits are the same.
      const double rn = (qq == 0.0 || fabs(n - 1.000) < 1e-10) ? 1.0 : iq;
      a.q0 *= rn; a.q1 *= rn; a.q2 *= rn; a.q3 *= rn;
    }
    a.wi = a.wi + dt * a.pqridot;                                                        // eRectEuler
    {                                                                                     // eAdamsBashforth2 on velocity
      const V3 a0 = a.uvwidot;
      a.vi = a.vi + dt * (1.5 * a0 - 0.5 * a.dqa0);
      a.dqa0 = a0;
    }
  }
}
// position state = a.ri, a.dqv0, a.dqv1, a.epa; v0 = the inertial velocity before this frame's update
FDM_DEV void fdm_propagate_pos(AcCore& a, Frame& f, const V3& v0, const double dt) {
  if (dt != 0.0) {                                                                        // eAdamsBashforth3 on position
    a.ri = a.ri + ((1 / 12.0) * dt) * (23.0 * v0 - 16.0 * a.dqv0 + 5.0 * a.dqv1);
    a.dqv1 = a.dqv0; a.dqv0 = v0;
  }
  a.epa += EARTH_OMEGA * dt;
  fm_sincos_small(a.epa, &f.sin_epa, &f.cos_epa);
  // vLocation = Ti2ec * vInertialPosition
  f.ecef = v3(f.cos_epa * a.ri.x + f.sin_epa * a.ri.y, -f.sin_epa * a.ri.x + f.cos_epa * a.ri.y, a.ri.z);
  M33 Tec2l;
  location_derived(f, Tec2l);
  // Ti2l = Tec2l * Ti2ec with Ti2ec = Rz(epa): only the first two columns mix (J/models/FGPropagate.cpp:475-496)
#pragma unroll
  for (int i = 0; i < 3; i++) {
    f.Ti2l.m[i][0] = Tec2l.m[i][0] * f.cos_epa - Tec2l.m[i][1] * f.sin_epa;
    f.Ti2l.m[i][1] = Tec2l.m[i][0] * f.sin_epa + Tec2l.m[i][1] * f.cos_epa;
    f.Ti2l.m[i][2] = Tec2l.m[i][2];
  }
}
// body matrices and velocities from the attitude / velocity of `rot` and the position / local frame of `pos`
// (ri_x, ri_y: inertial position components the earth-rate term needs)
FDM_DEV void fdm_propagate_combine(const AcCore& a, Props& p, Frame& f, const double ri_x, const double ri_y) {
  f.Ti2b = quat_T(a.q0, a.q1, a.q2, a.q3);
  f.Tl2b = mulABt(f.Ti2b, f.Ti2l);        // Ti2b * Tl2i
  // Omega = (0, 0, w): Omega x r = (-w y, w x, 0); Ti2b * Omega = w * third column of Ti2b
  f.uvw = mul(f.Ti2b, v3(a.vi.x + EARTH_OMEGA * ri_y, a.vi.y - EARTH_OMEGA * ri_x, a.vi.z));
  f.pqr = v3(a.wi.x - EARTH_OMEGA * f.Ti2b.m[0][2], a.wi.y - EARTH_OMEGA * f.Ti2b.m[1][2], a.wi.z - EARTH_OMEGA * f.Ti2b.m[2][2]);
  f.vel = mulT(f.Tl2b, f.uvw);
  // The FCS reads the Euler angles only as cos(pitch) * cos(roll) (fcs/n-pilot-z-correction), which is Tl2b(3,3); the
  // angles themselves are extracted once per interaction step in fdm_outputs.
  p.attitude_cos_pitch_cos_roll = f.Tl2b.m[2][2]; p.velocities_u_fps = f.uvw.x; p.velocities_v_fps = f.uvw.y;
}
FDM_DEV void fdm_stage_propagate(AcCore& a, Props& p, Frame& f, const double dt) {
  V3 v0;
  fdm_propagate_rot(a, dt, v0);
  fdm_propagate_pos(a, f, v0, dt);
  fdm_propagate_combine(a, p, f, a.ri.x, a.ri.y);
}

FDM_DEV void fdm_stage_gravity(Frame& f) {
  // ---------------- Inertial: J2 gravity in ECEF (J/models/FGInertial.cpp:193-211)
  {
    const double ir = fm_rcp(f.radius), sinLat = f.ecef.z * ir, adivr = EARTH_A * ir, preCommon = 1.5 * EARTH_J2 * adivr * adivr;
    const double xy = 1.0 - 5.0 * (sinLat * sinLat), z = 3.0 - 5.0 * (sinLat * sinLat), GMOverr2 = EARTH_GM * (ir * ir);
    f.grav.x = -GMOverr2 * ((1.0 + (preCommon * xy)) * f.ecef.x * ir);
    f.grav.y = -GMOverr2 * ((1.0 + (preCommon * xy)) * f.ecef.y * ir);
    f.grav.z = -GMOverr2 * ((1.0 + (preCommon * z)) * f.ecef.z * ir);
  }
}

FDM_DEV void fdm_stage_atmosphere(Props& p, Frame& f, const AtmoConst& ac) {
  // ---------------- Atmosphere at h = |r| - sea-level radius (J/models/FGPropagate.cpp:573-576, FGLocation.cpp:273-279)
  const double ecr = EARTH_B / EARTH_A;
  const double slr = (EARTH_A * ecr) * fm_rsqrt(1.0 - (1.0 - ecr * ecr) * f.cosLatGc * f.cosLatGc);
  f.h_asl = f.radius - slr;
  atmosphere_calculate(ac, f.h_asl, f.atm);
  p.atmosphere_density_altitude = f.atm.density_altitude;
}

FDM_DEV void fdm_stage_massbalance(AcCore& a, Frame& f) {
  // ---------------- MassBalance (J/models/FGMassBalance.cpp:181-260)
  {
    const double tw = ((a.tank0 + a.tank1) + a.tank2) + a.tank3;
    V3 tm = v3(0, 0, 0);
    tm = tm + a.tank0 * v3(K_TANK0_X, K_TANK0_Y, K_TANK0_Z);
    tm = tm + a.tank1 * v3(K_TANK1_X, K_TANK1_Y, K_TANK1_Z);
    tm = tm + a.tank2 * v3(K_TANK2_X, K_TANK2_Y, K_TANK2_Z);
    tm = tm + a.tank3 * v3(K_TANK3_X, K_TANK3_Y, K_TANK3_Z);
    M33 tankJ;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) tankJ.m[i][j] = 0.0;
    // tank inertia uses the previous frame's cg (LoadInputs runs before MassBalance::Run, J/FGFDMExec.cpp:563-571)
    add_pointmass_inertia(tankJ, a.cg, LBTOSLUG * a.tank0, K_TANK0_X, K_TANK0_Y, K_TANK0_Z);
    add_pointmass_inertia(tankJ, a.cg, LBTOSLUG * a.tank1, K_TANK1_X, K_TANK1_Y, K_TANK1_Z);
    add_pointmass_inertia(tankJ, a.cg, LBTOSLUG * a.tank2, K_TANK2_X, K_TANK2_Y, K_TANK2_Z);
    add_pointmass_inertia(tankJ, a.cg, LBTOSLUG * a.tank3, K_TANK3_X, K_TANK3_Y, K_TANK3_Z);
    const double pmw = K_PM0_W + K_PM1_W;
    const double Weight = K_emptywt + tw + pmw;
    f.Mass = LBTOSLUG * Weight;
    V3 pm = v3(0, 0, 0);
    pm = pm + K_PM0_W * v3(K_PM0_X, K_PM0_Y, K_PM0_Z);
    pm = pm + K_PM1_W * v3(K_PM1_X, K_PM1_Y, K_PM1_Z);
    const V3 num = (K_emptywt * v3(K_CG_X, K_CG_Y, K_CG_Z) + pm) + tm;
    const double rw = fm_rcp(Weight);
    f.cg = v3(num.x * rw, num.y * rw, num.z * rw);
    a.cg = f.cg;
    M33 J;
    if (!K_negated_crossproduct_inertia) {
      J.m[0][0] = K_ixx; J.m[0][1] = K_ixy; J.m[0][2] = -K_ixz; J.m[1][0] = K_ixy; J.m[1][1] = K_iyy; J.m[1][2] = K_iyz;
      J.m[2][0] = -K_ixz; J.m[2][1] = K_iyz; J.m[2][2] = K_izz;
    } else {
      J.m[0][0] = K_ixx; J.m[0][1] = -K_ixy; J.m[0][2] = K_ixz; J.m[1][0] = -K_ixy; J.m[1][1] = K_iyy; J.m[1][2] = -K_iyz;
      J.m[2][0] = K_ixz; J.m[2][1] = -K_iyz; J.m[2][2] = K_izz;
    }
    add_pointmass_inertia(J, f.cg, LBTOSLUG * K_emptywt, K_CG_X, K_CG_Y, K_CG_Z);
    M33 pmJ;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) pmJ.m[i][j] = 0.0;
    add_pointmass_inertia(pmJ, f.cg, LBTOSLUG * K_PM0_W, K_PM0_X, K_PM0_Y, K_PM0_Z);
    add_pointmass_inertia(pmJ, f.cg, LBTOSLUG * K_PM1_W, K_PM1_X, K_PM1_Y, K_PM1_Z);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) J.m[i][j] = (J.m[i][j] + pmJ.m[i][j]) + tankJ.m[i][j];
    f.J = J;
    const double Ixx = J.m[0][0], Iyy = J.m[1][1], Izz = J.m[2][2], Ixy = -J.m[0][1], Ixz = -J.m[0][2], Iyz = -J.m[1][2];
    double k1 = (Iyy * Izz - Iyz * Iyz), k2 = (Iyz * Ixz + Ixy * Izz), k3 = (Ixy * Iyz + Iyy * Ixz);
    const double denom = fm_rcp(Ixx * k1 - Ixy * k2 - Ixz * k3);
    k1 = k1 * denom; k2 = k2 * denom; k3 = k3 * denom;
    const double k4 = (Izz * Ixx - Ixz * Ixz) * denom, k5 = (Ixy * Ixz + Iyz * Ixx) * denom, k6 = (Ixx * Iyy - Ixy * Ixy) * denom;
    f.Jinv.m[0][0] = k1; f.Jinv.m[0][1] = k2; f.Jinv.m[0][2] = k3;
    f.Jinv.m[1][0] = k2; f.Jinv.m[1][1] = k4; f.Jinv.m[1][2] = k5;
    f.Jinv.m[2][0] = k3; f.Jinv.m[2][1] = k5; f.Jinv.m[2][2] = k6;
  }
}

// ---------------- Auxiliary (J/models/FGAuxiliary.cpp:134-231), in two halves so that the multi-warp frames can run the
// air-data half next to the atmosphere it depends on: `kin` needs the body velocities, rates and last frame's
// accelerations, `air` the true airspeed and the atmosphere.  Called back to back they are FGAuxiliary::Run.
FDM_DEV double fdm_airspeed(Frame& f) {   // Vt^2 and Vt from the body velocities (no wind: the reference never sets one)
  const double U = f.uvw.x, V = f.uvw.y, W = f.uvw.z;
  const double AeroU2 = U * U, AeroV2 = V * V, AeroW2 = W * W, mUW = AeroU2 + AeroW2;
  f.Vt2 = mUW + AeroV2;
  f.Vt = fm_sqrt0(f.Vt2);
  return mUW;
}
FDM_DEV void fdm_stage_aux_kin(const AcCore& a, Props& p, Frame& f, WindAxes& w) {
  double& sa = w.sa; double& ca = w.ca; double& sb = w.sb; double& cb = w.cb;
  // accelerations are the previous frame's
  const double U = f.uvw.x, V = f.uvw.y, W = f.uvw.z;
  const double mUW = fdm_airspeed(f);
  f.alpha = 0.0; f.beta = 0.0;
  sa = 0.0; ca = 1.0; sb = 0.0; cb = 1.0;
  if (f.Vt > 0.001) {
    // the wind->body matrix needs only sines and cosines of alpha and beta: ratios of the velocity components
    const double sUW = fm_sqrt0(mUW), iVt = fm_rcp(f.Vt);
    // beta = atan2(V, sUW) and alpha = atan2(W, U) from the sines and cosines the wind axes need anyway (fmath.cuh)
    sb = V * iVt; cb = sUW * iVt; f.beta = fm_angle_sc(sb, cb, V, sUW);
    if (mUW >= 1E-6) { const double iUW = fm_rcp(sUW); sa = W * iUW; ca = U * iUW; f.alpha = fm_angle_sc(sa, ca, W, U); }
  }
  const double Vground = fm_sqrt0(f.vel.x * f.vel.x + f.vel.y * f.vel.y);
  const V3 eye = structural_to_body(f.cg, K_EYEPOINT_X, K_EYEPOINT_Y, K_EYEPOINT_Z);
  V3 pa = a.bodyaccel + cross(a.pqridot, eye);
  pa = pa + cross(a.wi, cross(a.wi, eye));
  const double rg = 1.0 / G_ACCEL_REF;
  f.pilotN = v3(pa.x * rg, pa.y * rg, pa.z * rg);
  const V3 rp = structural_to_body(f.cg, K_AERORP_X, K_AERORP_Y, K_AERORP_Z);
  const double vMacz = f.Tl2b.m[0][2] * rp.x + f.Tl2b.m[1][2] * rp.y + f.Tl2b.m[2][2] * rp.z;  // (Tb2l * RPBody)(3)
  p.aero_alpha_rad = f.alpha; p.aero_alpha_deg = f.alpha * RADTODEG; p.aero_beta_rad = f.beta;
  p.velocities_vg_fps = Vground;
  p.velocities_p_aero_rad_sec = f.pqr.x; p.velocities_q_aero_rad_sec = f.pqr.y; p.velocities_r_aero_rad_sec = f.pqr.z;
  p.accelerations_n_pilot_y_norm = f.pilotN.y; p.accelerations_n_pilot_z_norm = f.pilotN.z;
  p.aero_h_b_mac_ft = (f.geodAlt - vMacz) * (1.0 / K_bw);
}
FDM_DEV void fdm_stage_aux_air(Props& p, Frame& f, const AtmoConst& ac) {   // needs f.Vt2, f.Vt, f.atm
  f.qbar = (0.5 * f.atm.rho) * f.Vt2;
  f.mach = fm_div(f.Vt, f.atm.a);
  if (fabs(f.mach) > 0.0) {
    const double qc = pitot_total_pressure(f.mach, f.atm.P) - f.atm.P;
    f.vcas = ac.StdDaySLsoundspeed * mach_from_impact_pressure(qc, ac.StdDaySLpressure);
  } else f.vcas = 0.0;
  p.aero_qbar_psf = f.qbar; p.velocities_mach = f.mach; p.velocities_vc_kts = f.vcas * FPSTOKTS;
}
FDM_DEV void fdm_stage_auxiliary(const AcCore& a, Props& p, Frame& f, const AtmoConst& ac, WindAxes& w) {
  fdm_stage_aux_kin(a, p, f, w);
  fdm_stage_aux_air(p, f, ac);
}

// turbine (J/models/propulsion/FGTurbine.cpp:107-272): engine state = a.N1, a.N2, a.N2norm, a.FF; `starved` is the flag
// ConsumeFuel left last frame, `augmentation` the afterburner latch; returns the thrust
FDM_DEV double fdm_stage_engine(AcCore& a, const Props& p, const Atmo& atm, const double qbar, const double* __restrict__ T,
                                const AtmoConst& ac, const double dt, const bool starved, bool& augmentation) {
  double idleT, milT, augT;
  f16_engine_tables(p, T, idleT, milT, augT);
  double ThrottlePos = p.fcs_throttle_pos_norm, AugmentCmd;
  if (ThrottlePos > 1.0) { AugmentCmd = ThrottlePos - 1.0; ThrottlePos -= AugmentCmd; } else AugmentCmd = 0.0;
  double thrust;
  const double N1_factor = K_ENG_maxn1 - K_ENG_idlen1, N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
  auto seek = [dt](double v, double target, double accel, double decel) {
    if (v > target) { v -= dt * decel; if (v < target) v = target; }
    else if (v < target) { v += dt * accel; if (v > target) v = target; }
    return v;
  };
  if (dt == 0.0) {  // tpTrim (:341-372)
    const double idlethrust = K_ENG_milthrust * idleT, milthrust = (K_ENG_milthrust - idlethrust) * milT;
    const double N2t = K_ENG_idlen2 + ThrottlePos * N2_factor, N2n = (N2t - K_ENG_idlen2) / N2_factor;
    thrust = (idlethrust + (milthrust * N2n * N2n)) * (1.0 - 0.0);
    if (AugmentCmd > 0.0) { const double tdiff = (K_ENG_maxthrust * augT) - thrust; thrust += (tdiff * AugmentCmd); }
  } else if (starved) {  // tpOff (:172-194)
    a.FF = seek(a.FF, 0, 1000.0, 10000.0);
    a.N1 = seek(a.N1, qbar / 10.0, a.N1 / 2.0 + 0.1, a.N1 / 2.0);
    a.N2 = seek(a.N2, qbar / 15.0, a.N2 / 2.0 + 0.1, a.N2 / 2.0);
    augmentation = false;
    thrust = 0.0;
  } else {  // tpRun (:196-272)
    const double idlethrust = K_ENG_milthrust * idleT, milthrust = (K_ENG_milthrust - idlethrust) * milT;
    const double sigma = atm.rho * ac.invSLdensity;
    const double n0 = a.N2norm + 0.1, n = n0 < 1.0 ? n0 : 1.0;   // fmin without its NaN handling
    const double sden = (1 + 3 * (1 - n) * (1 - n) * (1 - n) + (1 - sigma));
    const double dbase = 90.0 / (K_ENG_bypassratio + 3.0);
    const double rden = fm_rcp(sden);   // one reciprocal for the three spool rates
    const double up = (1.0 * dbase) * rden, dn2 = (3.0 * dbase) * rden, dn1 = (2.4 * dbase) * rden;
    a.N2 = seek(a.N2, K_ENG_idlen2 + ThrottlePos * N2_factor, up, dn2);
    a.N1 = seek(a.N1, K_ENG_idlen1 + ThrottlePos * N1_factor, up, dn1);
    a.N2norm = (a.N2 - K_ENG_idlen2) * (1.0 / N2_factor);
    thrust = idlethrust + (milthrust * a.N2norm * a.N2norm);
    if (!augmentation) {
      const double tsfc = K_ENG_tsfc * fm_sqrt(atm.T * (1.0 / 389.7)) * (0.84 + (1 - a.N2norm) * (1 - a.N2norm));
      a.FF = seek(a.FF, thrust * tsfc, 1000.0, 10000.0);
      if (a.FF < K_ENG_idleff) a.FF = K_ENG_idleff;
    }
    if (AugmentCmd > 0.0) {
      augmentation = true;
      const double tdiff = (K_ENG_maxthrust * augT) - thrust;
      thrust += (tdiff * AugmentCmd);
      a.FF = seek(a.FF, thrust * K_ENG_atsfc, 5000.0, 10000.0);
    } else augmentation = false;
  }
  return thrust;
}

// FGPropulsion::ConsumeFuel (J/models/FGPropulsion.cpp:164-260): equal split over the tanks that still hold fuel;
// returns the Starved flag the engine sees NEXT frame
FDM_DEV bool fdm_stage_consume_fuel(AcCore& a, const double dt, const bool starved, const bool trim_fuel_freeze) {
  // ConsumeFuel: equal split over tanks that still hold fuel; Starved takes effect next frame
  bool starved_next = starved;
  if (!trim_fuel_freeze) {
    const int n_with = (a.tank0 > 0.0) + (a.tank1 > 0.0) + (a.tank2 > 0.0) + (a.tank3 > 0.0);
    starved_next = (n_with == 0);
    if (n_with > 0) {
      // n_with is 1..4: its reciprocal is a select, not a division
      const double inv_n = n_with == 2 ? 0.5 : (n_with == 4 ? 0.25 : (n_with == 1 ? 1.0 : 1.0 / 3.0));
      const double per = ((a.FF * (1.0 / 3600.0)) * dt) * inv_n;
      auto drain = [per](double& c) { if (c > 0.0) { if (c - per >= 0.0) c -= per; else c = 0.0; } };
      drain(a.tank0); drain(a.tank1); drain(a.tank2); drain(a.tank3);
    }
  }
  return starved_next;
}

// FGAerodynamics wind->body + FGAircraft sums + FGAccelerations; c[6] = DRAG, SIDE, LIFT, ROLL, PITCH, YAW build-ups
FDM_DEV void fdm_stage_accelerations(AcCore& a, Frame& f, const WindAxes& w, const double c[6]) {
  const double sa = w.sa, ca = w.ca, sb = w.sb, cb = w.cb;
  // wind -> body (J/models/FGAerodynamics.cpp:205-214): drag and lift flip sign, F_b = Tw2b * F_w
  const double fw0 = -c[0], fw1 = c[1], fw2 = -c[2];
  V3 Fa;
  Fa.x = (ca * cb) * fw0 + (-ca * sb) * fw1 + (-sa) * fw2;
  Fa.y = sb * fw0 + cb * fw1 + 0.0 * fw2;
  Fa.z = (sa * cb) * fw0 + (-sa * sb) * fw1 + ca * fw2;
  const V3 rp = structural_to_body(f.cg, K_AERORP_X, K_AERORP_Y, K_AERORP_Z);
  const V3 Ma = v3(c[3], c[4], c[5]) + cross(rp, Fa);
  // thruster (J/models/propulsion/FGForce.cpp:78-91)
  const V3 Fp = v3(f.thrust, 0.0, 0.0);
  const V3 Mp = cross(structural_to_body(f.cg, K_THRUSTER_X, K_THRUSTER_Y, K_THRUSTER_Z), Fp);
  const V3 F = Fa + Fp, M = Ma + Mp;
  // Accelerations (J/models/FGAccelerations.cpp:138-207)
  a.pqridot = mul(f.Jinv, M - cross(a.wi, mul(f.J, a.wi)));
  const double rm = fm_rcp(f.Mass);
  a.bodyaccel = v3(F.x * rm, F.y * rm, F.z * rm);
  // vUVWidot = Tb2i * vBodyAccel + Tec2i * vGravAccel
  const V3 gi = v3(f.cos_epa * f.grav.x - f.sin_epa * f.grav.y, f.sin_epa * f.grav.x + f.cos_epa * f.grav.y, f.grav.z);
  a.uvwidot = mulT(f.Ti2b, a.bodyaccel) + gi;
}

#ifdef ACS_FRAME_PROFILE
// tuning builds only: cycles per stage of thread 0 of block 0, accumulated (acs_debug_frame_profile)
__device__ long long g_frame_prof[16];
#define FPROF(i) if (fprof_) { const long long n_ = clock64(); g_frame_prof[i] += n_ - fpc_; fpc_ = n_; }
#else
#define FPROF(i)
#endif
FDM_DEV void fdm_frame(AcCore& a, Props& p, FcsState& s, Frame& f, const double* __restrict__ T, const AtmoConst& ac,
                       const double dt, const double fcs_dt, const bool trim_fuel_freeze) {
#ifdef ACS_FRAME_PROFILE
  const bool fprof_ = blockIdx.x == 0 && threadIdx.x == 0 && dt != 0.0;
  long long fpc_ = clock64();
#endif
  fdm_stage_propagate(a, p, f, dt);
  FPROF(0)
  fdm_stage_gravity(f);
  FPROF(1)
  fdm_stage_atmosphere(p, f, ac);
  FPROF(2)
  f16_fcs(p, s, T, fcs_dt);
  FPROF(3)
  fdm_stage_massbalance(a, f);
  FPROF(4)
  WindAxes w;
  fdm_stage_auxiliary(a, p, f, ac, w);
  FPROF(5)
  {
    const int flags = (int)a.engflags;
    bool augmentation = flags & 2;
    f.thrust = fdm_stage_engine(a, p, f.atm, f.qbar, T, ac, dt, flags & 1, augmentation);
    const bool starved_next = fdm_stage_consume_fuel(a, dt, flags & 1, trim_fuel_freeze);
    a.engflags = (double)((starved_next ? 1 : 0) | (augmentation ? 2 : 0));
  }
  FPROF(6)
  double c[6];
#ifdef ACS_FRAME_PROFILE
  f16_aero<1>(p, T, 2 * f.Vt, c); FPROF(7)
  f16_aero<2>(p, T, 2 * f.Vt, c); FPROF(8)
  f16_aero<4>(p, T, 2 * f.Vt, c); FPROF(9)
  f16_aero<8>(p, T, 2 * f.Vt, c); FPROF(10)
  f16_aero<16>(p, T, 2 * f.Vt, c); FPROF(11)
  f16_aero<32>(p, T, 2 * f.Vt, c); FPROF(12)
#else
  f16_aero<63>(p, T, 2 * f.Vt, c);
#endif
  fdm_stage_accelerations(a, f, w, c);
  FPROF(13)
}

// ============================================================== lean frame (throughput kernel)
// fdm_frame keeps its whole scratch (three 3x3 matrices, the inertia tensor and its inverse, the geodetic quantities, the
// local-frame velocities ...) alive until the frame ends, because its caller reads outputs from it afterwards: ~145
// doubles live through the flight-control / engine / aerodynamics stages, i.e. 290 registers for a 255-register thread
// (ncu: 73 STL + 90 LDL per frame, 15 % long-scoreboard stalls).  The lean frame evaluates the SAME expressions in the same
// order but hands nothing back except the handful of scalars that cannot be recomputed from the carried state
// (FrameKeep); what the per-step logic reads is recomputed ONCE per interaction step from the final state by
// fdm_refresh (position and attitude do not change between Propagate and the end of a frame):
//   * body / local matrices, geodetic quantities, velocities: recomputed after the K loop (fdm_refresh);
//   * the inertia tensor and its inverse: built where Accelerations consumes them (from the tank contents and the
//     previous frame's cg, as MassBalance does), not held from MassBalance through the aerodynamics;
//   * Ti2b: rebuilt from the quaternion in Accelerations; gravity: rotated to the inertial frame at once (3 doubles
//     instead of gravity + sin / cos of the earth rotation angle).
// opaque(): the optimiser must not merge a recomputation with the original (that would re-create the long live range).
#ifdef __CUDACC__
FDM_DEV double opaque(double x) { asm volatile("" : "+d"(x)); return x; }
#else
FDM_DEV double opaque(double x) { return x; }     // host build of this header (tests/native/fdm_host.cpp)
#endif

struct FrameKeep { double pilot_nx, vcas, beta, thrust; };

// MassBalance, first half: weight, mass, cg (J/models/FGMassBalance.cpp:181-260); the inertia half is fdm_inertia
FDM_DEV void fdm_mass_cg(const AcCore& a, double& Mass, V3& cg) {
  const double tw = ((a.tank0 + a.tank1) + a.tank2) + a.tank3;
  V3 tm = v3(0, 0, 0);
  tm = tm + a.tank0 * v3(K_TANK0_X, K_TANK0_Y, K_TANK0_Z);
  tm = tm + a.tank1 * v3(K_TANK1_X, K_TANK1_Y, K_TANK1_Z);
  tm = tm + a.tank2 * v3(K_TANK2_X, K_TANK2_Y, K_TANK2_Z);
  tm = tm + a.tank3 * v3(K_TANK3_X, K_TANK3_Y, K_TANK3_Z);
  const double pmw = K_PM0_W + K_PM1_W;
  const double Weight = K_emptywt + tw + pmw;
  Mass = LBTOSLUG * Weight;
  V3 pm = v3(0, 0, 0);
  pm = pm + K_PM0_W * v3(K_PM0_X, K_PM0_Y, K_PM0_Z);
  pm = pm + K_PM1_W * v3(K_PM1_X, K_PM1_Y, K_PM1_Z);
  const V3 num = (K_emptywt * v3(K_CG_X, K_CG_Y, K_CG_Z) + pm) + tm;
  const double rw = fm_rcp(Weight);
  cg = v3(num.x * rw, num.y * rw, num.z * rw);
}
// MassBalance, second half: inertia tensor about the cg and its inverse.  tanks = the contents MassBalance saw (before this
// frame's fuel burn), cg_prev = the previous frame's cg (tank inertia lags, J/FGFDMExec.cpp:563-571), cg = this frame's
FDM_DEV void fdm_inertia(const double t0, const double t1, const double t2, const double t3, const V3& cg_prev, const V3& cg, M33& Jout, M33& Jinv) {
  M33 tankJ;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) tankJ.m[i][j] = 0.0;
  add_pointmass_inertia(tankJ, cg_prev, LBTOSLUG * t0, K_TANK0_X, K_TANK0_Y, K_TANK0_Z);
  add_pointmass_inertia(tankJ, cg_prev, LBTOSLUG * t1, K_TANK1_X, K_TANK1_Y, K_TANK1_Z);
  add_pointmass_inertia(tankJ, cg_prev, LBTOSLUG * t2, K_TANK2_X, K_TANK2_Y, K_TANK2_Z);
  add_pointmass_inertia(tankJ, cg_prev, LBTOSLUG * t3, K_TANK3_X, K_TANK3_Y, K_TANK3_Z);
  M33 J;
  if (!K_negated_crossproduct_inertia) {
    J.m[0][0] = K_ixx; J.m[0][1] = K_ixy; J.m[0][2] = -K_ixz; J.m[1][0] = K_ixy; J.m[1][1] = K_iyy; J.m[1][2] = K_iyz;
    J.m[2][0] = -K_ixz; J.m[2][1] = K_iyz; J.m[2][2] = K_izz;
  } else {
    J.m[0][0] = K_ixx; J.m[0][1] = -K_ixy; J.m[0][2] = K_ixz; J.m[1][0] = -K_ixy; J.m[1][1] = K_iyy; J.m[1][2] = -K_iyz;
    J.m[2][0] = K_ixz; J.m[2][1] = -K_iyz; J.m[2][2] = K_izz;
  }
  add_pointmass_inertia(J, cg, LBTOSLUG * K_emptywt, K_CG_X, K_CG_Y, K_CG_Z);
  M33 pmJ;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) pmJ.m[i][j] = 0.0;
  add_pointmass_inertia(pmJ, cg, LBTOSLUG * K_PM0_W, K_PM0_X, K_PM0_Y, K_PM0_Z);
  add_pointmass_inertia(pmJ, cg, LBTOSLUG * K_PM1_W, K_PM1_X, K_PM1_Y, K_PM1_Z);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J.m[i][j] = (J.m[i][j] + pmJ.m[i][j]) + tankJ.m[i][j];
  Jout = J;
  const double Ixx = J.m[0][0], Iyy = J.m[1][1], Izz = J.m[2][2], Ixy = -J.m[0][1], Ixz = -J.m[0][2], Iyz = -J.m[1][2];
  double k1 = (Iyy * Izz - Iyz * Iyz), k2 = (Iyz * Ixz + Ixy * Izz), k3 = (Ixy * Iyz + Iyy * Ixz);
  const double denom = fm_rcp(Ixx * k1 - Ixy * k2 - Ixz * k3);
  k1 = k1 * denom; k2 = k2 * denom; k3 = k3 * denom;
  const double k4 = (Izz * Ixx - Ixz * Ixz) * denom, k5 = (Ixy * Ixz + Iyz * Ixx) * denom, k6 = (Ixx * Iyy - Ixy * Ixy) * denom;
  Jinv.m[0][0] = k1; Jinv.m[0][1] = k2; Jinv.m[0][2] = k3;
  Jinv.m[1][0] = k2; Jinv.m[1][1] = k4; Jinv.m[1][2] = k5;
  Jinv.m[2][0] = k3; Jinv.m[2][1] = k5; Jinv.m[2][2] = k6;
}

// AFTER_PROPAGATE: called with the frame scratch once Propagate and Atmosphere are done (the env layer publishes position / velocity to
// the other lanes' missiles from there); everything it does not consume dies at that point.
template <class AfterPropagate>
FDM_DEV void fdm_frame_lean(AcCore& a, Props& p, FcsState& s, FrameKeep& keep, const double* __restrict__ T, const AtmoConst& ac,
                            const double dt, const double fcs_dt, AfterPropagate after_propagate) {
  V3 gi;
  double twoVt, Mass, qbar;
  V3 cg, cg_prev;
  double tk0, tk1, tk2, tk3;
  WindAxes w;
  Atmo atm;
  {
    Frame f;
    fdm_stage_propagate(a, p, f, dt);
    fdm_stage_gravity(f);
    // Tec2i * gravity, formed here so that neither the ECEF gravity nor sin / cos(epa) outlive this block
    gi = v3(f.cos_epa * f.grav.x - f.sin_epa * f.grav.y, f.sin_epa * f.grav.x + f.cos_epa * f.grav.y, f.grav.z);
    fdm_stage_atmosphere(p, f, ac);
    after_propagate(f);           // position, velocities and h_asl of this frame are final
    f16_fcs(p, s, T, fcs_dt);
    tk0 = a.tank0; tk1 = a.tank1; tk2 = a.tank2; tk3 = a.tank3;
    cg_prev = a.cg;
    fdm_mass_cg(a, Mass, cg);
    a.cg = cg; f.cg = cg;
    fdm_stage_auxiliary(a, p, f, ac, w);
    twoVt = 2 * f.Vt; qbar = f.qbar; atm = f.atm;
    keep.pilot_nx = f.pilotN.x; keep.vcas = f.vcas; keep.beta = f.beta;
  }
  double thrust;
  {
    const int flags = (int)a.engflags;
    bool augmentation = flags & 2;
    thrust = fdm_stage_engine(a, p, atm, qbar, T, ac, dt, flags & 1, augmentation);
    const bool starved_next = fdm_stage_consume_fuel(a, dt, flags & 1, false);
    a.engflags = (double)((starved_next ? 1 : 0) | (augmentation ? 2 : 0));
  }
  keep.thrust = thrust;
  double c[6];
  f16_aero<63>(p, T, twoVt, c);
  {
    // FGAerodynamics wind->body + FGAircraft sums + FGAccelerations, as fdm_stage_accelerations
    const double sa = w.sa, ca = w.ca, sb = w.sb, cb = w.cb;
    const double fw0 = -c[0], fw1 = c[1], fw2 = -c[2];
    V3 Fa;
    Fa.x = (ca * cb) * fw0 + (-ca * sb) * fw1 + (-sa) * fw2;
    Fa.y = sb * fw0 + cb * fw1 + 0.0 * fw2;
    Fa.z = (sa * cb) * fw0 + (-sa * sb) * fw1 + ca * fw2;
    const V3 rp = structural_to_body(cg, K_AERORP_X, K_AERORP_Y, K_AERORP_Z);
    const V3 Ma = v3(c[3], c[4], c[5]) + cross(rp, Fa);
    const V3 Fp = v3(thrust, 0.0, 0.0);
    const V3 Mp = cross(structural_to_body(cg, K_THRUSTER_X, K_THRUSTER_Y, K_THRUSTER_Z), Fp);
    const V3 F = Fa + Fp, M = Ma + Mp;
    M33 J, Jinv;
    fdm_inertia(tk0, tk1, tk2, tk3, cg_prev, cg, J, Jinv);
    a.pqridot = mul(Jinv, M - cross(a.wi, mul(J, a.wi)));
    const double rm = fm_rcp(Mass);
    a.bodyaccel = v3(F.x * rm, F.y * rm, F.z * rm);
    const M33 Ti2b = quat_T(opaque(a.q0), opaque(a.q1), opaque(a.q2), opaque(a.q3));
    a.uvwidot = mulT(Ti2b, a.bodyaccel) + gi;
  }
}

// The frame scratch the outputs and the publication need, recomputed from the state a frame left behind (position,
// attitude and velocities do not change after Propagate; dt = 0 evaluates the derived quantities without integrating).
FDM_DEV void fdm_refresh(const AcCore& a, const Props& p, const FrameKeep& keep, Frame& f) {
  AcCore t = a;                 // fdm_propagate_pos adds EARTH_OMEGA * 0 to epa: exact
  Props pp = p;
  fdm_propagate_pos(t, f, t.vi, 0.0);
  fdm_propagate_combine(t, pp, f, t.ri.x, t.ri.y);
  const double ecr = EARTH_B / EARTH_A;
  const double slr = (EARTH_A * ecr) * fm_rsqrt(1.0 - (1.0 - ecr * ecr) * f.cosLatGc * f.cosLatGc);
  f.h_asl = f.radius - slr;
  f.vcas = keep.vcas; f.pilotN = v3(keep.pilot_nx, p.accelerations_n_pilot_y_norm, p.accelerations_n_pilot_z_norm);
  f.alpha = p.aero_alpha_rad; f.beta = keep.beta; f.mach = p.velocities_mach; f.thrust = keep.thrust;
}

// outputs of the frame that has just run (what the reference reads back through get_property_value)
FDM_DEV void fdm_outputs(const AcCore& a, const Frame& f, AcOut& o) {
  const double lon = (f.rxy == 0.0) ? 0.0 : fm_atan2(f.ecef.y, f.ecef.x);
  o.lon_deg = lon * RADTODEG;
  const double sgn = f.ecef.z < 0.0 ? -1.0 : 1.0;
  o.lat_geod_deg = (sgn * fm_atan2(f.gd_s1, f.gd_cc)) * RADTODEG;      // atan(s1 / cc), both >= 0
  o.h_sl_ft = f.h_asl;
  {  // Euler angles through the local quaternion, as FGPropagate::GetEuler does (qAttitudeLocal = Tl2b.GetQuaternion())
    double ql[4];
    mat_quat(f.Tl2b, ql);
    const M33 mT = quat_T(ql[0], ql[1], ql[2], ql[3]);
    mat_euler(mT, o.roll, o.pitch, o.heading);
  }
  o.vn = f.vel.x; o.ve = f.vel.y; o.vd = f.vel.z; o.u = f.uvw.x; o.v = f.uvw.y; o.w = f.uvw.z;
  o.vc_fps = f.vcas; o.npx = f.pilotN.x; o.npy = f.pilotN.y; o.npz = f.pilotN.z;
  o.p = f.pqr.x; o.q = f.pqr.y; o.r = f.pqr.z; o.eci_vmag = mag(a.vi);
  o.geod_alt_ft = f.geodAlt; o.alpha = f.alpha; o.beta = f.beta; o.mach = f.mach; o.thrust = f.thrust;
}

// ============================================================== reset: FGFDMExec load + IC + RunIC + engine start
// Follows AircraftSimulator.reload (reference envs/JSBSim/core/simulatior.py:152-190); see oracle F16::reset.
struct IcParams { double lon_deg, lat_geod_deg, h_sl_ft, psi_deg, u, v, w, p, q, r, phi_deg, theta_deg; };

FDM_DEV void fdm_reset(AcCore& a, Props& p, FcsState& s, Frame& f, const double* __restrict__ T, const AtmoConst& ac,
                       const IcParams& ic, const double fcs_dt) {
  f16_props_init(p, s);
  a.N1 = a.N2 = a.N2norm = a.FF = 0.0; a.engflags = 0.0;
  a.tank0 = K_TANK0_CONTENTS; a.tank1 = K_TANK1_CONTENTS; a.tank2 = K_TANK2_CONTENTS; a.tank3 = K_TANK3_CONTENTS;
  a.cg = v3(0, 0, 0); a.sim_time = 0.0; a.epa = 0.0;
  a.pqridot = a.uvwidot = a.bodyaccel = v3(0, 0, 0);
  // geodetic altitude such that |r| - SLR == h (J/initialization/FGInitialCondition.cpp:749-815, setgeod branch)
  const double lon = ic.lon_deg * DEGTORAD, geodLatitude = ic.lat_geod_deg * DEGTORAD, alt = ic.h_sl_ft;
  const double aa = EARTH_A, bb = EARTH_B, e2 = 1.0 - bb * bb / (aa * aa);
  double cosGeodLat, sinGeodLat;
  sincos(geodLatitude, &sinGeodLat, &cosGeodLat);
  const double N = aa / sqrt(1 - e2 * sinGeodLat * sinGeodLat);
  double geodAlt, n = e2, prev_n = 1.0;
  int iter = 0;
  if (cosGeodLat > fabs(sinGeodLat)) {
    const double tanGeodLat = sinGeodLat / cosGeodLat, x0 = N * e2 * cosGeodLat;
    double x = 0.0;
    while (fabs(n - prev_n) > 1E-15 && iter < 10) {
      const double tanLat = (1 - n) * tanGeodLat, cos2Lat = 1. / (1. + tanLat * tanLat);
      const double slr = bb / sqrt(1. - e2 * cos2Lat), R = slr + alt;
      x = R * sqrt(cos2Lat); prev_n = n; n = x0 / x; iter++;
    }
    geodAlt = x / cosGeodLat - N;
  } else {
    const double cotanGeodLat = cosGeodLat / sinGeodLat, z0 = N * e2 * sinGeodLat;
    double z = 0.0;
    while (fabs(n - prev_n) > 1E-15 && iter < 10) {
      const double cotanLat = cotanGeodLat / (1 - n), sin2Lat = 1. / (1. + cotanLat * cotanLat), cos2Lat = 1. - sin2Lat;
      const double slr = bb / sqrt(1. - e2 * cos2Lat), R = slr + alt;
      z = R * (cotanLat < 0.0 ? -1.0 : 1.0) * sqrt(sin2Lat); prev_n = n; n = z0 / (z0 + z); iter++;
    }
    geodAlt = z / sinGeodLat - N * (1 - e2);
  }
  {  // FGLocation::SetPositionGeodetic (J/math/FGLocation.cpp:247-258); epa = 0 => ECI == ECEF
    const double RN = aa / sqrt(1.0 - e2 * sinGeodLat * sinGeodLat);
    double sl, cl;
    sincos(lon, &sl, &cl);
    a.ri = v3((RN + geodAlt) * cosGeodLat * cl, (RN + geodAlt) * cosGeodLat * sl, ((1 - e2) * RN + geodAlt) * sinGeodLat);
  }
  // FGPropagate::SetInitialState (J/models/FGPropagate.cpp:143-186)
  f.sin_epa = 0.0; f.cos_epa = 1.0;
  f.ecef = a.ri;
  M33 Tec2l;
  location_derived(f, Tec2l);      // Ti2l == Tec2l at epa = 0
  double qo[4];
  {  // FGQuaternion(phi, tht, psi) (J/math/FGQuaternion.cpp:106-133)
    double sth, cth, sps, cps, sph, cph;
    sincos(0.5 * ic.theta_deg * DEGTORAD, &sth, &cth);
    sincos(0.5 * ic.psi_deg * DEGTORAD, &sps, &cps);
    sincos(0.5 * ic.phi_deg * DEGTORAD, &sph, &cph);
    const double CC = cph * cth, CS = cph * sth, SS = sph * sth, SC = sph * cth;
    qo[0] = CC * cps + SS * sps; qo[1] = SC * cps - CS * sps; qo[2] = CS * cps + SC * sps; qo[3] = CC * sps - SS * cps;
    const double nn = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
    if (!(nn == 0.0 || fabs(nn - 1.000) < 1e-10)) { const double rn = 1.0 / nn; qo[0] *= rn; qo[1] *= rn; qo[2] *= rn; qo[3] *= rn; }
  }
  const M33 icTl2b = quat_T(qo[0], qo[1], qo[2], qo[3]);
  const V3 uvw_ned = mulT(icTl2b, v3(ic.u, ic.v, ic.w));
  const V3 icUVW = mul(icTl2b, uvw_ned);
  double qi[4];
  M33 Ti2ecI;  // identity rotation (epa = 0): Ti2l = Tec2l * I, evaluated with the same products as the oracle
  Ti2ecI.m[0][0] = 1.0; Ti2ecI.m[0][1] = 0.0; Ti2ecI.m[0][2] = 0.0; Ti2ecI.m[1][0] = -0.0; Ti2ecI.m[1][1] = 1.0; Ti2ecI.m[1][2] = 0.0;
  Ti2ecI.m[2][0] = 0.0; Ti2ecI.m[2][1] = 0.0; Ti2ecI.m[2][2] = 1.0;
  const M33 Ti2l = mul(Tec2l, Ti2ecI);
  mat_quat(Ti2l, qi);
  // qAttitudeECI = Ti2l.GetQuaternion() * qAttitudeLocal (Hamilton product)
  a.q0 = qi[0] * qo[0] - qi[1] * qo[1] - qi[2] * qo[2] - qi[3] * qo[3];
  a.q1 = qi[0] * qo[1] + qi[1] * qo[0] + qi[2] * qo[3] - qi[3] * qo[2];
  a.q2 = qi[0] * qo[2] - qi[1] * qo[3] + qi[2] * qo[0] + qi[3] * qo[1];
  a.q3 = qi[0] * qo[3] + qi[1] * qo[2] - qi[2] * qo[1] + qi[3] * qo[0];
  const M33 Ti2b = quat_T(a.q0, a.q1, a.q2, a.q3);
  const V3 Omega = v3(0.0, 0.0, EARTH_OMEGA);
  a.wi = v3(ic.p, ic.q, ic.r) + mul(Ti2b, Omega);
  a.vi = mulT(Ti2b, icUVW) + cross(Omega, a.ri);
  // RunIC: two suspended-integration passes (Initialize()'s Run + RunIC's Run), then InitializeDerivatives
  // (one copy of the frame code, run twice: the reset kernel executes cold, its cost is instruction fetch)
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) fdm_frame(a, p, s, f, T, ac, 0.0, fcs_dt, false);
  a.dqv0 = a.vi; a.dqv1 = a.vi; a.dqa0 = a.uvwidot;
  // engine.init_running(): N2 = IdleN2 + ThrottlePos*N2_factor with the throttle still at its IC value
  {
    double ThrottlePos = p.fcs_throttle_pos_norm;
    if (ThrottlePos > 1.0) ThrottlePos = 1.0;
    // NOTE FGTurbine::InitRunning uses the member ThrottlePos latched by the last Calculate()
    a.N2 = K_ENG_idlen2 + ThrottlePos * (K_ENG_maxn2 - K_ENG_idlen2);
    a.N1 = K_ENG_idlen1 + ThrottlePos * (K_ENG_maxn1 - K_ENG_idlen1);
  }
  // propulsion.get_steady_state(): march the running engine with dt = 0.5 s until the thrust is steady
  {
    double idleT, milT, augT;
    f16_engine_tables(p, T, idleT, milT, augT);
    double ThrottlePos = p.fcs_throttle_pos_norm, AugmentCmd;
    if (ThrottlePos > 1.0) { AugmentCmd = ThrottlePos - 1.0; ThrottlePos -= AugmentCmd; } else AugmentCmd = 0.0;
    const double tdt = 0.5;
    auto seek = [tdt](double v, double target, double accel, double decel) {
      if (v > target) { v -= tdt * decel; if (v < target) v = target; }
      else if (v < target) { v += tdt * accel; if (v > target) v = target; }
      return v;
    };
    const double N1_factor = K_ENG_maxn1 - K_ENG_idlen1, N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
    const double idlethrust = K_ENG_milthrust * idleT, milthrust = (K_ENG_milthrust - idlethrust) * milT;
    const double sigma = f.atm.rho / ac.SLdensity, dbase = 90.0 / (K_ENG_bypassratio + 3.0);
    const double IdleFF = K_ENG_idleff;
    bool augmentation = false;
    double currentThrust = 0, lastThrust = -1;
    int steady_count = 0, j = 0;
    bool steady = false, pAug = false;
    double pN1 = -1.0, pN2 = -1.0, pFF = -1.0;
    while (!steady && j < 6000) {
      const double n0 = a.N2norm + 0.1, n = n0 < 1.0 ? n0 : 1.0;   // fmin without its NaN handling
      const double sden = (1 + 3 * (1 - n) * (1 - n) * (1 - n) + (1 - sigma));
      const double up = (1.0 * dbase) / sden, dn2 = (3.0 * dbase) / sden, dn1 = (2.4 * dbase) / sden;
      a.N2 = seek(a.N2, K_ENG_idlen2 + ThrottlePos * N2_factor, up, dn2);
      a.N1 = seek(a.N1, K_ENG_idlen1 + ThrottlePos * N1_factor, up, dn1);
      a.N2norm = (a.N2 - K_ENG_idlen2) / N2_factor;
      double thrust = idlethrust + (milthrust * a.N2norm * a.N2norm);
      if (!augmentation) {
        const double tsfc = K_ENG_tsfc * sqrt(f.atm.T / 389.7) * (0.84 + (1 - a.N2norm) * (1 - a.N2norm));
        a.FF = seek(a.FF, thrust * tsfc, 1000.0, 10000.0);
        if (a.FF < IdleFF) a.FF = IdleFF;
      }
      if (AugmentCmd > 0.0) {
        augmentation = true;
        const double tdiff = (K_ENG_maxthrust * augT) - thrust;
        thrust += (tdiff * AugmentCmd);
        a.FF = seek(a.FF, thrust * K_ENG_atsfc, 5000.0, 10000.0);
      } else augmentation = false;
      lastThrust = currentThrust; currentThrust = thrust;
      if (fabs(lastThrust - currentThrust) < 0.0001) { steady_count++; if (steady_count > 120) steady = true; }
      else steady_count = 0;
      j++;
      // Exact shortcut: once an iteration leaves the whole engine state bit-identical, every further iteration is the
      // same map on the same state, so the reference's count to 120 equal thrusts ends with exactly this state.
      if (a.N1 == pN1 && a.N2 == pN2 && a.FF == pFF && augmentation == pAug && lastThrust == currentThrust) break;
      pN1 = a.N1; pN2 = a.N2; pFF = a.FF; pAug = augmentation;
    }
    f.thrust = currentThrust;
    a.engflags = (double)(augmentation ? 2 : 0);
  }
}
