// =====================================================================================
// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
// executed by the product path (aircombat_selfplay_b200/); only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY UNPINNED (except the atmosphere model, which reproduces the real-JSBSim reference data of
// envs/JSBSim/data/tests/TestDensityAltitude.py / TestPressureAltitude.py: tests/test_oracle_fdm.py): the reference's FDM is the pip wheel jsbsim==1.1.6, which is not installable
// here (no network), and the vendored JSBSim sources under /root/reference/envs/JSBSim/data/src
// ship without headers, so the real FDM cannot be built or run in this container.  This file is a
// scalar fp64 restatement of exactly the code path the reference exercises for the F-16
// (sim_freq = 60 Hz), following the vendored sources file by file; each block cites the file:line
// it follows.  "J/" below = /root/reference/envs/JSBSim/data/src/.
//
// Execution model (deliberately different from the CUDA kernels): JSBSim-like.  A flat property
// array P[] stands in for the property tree; the flight-control components and aerodynamic
// coefficient functions are *interpreted* from the IR emitted by modelc/gen_oracle.py
// (oracle/gen/f16_ir.inc), the way FGFCS / FGAerodynamics walk their component lists; the core
// models publish their tied properties into P[] when they run.  The CUDA side instead compiles the
// same IR to straight-line code, so a parity failure isolates either back end.
// =====================================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <algorithm>
#include <string>
#include <vector>

namespace orc {

// ---------------------------------------------------------------- IR structs (filled by gen/f16_ir.inc)
struct Table { int dim; int row_var, col_var; int nrows, ncols; const double* rk; const double* ck; const double* v; };
struct ValRef { int is_prop; double value; int prop; double sign; };
struct Cond { int prop; int op; ValRef rhs; };
struct Test { int logic; ValRef value; int ncond; const Cond* conds; };
struct Factor { int kind; int prop; double value; const Table* table; };
struct Comp {
  int type; const char* name; int n_in; int in_prop[4]; double in_sign[4]; int n_out; int out_prop[3];
  int has_clip; double clip_min, clip_max; double gain; const Table* table;
  double in_min, in_max, out_min, out_max; int zero_centered; double bias; double kp, ki, kd; int int_type;
  int has_trigger; int trig_prop; double trig_sign; int noscale; int ndet; const double* detents; const double* times;
  ValRef defval; int ntests; const Test* tests; int nfac; const Factor* factors;
};
struct AeroFn { const char* name; int axis_or_prop; int nfac; const Factor* factors; };
enum { C_SWITCH = 0, C_PURE_GAIN, C_SCHED_GAIN, C_AEROSURF, C_SUMMER, C_PID, C_KINEMATIC, C_FCSFUNC };

#include "gen/f16_ir.inc"

// ---------------------------------------------------------------- unit constants (FGJSBBase.h, header-only upstream)
static const double radtodeg = 180.0 / M_PI, degtorad = M_PI / 180.0;
static const double fttom = 0.3048, inchtoft = 1.0 / 12.0;
static const double slugtolb = 32.174049, lbtoslug = 1.0 / slugtolb;
static const double kgtoslug = 0.06852168;
static const double ktstofps = 1.68781, fpstokts = 1.0 / ktstofps;

// ---------------------------------------------------------------- tiny linear algebra
struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline V3 operator/(const V3& a, double s) { double t = 1.0 / s; return V3(a.x * t, a.y * t, a.z * t); }  // FGColumnVector3::operator/ multiplies by the reciprocal (J/math/FGColumnVector3.cpp:92-98)
// FGColumnVector3 operator*(V3,V3) is the CROSS product (J/math/FGColumnVector3.h, header-only)
static inline V3 cross(const V3& a, const V3& b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double mag(const V3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

struct M33 {
  double m[3][3];
  M33() { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r][c] = 0.0; }
  M33(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    m[0][0] = a; m[0][1] = b; m[0][2] = c; m[1][0] = d; m[1][1] = e; m[1][2] = f; m[2][0] = g; m[2][1] = h; m[2][2] = i;
  }
};
static inline M33 T(const M33& a) { M33 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// J/math/FGMatrix33.cpp:380 operator*(M33): plain row-by-column sums, left-to-right
static inline M33 operator*(const M33& a, const M33& b) {
  M33 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
static inline V3 operator*(const M33& a, const V3& v) {
  return V3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
static inline M33 operator+(const M33& a, const M33& b) { M33 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }

struct Quat { double q[4]; };
// J/math/FGQuaternion.cpp:170-182
static inline void quat_normalize(Quat& q) {
  double n = std::sqrt(q.q[0] * q.q[0] + q.q[1] * q.q[1] + q.q[2] * q.q[2] + q.q[3] * q.q[3]);
  if (n == 0.0 || std::fabs(n - 1.000) < 1e-10) return;
  double rn = 1.0 / n;
  for (int i = 0; i < 4; i++) q.q[i] *= rn;
}
// J/math/FGQuaternion.cpp:106-133 InitializeFromEulerAngles
static inline Quat quat_from_euler(double phi, double tht, double psi) {
  double thtd2 = 0.5 * tht, psid2 = 0.5 * psi, phid2 = 0.5 * phi;
  double Sthtd2 = std::sin(thtd2), Spsid2 = std::sin(psid2), Sphid2 = std::sin(phid2);
  double Cthtd2 = std::cos(thtd2), Cpsid2 = std::cos(psid2), Cphid2 = std::cos(phid2);
  double CC = Cphid2 * Cthtd2, CS = Cphid2 * Sthtd2, SS = Sphid2 * Sthtd2, SC = Sphid2 * Cthtd2;
  Quat q;
  q.q[0] = CC * Cpsid2 + SS * Spsid2;
  q.q[1] = SC * Cpsid2 - CS * Spsid2;
  q.q[2] = CS * Cpsid2 + SC * Spsid2;
  q.q[3] = CC * Spsid2 - SS * Cpsid2;
  quat_normalize(q);
  return q;
}
// FGQuaternion::operator* (header-only upstream): Hamilton product
static inline Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.q[0] = a.q[0] * b.q[0] - a.q[1] * b.q[1] - a.q[2] * b.q[2] - a.q[3] * b.q[3];
  r.q[1] = a.q[0] * b.q[1] + a.q[1] * b.q[0] + a.q[2] * b.q[3] - a.q[3] * b.q[2];
  r.q[2] = a.q[0] * b.q[2] - a.q[1] * b.q[3] + a.q[2] * b.q[0] + a.q[3] * b.q[1];
  r.q[3] = a.q[0] * b.q[3] + a.q[1] * b.q[2] - a.q[2] * b.q[1] + a.q[3] * b.q[0];
  return r;
}
// J/math/FGQuaternion.cpp:187-216 (transformation matrix part of ComputeDerivedUnconditional)
static inline M33 quat_T(const Quat& Q) {
  double q0 = Q.q[0], q1 = Q.q[1], q2 = Q.q[2], q3 = Q.q[3];
  double q0q0 = q0 * q0, q1q1 = q1 * q1, q2q2 = q2 * q2, q3q3 = q3 * q3;
  double q0q1 = q0 * q1, q0q2 = q0 * q2, q0q3 = q0 * q3, q1q2 = q1 * q2, q1q3 = q1 * q3, q2q3 = q2 * q3;
  return M33(q0q0 + q1q1 - q2q2 - q3q3, 2.0 * (q1q2 + q0q3), 2.0 * (q1q3 - q0q2),
             2.0 * (q1q2 - q0q3), q0q0 - q1q1 + q2q2 - q3q3, 2.0 * (q2q3 + q0q1),
             2.0 * (q1q3 + q0q2), 2.0 * (q2q3 - q0q1), q0q0 - q1q1 - q2q2 + q3q3);
}
// J/math/FGQuaternion.cpp:158-166 GetQDot
static inline Quat quat_dot(const Quat& Q, const V3& w) {
  Quat r;
  r.q[0] = -0.5 * (Q.q[1] * w.x + Q.q[2] * w.y + Q.q[3] * w.z);
  r.q[1] = 0.5 * (Q.q[0] * w.x - Q.q[3] * w.y + Q.q[2] * w.z);
  r.q[2] = 0.5 * (Q.q[3] * w.x + Q.q[0] * w.y - Q.q[1] * w.z);
  r.q[3] = 0.5 * (-Q.q[2] * w.x + Q.q[1] * w.y + Q.q[0] * w.z);
  return r;
}
// J/math/FGMatrix33.cpp:106-154 GetQuaternion (data[] there is column-major: data[3]=m12, data[1]=m21 ...)
static inline Quat mat_quat(const M33& a) {
  const double m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2], m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2],
               m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
  double t[4] = {1.0 + m11 + m22 + m33, 1.0 + m11 - m22 - m33, 1.0 - m11 + m22 - m33, 1.0 - m11 - m22 + m33};
  int idx = 0;
  for (int i = 1; i < 4; i++) if (t[i] > t[idx]) idx = i;
  Quat Q; Q.q[0] = 1; Q.q[1] = Q.q[2] = Q.q[3] = 0;
  switch (idx) {
    case 0: Q.q[0] = 0.50 * std::sqrt(t[0]); Q.q[1] = 0.25 * (m23 - m32) / Q.q[0]; Q.q[2] = 0.25 * (m31 - m13) / Q.q[0]; Q.q[3] = 0.25 * (m12 - m21) / Q.q[0]; break;
    case 1: Q.q[1] = 0.50 * std::sqrt(t[1]); Q.q[0] = 0.25 * (m23 - m32) / Q.q[1]; Q.q[2] = 0.25 * (m12 + m21) / Q.q[1]; Q.q[3] = 0.25 * (m31 + m13) / Q.q[1]; break;
    case 2: Q.q[2] = 0.50 * std::sqrt(t[2]); Q.q[0] = 0.25 * (m31 - m13) / Q.q[2]; Q.q[1] = 0.25 * (m12 + m21) / Q.q[2]; Q.q[3] = 0.25 * (m23 + m32) / Q.q[2]; break;
    case 3: Q.q[3] = 0.50 * std::sqrt(t[3]); Q.q[0] = 0.25 * (m12 - m21) / Q.q[3]; Q.q[1] = 0.25 * (m13 + m31) / Q.q[3]; Q.q[2] = 0.25 * (m23 + m32) / Q.q[3]; break;
  }
  return Q;
}
// J/math/FGMatrix33.cpp:159-192 GetEuler -> (phi, theta, psi)
static inline V3 mat_euler(const M33& a) {
  V3 e; bool lock = false;
  double m13 = a.m[0][2];
  if (m13 <= -1.0) { e.y = 0.5 * M_PI; lock = true; }
  else if (1.0 <= m13) { e.y = -0.5 * M_PI; lock = true; }
  else e.y = std::asin(-m13);
  if (lock) e.x = std::atan2(-a.m[2][1], a.m[1][1]);
  else e.x = std::atan2(a.m[1][2], a.m[2][2]);
  if (lock) e.z = 0.0;
  else { double psi = std::atan2(a.m[0][1], a.m[0][0]); if (psi < 0.0) psi += 2 * M_PI; e.z = psi; }
  return e;
}

static inline double Constrain(double lo, double v, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
// FGJSBBase::EqualToRoundoff (header-only upstream)
static inline bool EqualToRoundoff(double a, double b) {
  double eps = 2.0 * DBL_EPSILON;
  return std::fabs(a - b) <= eps * std::max(std::fabs(a), std::fabs(b));
}
static inline double sign(double x) { return x < 0.0 ? -1.0 : 1.0; }

// ---------------------------------------------------------------- FGTable lookups (J/math/FGTable.cpp:443-517)
// The upstream search starts from the previous row index; the bracketing row it ends on is the same
// wherever it starts except when the key equals a breakpoint exactly, where the two candidates give the
// same value up to one rounding.  The stateless rule used here (and in the CUDA code): the first row r>=2
// whose key is >= the lookup key (what upstream finds on a fresh table, lastRowIndex = 2).
static double table1(const Table* t, double key) {
  const int n = t->nrows;
  if (key <= t->rk[0]) return t->v[0];
  if (key >= t->rk[n - 1]) return t->v[n - 1];
  int r = 1;  // 0-based index of the upper row
  while (r < n - 1 && t->rk[r] < key) r++;
  double span = t->rk[r] - t->rk[r - 1], factor;
  if (span != 0.0) { factor = (key - t->rk[r - 1]) / span; if (factor > 1.0) factor = 1.0; }
  else factor = 1.0;
  return factor * (t->v[r] - t->v[r - 1]) + t->v[r - 1];
}
static double table2(const Table* t, double rowKey, double colKey) {
  const int nr = t->nrows, nc = t->ncols;
  int r = 1, c = 1;
  while (r < nr - 1 && t->rk[r] < rowKey) r++;
  while (c < nc - 1 && t->ck[c] < colKey) c++;
  double rF = (rowKey - t->rk[r - 1]) / (t->rk[r] - t->rk[r - 1]);
  double cF = (colKey - t->ck[c - 1]) / (t->ck[c] - t->ck[c - 1]);
  if (rF > 1.0) rF = 1.0; else if (rF < 0.0) rF = 0.0;
  if (cF > 1.0) cF = 1.0; else if (cF < 0.0) cF = 0.0;
#define TV(rr, cc) t->v[(rr) * nc + (cc)]
  double col1 = rF * (TV(r, c - 1) - TV(r - 1, c - 1)) + TV(r - 1, c - 1);
  double col2 = rF * (TV(r, c) - TV(r - 1, c)) + TV(r - 1, c);
#undef TV
  return col1 + cF * (col2 - col1);
}

// ---------------------------------------------------------------- ISA-1976 (J/models/FGAtmosphere.cpp, atmosphere/FGStandardAtmosphere.cpp)
struct StdAtmosphere {
  // header constants (FGAtmosphere.h / FGStandardAtmosphere.h upstream; pinned by data/tests/TestStdAtmosphere.py:49-70)
  double Rstar, Mair, g0, Reng, Rdry, SHRatio, EarthRadius;
  double StdDaySLtemperature, StdDaySLpressure, StdDaySLsoundspeed;
  static const int NR = 9;
  double H[NR], Tt[NR];
  double LapseRates[NR - 1], PressureBreakpoints[NR], StdDensityBreakpoints[NR];
  double StdLapseRates[NR - 1], StdPressureBreakpoints[NR];   // frozen at construction (:118-135); the inverse functions use these
  double TemperatureBias;                                     // atmosphere/delta-T (:343-353); the reference never sets it: 0
  double SLdensity, SLpressure, SLtemperature, SLsoundspeed;
  // outputs of Calculate()
  double Temperature, Pressure, Density, Soundspeed, DensityAltitude, PressureAltitude, Viscosity, KinematicViscosity;

  StdAtmosphere() {
    Rstar = 8.31432 * kgtoslug / (1.8 * (fttom * fttom));  // KelvinToRankine(x) = 1.8*x
    Mair = 28.9645 * kgtoslug / 1000.0;
    g0 = 9.80665 / fttom;
    Reng = Rstar / Mair; Rdry = Rstar / Mair;
    SHRatio = 1.4;
    EarthRadius = 6356766.0 / fttom;
    StdDaySLtemperature = 518.67; StdDaySLpressure = 2116.228;
    StdDaySLsoundspeed = std::sqrt(SHRatio * Reng * StdDaySLtemperature);
    // J/models/atmosphere/FGStandardAtmosphere.cpp:91-99 (geopotential ft, deg R)
    const double h[NR] = {0.0000, 36089.2388, 65616.7979, 104986.8766, 154199.4751, 167322.8346, 232939.6325, 278385.8268, 298556.4304};
    const double t[NR] = {518.67, 389.97, 389.97, 411.57, 487.17, 487.17, 386.37, 336.5028, 336.5028};
    for (int i = 0; i < NR; i++) { H[i] = h[i]; Tt[i] = t[i]; }
    // CalculateLapseRates (:405-417), TemperatureDeltaGradient = 0
    TemperatureBias = 0.0;
    for (int b = 0; b < NR - 1; b++) { LapseRates[b] = (Tt[b + 1] - Tt[b]) / (H[b + 1] - H[b]) - 0.0; StdLapseRates[b] = LapseRates[b]; }
    CalculatePressureBreakpoints(StdDaySLpressure);
    for (int i = 0; i < NR; i++) StdPressureBreakpoints[i] = PressureBreakpoints[i];
    for (int i = 0; i < NR; i++) StdDensityBreakpoints[i] = StdPressureBreakpoints[i] / (Rdry * Tt[i]);  // :457-462
    SLtemperature = Tt[0]; SLpressure = StdDaySLpressure; SLdensity = SLpressure / (Rdry * SLtemperature);
    SLsoundspeed = std::sqrt(SHRatio * Rdry * SLtemperature);
    Calculate(0.0);
  }
  // CalculatePressureBreakpoints (:419-440); TemperatureDeltaGradient = 0
  void CalculatePressureBreakpoints(double SLpress) {
    PressureBreakpoints[0] = SLpress;
    for (int b = 0; b < NR - 1; b++) {
      double BaseTemp = Tt[b], deltaH = H[b + 1] - H[b], Tmb = BaseTemp + TemperatureBias + (H[NR - 1] - H[b]) * 0.0;
      if (LapseRates[b] != 0.00) {
        double Lmb = LapseRates[b], Exp = g0 / (Rdry * Lmb), factor = Tmb / (Tmb + Lmb * deltaH);
        PressureBreakpoints[b + 1] = PressureBreakpoints[b] * std::pow(factor, Exp);
      } else PressureBreakpoints[b + 1] = PressureBreakpoints[b] * std::exp(-g0 * deltaH / (Rdry * Tmb));
    }
  }
  // SetTemperatureBias(eRankine, t) (:343-353) -- what writing the property atmosphere/delta-T does.  Only the known-answer
  // tests of JSBSim's own suite use it (tests/test_oracle_fdm.py); the F-16 path keeps the bias at 0.
  void SetTemperatureBias(double t_R) {
    TemperatureBias = t_R;
    CalculatePressureBreakpoints(SLpressure);
    SLtemperature = GetTemperature(0.0);
    SLsoundspeed = std::sqrt(SHRatio * Reng * SLtemperature); SLdensity = SLpressure / (Reng * SLtemperature);
  }
  double GeopotentialAltitude(double h) const { return (h * EarthRadius) / (EarthRadius + h); }
  double GeometricAltitude(double H_) const { return (H_ * EarthRadius) / (EarthRadius - H_); }
  double TempTable(double key) const {  // FGTable 1-D on the temperature table
    if (key <= H[0]) return Tt[0];
    if (key >= H[NR - 1]) return Tt[NR - 1];
    int r = 1; while (r < NR - 1 && H[r] < key) r++;
    double span = H[r] - H[r - 1], f;
    if (span != 0.0) { f = (key - H[r - 1]) / span; if (f > 1.0) f = 1.0; } else f = 1.0;
    return f * (Tt[r] - Tt[r - 1]) + Tt[r - 1];
  }
  // :244-271 GetTemperature (gradient = 0)
  double GetTemperature(double altitude) const {
    double GeoPotAlt = GeopotentialAltitude(altitude), Tm;
    if (GeoPotAlt >= 0.0) Tm = TempTable(GeoPotAlt);
    else Tm = TempTable(0.0) + GeoPotAlt * LapseRates[0];
    Tm += TemperatureBias;   // :263
    return Tm;
  }
  // :191-227 GetPressure
  double GetPressure(double altitude) const {
    double GeoPotAlt = GeopotentialAltitude(altitude);
    double BaseAlt = H[0];
    int b;
    for (b = 0; b < NR - 2; ++b) { double testAlt = H[b + 1]; if (GeoPotAlt < testAlt) break; BaseAlt = testAlt; }
    double Tmb = GetTemperature(GeometricAltitude(BaseAlt));
    double deltaH = GeoPotAlt - BaseAlt, Lmb = LapseRates[b];
    if (Lmb != 0.0) { double Exp = g0 / (Rdry * Lmb), factor = Tmb / (Tmb + Lmb * deltaH); return PressureBreakpoints[b] * std::pow(factor, Exp); }
    return PressureBreakpoints[b] * std::exp(-g0 * deltaH / (Rdry * Tmb));
  }
  double GetDensity(double altitude) const { return GetPressure(altitude) / (Reng * GetTemperature(altitude)); }
  double GetSoundSpeed(double altitude) const { return std::sqrt(SHRatio * Reng * GetTemperature(altitude)); }
  // :464-492 CalculateDensityAltitude
  double CalculateDensityAltitude(double density) const {
    int b = 0;
    for (; b < NR - 2; b++) if (density >= StdDensityBreakpoints[b + 1]) break;
    double Tmb = Tt[b], Hb = H[b], Lmb = StdLapseRates[b], pb = StdDensityBreakpoints[b], da;
    if (Lmb != 0.0) { double Exp = -1.0 / (1.0 + g0 / (Rdry * Lmb)); da = Hb + (Tmb / Lmb) * (std::pow(density / pb, Exp) - 1); }
    else { double Factor = -Rdry * Tmb / g0; da = Hb + Factor * std::log(density / pb); }
    return GeometricAltitude(da);
  }
  double CalculatePressureAltitude(double pressure) const {
    int b = 0;
    for (; b < NR - 2; b++) if (pressure >= StdPressureBreakpoints[b + 1]) break;
    double Tmb = Tt[b], Hb = H[b], Lmb = StdLapseRates[b], Pb = StdPressureBreakpoints[b], pa;
    if (Lmb != 0.00) { double Exp = -Rdry * Lmb / g0; pa = Hb + (Tmb / Lmb) * (std::pow(pressure / Pb, Exp) - 1); }
    else { double Factor = -Rdry * Tmb / g0; pa = Hb + Factor * std::log(pressure / Pb); }
    return GeometricAltitude(pa);
  }
  // J/models/FGAtmosphere.cpp:107-131 Calculate (no overrides); vapour fraction stays 0 => Reng = Rdry
  void Calculate(double altitude) {
    Temperature = GetTemperature(altitude);
    Pressure = GetPressure(altitude);
    Density = GetDensity(altitude);
    Soundspeed = std::sqrt(SHRatio * Reng * Temperature);
    PressureAltitude = CalculatePressureAltitude(Pressure);
    DensityAltitude = CalculateDensityAltitude(Density);
    Viscosity = 2.269690E-08 * std::pow(Temperature, 1.5) / (198.72 + Temperature);
    KinematicViscosity = Viscosity / Density;
  }
};

// J/FGJSBBase.cpp:245-296
static double PitotTotalPressure(double mach, double p) {
  if (mach < 0) return p;
  if (mach < 1) return p * std::pow((1 + 0.2 * mach * mach), 3.5);
  return p * 166.92158009316827 * std::pow(mach, 7.0) / std::pow(7 * mach * mach - 1, 2.5);
}
static double MachFromImpactPressure(double qc, double p) {
  double A = qc / p + 1;
  double M = std::sqrt(5.0 * (std::pow(A, 1. / 3.5) - 1));
  if (M > 1.0) for (unsigned int i = 0; i < 10; i++) M = 0.8812848543473311 * std::sqrt(A * std::pow(1 - 1.0 / (7.0 * M * M), 2.5));
  return M;
}

// ---------------------------------------------------------------- FGLocation (J/math/FGLocation.cpp)
struct Location {
  V3 ec;            // ECEF, ft
  double a, ecc, ec2, e2, c;  // ellipse (SetEllipse :262-271): ec=b/a, ec2, e2=1-ec2, c=a*e2
  // derived
  double lon, lat, radius, geodLat, geodAlt;
  M33 Tec2l, Tl2ec;
  void SetEllipse(double semimajor, double semiminor) { a = semimajor; ecc = semiminor / a; ec2 = ecc * ecc; e2 = 1.0 - ec2; c = a * e2; }
  // :247-258
  void SetPositionGeodetic(double lon_, double lat_, double height) {
    double slat = std::sin(lat_), clat = std::cos(lat_);
    double RN = a / std::sqrt(1.0 - e2 * slat * slat);
    ec.x = (RN + height) * clat * std::cos(lon_);
    ec.y = (RN + height) * clat * std::sin(lon_);
    ec.z = ((1 - e2) * RN + height) * slat;
    ComputeDerived();
  }
  void SetRadius(double r) { double rold = mag(ec); if (rold == 0.0) ec.x = r; else { double s = r / rold; ec = V3(ec.x * s, ec.y * s, ec.z * s); } ComputeDerived(); }
  // :273-279
  double GetSeaLevelRadius() const { double cosLat = std::cos(lat); return a * ecc / std::sqrt(1.0 - e2 * cosLat * cosLat); }
  // :283-370 ComputeDerivedUnconditional (ellipse set)
  void ComputeDerived() {
    radius = mag(ec);
    double rxy = std::sqrt(ec.x * ec.x + ec.y * ec.y);
    double sinLon, cosLon;
    if (rxy == 0.0) { sinLon = 0.0; cosLon = 1.0; lon = 0.0; }
    else { sinLon = ec.y / rxy; cosLon = ec.x / rxy; lon = std::atan2(ec.y, ec.x); }
    double sinLat, cosLat;
    if (radius == 0.0) { lat = 0.0; sinLat = 0.0; cosLat = 1.0; geodLat = 0.0; geodAlt = -a; }
    else {
      lat = std::atan2(ec.z, rxy);
      double s0 = std::fabs(ec.z), zc = ecc * s0, c0 = ecc * rxy, c02 = c0 * c0, s02 = s0 * s0, a02 = c02 + s02;
      double a0 = std::sqrt(a02), a03 = a02 * a0;
      double s1 = zc * a03 + c * s02 * s0, c1 = rxy * a03 - c * c02 * c0, cs0c0 = c * c0 * s0;
      double b0 = 1.5 * cs0c0 * ((rxy * s0 - zc * c0) * a0 - cs0c0);
      s1 = s1 * a03 - b0 * s0;
      double cc = ecc * (c1 * a03 - b0 * c0);
      geodLat = sign(ec.z) * std::atan(s1 / cc);
      double s12 = s1 * s1, cc2 = cc * cc, norm = std::sqrt(s12 + cc2);
      cosLat = cc / norm; sinLat = sign(ec.z) * s1 / norm;
      geodAlt = (rxy * cc + s0 * s1 - a * std::sqrt(ec2 * s12 + cc2)) / norm;
    }
    Tec2l = M33(-cosLon * sinLat, -sinLon * sinLat, cosLat, -sinLon, cosLon, 0.0, -cosLon * cosLat, -sinLon * cosLat, -sinLat);
    Tl2ec = T(Tec2l);
  }
};

// ---------------------------------------------------------------- the F-16 FDM
struct PidState { double Input_prev, Input_prev2, I_out_total; };

// ---- FCS component steps as free functions: the F-16's component list (fcs_run) and the known-answer tests ported from
// JSBSim's own test-suite (tests/test_oracle_fdm.py <- envs/JSBSim/data/tests/TestKinematic.py, TestIntegrators.py) run the
// SAME code.
// J/models/flight_control/FGPID.cpp:154-214 (non-"standard" form, no pvdot)
inline double pid_step(PidState& s, double Input, double test, double kp, double ki, double kd, int int_type, double dt) {
  double I_out_delta = 0.0;
  double Dval = (Input - s.Input_prev) / dt;
  if (std::fabs(test) < 0.000001) {
    switch (int_type) { case 1: I_out_delta = Input; break; case 2: I_out_delta = 0.5 * (Input + s.Input_prev); break;
      case 3: I_out_delta = 1.5 * Input - 0.5 * s.Input_prev; break; case 4: I_out_delta = (23.0 * Input - 16.0 * s.Input_prev + 5.0 * s.Input_prev2) / 12.0; break; default: I_out_delta = 0.0; }
  }
  if (test < 0.0) s.I_out_total = 0.0;
  s.I_out_total += ki * dt * I_out_delta;
  const double Output = kp * Input + s.I_out_total + kd * Dval;
  s.Input_prev2 = test < 0.0 ? 0.0 : s.Input_prev; s.Input_prev = Input;
  return Output;
}
// J/models/flight_control/FGKinemat.cpp:99-157; Input is the (signed) input property, Output the component's current output
inline double kinemat_step(const double* detents, const double* times, int ndet, bool noscale, bool trim_status, double Input, double Output, double dt) {
  double dt0 = dt;
  if (!noscale) Input *= detents[ndet - 1];
  Input = Constrain(detents[0], Input, detents[ndet - 1]);
  if (trim_status) return Input;
  while (dt0 > 0.0 && !EqualToRoundoff(Input, Output)) {
    int ind;
    for (ind = 1; ind < ndet && ((Input < Output) ? detents[ind] < Output : detents[ind] <= Output); ++ind) {}
    if (ind >= ndet) ind = ndet - 1;  // upstream reads past the end here; unreachable while Output stays inside the detents
    if (times[ind] <= 0.0) { Output = Input; break; }
    double Rate = (detents[ind] - detents[ind - 1]) / times[ind];
    double ThisInput = Constrain(detents[ind - 1], Input, detents[ind]);
    double ThisDt = std::fabs((ThisInput - Output) / Rate);
    if (dt0 < ThisDt) { ThisDt = dt0; if (Output < Input) Output += ThisDt * Rate; else Output -= ThisDt * Rate; }
    else Output = ThisInput;
    dt0 -= ThisDt;
  }
  return Output;
}

struct F16 {
  // ---- executive (J/FGFDMExec.cpp)
  double dT, saved_dT, sim_time;
  double fcs_dt;  // FGFCSComponent::dt, latched at load_model time (J/models/flight_control/FGFCSComponent.cpp:58):
                  // the reference calls set_dt() only AFTER load_model (core/simulatior.py:167-169), so the
                  // components keep FGFDMExec's constructor default 1/120 s (J/FGFDMExec.cpp:96).
  bool trim_status;
  // ---- planet (J/models/FGInertial.cpp:55-61)
  double GM, J2, a_ft, b_ft; V3 Omega; double gAccelReference;
  // ---- Propagate state (J/models/FGPropagate.cpp)
  Location loc;
  V3 vInertialPosition, vInertialVelocity, vPQRi, vPQR, vUVW, vVel;
  Quat qAttitudeECI, qAttitudeLocal, vQtrndot;
  double epa;
  V3 dqPQRidot[5], dqUVWidot[5], dqInertialVelocity[5]; Quat dqQtrndot[5];
  M33 Ti2ec, Tec2i, Tl2ec, Tec2l, Ti2l, Tl2i, Ti2b, Tb2i, Tl2b, Tb2l, Tec2b, Tb2ec;
  V3 euler; double sinEuler[3], cosEuler[3];
  // ---- Inertial
  V3 vGravAccel;
  // ---- Atmosphere
  StdAtmosphere atm;
  // ---- property array + FCS component state
  double P[N_PROPS];
  std::vector<PidState> pid; std::vector<double> compOutput;
  // ---- MassBalance
  double Weight, Mass, EmptyWeight; V3 vbaseXYZcg, vXYZcg, vLastXYZcg; M33 baseJ, mJ, mJinv;
  // ---- Auxiliary
  double alpha, beta, Vt, qbar, Mach, vcas, veas, Vground, pt, tat, tatc, hoverbmac, psigt, gamma_;
  V3 vAeroPQR, vAeroUVW, vPilotAccel, vPilotAccelN, vNcg, vEulerRates; M33 mTw2b, mTb2w;
  // ---- Propulsion / turbine / tanks
  double tank[8]; int ntanks;
  double N1, N2, N2norm, FuelFlow_pph, ThrottlePos, AugmentCmd, correctedTSFC, NozzlePosition, EPR, IdleFF, N1_factor, N2_factor;
  bool Running, Cutoff, Starter, Starved, Augmentation; int phase;  // tpOff=0,tpRun,tpSpinUp,tpStart,tpStall,tpSeize,tpTrim
  double idleThrustVal, milThrustVal, augThrustVal, Thrust, PropTotalDeltaT, FuelUsedLbs;
  V3 propForces, propMoments; M33 tankJ; double TanksWeight; V3 TanksMoment;
  // ---- Aerodynamics
  V3 aeroForces, aeroMoments, vFw; double clsq, bi2vel, ci2vel;
  // ---- Aircraft, Accelerations
  V3 acForces, acMoments, vPQRidot, vPQRdot, vUVWidot, vUVWdot, vBodyAccel;

  enum { tpOff = 0, tpRun, tpSpinUp, tpStart, tpStall, tpSeize, tpTrim };

  F16(double dt, double fcs_component_dt) {
    dT = dt; saved_dT = dt; sim_time = 0; fcs_dt = fcs_component_dt; trim_status = false;
    GM = 14.0764417572E15; J2 = 1.08262982E-03; a_ft = 20925646.32546; b_ft = 20855486.5951;
    Omega = V3(0, 0, 0.00007292115); gAccelReference = 9.80665 / fttom;
    pid.assign(N_FCS_COMPS, PidState{0, 0, 0}); compOutput.assign(N_FCS_COMPS, 0.0);
    ntanks = K_NTANKS;
    loc.SetEllipse(a_ft, b_ft);
  }

  // ============================================================== reset == FGFDMExec load + IC + RunIC
  // Follows AircraftSimulator.reload (reference envs/JSBSim/core/simulatior.py:152-190):
  // new FGFDMExec -> load_model -> set_dt -> default ICs + yaml init_state -> run_ic() ->
  // engine.init_running() -> propulsion.get_steady_state().
  void model_init() {
    // fresh FGFDMExec: every model at its constructor/InitModel state
    sim_time = 0.0; trim_status = false;
    for (int i = 0; i < N_PROPS; i++) P[i] = 0.0;
    P[P_gear_gear_cmd_norm] = 1.0; P[P_gear_gear_pos_norm] = 1.0;  // J/models/FGFCS.cpp:81 "default to gear down"
    P[P_metrics_Sw_sqft] = K_Sw; P[P_metrics_bw_ft] = K_bw; P[P_metrics_cbarw_ft] = K_cbarw;
    for (auto& s : pid) s = PidState{0, 0, 0};
    for (auto& o : compOutput) o = 0.0;
    // mass balance (J/models/FGMassBalance.cpp:96-128,130-178)
    double bixx = K_ixx, biyy = K_iyy, bizz = K_izz, bixy = K_ixy, bixz = K_ixz, biyz = K_iyz;
    if (!K_negated_crossproduct_inertia) baseJ = M33(bixx, bixy, -bixz, bixy, biyy, biyz, -bixz, biyz, bizz);
    else baseJ = M33(bixx, -bixy, bixz, -bixy, biyy, -biyz, bixz, -biyz, bizz);
    EmptyWeight = K_emptywt; vbaseXYZcg = V3(K_CG[0], K_CG[1], K_CG[2]);
    vXYZcg = V3(); vLastXYZcg = V3();
    for (int i = 0; i < ntanks; i++) tank[i] = K_TANK_CONTENTS[i];
    // FGMassBalance::Load computes Weight once (cg stays 0 until the first Run)
    Weight = EmptyWeight + tanks_weight() + pm_weight(); Mass = lbtoslug * Weight;
    // auxiliary (J/models/FGAuxiliary.cpp:60-118)
    alpha = beta = Vt = qbar = Mach = vcas = veas = Vground = psigt = gamma_ = hoverbmac = 0.0;
    vPilotAccel = vPilotAccelN = vAeroPQR = vAeroUVW = vNcg = vEulerRates = V3();
    // turbine (J/models/propulsion/FGTurbine.cpp:60-101 ctor + ResetToIC, :425-499 Load)
    N1 = N2 = N2norm = 0.0; FuelFlow_pph = 0.0; AugmentCmd = 0.0; NozzlePosition = 1.0; EPR = 1.0;
    Running = false; Cutoff = true; Starter = false; Starved = false; Augmentation = false; phase = tpOff;
    N1_factor = K_ENG_maxn1 - K_ENG_idlen1; N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
    IdleFF = std::pow(K_ENG_milthrust, 0.2) * 107.0;
    ThrottlePos = 0.0; Thrust = 0.0; FuelUsedLbs = 0.0; correctedTSFC = 0.0;
    propForces = propMoments = V3(); tankJ = M33();
    aeroForces = aeroMoments = vFw = V3(); clsq = bi2vel = ci2vel = 0.0;
    acForces = acMoments = vPQRidot = vPQRdot = vUVWidot = vUVWdot = vBodyAccel = V3();
    vGravAccel = V3();
    atm = StdAtmosphere();
    epa = 0.0;
  }

  // Initial-condition setters reduced to the family the reference uses: geodetic
  // latitude, longitude, altitude ASL, Euler angles, body velocities and body rates, no wind, no
  // climb-rate override.  Altitude solve: J/initialization/FGInitialCondition.cpp:749-815 (setgeod).
  void reset(double lon_deg, double lat_geod_deg, double h_sl_ft, double psi_deg, double u, double v, double w,
             double p, double q, double r, double phi_deg, double theta_deg) {
    model_init();
    double lon = lon_deg * degtorad, geodLatitude = lat_geod_deg * degtorad, alt = h_sl_ft;
    {
      double a = a_ft, b = b_ft, e2 = 1.0 - b * b / (a * a);
      double cosGeodLat = std::cos(geodLatitude), sinGeodLat = std::sin(geodLatitude);
      double N = a / std::sqrt(1 - e2 * sinGeodLat * sinGeodLat);
      double geodAlt = 0.0, n = e2, prev_n = 1.0; int iter = 0;
      if (cosGeodLat > std::fabs(sinGeodLat)) {
        double tanGeodLat = sinGeodLat / cosGeodLat, x0 = N * e2 * cosGeodLat, x = 0.0;
        while (std::fabs(n - prev_n) > 1E-15 && iter < 10) {
          double tanLat = (1 - n) * tanGeodLat, cos2Lat = 1. / (1. + tanLat * tanLat);
          double slr = b / std::sqrt(1. - e2 * cos2Lat), R = slr + alt;
          x = R * std::sqrt(cos2Lat); prev_n = n; n = x0 / x; iter++;
        }
        geodAlt = x / cosGeodLat - N;
      } else {
        double cotanGeodLat = cosGeodLat / sinGeodLat, z0 = N * e2 * sinGeodLat, z = 0.0;
        while (std::fabs(n - prev_n) > 1E-15 && iter < 10) {
          double cotanLat = cotanGeodLat / (1 - n), sin2Lat = 1. / (1. + cotanLat * cotanLat), cos2Lat = 1. - sin2Lat;
          double slr = b / std::sqrt(1. - e2 * cos2Lat), R = slr + alt;
          z = R * sign(cotanLat) * std::sqrt(sin2Lat); prev_n = n; n = z0 / (z0 + z); iter++;
        }
        geodAlt = z / sinGeodLat - N * (1 - e2);
      }
      loc.SetPositionGeodetic(lon, geodLatitude, geodAlt);
    }
    // orientation / velocities (FGInitialCondition::SetEulerAngleRadIC :446-467, SetBodyVelFpsIC :473-490)
    Quat orientation = quat_from_euler(phi_deg * degtorad, theta_deg * degtorad, psi_deg * degtorad);
    M33 icTl2b = quat_T(orientation), icTb2l = T(icTl2b);
    V3 vUVW_NED = icTb2l * V3(u, v, w);
    V3 icUVW = icTl2b * vUVW_NED;  // GetUVWFpsIC
    V3 icPQR(p, q, r);

    // ---- FGFDMExec::RunIC (J/FGFDMExec.cpp:636-669): SuspendIntegration; Initialize(IC){SetInitialState; Run()}; Run();
    saved_dT = dT; dT = 0.0;
    // FGPropagate::SetInitialState (J/models/FGPropagate.cpp:143-186)
    epa = 0.0;
    Ti2ec = M33(std::cos(epa), std::sin(epa), 0.0, -std::sin(epa), std::cos(epa), 0.0, 0.0, 0.0, 1.0);
    Tec2i = T(Ti2ec);
    vInertialPosition = Tec2i * loc.ec;
    UpdateLocationMatrices();
    qAttitudeLocal = orientation;
    qAttitudeECI = quat_mul(mat_quat(Ti2l), qAttitudeLocal);
    UpdateBodyMatrices();
    vUVW = icUVW;
    vVel = Tb2l * vUVW;
    vPQR = icPQR;
    vPQRi = vPQR + Ti2b * Omega;
    vInertialVelocity = Tb2i * vUVW + cross(Omega, vInertialPosition);
    vQtrndot = quat_dot(qAttitudeECI, vPQRi);
    update_euler();
    publish_propagate();
    Run();
    Run();
    // InitializeDerivatives (:190-196)
    for (int i = 0; i < 5; i++) { dqPQRidot[i] = vPQRidot; dqUVWidot[i] = vUVWidot; dqInertialVelocity[i] = vInertialVelocity; dqQtrndot[i] = vQtrndot; }
    dT = saved_dT;  // ResumeIntegration
    // ---- engine.init_running() (J/models/propulsion/FGTurbine.cpp:604-616)
    {
      double keep = dT; dT = 0.0;  // SuspendIntegration
      Cutoff = false; Running = true;
      N1_factor = K_ENG_maxn1 - K_ENG_idlen1; N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
      N2 = K_ENG_idlen2 + ThrottlePos * N2_factor; N1 = K_ENG_idlen1 + ThrottlePos * N1_factor;
      PropTotalDeltaT = 0.0;  // in.TotalDeltaT still holds the suspended value loaded by the last Run()
      turbine_calculate();
      dT = keep;
      phase = tpRun;
    }
    // ---- propulsion.get_steady_state() (J/models/FGPropulsion.cpp:262-308)
    {
      double currentThrust = 0, lastThrust = -1; int steady_count = 0, j = 0; bool steady = false;
      bool TrimMode = trim_status;
      propForces = propMoments = V3();
      trim_status = true;
      PropTotalDeltaT = 0.5;
      while (!steady && j < 6000) {
        turbine_calculate();
        lastThrust = currentThrust; currentThrust = Thrust;
        if (std::fabs(lastThrust - currentThrust) < 0.0001) { steady_count++; if (steady_count > 120) steady = true; }
        else steady_count = 0;
        j++;
      }
      thruster_forces();
      trim_status = TrimMode;
      PropTotalDeltaT = dT;
    }
  }

  double tanks_weight() const { double w = 0; for (int i = 0; i < ntanks; i++) w += tank[i]; return w; }
  double pm_weight() const { double w = 0; for (int i = 0; i < K_NPM; i++) w += K_PM_W[i]; return w; }

  // J/models/FGMassBalance.cpp:347-374
  V3 StructuralToBody(const V3& r) const { return V3(inchtoft * (vXYZcg.x - r.x), inchtoft * (r.y - vXYZcg.y), inchtoft * (vXYZcg.z - r.z)); }
  // FGMassBalance::GetPointmassInertia (header-only upstream)
  M33 GetPointmassInertia(double mass_sl, const V3& r) const {
    V3 v = StructuralToBody(r); V3 sv = mass_sl * v;
    double xx = sv.x * v.x, yy = sv.y * v.y, zz = sv.z * v.z, xy = -sv.x * v.y, xz = -sv.x * v.z, yz = -sv.y * v.z;
    return M33(yy + zz, xy, xz, xy, xx + zz, yz, xz, yz, xx + yy);
  }

  // J/models/FGPropagate.cpp:475-496
  void UpdateLocationMatrices() { Tl2ec = loc.Tl2ec; Tec2l = T(Tl2ec); Ti2l = Tec2l * Ti2ec; Tl2i = T(Ti2l); }
  void UpdateBodyMatrices() { Ti2b = quat_T(qAttitudeECI); Tb2i = T(Ti2b); Tl2b = Ti2b * Tl2i; Tb2l = T(Tl2b); Tec2b = Ti2b * Tec2i; Tb2ec = T(Tec2b); }
  void update_euler() {
    // FGPropagate::GetEuler -> qAttitudeLocal.GetEuler(): matrix rebuilt from the local quaternion
    M33 mT = quat_T(qAttitudeLocal);
    euler = mat_euler(mT);
    sinEuler[0] = std::sin(euler.x); sinEuler[1] = -mT.m[0][2]; sinEuler[2] = std::sin(euler.z);
    cosEuler[0] = std::cos(euler.x); cosEuler[1] = std::cos(euler.y); cosEuler[2] = std::cos(euler.z);
  }
  double GetAltitudeASL() const { return loc.radius - loc.GetSeaLevelRadius(); }
  void publish_propagate() {
    P[P_attitude_pitch_rad] = euler.y; P[P_attitude_roll_rad] = euler.x;
    P[P_velocities_u_fps] = vUVW.x; P[P_velocities_v_fps] = vUVW.y;
  }

  // ============================================================== FGPropagate::Run (J/models/FGPropagate.cpp:218-297)
  void propagate_run() {
    double dt = dT;
    if (dT != 0.0) {  // !IntegrationSuspended()
      // Integrate(): push_front(current derivative), pop_back (:336-372)
      for (int i = 4; i > 0; i--) dqQtrndot[i] = dqQtrndot[i - 1];
      dqQtrndot[0] = vQtrndot;
      for (int i = 0; i < 4; i++) qAttitudeECI.q[i] += dt * dqQtrndot[0].q[i];  // eRectEuler
      quat_normalize(qAttitudeECI);
      for (int i = 4; i > 0; i--) dqPQRidot[i] = dqPQRidot[i - 1];
      dqPQRidot[0] = vPQRidot;
      vPQRi = vPQRi + dt * dqPQRidot[0];  // eRectEuler
      for (int i = 4; i > 0; i--) dqInertialVelocity[i] = dqInertialVelocity[i - 1];
      dqInertialVelocity[0] = vInertialVelocity;
      vInertialPosition = vInertialPosition + ((1 / 12.0) * dt) * (23.0 * dqInertialVelocity[0] - 16.0 * dqInertialVelocity[1] + 5.0 * dqInertialVelocity[2]);  // AB3
      for (int i = 4; i > 0; i--) dqUVWidot[i] = dqUVWidot[i - 1];
      dqUVWidot[0] = vUVWidot;
      vInertialVelocity = vInertialVelocity + dt * (1.5 * dqUVWidot[0] - 0.5 * dqUVWidot[1]);  // AB2
    }
    epa += Omega.z * dt;
    double cos_epa = std::cos(epa), sin_epa = std::sin(epa);
    Ti2ec = M33(cos_epa, sin_epa, 0.0, -sin_epa, cos_epa, 0.0, 0.0, 0.0, 1.0);
    Tec2i = T(Ti2ec);
    loc.ec = Ti2ec * vInertialPosition; loc.ComputeDerived();
    UpdateLocationMatrices();
    UpdateBodyMatrices();
    vUVW = Ti2b * (vInertialVelocity - cross(Omega, vInertialPosition));  // CalculateUVW :329-332
    vPQR = vPQRi - Ti2b * Omega;
    vQtrndot = quat_dot(qAttitudeECI, vPQRi);
    qAttitudeLocal = mat_quat(Tl2b);
    vVel = Tb2l * vUVW;
    update_euler();
    publish_propagate();
  }

  // ============================================================== FGInertial::Run (J/models/FGInertial.cpp:127-145,193-211)
  void inertial_run() {
    double r = loc.radius, sinLat = std::sin(loc.lat), adivr = a_ft / r, preCommon = 1.5 * J2 * adivr * adivr;
    double xy = 1.0 - 5.0 * (sinLat * sinLat), z = 3.0 - 5.0 * (sinLat * sinLat), GMOverr2 = GM / (r * r);
    vGravAccel.x = -GMOverr2 * ((1.0 + (preCommon * xy)) * loc.ec.x / r);
    vGravAccel.y = -GMOverr2 * ((1.0 + (preCommon * xy)) * loc.ec.y / r);
    vGravAccel.z = -GMOverr2 * ((1.0 + (preCommon * z)) * loc.ec.z / r);
  }

  // ============================================================== FGFCS::Run (J/models/FGFCS.cpp:153-178) + components
  double valref(const ValRef& v) const { return v.is_prop ? v.sign * P[v.prop] : v.value; }
  void set_prop(int pidx, double val) {
    P[pidx] = val;
    // tied setters with unit aliases (J/models/FGFCS.cpp:182-290 SetD*Pos): only the speedbrake pair is
    // reached by the F-16 channels (kinematic writes -deg, aero reads -rad).
    if (pidx == P_fcs_speedbrake_pos_deg) P[P_fcs_speedbrake_pos_rad] = val * degtorad;
    else if (pidx == P_fcs_speedbrake_pos_rad) P[P_fcs_speedbrake_pos_deg] = val * radtodeg;
  }
  double eval_factors(const Factor* f, int n) const {
    // FGFunction product (J/math/FGFunction.cpp): temp = first; temp *= next ... in document order
    double temp = 1.0;
    for (int i = 0; i < n; i++) {
      double x;
      switch (f[i].kind) {
        case 0: x = P[f[i].prop]; break;
        case 1: x = f[i].value; break;
        case 2: x = (f[i].table->dim == 1) ? table1(f[i].table, P[f[i].table->row_var]) : table2(f[i].table, P[f[i].table->row_var], P[f[i].table->col_var]); break;
        case 3: x = std::cos(P[f[i].prop]); break;
        default: x = std::sin(P[f[i].prop]); break;
      }
      temp = (i == 0) ? x : temp * x;
    }
    return temp;
  }
  void fcs_run() {
    P[P_fcs_throttle_pos_norm] = P[P_fcs_throttle_cmd_norm];  // ThrottlePos[i] = ThrottleCmd[i]
    const double dt = fcs_dt;
    for (int ci = 0; ci < N_FCS_COMPS; ci++) {
      const Comp& c = FCS_COMPS[ci];
      double Output = compOutput[ci], Input = 0.0;
      switch (c.type) {
        case C_SWITCH: {  // J/models/flight_control/FGSwitch.cpp:125-152
          bool pass = false; double default_output = valref(c.defval);
          for (int t = 0; t < c.ntests && !pass; t++) {
            const Test& te = c.tests[t];
            bool res = (te.logic == 0);
            for (int k = 0; k < te.ncond; k++) {  // J/math/FGCondition.cpp:167-221
              const Cond& cd = te.conds[k]; double lhs = P[cd.prop], rhs = valref(cd.rhs); bool b;
              switch (cd.op) { case 0: b = lhs < rhs; break; case 1: b = lhs <= rhs; break; case 2: b = lhs > rhs; break;
                               case 3: b = lhs >= rhs; break; case 4: b = lhs == rhs; break; default: b = lhs != rhs; break; }
              if (te.logic == 0) { if (!b) res = false; } else { if (b) res = true; }
            }
            if (res) { pass = true; Output = valref(te.value); }
          }
          if (!pass) Output = default_output;
        } break;
        case C_PURE_GAIN: Input = c.in_sign[0] * P[c.in_prop[0]]; Output = c.gain * Input; break;  // FGGain.cpp:138-172
        case C_SCHED_GAIN: { Input = c.in_sign[0] * P[c.in_prop[0]]; double SchedGain = table1(c.table, P[c.table->row_var]); Output = c.gain * SchedGain * Input; } break;
        case C_AEROSURF:
          Input = c.in_sign[0] * P[c.in_prop[0]];
          if (c.zero_centered) { if (Input == 0.0) Output = 0.0; else if (Input > 0) Output = (Input / c.in_max) * c.out_max; else Output = (Input / c.in_min) * c.out_min; }
          else Output = c.out_min + ((Input - c.in_min) / (c.in_max - c.in_min)) * (c.out_max - c.out_min);
          Output *= c.gain;
          break;
        case C_SUMMER: Output = 0.0; for (int k = 0; k < c.n_in; k++) Output += c.in_sign[k] * P[c.in_prop[k]]; Output += c.bias; break;  // FGSummer.cpp:72-84
        case C_PID: {  // pid_step above
          Input = c.in_sign[0] * P[c.in_prop[0]];
          double test = 0.0; if (c.has_trigger) test = c.trig_sign * P[c.trig_prop];
          Output = pid_step(pid[ci], Input, test, c.kp, c.ki, c.kd, c.int_type, dt);
        } break;
        case C_KINEMATIC:  // kinemat_step above
          Input = c.in_sign[0] * P[c.in_prop[0]];
          Output = kinemat_step(c.detents, c.times, c.ndet, c.noscale, trim_status, Input, P[c.out_prop[0]], dt);
          break;
        case C_FCSFUNC: Output = eval_factors(c.factors, c.nfac); if (c.n_in > 0) { Input = c.in_sign[0] * P[c.in_prop[0]]; Output *= Input; } break;  // FGFCSFunction.cpp:73-84
      }
      if (c.has_clip) Output = Constrain(c.clip_min, Output, c.clip_max);  // FGFCSComponent::Clip :266-290
      compOutput[ci] = Output;
      for (int k = 0; k < c.n_out; k++) set_prop(c.out_prop[k], Output);  // SetOutput :232-236
    }
  }

  // ============================================================== FGMassBalance::Run (J/models/FGMassBalance.cpp:181-260)
  void massbalance_run() {
    // LoadInputs(eMassBalance) (J/FGFDMExec.cpp:563-571): tank aggregates, tank inertia with the OLD cg
    TanksWeight = tanks_weight();
    TanksMoment = V3();
    for (int i = 0; i < ntanks; i++) TanksMoment = TanksMoment + tank[i] * V3(K_TANK_LOC[i][0], K_TANK_LOC[i][1], K_TANK_LOC[i][2]);
    tankJ = M33();
    for (int i = 0; i < ntanks; i++) tankJ = tankJ + GetPointmassInertia(lbtoslug * tank[i], V3(K_TANK_LOC[i][0], K_TANK_LOC[i][1], K_TANK_LOC[i][2]));
    Weight = EmptyWeight + TanksWeight + pm_weight() + 0.0 * slugtolb + 0.0;
    Mass = lbtoslug * Weight;
    V3 pmMoment;
    for (int i = 0; i < K_NPM; i++) pmMoment = pmMoment + K_PM_W[i] * V3(K_PM_LOC[i][0], K_PM_LOC[i][1], K_PM_LOC[i][2]);
    vXYZcg = (EmptyWeight * vbaseXYZcg + pmMoment + TanksMoment + V3()) / Weight;
    if (mag(vLastXYZcg) == 0.0) vLastXYZcg = vXYZcg;
    vLastXYZcg = vXYZcg;
    mJ = baseJ;
    mJ = mJ + GetPointmassInertia(lbtoslug * EmptyWeight, vbaseXYZcg);
    M33 pmJ;
    for (int i = 0; i < K_NPM; i++) pmJ = pmJ + GetPointmassInertia(lbtoslug * K_PM_W[i], V3(K_PM_LOC[i][0], K_PM_LOC[i][1], K_PM_LOC[i][2]));
    mJ = mJ + pmJ; mJ = mJ + tankJ;
    double Ixx = mJ.m[0][0], Iyy = mJ.m[1][1], Izz = mJ.m[2][2], Ixy = -mJ.m[0][1], Ixz = -mJ.m[0][2], Iyz = -mJ.m[1][2];
    double k1 = (Iyy * Izz - Iyz * Iyz), k2 = (Iyz * Ixz + Ixy * Izz), k3 = (Ixy * Iyz + Iyy * Ixz);
    double denom = 1.0 / (Ixx * k1 - Ixy * k2 - Ixz * k3);
    k1 = k1 * denom; k2 = k2 * denom; k3 = k3 * denom;
    double k4 = (Izz * Ixx - Ixz * Ixz) * denom, k5 = (Ixy * Ixz + Iyz * Ixx) * denom, k6 = (Ixx * Iyy - Ixy * Ixy) * denom;
    mJinv = M33(k1, k2, k3, k2, k4, k5, k3, k5, k6);
  }

  // ============================================================== FGAuxiliary::Run (J/models/FGAuxiliary.cpp:134-231)
  void auxiliary_run() {
    vEulerRates.y = vPQR.y * cosEuler[0] - vPQR.z * sinEuler[0];
    if (cosEuler[1] != 0.0) { vEulerRates.z = (vPQR.y * sinEuler[0] + vPQR.z * cosEuler[0]) / cosEuler[1]; vEulerRates.x = vPQR.x + vEulerRates.z * sinEuler[1]; }
    vAeroPQR = vPQR;   // no turbulence
    vAeroUVW = vUVW;   // no wind (Tl2b * 0)
    alpha = beta = 0;
    double AeroU2 = vAeroUVW.x * vAeroUVW.x, AeroV2 = vAeroUVW.y * vAeroUVW.y, AeroW2 = vAeroUVW.z * vAeroUVW.z;
    double mUW = AeroU2 + AeroW2, Vt2 = mUW + AeroV2;
    Vt = std::sqrt(Vt2);
    if (Vt > 0.001) { beta = std::atan2(vAeroUVW.y, std::sqrt(mUW)); if (mUW >= 1E-6) alpha = std::atan2(vAeroUVW.z, vAeroUVW.x); }
    double ca = std::cos(alpha), sa = std::sin(alpha), cb = std::cos(beta), sb = std::sin(beta);  // UpdateWindMatrices :247-268
    mTw2b = M33(ca * cb, -ca * sb, -sa, sb, cb, 0.0, sa * cb, -sa * sb, ca);
    mTb2w = T(mTw2b);
    double densityD2 = 0.5 * atm.Density;
    qbar = densityD2 * Vt2;
    Mach = Vt / atm.Soundspeed;
    Vground = std::sqrt(vVel.x * vVel.x + vVel.y * vVel.y);
    psigt = std::atan2(vVel.y, vVel.x); if (psigt < 0.0) psigt += 2 * M_PI;
    gamma_ = std::atan2(-vVel.z, Vground);
    tat = atm.Temperature * (1 + 0.2 * Mach * Mach);
    tatc = (tat - 491.67) / 1.8;  // RankineToCelsius
    pt = PitotTotalPressure(Mach, atm.Pressure);
    if (std::fabs(Mach) > 0.0) {
      // VcalibratedFromMach (J/FGJSBBase.cpp:289-296)
      double qc = PitotTotalPressure(Mach, atm.Pressure) - atm.Pressure;
      vcas = atm.StdDaySLsoundspeed * MachFromImpactPressure(qc, atm.StdDaySLpressure);
      veas = std::sqrt(2 * qbar / atm.SLdensity);
    } else vcas = veas = 0.0;
    vNcg = vBodyAccel / gAccelReference;
    V3 ToEyePt = StructuralToBody(V3(K_EYEPOINT[0], K_EYEPOINT[1], K_EYEPOINT[2]));
    vPilotAccel = vBodyAccel + cross(vPQRidot, ToEyePt);
    vPilotAccel = vPilotAccel + cross(vPQRi, cross(vPQRi, ToEyePt));
    vPilotAccelN = vPilotAccel / gAccelReference;
    V3 vMac = Tb2l * StructuralToBody(V3(K_AERORP[0], K_AERORP[1], K_AERORP[2]));
    double DistanceAGL = loc.geodAlt - 0.0;  // FGDefaultGroundCallback, terrain elevation 0
    hoverbmac = (DistanceAGL - vMac.z) / K_bw;
    // publish tied properties
    P[P_aero_alpha_rad] = alpha; P[P_aero_alpha_deg] = alpha * radtodeg; P[P_aero_beta_rad] = beta;
    P[P_aero_qbar_psf] = qbar; P[P_velocities_mach] = Mach; P[P_velocities_vc_kts] = vcas * fpstokts; P[P_velocities_vg_fps] = Vground;
    P[P_velocities_p_aero_rad_sec] = vAeroPQR.x; P[P_velocities_q_aero_rad_sec] = vAeroPQR.y; P[P_velocities_r_aero_rad_sec] = vAeroPQR.z;
    P[P_accelerations_n_pilot_y_norm] = vPilotAccelN.y; P[P_accelerations_n_pilot_z_norm] = vPilotAccelN.z;
    P[P_aero_h_b_mac_ft] = hoverbmac;
  }

  // ============================================================== FGPropulsion::Run (J/models/FGPropulsion.cpp:113-160) + FGTurbine
  double Seek(double v, double target, double accel, double decel) const {  // FGTurbine.cpp:400-410
    if (v > target) { v -= PropTotalDeltaT * decel; if (v < target) v = target; }
    else if (v < target) { v += PropTotalDeltaT * accel; if (v > target) v = target; }
    return v;
  }
  double SpoolUp(double factor) const {  // FGSpoolUp (header-only upstream; pinned by data/tests/TestTurbine.py:36-41)
    double delay = factor * 90.0 / (K_ENG_bypassratio + 3.0);
    double n = std::min(1.0, N2norm + 0.1);
    return delay / (1 + 3 * (1 - n) * (1 - n) * (1 - n) + (1 - atm.Density / atm.SLdensity));
  }
  void turbine_calculate() {  // FGTurbine::Calculate :107-170
    double thrust;
    // RunPreFunctions: the three thrust tables on (velocities/mach, atmosphere/density-altitude)
    idleThrustVal = table2(ENG_IdleThrust, P[P_velocities_mach], P[P_atmosphere_density_altitude]);
    milThrustVal = table2(ENG_MilThrust, P[P_velocities_mach], P[P_atmosphere_density_altitude]);
    augThrustVal = table2(ENG_AugThrust, P[P_velocities_mach], P[P_atmosphere_density_altitude]);
    ThrottlePos = P[P_fcs_throttle_pos_norm];
    if (ThrottlePos > 1.0) { AugmentCmd = ThrottlePos - 1.0; ThrottlePos -= AugmentCmd; } else AugmentCmd = 0.0;
    if ((phase == tpTrim) && (PropTotalDeltaT > 0)) {
      if (Running && !Starved) { phase = tpRun; N1_factor = K_ENG_maxn1 - K_ENG_idlen1; N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
        N2 = K_ENG_idlen2 + ThrottlePos * N2_factor; N1 = K_ENG_idlen1 + ThrottlePos * N1_factor; Cutoff = false; }
      else { phase = tpOff; Cutoff = true; }
    }
    if (!Running && Cutoff && Starter) { if (phase == tpOff) phase = tpSpinUp; }
    if ((Starter == true) || (qbar > 30.0)) { if (!Running && !Cutoff && (N2 > 15.0)) phase = tpStart; }
    if (Cutoff && (phase != tpSpinUp)) phase = tpOff;
    if (PropTotalDeltaT == 0) phase = tpTrim;
    if (Starved) phase = tpOff;
    switch (phase) {
      case tpOff: thrust = turbine_off(); break;
      case tpRun: thrust = turbine_run(); break;
      case tpTrim: thrust = turbine_trim(); break;
      default: std::fprintf(stderr, "oracle: turbine phase %d not on the reference path\n", phase); std::abort();
    }
    Thrust = std::cos(0.0) * thrust;  // FGThruster::Calculate, ReverserAngle = 0
  }
  double turbine_off() {  // :172-194
    Running = false;
    FuelFlow_pph = Seek(FuelFlow_pph, 0, 1000.0, 10000.0);
    N1 = Seek(N1, qbar / 10.0, N1 / 2.0 + 0.1, N1 / 2.0);
    N2 = Seek(N2, qbar / 15.0, N2 / 2.0 + 0.1, N2 / 2.0);
    NozzlePosition = Seek(NozzlePosition, 1.0, 0.8, 0.8); EPR = Seek(EPR, 1.0, 0.2, 0.2);
    Augmentation = false;
    return 0.0;
  }
  double turbine_run() {  // :196-272
    double idlethrust = K_ENG_milthrust * idleThrustVal;
    double milthrust = (K_ENG_milthrust - idlethrust) * milThrustVal;
    Running = true; Starter = false;
    N1_factor = K_ENG_maxn1 - K_ENG_idlen1; N2_factor = K_ENG_maxn2 - K_ENG_idlen2;
    N2 = Seek(N2, K_ENG_idlen2 + ThrottlePos * N2_factor, SpoolUp(1.0), SpoolUp(3.0));
    N1 = Seek(N1, K_ENG_idlen1 + ThrottlePos * N1_factor, SpoolUp(1.0), SpoolUp(2.4));
    N2norm = (N2 - K_ENG_idlen2) / N2_factor;
    double thrust = idlethrust + (milthrust * N2norm * N2norm);
    if (!Augmentation) {
      // FGSimplifiedTSFC (header-only upstream, NOT pinned in the tree): tsfc*sqrt(T/389.7)*(0.84+(1-N2norm)^2)
      correctedTSFC = K_ENG_tsfc * std::sqrt(atm.Temperature / 389.7) * (0.84 + (1 - N2norm) * (1 - N2norm));
      FuelFlow_pph = Seek(FuelFlow_pph, thrust * correctedTSFC, 1000.0, 10000.0);
      if (FuelFlow_pph < IdleFF) FuelFlow_pph = IdleFF;
      NozzlePosition = Seek(NozzlePosition, 1.0 - N2norm, 0.8, 0.8);
      thrust = thrust * (1.0 - 0.0);  // BleedDemand = 0
      EPR = 1.0 + thrust / K_ENG_milthrust;
    }
    if (K_ENG_augmethod == 1) { if ((ThrottlePos > 0.99) && (N2 > 97.0)) Augmentation = true; else Augmentation = false; }
    if ((K_ENG_augmented == 1) && Augmentation && (K_ENG_augmethod < 2)) {
      thrust = augThrustVal * K_ENG_maxthrust;
      FuelFlow_pph = Seek(FuelFlow_pph, thrust * K_ENG_atsfc, 5000.0, 10000.0);
      NozzlePosition = Seek(NozzlePosition, 1.0, 0.8, 0.8);
    }
    if (K_ENG_augmethod == 2) {
      if (AugmentCmd > 0.0) {
        Augmentation = true;
        double tdiff = (K_ENG_maxthrust * augThrustVal) - thrust;
        thrust += (tdiff * AugmentCmd);
        FuelFlow_pph = Seek(FuelFlow_pph, thrust * K_ENG_atsfc, 5000.0, 10000.0);
        NozzlePosition = Seek(NozzlePosition, 1.0, 0.8, 0.8);
      } else Augmentation = false;
    }
    if (Cutoff) phase = tpOff;
    if (Starved) phase = tpOff;
    return thrust;
  }
  double turbine_trim() {  // :341-372
    double idlethrust = K_ENG_milthrust * idleThrustVal;
    double milthrust = (K_ENG_milthrust - idlethrust) * milThrustVal;
    double N2_ = K_ENG_idlen2 + ThrottlePos * N2_factor;
    double N2norm_ = (N2_ - K_ENG_idlen2) / N2_factor;
    double thrust = (idlethrust + (milthrust * N2norm_ * N2norm_)) * (1.0 - 0.0);
    if (K_ENG_augmethod == 1) { if ((ThrottlePos > 0.99) && (N2_ > 97.0)) Augmentation = true; else Augmentation = false; }
    if ((K_ENG_augmented == 1) && Augmentation && (K_ENG_augmethod < 2)) thrust = K_ENG_maxthrust * augThrustVal;
    if (K_ENG_augmethod == 2) { if (AugmentCmd > 0.0) { double tdiff = (K_ENG_maxthrust * augThrustVal) - thrust; thrust += (tdiff * AugmentCmd); } }
    return thrust;
  }
  void thruster_forces() {  // FGForce::GetBodyForces (J/models/propulsion/FGForce.cpp:78-91), mT = I
    V3 vFb(Thrust, 0.0, 0.0);
    V3 vDXYZ = StructuralToBody(V3(K_THRUSTER_LOC[0], K_THRUSTER_LOC[1], K_THRUSTER_LOC[2]));
    propForces = vFb; propMoments = V3() + cross(vDXYZ, vFb);
  }
  void propulsion_run() {
    PropTotalDeltaT = dT;
    turbine_calculate();
    // ConsumeFuel (J/models/FGPropulsion.cpp:164-260): equal split over feed tanks that still hold fuel
    if (!trim_status) {
      int TanksWithFuel = 0; int feed[8];
      for (int i = 0; i < ntanks; i++) if (tank[i] > 0.0) feed[TanksWithFuel++] = i;  // all tanks priority 1, selected, unusable = 0
      Starved = (TanksWithFuel == 0);
      if (!Starved) {
        double FuelFlowRate = FuelFlow_pph / 3600.0, FuelExpended = FuelFlowRate * PropTotalDeltaT;  // CalcFuelNeed :381-387
        FuelUsedLbs += FuelExpended;
        double per = FuelExpended / TanksWithFuel;
        for (int k = 0; k < TanksWithFuel; k++) {  // FGTank::Drain (J/models/propulsion/FGTank.cpp:282-296)
          int i = feed[k]; double remaining = tank[i] - per;
          if (remaining >= 0.0) tank[i] -= per; else if (tank[i] > 0.0) tank[i] = 0.0;
        }
      }
    }
    thruster_forces();
  }

  // ============================================================== FGAerodynamics::Run (J/models/FGAerodynamics.cpp:132-300)
  void aerodynamics_run() {
    const double twovel = 2 * Vt;
    if (qbar > 1.0) { clsq = vFw.z / (K_Sw * qbar); clsq *= clsq; }
    for (int i = 0; i < N_AERO_PRE; i++) P[AERO_PRE[i].axis_or_prop] = eval_factors(AERO_PRE[i].factors, AERO_PRE[i].nfac);  // RunPreFunctions
    if (twovel != 0) { bi2vel = K_bw / twovel; ci2vel = K_cbarw / twovel; }
    P[P_aero_bi2vel] = bi2vel; P[P_aero_ci2vel] = ci2vel;
    double f[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < N_AERO_FNS; i++) f[AERO_FNS[i].axis_or_prop] += eval_factors(AERO_FNS[i].factors, AERO_FNS[i].nfac);
    V3 vFnative(f[0], f[1], f[2]);
    vFnative.x *= -1; vFnative.z *= -1;  // atWind
    aeroForces = mTw2b * vFnative;
    V3 RPBody = StructuralToBody(V3(K_AERORP[0], K_AERORP[1], K_AERORP[2]));
    V3 vDXYZcg(RPBody.x - 0.0, RPBody.y + 0.0, RPBody.z - 0.0);
    V3 vMomentsMRC(f[3], f[4], f[5]);  // moments given in body axes
    aeroMoments = vMomentsMRC + cross(vDXYZcg, aeroForces);
    vFw = mTb2w * aeroForces; vFw.x *= -1; vFw.z *= -1;
  }

  // ============================================================== FGAccelerations::Run (J/models/FGAccelerations.cpp:109-207)
  void accelerations_run() {
    acForces = aeroForces + propForces; acMoments = aeroMoments + propMoments;  // FGAircraft::Run (ground/external/buoyant = 0 in flight)
    vPQRidot = mJinv * (acMoments - cross(vPQRi, mJ * vPQRi));
    vPQRdot = vPQRidot - cross(vPQRi, Ti2b * Omega);
    vBodyAccel = acForces / Mass;
    vUVWdot = vBodyAccel - cross(vPQR + 2.0 * (Ti2b * Omega), vUVW);
    vUVWdot = vUVWdot - Ti2b * cross(Omega, cross(Omega, vInertialPosition));
    vUVWdot = vUVWdot + Tec2b * vGravAccel;
    vUVWidot = Tb2i * vBodyAccel + Tec2i * vGravAccel;
  }

  // ============================================================== FGFDMExec::Run (J/FGFDMExec.cpp:407-431), model order :222-236
  void Run() {
    if (dT != 0.0) sim_time += dT;  // IncrTime :196-203
    propagate_run();
    inertial_run();
    atm.Calculate(GetAltitudeASL());
    P[P_atmosphere_density_altitude] = atm.DensityAltitude;
    fcs_run();
    massbalance_run();
    auxiliary_run();
    propulsion_run();
    aerodynamics_run();
    accelerations_run();
  }
};

}  // namespace orc

// ---------------------------------------------------------------------------------- C API for ctypes
using orc::F16;
static const char* SNAP_NAMES[] = {
    // core propagate state
    "q0", "q1", "q2", "q3", "wi_x", "wi_y", "wi_z", "ri_x", "ri_y", "ri_z", "vi_x", "vi_y", "vi_z", "epa",
    "dqv0_x", "dqv0_y", "dqv0_z", "dqv1_x", "dqv1_y", "dqv1_z", "dqa0_x", "dqa0_y", "dqa0_z",
    "pqridot_x", "pqridot_y", "pqridot_z", "uvwidot_x", "uvwidot_y", "uvwidot_z", "bodyaccel_x", "bodyaccel_y", "bodyaccel_z",
    // engine / fuel / mass
    "N1", "N2", "N2norm", "FuelFlow_pph", "tank0", "tank1", "tank2", "tank3", "cg_x", "cg_y", "cg_z", "sim_time",
    // outputs read by the reference's python layer (Catalog)
    "lon_deg", "lat_geod_deg", "h_sl_ft", "roll_rad", "pitch_rad", "heading_rad", "psi_deg",
    "v_north_fps", "v_east_fps", "v_down_fps", "u_fps", "v_fps", "w_fps", "vc_fps",
    "n_pilot_x", "n_pilot_y", "n_pilot_z", "p_rad_sec", "q_rad_sec", "r_rad_sec", "eci_velocity_mag_fps",
    // diagnostics
    "alpha_rad", "beta_rad", "mach", "qbar", "vt_fps", "thrust_lbs", "mass_slugs", "geod_alt_ft",
    "fx", "fy", "fz", "mx", "my", "mz", "temperature_R", "pressure_psf", "density", "density_altitude",
    // turbine flags as the CUDA state packs them: bit 0 Starved, bit 1 Augmentation
    "engflags",
};
static const int N_SNAP = sizeof(SNAP_NAMES) / sizeof(SNAP_NAMES[0]);

extern "C" {
void* orc_fdm_create(double dt, double fcs_dt) { return new F16(dt, fcs_dt); }
void orc_fdm_destroy(void* h) { delete (F16*)h; }
void orc_fdm_reset(void* h, double lon_deg, double lat_geod_deg, double h_sl_ft, double psi_deg, double u, double v, double w,
                   double p, double q, double r, double phi_deg, double theta_deg) {
  ((F16*)h)->reset(lon_deg, lat_geod_deg, h_sl_ft, psi_deg, u, v, w, p, q, r, phi_deg, theta_deg);
}
void orc_fdm_set_controls(void* h, double aileron, double elevator, double rudder, double throttle) {
  F16* f = (F16*)h;
  f->P[orc::P_fcs_aileron_cmd_norm] = aileron; f->P[orc::P_fcs_elevator_cmd_norm] = elevator;
  f->P[orc::P_fcs_rudder_cmd_norm] = rudder; f->P[orc::P_fcs_throttle_cmd_norm] = throttle;
}
void orc_fdm_run(void* h, int nframes) { F16* f = (F16*)h; for (int i = 0; i < nframes; i++) f->Run(); }
int orc_fdm_n_snapshot() { return N_SNAP; }
const char* orc_fdm_snapshot_name(int i) { return SNAP_NAMES[i]; }
void orc_fdm_snapshot(void* h, double* o) {
  F16* f = (F16*)h; int k = 0;
  for (int i = 0; i < 4; i++) o[k++] = f->qAttitudeECI.q[i];
  o[k++] = f->vPQRi.x; o[k++] = f->vPQRi.y; o[k++] = f->vPQRi.z;
  o[k++] = f->vInertialPosition.x; o[k++] = f->vInertialPosition.y; o[k++] = f->vInertialPosition.z;
  o[k++] = f->vInertialVelocity.x; o[k++] = f->vInertialVelocity.y; o[k++] = f->vInertialVelocity.z; o[k++] = f->epa;
  for (int j = 0; j < 2; j++) { o[k++] = f->dqInertialVelocity[j].x; o[k++] = f->dqInertialVelocity[j].y; o[k++] = f->dqInertialVelocity[j].z; }
  o[k++] = f->dqUVWidot[0].x; o[k++] = f->dqUVWidot[0].y; o[k++] = f->dqUVWidot[0].z;
  o[k++] = f->vPQRidot.x; o[k++] = f->vPQRidot.y; o[k++] = f->vPQRidot.z;
  o[k++] = f->vUVWidot.x; o[k++] = f->vUVWidot.y; o[k++] = f->vUVWidot.z;
  o[k++] = f->vBodyAccel.x; o[k++] = f->vBodyAccel.y; o[k++] = f->vBodyAccel.z;
  o[k++] = f->N1; o[k++] = f->N2; o[k++] = f->N2norm; o[k++] = f->FuelFlow_pph;
  for (int i = 0; i < 4; i++) o[k++] = f->tank[i];
  o[k++] = f->vXYZcg.x; o[k++] = f->vXYZcg.y; o[k++] = f->vXYZcg.z; o[k++] = f->sim_time;
  o[k++] = f->loc.lon * orc::radtodeg; o[k++] = f->loc.geodLat * orc::radtodeg; o[k++] = f->GetAltitudeASL();
  o[k++] = f->euler.x; o[k++] = f->euler.y; o[k++] = f->euler.z; o[k++] = f->euler.z * orc::radtodeg;
  o[k++] = f->vVel.x; o[k++] = f->vVel.y; o[k++] = f->vVel.z; o[k++] = f->vUVW.x; o[k++] = f->vUVW.y; o[k++] = f->vUVW.z; o[k++] = f->vcas;
  o[k++] = f->vPilotAccelN.x; o[k++] = f->vPilotAccelN.y; o[k++] = f->vPilotAccelN.z;
  o[k++] = f->vPQR.x; o[k++] = f->vPQR.y; o[k++] = f->vPQR.z; o[k++] = orc::mag(f->vInertialVelocity);
  o[k++] = f->alpha; o[k++] = f->beta; o[k++] = f->Mach; o[k++] = f->qbar; o[k++] = f->Vt; o[k++] = f->Thrust; o[k++] = f->Mass; o[k++] = f->loc.geodAlt;
  o[k++] = f->acForces.x; o[k++] = f->acForces.y; o[k++] = f->acForces.z; o[k++] = f->acMoments.x; o[k++] = f->acMoments.y; o[k++] = f->acMoments.z;
  o[k++] = f->atm.Temperature; o[k++] = f->atm.Pressure; o[k++] = f->atm.Density; o[k++] = f->atm.DensityAltitude;
  o[k++] = (double)((f->Starved ? 1 : 0) | (f->Augmentation ? 2 : 0));
  if (k != N_SNAP) { std::fprintf(stderr, "oracle: snapshot size mismatch %d vs %d\n", k, N_SNAP); std::abort(); }
}
int orc_fdm_n_props() { return orc::N_PROPS; }
const char* orc_fdm_prop_name(int i) { return orc::PROP_NAMES[i]; }
void orc_fdm_get_props(void* h, double* o) { std::memcpy(o, ((F16*)h)->P, sizeof(double) * orc::N_PROPS); }
void orc_fdm_set_props(void* h, const double* o) { std::memcpy(((F16*)h)->P, o, sizeof(double) * orc::N_PROPS); }
int orc_fdm_n_comps() { return orc::N_FCS_COMPS; }
const char* orc_fdm_comp_name(int i) { return orc::FCS_COMPS[i].name; }
int orc_fdm_comp_type(int i) { return orc::FCS_COMPS[i].type; }
// pid state: [n_comps][3]
void orc_fdm_get_pid(void* h, double* o) { F16* f = (F16*)h; for (int i = 0; i < orc::N_FCS_COMPS; i++) { o[3 * i] = f->pid[i].Input_prev; o[3 * i + 1] = f->pid[i].Input_prev2; o[3 * i + 2] = f->pid[i].I_out_total; } }
// standalone closed-form helpers (used by the known-answer tests)
void orc_atmosphere_biased(double h_ft, double delta_T_R, double* out) { orc::StdAtmosphere a; a.SetTemperatureBias(delta_T_R); a.Calculate(h_ft); out[0] = a.Temperature; out[1] = a.Pressure; out[2] = a.Density; out[3] = a.Soundspeed; out[4] = a.DensityAltitude; out[5] = a.PressureAltitude; }
void orc_atmosphere(double h_ft, double* out) { orc::StdAtmosphere a; a.Calculate(h_ft); out[0] = a.Temperature; out[1] = a.Pressure; out[2] = a.Density; out[3] = a.Soundspeed; out[4] = a.DensityAltitude; out[5] = a.PressureAltitude; }
double orc_vcas_from_mach(double mach, double p) { orc::StdAtmosphere a; double qc = orc::PitotTotalPressure(mach, p) - p; return a.StdDaySLsoundspeed * orc::MachFromImpactPressure(qc, a.StdDaySLpressure); }
double orc_kinemat(const double* detents, const double* times, int ndet, int noscale, double input, double output, double dt) {
  return orc::kinemat_step(detents, times, ndet, noscale != 0, false, input, output, dt);
}
// state = {Input_prev, Input_prev2, I_out_total}, updated in place
double orc_pid(double* state, double input, double test, double kp, double ki, double kd, int int_type, double dt) {
  orc::PidState s{state[0], state[1], state[2]};
  const double out = orc::pid_step(s, input, test, kp, ki, kd, int_type, dt);
  state[0] = s.Input_prev; state[1] = s.Input_prev2; state[2] = s.I_out_total;
  return out;
}
// J2 gravity in ECEF at an ECEF position (ft, ft/s^2): FGInertial::GetGravityJ2 (J/models/FGInertial.cpp:193-211)
void orc_gravity(double x, double y, double z, double* out) {
  F16 f(1.0 / 60.0, 1.0 / 120.0);
  f.loc.ec = orc::V3(x, y, z); f.loc.ComputeDerived();
  f.inertial_run();
  out[0] = f.vGravAccel.x; out[1] = f.vGravAccel.y; out[2] = f.vGravAccel.z;
}
// mass properties of the last frame: out = {Weight lb, Mass slug, cg x y z (in), J[9], Jinv[9]}
void orc_fdm_mass(void* h, double* out) {
  F16* f = (F16*)h; int k = 0;
  out[k++] = f->Weight; out[k++] = f->Mass; out[k++] = f->vXYZcg.x; out[k++] = f->vXYZcg.y; out[k++] = f->vXYZcg.z;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[k++] = f->mJ.m[i][j];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[k++] = f->mJinv.m[i][j];
}
void orc_fdm_set_tanks(void* h, const double* t) { F16* f = (F16*)h; for (int i = 0; i < f->ntanks; i++) f->tank[i] = t[i]; }
void orc_geodetic(double x, double y, double z, double* out) { orc::Location l; l.SetEllipse(20925646.32546, 20855486.5951); l.ec = orc::V3(x, y, z); l.ComputeDerived(); out[0] = l.lon; out[1] = l.lat; out[2] = l.geodLat; out[3] = l.geodAlt; out[4] = l.radius; out[5] = l.GetSeaLevelRadius(); }
}
