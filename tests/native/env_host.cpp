// TEST INFRASTRUCTURE: the PRODUCT's env-layer building blocks -- csrc/env_core.cuh (+ csrc/fmath.cuh): keyed RNG, WGS-84 local
// frames, AO / TA / R, the proportional-navigation missile step, the posture reward shaping -- compiled for the HOST, so that
// `pytest -m "not gpu"` checks the device arithmetic of the missile / geometry path against the CPU oracle (oracle/env_oracle.py,
// itself pinned by golden episodes of the reference's own Python) without a GPU.  Nothing here ships (no CPU path exists in the
// simulator); tests/test_env_host.py compiles this file into a temporary directory.  See tests/native/fdm_host.cpp for what
// differs from the device build (emulated reciprocal seeds, no FMA contraction).
#include <cstdint>
#include <cstring>
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static
struct HostDim3 { unsigned x, y, z; };
static HostDim3 threadIdx;          // named by one prefetch helper that the host build never calls
#include "../../include/acs.h"
#include "env_core.cuh"

extern "C" {
double eh_u01(uint64_t seed, int64_t env, int64_t purpose, int64_t a, int64_t b, int64_t c) { return env_u01(seed, env, purpose, a, b, c); }
void eh_lla2neu(const double* origin, const double* lla, double* neu) {       // origin / lla = (lon, lat, alt)
  const GeoOrigin o = geo_origin(origin[0], origin[1], origin[2]);
  lla2neu(o, lla[0], lla[1], lla[2], neu[0], neu[1], neu[2]);
}
double eh_neu2alt(const double* origin, const double* neu) {
  const GeoOrigin o = geo_origin(origin[0], origin[1], origin[2]);
  return neu2alt(o, neu[0], neu[1], neu[2]);
}
void eh_ao_ta_r(const double* ego, const double* enm, int two_d, double* out) {
  Feat a = {ego[0], ego[1], ego[2], ego[3], ego[4], ego[5]}, b = {enm[0], enm[1], enm[2], enm[3], enm[4], enm[5]};
  const AoTaR g = get_ao_ta_r(a, b, two_d != 0);
  out[0] = g.AO; out[1] = g.TA; out[2] = g.R; out[3] = g.side;
}
// one MissileSimulator.run() without the fuze / miss bookkeeping: t += dt, _guidance, _state_trans (what the kernels do per substep)
// st: pn pe pu vn ve vu theta phi alt t m dtheta dphi (13 doubles, in / out); tg: target feature; out: ny nz distance
void eh_missile_step(int kind, const double* origin, double* st, const double* tg, double dt, double* out) {
  const GeoOrigin o = geo_origin(origin[0], origin[1], origin[2]);
  const MissileParams pr = missile_params(kind);
  Missile m;
  std::memset(&m, 0, sizeof(m));
  m.pn = st[0]; m.pe = st[1]; m.pu = st[2]; m.vn = st[3]; m.ve = st[4]; m.vu = st[5]; m.theta = st[6]; m.phi = st[7]; m.alt = st[8];
  m.t = st[9]; m.m = st[10]; m.dtheta = st[11]; m.dphi = st[12];
  m.st = sin(m.theta); m.ct = cos(m.theta);             // the kernels cache these from the previous update
  const Feat t = {tg[0], tg[1], tg[2], tg[3], tg[4], tg[5]};
  m.t += dt;
  double ny, nz, dist;
  missile_guidance(m, pr, t, ny, nz, dist);
  missile_state_trans(m, pr, o, ny, nz, dt);
  st[0] = m.pn; st[1] = m.pe; st[2] = m.pu; st[3] = m.vn; st[4] = m.ve; st[5] = m.vu; st[6] = m.theta; st[7] = m.phi; st[8] = m.alt;
  st[9] = m.t; st[10] = m.m; st[11] = m.dtheta; st[12] = m.dphi;
  out[0] = ny; out[1] = nz; out[2] = dist;
}
void eh_posture(int ov, int rv, double AO, double TA, double R, double td, double* out) {
  out[0] = posture_orientation(ov, AO, TA); out[1] = posture_range(rv, R, td);
}
double eh_delta_heading(double target_deg, double psi_deg) { return delta_heading_deg(target_deg, psi_deg); }
}
