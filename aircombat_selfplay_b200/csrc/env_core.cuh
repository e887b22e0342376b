// env_core.cuh -- device-side building blocks of the env-step layer (sm_100a): keyed RNG, WGS-84 local-frame
// transforms, AO/TA/R geometry, the proportional-navigation missile, reward shaping functions.
//
// Replaces the reference's per-env Python (E/ = reference envs/JSBSim/): E/utils/utils.py (LLA2NEU, NEU2LLA,
// get_AO_TA_R), E/core/simulatior.py (MissileSimulator, ChaffSimulator), E/reward_functions/*.py.  pymap3d (an
// unpinned third-party dependency of the reference) is restated from its published algorithm.
#pragma once
#include <stdint.h>
#include <math.h>

#define ENV_DEV __device__ __forceinline__

// The missile / chaff path takes its quotients, roots and elementary functions from fmath.cuh (guard-free Newton
// sequences, coefficient tables in constant memory) like the FDM frame; -DACS_IEEE_MATH_MISSILE (or -DACS_IEEE_MATH)
// restores `/`, sqrt() and libdevice on this path alone (tuning builds).
#include "fmath.cuh"
#if defined(ACS_IEEE_MATH_MISSILE) && !defined(ACS_IEEE_MATH)
ENV_DEV double em_div(double a, double b) { return a / b; }
ENV_DEV double em_rcp(double b) { return 1.0 / b; }
ENV_DEV double em_sqrt(double x) { return sqrt(x); }
ENV_DEV double em_sqrt0(double x) { return sqrt(x); }
ENV_DEV double em_rsqrt(double x) { return 1.0 / sqrt(x); }
ENV_DEV double em_sin(double x) { return sin(x); }
ENV_DEV double em_exp(double x) { return exp(x); }
ENV_DEV void em_sincos(double x, double* s, double* c) { sincos(x, s, c); }
ENV_DEV double em_cos(double x) { return cos(x); }
ENV_DEV double em_acos(double x) { return acos(x); }
ENV_DEV double em_tanh(double x) { return tanh(x); }
ENV_DEV double em_atanh(double x) { return atanh(x); }
#else
ENV_DEV double em_div(double a, double b) { return fm_div(a, b); }
ENV_DEV double em_rcp(double b) { return fm_rcp(b); }
ENV_DEV double em_sqrt(double x) { return fm_sqrt(x); }
ENV_DEV double em_sqrt0(double x) { return fm_sqrt0(x); }
ENV_DEV double em_rsqrt(double x) { return fm_rsqrt(x); }
ENV_DEV double em_sin(double x) { return fm_sin(x); }
ENV_DEV double em_exp(double x) { return fm_exp(x); }
ENV_DEV void em_sincos(double x, double* s, double* c) { fm_sincos(x, s, c); }
ENV_DEV double em_cos(double x) { return fm_cos(x); }
ENV_DEV double em_acos(double x) { return fm_acos(x); }
ENV_DEV double em_tanh(double x) { return fm_tanh(x); }
ENV_DEV double em_atanh(double x) { return fm_atanh(x); }
#endif
// fmax / fmin without their NaN handling (sm_100a expands them into compare + selects + fix-up)
ENV_DEV double env_max(double a, double b) { return a > b ? a : b; }
ENV_DEV double env_min(double a, double b) { return a < b ? a : b; }

// ----------------------------------------------------------------------------- arenas (structure of arrays)
// per aircraft doubles
enum {
  AD_POS_N, AD_POS_E, AD_POS_U,        // AircraftSimulator._position (north, east, up) [m]          simulatior.py:245
  AD_VEL_N, AD_VEL_E, AD_VEL_D,        // _velocity: v_north, v_east, v_down [m/s], catalog-clipped  simulatior.py:253
  AD_H_SL_M,                           // position_h_sl_m, clipped [-500, 26000]                      catalog.py:292-338
  AD_U_MPS, AD_V_MPS, AD_W_MPS, AD_VC_MPS,
  AD_BLOODS,
  AD_HR_ROLL, AD_HR_P, AD_HR_Q,        // HeadingReward last roll/p/q                                 heading_reward.py
  AD_CH_N, AD_CH_E, AD_CH_U, AD_CH_T,  // the aircraft's current chaff group                          simulatior.py:327-391
  AD_PRE_REWARD0,                      // BaseRewardFunction.pre_rewards[agent] per reward            reward_function_base.py:56
  N_AD = AD_PRE_REWARD0 + ACS_MAX_REWARDS
};
// per aircraft ints
enum {
  AI_STATUS, AI_DIE_FLAG, AI_REM_MISSILES, AI_REM_9M, AI_REM_120B, AI_REM_GUN, AI_REM_CHAFF, AI_LAST_SHOOT_TIME,
  AI_LAST_SHOT_SLOT, AI_N_LAUNCHED, AI_LOCK_LO, AI_LOCK_HI, AI_LOCK_N, AI_PRE_REMAINING, AI_SHOOT, AI_CH_STATE,
  AI_CH_COUNT, N_AI
};
// per env doubles
enum {
  ED_TGT_HEADING, ED_TGT_ALT, ED_TGT_VEL, ED_CHECK_TIME, ED_CG_PREV0, ED_CG_PREV1,
  ED_TT_PREV0, ED_WD_PREV0 = ED_TT_PREV0 + ACS_MAX_AGENTS, N_ED = ED_WD_PREV0 + ACS_MAX_AGENTS
};
// per env ints
enum { EI_CURRENT_STEP, EI_EPISODE, EI_SUBSTEP_COUNT, EI_TURN_COUNTS, EI_PMV_REF, EI_CG_VALID, EI_TT_VALID, EI_WD_VALID,
       EI_ORDER_SEQ, EI_BORN_SEQ, EI_FAULTS, EI_DEFERRED, EI_STAGE, EI_CUR_BITS, EI_CUR_COUNT, N_EI };
// per missile slot doubles / ints
enum { MD_POS_N, MD_POS_E, MD_POS_U, MD_VEL_N, MD_VEL_E, MD_VEL_U, MD_THETA, MD_PHI, MD_ALT, MD_T, MD_M, MD_DTHETA,
       MD_DPHI, MD_D_PREV, MD_SIN_THETA, MD_COS_THETA, N_MD };
enum { MI_STATUS, MI_KIND, MI_TARGET, MI_CONSEC, MI_ORDER, MI_BORN, MI_KEYN, MI_DETACHED, N_MI };

static const char* const AD_NAMES[] = {"pos_n", "pos_e", "pos_u", "vel_n", "vel_e", "vel_d", "h_sl_m", "u_mps", "v_mps", "w_mps",
  "vc_mps", "bloods", "hr_roll", "hr_p", "hr_q", "chaff_n", "chaff_e", "chaff_u", "chaff_t", "pre_reward0", "pre_reward1",
  "pre_reward2", "pre_reward3", "pre_reward4", "pre_reward5", "pre_reward6", "pre_reward7", "pre_reward8", "pre_reward9",
  "pre_reward10", "pre_reward11"};
static const char* const AI_NAMES[] = {"status", "die_flag", "rem_missiles", "rem_9m", "rem_120b", "rem_gun", "rem_chaff",
  "last_shoot_time", "last_shot_slot", "n_launched", "lock_lo", "lock_hi", "lock_n", "pre_remaining", "shoot", "chaff_state",
  "chaff_count"};
static const char* const ED_NAMES[] = {"tgt_heading_deg", "tgt_altitude_ft", "tgt_velocity_mps", "check_time", "cg_prev_ao",
  "cg_prev_ta", "tt_prev0", "tt_prev1", "tt_prev2", "tt_prev3", "tt_prev4", "tt_prev5", "tt_prev6", "tt_prev7", "wd_prev0",
  "wd_prev1", "wd_prev2", "wd_prev3", "wd_prev4", "wd_prev5", "wd_prev6", "wd_prev7"};
static const char* const EI_NAMES[] = {"current_step", "episode", "substep_count", "turn_counts", "pmv_ref", "cg_valid", "tt_valid",
  "wd_valid", "order_seq", "born_seq", "faults", "deferred", "stage", "curriculum_record", "curriculum_count"};
static const char* const MD_NAMES[] = {"pos_n", "pos_e", "pos_u", "vel_n", "vel_e", "vel_u", "theta", "phi", "alt", "t", "m",
  "dtheta", "dphi", "d_prev", "sin_theta", "cos_theta"};
static const char* const MI_NAMES[] = {"status", "kind", "target", "consec", "order", "born", "keyn", "detached"};
static_assert(sizeof(AD_NAMES) / sizeof(AD_NAMES[0]) == N_AD, "AD names");
static_assert(sizeof(AI_NAMES) / sizeof(AI_NAMES[0]) == N_AI, "AI names");
static_assert(sizeof(ED_NAMES) / sizeof(ED_NAMES[0]) == N_ED, "ED names");
static_assert(sizeof(EI_NAMES) / sizeof(EI_NAMES[0]) == N_EI, "EI names");
static_assert(sizeof(MD_NAMES) / sizeof(MD_NAMES[0]) == N_MD, "MD names");
static_assert(sizeof(MI_NAMES) / sizeof(MI_NAMES[0]) == N_MI, "MI names");

enum { ST_ALIVE = 0, ST_CRASH = 1, ST_SHOTDOWN = 2 };                     // AircraftSimulator status, simulatior.py:93-95
enum { MS_INACTIVE = -1, MS_LAUNCHED = 0, MS_HIT = 1, MS_MISS = 2 };      // MissileSimulator status, simulatior.py:395-398
enum { CH_NONE = 0, CH_ACTIVE = 1, CH_DONE = 2 };
enum { RNG_RESET = 1, RNG_HEADING = 2, RNG_CHAFF = 3 };

struct EnvView {
  double* fdm;   // [N_STATE][rows]
  double* out;   // [FDM_N_OUT][rows]
  double* ad; int* ai;     // [N_AD][rows], [N_AI][rows]
  double* ed; int* ei;     // [N_ED][B],   [N_EI][B]
  double* md; int* mi;     // [N_MD][rows*S], [N_MI][rows*S]
  int B, A, S, rows;
  double* traj;  // [K][6][rows] position / velocity every aircraft published after each substep of the current step, for
                 // the envs whose missiles are integrated by k_env_missiles (nullptr: the task has no missiles)
  double* snap;  // [K][N_STATE + 4][rows] FDM state (+ the frame's kept scalars) after each substep, written only for aircraft a
                 // missile may hit within the step: the hit restores the state of the substep it happened in
};
// Reset template of a handle (acs.cu, build_reset_template).  Every task but the heading task resets an env to the same
// state each time (fixed per-lane initial conditions, no random draw), so reset() is run once on a one-env arena `t`.
// full != 0: the words reset() writes (found by running it over two different fill patterns) are kept packed --
// v64[i] / v32[i] is the value of word d64[i] / d32[i] = arena << 24 | field << 12 | index within the env -- and a reset
// is a warp-cooperative scatter of those words plus the reset observation (reset_copy_warp).
struct ResetTpl {
  EnvView t;            // t.fdm == nullptr: no template.  full == 0: only the FDM reload (fdm / out / derived ad fields) is used
  const double* obs;    // [A][obs_dim] reset observation
  const double* v64; const int* d64; int n64;   // arenas 0 fdm, 1 out, 2 ad, 4 ed, 6 md
  const int* v32; const int* d32; int n32;      // arenas 3 ai, 5 ei, 7 mi
  int full;
  const char* base;     // the template is one contiguous block [base, base + bytes)
  int bytes;
  // curriculum stages (acs_env_set_stage_init_states): stage s > 0 keeps its own values -- reset observation, v64, v32 -- at
  // stage_base + s * stage_stride (+ 0, stage_off64, stage_off32); the word lists d64 / d32 are those of stage 0
  const char* stage_base; int stage_stride, stage_off64, stage_off32, n_stages;
};
// The template is a few KB read by the few warps that reset an env, i.e. cold; a warp asks for all of its lines at once.
__device__ __forceinline__ void tpl_prefetch(const ResetTpl& tp) {
  for (int o = (threadIdx.x & 31) * 128; o < tp.bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp.base + o));
}
#define AD(v, f, row) (v).ad[(size_t)(f) * (v).rows + (row)]
#define AI(v, f, row) (v).ai[(size_t)(f) * (v).rows + (row)]
#define ED(v, f, env) (v).ed[(size_t)(f) * (v).B + (env)]
#define EI(v, f, env) (v).ei[(size_t)(f) * (v).B + (env)]
#define MD(v, f, m) (v).md[(size_t)(f) * (v).rows * (v).S + (m)]
#define MI(v, f, m) (v).mi[(size_t)(f) * (v).rows * (v).S + (m)]
#define OUTF(v, f, row) (v).out[(size_t)(f) * (v).rows + (row)]

// indices into the FDM output arena (order of FDM_OUT_FIELDS in fdm_core.cuh)
enum { O_LON, O_LAT, O_H_SL_FT, O_ROLL, O_PITCH, O_HEADING, O_VN, O_VE, O_VD, O_U, O_V, O_W, O_VC, O_NPX, O_NPY, O_NPZ, O_P, O_Q,
       O_R, O_ECI_VMAG, O_GEOD_ALT, O_ALPHA, O_BETA, O_MACH, O_THRUST };

ENV_DEV double env_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ----------------------------------------------------------------------------- keyed counter RNG
// Replaces np.random.rand() (E/envs/env_base.py:153), env.np_random.uniform (E/envs/singlecontrol_env.py:35-37,
// E/termination_conditions/unreach_heading.py:45-47).  Stateless: a splitmix64 hash chain over the key.
ENV_DEV uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ENV_DEV double env_u01(uint64_t seed, int64_t env, int64_t purpose, int64_t a, int64_t b, int64_t c) {
  uint64_t h = splitmix64(seed);
  h = splitmix64(h ^ (uint64_t)env);
  h = splitmix64(h ^ (uint64_t)purpose);
  h = splitmix64(h ^ (uint64_t)a);
  h = splitmix64(h ^ (uint64_t)b);
  h = splitmix64(h ^ (uint64_t)c);
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

// ----------------------------------------------------------------------------- WGS-84 local frames (pymap3d restated)
static constexpr double WGS84_A = 6378137.0;
static constexpr double WGS84_F = 1.0 / 298.257223563;
static constexpr double WGS84_B = WGS84_A * (1.0 - WGS84_F);
static constexpr double ENV_DEG2RAD = 3.14159265358979323846 / 180.0;

ENV_DEV void geodetic2ecef(double lat_deg, double lon_deg, double alt, double& x, double& y, double& z) {
  double sla, cla, slo, clo;
  em_sincos(lat_deg * ENV_DEG2RAD, &sla, &cla);
  em_sincos(lon_deg * ENV_DEG2RAD, &slo, &clo);
  const double ac = WGS84_A * cla, bs = WGS84_B * sla;
  const double n = (WGS84_A * WGS84_A) * em_rsqrt(ac * ac + bs * bs);      // a^2 / hypot(a cos, b sin); no overflow at 6e6 m
  x = (n + alt) * cla * clo;
  y = (n + alt) * cla * slo;
  z = (n * ((WGS84_B / WGS84_A) * (WGS84_B / WGS84_A)) + alt) * sla;
}
// precomputed once per kernel: the battle-field origin
struct GeoOrigin { double x0, y0, z0, sla, cla, slo, clo; };
ENV_DEV GeoOrigin geo_origin(double lon0, double lat0, double alt0) {
  GeoOrigin o;
  geodetic2ecef(lat0, lon0, alt0, o.x0, o.y0, o.z0);
  em_sincos(lat0 * ENV_DEG2RAD, &o.sla, &o.cla);
  em_sincos(lon0 * ENV_DEG2RAD, &o.slo, &o.clo);
  return o;
}
// LLA2NEU (E/utils/utils.py:30-41): pymap3d.geodetic2ned -> (north, east, up)
ENV_DEV void lla2neu(const GeoOrigin& o, double lon, double lat, double alt, double& north, double& east, double& up) {
  double x, y, z;
  geodetic2ecef(lat, lon, alt, x, y, z);
  const double u = x - o.x0, v = y - o.y0, w = z - o.z0;
  const double t = o.clo * u + o.slo * v;
  east = -o.slo * u + o.clo * v;
  up = o.cla * t + o.sla * w;
  north = -o.sla * t + o.cla * w;
}
// altitude component of NEU2LLA (E/utils/utils.py:44-55, pymap3d.ned2geodetic): enu2ecef, then the geodetic height.
// Only the height is consumed on the hot path (missile air density, rho = 1.225 exp(-h / 9300)); it is computed with
// Fukushima's (2006) trig-free Halley step -- the same closed form JSBSim's FGLocation uses -- which agrees with the
// iteration pymap3d uses (You 2000, restated in the oracle) to < 1e-8 m at flight altitudes.
ENV_DEV double neu2alt(const GeoOrigin& o, double n, double e, double u) {
  const double t = o.cla * u - o.sla * n;
  const double w = o.sla * u + o.cla * n;
  const double uu = o.clo * t - o.slo * e;
  const double vv = o.slo * t + o.clo * e;
  const double x = o.x0 + uu, y = o.y0 + vv, z = o.z0 + w;
  const double a = WGS84_A, ec = WGS84_B / WGS84_A, ec2 = ec * ec, e2 = 1.0 - ec2, c = a * e2;
  const double rxy = em_sqrt(x * x + y * y);      // (never on the polar axis: fmath.cuh, operands > 0)
  const double s0 = fabs(z), zc = ec * s0, c0 = ec * rxy, c02 = c0 * c0, s02 = s0 * s0, a02 = c02 + s02;
  const double a0 = em_sqrt(a02), a03 = a02 * a0;
  double s1 = zc * a03 + c * s02 * s0;
  const double c1 = rxy * a03 - c * c02 * c0, cs0c0 = c * c0 * s0;
  const double b0 = 1.5 * cs0c0 * ((rxy * s0 - zc * c0) * a0 - cs0c0);
  s1 = s1 * a03 - b0 * s0;
  const double cc = ec * (c1 * a03 - b0 * c0);
  const double s12 = s1 * s1, cc2 = cc * cc;
  return (rxy * cc + s0 * s1 - a * em_sqrt(ec2 * s12 + cc2)) * em_rsqrt(s12 + cc2);
}

// ----------------------------------------------------------------------------- AO / TA / R (E/utils/utils.py:58-103)
struct Feat { double n, e, u, vn, ve, vd; };
struct AoTaR { double AO, TA, R, side; };
ENV_DEV AoTaR get_ao_ta_r(const Feat& ego, const Feat& enm, bool two_d) {
  const double dx = enm.n - ego.n, dy = enm.e - ego.e, dz = enm.u - ego.u;
  double ego_v, enm_v, R, p1, p2;
  if (two_d) {
    ego_v = em_sqrt0(ego.vn * ego.vn + ego.ve * ego.ve);
    enm_v = em_sqrt0(enm.vn * enm.vn + enm.ve * enm.ve);
    R = em_sqrt0(dx * dx + dy * dy);
    p1 = dx * ego.vn + dy * ego.ve;
    p2 = dx * enm.vn + dy * enm.ve;
  } else {
    ego_v = em_sqrt0(ego.vn * ego.vn + ego.ve * ego.ve + ego.vd * ego.vd);
    enm_v = em_sqrt0(enm.vn * enm.vn + enm.ve * enm.ve + enm.vd * enm.vd);
    R = em_sqrt0(dx * dx + dy * dy + dz * dz);
    p1 = dx * ego.vn + dy * ego.ve + dz * ego.vd;
    p2 = dx * enm.vn + dy * enm.ve + dz * enm.vd;
  }
  AoTaR g;
  g.AO = em_acos(env_clip(em_div(p1, R * ego_v + 1e-8), -1.0, 1.0));
  g.TA = em_acos(env_clip(em_div(p2, R * enm_v + 1e-8), -1.0, 1.0));
  g.R = R;
  const double cr = ego.vn * dy - ego.ve * dx;
  g.side = (double)((cr > 0) - (cr < 0));
  return g;
}

// ----------------------------------------------------------------------------- missile (E/core/simulatior.py:393-608)
struct MissileParams { double g, t_max, t_thrust, Isp, Length, Diameter, cD, m0, dm, K, nyz_max, Rc, v_min, inv_t_max; };
// kind 0: base class numbers (AIM-9L, :420-433); kind 1: AIM_9M / AIM_120B subclasses, both AIM-120B numbers (:663-712)
ENV_DEV MissileParams missile_params(int kind) {
  MissileParams p;
  p.g = 9.81; p.dm = 6.0; p.v_min = 150.0;
  if (kind == 0) { p.t_max = 60.0; p.inv_t_max = 1.0 / 60.0; p.t_thrust = 3.0; p.Isp = 120.0; p.Length = 2.87; p.Diameter = 0.127; p.cD = 0.4; p.m0 = 84.0; p.K = 3.0; p.nyz_max = 30.0; p.Rc = 300.0; }
  else { p.t_max = 27.22; p.inv_t_max = 1.0 / 27.22; p.t_thrust = 1.4; p.Isp = 1837.0; p.Length = 3.66; p.Diameter = 0.18; p.cD = 0.02; p.m0 = 152.0; p.K = 5.0; p.nyz_max = 50.0; p.Rc = 5.0; }
  return p;
}
struct Missile {
  double pn, pe, pu, vn, ve, vu, theta, phi, alt, t, m, dtheta, dphi, d_prev, st, ct;   // st, ct = sin/cos(theta), cached
  int status, kind, target, consec;
};
ENV_DEV void missile_load(const EnvView& v, int mid, Missile& m) {
  m.pn = MD(v, MD_POS_N, mid); m.pe = MD(v, MD_POS_E, mid); m.pu = MD(v, MD_POS_U, mid);
  m.vn = MD(v, MD_VEL_N, mid); m.ve = MD(v, MD_VEL_E, mid); m.vu = MD(v, MD_VEL_U, mid);
  m.theta = MD(v, MD_THETA, mid); m.phi = MD(v, MD_PHI, mid); m.alt = MD(v, MD_ALT, mid); m.t = MD(v, MD_T, mid);
  m.m = MD(v, MD_M, mid); m.dtheta = MD(v, MD_DTHETA, mid); m.dphi = MD(v, MD_DPHI, mid); m.d_prev = MD(v, MD_D_PREV, mid);
  m.st = MD(v, MD_SIN_THETA, mid); m.ct = MD(v, MD_COS_THETA, mid);
  m.status = MI(v, MI_STATUS, mid); m.kind = MI(v, MI_KIND, mid); m.target = MI(v, MI_TARGET, mid); m.consec = MI(v, MI_CONSEC, mid);
}
ENV_DEV void missile_store(const EnvView& v, int mid, const Missile& m) {
  MD(v, MD_POS_N, mid) = m.pn; MD(v, MD_POS_E, mid) = m.pe; MD(v, MD_POS_U, mid) = m.pu;
  MD(v, MD_VEL_N, mid) = m.vn; MD(v, MD_VEL_E, mid) = m.ve; MD(v, MD_VEL_U, mid) = m.vu;
  MD(v, MD_THETA, mid) = m.theta; MD(v, MD_PHI, mid) = m.phi; MD(v, MD_ALT, mid) = m.alt; MD(v, MD_T, mid) = m.t;
  MD(v, MD_M, mid) = m.m; MD(v, MD_DTHETA, mid) = m.dtheta; MD(v, MD_DPHI, mid) = m.dphi; MD(v, MD_D_PREV, mid) = m.d_prev;
  MD(v, MD_SIN_THETA, mid) = m.st; MD(v, MD_COS_THETA, mid) = m.ct;
  MI(v, MI_STATUS, mid) = m.status; MI(v, MI_CONSEC, mid) = m.consec;
}
// distance to the target: the R returned by _guidance (:563)
ENV_DEV double missile_distance(const Missile& m, const Feat& tg) {
  const double ax = m.pn - tg.n, ay = m.pe - tg.e, az = tg.u - m.pu;
  return em_sqrt0(ax * ax + ay * ay + az * az);
}
// _guidance (:556-576): proportional navigation; returns clipped (ny, nz) and the distance
ENV_DEV void missile_guidance(const Missile& m, const MissileParams& pr, const Feat& tg, double& ny, double& nz, double& dist) {
  // divisions by constants are multiplications by their reciprocals, quotients and roots the guard-free sequences of
  // fmath.cuh (<= 1 ulp from the reference's operations each)
  const double v_m = em_sqrt(m.vn * m.vn + m.ve * m.ve + m.vu * m.vu);
  // theta_m = arcsin(dz/v) enters only through its cosine (:569-570): cos(arcsin(x)) = |v_xy| / |v|
  const double ex = m.pn - tg.n, ey = m.pe - tg.e;
  const double Rxy = em_sqrt(ex * ex + ey * ey);
  const double ez = tg.u - m.pu;
  const double Rxyz = em_sqrt(ex * ex + ey * ey + ez * ez);
  const double dxt = tg.n - m.pn, dyt = tg.e - m.pe, dzt = tg.u - m.pu;
  const double dvx = tg.vn - m.vn, dvy = tg.ve - m.ve, dvz = tg.vd - m.vu;
  const double dbeta = em_div(dvy * dxt - dvx * dyt, Rxy * Rxy);
  const double deps = em_div(dvz * (Rxy * Rxy) - dzt * (dxt * dvx + dyt * dvy), (Rxyz * Rxyz) * Rxy);
  const double K0 = pr.K * (pr.t_max - m.t) * pr.inv_t_max, K = K0 > 0.0 ? K0 : 0.0;
  const double ct = em_div(em_sqrt0(m.vn * m.vn + m.ve * m.ve), v_m);
  const double Kvg = K * v_m * (1.0 / 9.81);       // pr.g is 9.81 for every missile kind
  ny = env_clip(Kvg * ct * dbeta, -pr.nyz_max, pr.nyz_max);
  nz = env_clip(Kvg * deps + ct, -pr.nyz_max, pr.nyz_max);
  dist = Rxyz;
}
// _state_trans (:578-608)
ENV_DEV void missile_state_trans(Missile& m, const MissileParams& pr, const GeoOrigin& org, double ny, double nz, double dt) {
  m.pn = m.pn + dt * m.vn; m.pe = m.pe + dt * m.ve; m.pu = m.pu + dt * m.vu;
  m.alt = neu2alt(org, m.pn, m.pe, m.pu);
  double v = em_sqrt(m.vn * m.vn + m.ve * m.ve + m.vu * m.vu);
  double theta = m.theta, phi = m.phi;
  const double Isp = m.t < pr.t_thrust ? pr.Isp : 0.0;
  const double T = pr.g * Isp * pr.dm;
  double S0 = 3.14159265358979323846 * ((pr.Diameter / 2) * (pr.Diameter / 2));
  const double sdt = em_sin(m.dtheta), sdp = em_sin(m.dphi);
  S0 += em_sqrt0(sdt * sdt + sdp * sdp) * pr.Diameter * pr.Length;
  const double rho = 1.225 * em_exp(m.alt * (-1.0 / 9300.0));
  const double D = 0.5 * pr.cD * S0 * rho * (v * v);
  const double nx = em_div(T - D, m.m * pr.g);
  double st = m.st, ct = m.ct;                 // sin/cos of the current pitch angle, kept from the previous update
  const double dv = pr.g * (nx - st);
  const double gv = pr.g * em_rcp(v);
  m.dphi = gv * em_div(ny, ct);
  m.dtheta = gv * (nz - ct);
  v += dt * dv;
  phi += dt * m.dphi;
  theta += dt * m.dtheta;
  double sp, cp;
  em_sincos(theta, &st, &ct);
  em_sincos(phi, &sp, &cp);
  m.vn = v * ct * cp; m.ve = v * ct * sp; m.vu = v * st;
  m.theta = theta; m.phi = phi; m.st = st; m.ct = ct;
  if (m.t < pr.t_thrust) m.m = m.m - dt * pr.dm;
}

// ----------------------------------------------------------------------------- reward shaping (E/reward_functions/posture_reward.py:51-75)
ENV_DEV double posture_orientation(int version, double AO, double TA) {
  // quotients by constants as reciprocal multiplies, tanh / atanh / exp from fmath.cuh: <= 1e-15 from the reference's
  // expressions, compared at 1e-6
  const double PI = 3.14159265358979323846, I2PI = 1.0 / (2 * PI);
  const double ta_term = em_atanh(1.0 - env_max((2 / PI) * TA, 1e-4)) * I2PI;
  if (version == 0) return (1.0 - em_tanh(9 * (AO - PI / 9))) * (1.0 / 3.0) + 1 / 3.0 + env_min(ta_term, 0.0) + 0.5;
  if (version == 1) return (1.0 - em_tanh(2 * (AO - PI / 2))) * 0.5 * ta_term + 0.5;
  return em_rcp((50 / PI) * AO + 2) + 1 / 2.0 + env_min(ta_term, 0.0) + 0.5;
}
ENV_DEV double posture_range(int version, double R, double td) {
  if (version == 0) return em_div(em_exp(-((R - td) * (R - td)) * 0.004), 1.0 + em_exp(-(R - td + 2) * 2));
  if (version == 1 || version == 2) {
    const double base = env_clip(em_div(1.2 * env_min(em_exp(-(R - td) * 0.21), 1.0), 1.0 + em_exp(-(R - td + 1) * 0.8)), 0.3, 1.0);
    if (version == 1) return base;
    const double sg = (double)((7 - R > 0) - (7 - R < 0));
    return env_max(base, sg);
  }
  return 1.0 * (R < 5) + (R >= 5) * env_clip(-0.032 * (R * R) + 0.284 * R + 0.38, 0.0, 1.0) + env_clip(em_exp(-0.16 * R), 0.0, 0.2);
}
// delta heading in [-180, 180] as the catalog derives it (E/core/catalog.py update_delta_heading) and clips it
ENV_DEV double delta_heading_deg(double target_deg, double psi_deg) {
  double ang = fmod(target_deg - psi_deg, 360.0);
  if (ang < 0) ang += 360.0;   // python % returns a non-negative result for a positive modulus
  if (ang > 180) ang -= 360;
  return env_clip(ang, -180.0, 180.0);
}
