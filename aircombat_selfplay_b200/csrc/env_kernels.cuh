// env_kernels.cuh -- the env-layer kernels (sm_100a).  The G = 1/2/4/8 lanes of one environment are adjacent inside a warp, so
// cross-aircraft exchange is shared memory + __syncwarp on the group's lane mask:
//
//   k_env_substeps[_split|_split3|_split4]
//                    actions -> controls, then the agent_interaction_steps substep loop with the FDM state in registers:
//                    aircraft run(), missile run() (PN guidance + fused proximity fuze), chaff run(), chaff x missile test
//                    (E/envs/env_base.py:131-154, E/core/simulatior.py:210-229,520-533).  One thread per aircraft, or -- for
//                    batches too small to fill the machine -- two / three / four threads in different warps per aircraft
//                    running the stages of one frame concurrently (acs.cu picks by batch size; same expressions in all)
//   k_env_post       task.step (weapon launches), get_obs, get_reward, get_termination, packing
//                    (E/envs/env_base.py:155-173, E/envs/multiplecombat_env.py:161-182), and -- fused -- the auto-reset of
//                    every env whose agents are all done, as a scatter of the handle's reset template
//   k_env_reset_fdm, k_env_reset_task
//                    masked per-env reset: sim.reload() for every aircraft, task.reset, reward resets, get_obs
//                    (E/envs/env_base.py:98-113); with the mask = "all agents done" this is the VecEnv auto-reset
//                    (R/envs/env_wrappers.py:191-204).  Used by explicit reset(), to build the reset template, and by
//                    the heading task (random initial conditions: no template)
#pragma once

static constexpr int F_SIM_TIME = FDM_N_CORE - 1;  // "sim_time" is the last core field
static constexpr int F_CMD0 = FDM_N_CORE;          // fcs/{aileron,elevator,rudder,throttle}-cmd-norm are the first carried fields

struct Lane {
  int tid, env, lane, row, gbase;
  unsigned gmask;
  bool valid;
};
// lg = log2 of the lanes per env (1, 2, 4 or 8 lanes)
ENV_DEV Lane lane_setup(const EnvView& v, const int lg) {
  Lane L;
  const int G = 1 << lg;
  L.tid = threadIdx.x;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  L.env = gid >> lg;
  L.lane = gid & (G - 1);
  L.gbase = L.tid - L.lane;
  L.valid = (L.env < v.B) && (L.lane < v.A);
  L.row = L.env * v.A + L.lane;
  const unsigned wl = threadIdx.x & 31;
  L.gmask = ((1u << G) - 1u) << (wl & ~(unsigned)(G - 1));
  return L;
}

// what every aircraft publishes to the other lanes of its env
struct PubAc {
  Feat f;            // position (n, e, u) and velocity (vn, ve, vd), as AircraftSimulator.get_position / get_velocity
  double h, u_mps, bloods;
  int status;
};
struct PubChaff { double n, e, u; int state, count; };

ENV_DEV bool same_team(const AcsTaskConfig& c, int a, int b) { return (a < c.n_ego) == (b < c.n_ego); }

// derived catalog properties of an aircraft from the FDM outputs of its last frame (E/core/catalog.py:292-338: the
// *_mps / h_sl_m properties are stored through set_property_value, hence clipped)
ENV_DEV void derive_aircraft(const AcOut& o, const GeoOrigin& org, PubAc& p, double& v_mps, double& w_mps, double& vc_mps) {
  p.h = env_clip(o.h_sl_ft * 0.3048, -500.0, 26000.0);
  lla2neu(org, o.lon_deg, o.lat_geod_deg, p.h, p.f.n, p.f.e, p.f.u);
  p.f.vn = env_clip(o.vn * 0.3048, -700.0, 700.0);
  p.f.ve = env_clip(o.ve * 0.3048, -700.0, 700.0);
  p.f.vd = env_clip(o.vd * 0.3048, -700.0, 700.0);
  p.u_mps = env_clip(o.u * 0.3048, -700.0, 700.0);
  v_mps = env_clip(o.v * 0.3048, -700.0, 700.0);
  w_mps = env_clip(o.w * 0.3048, -700.0, 700.0);
  vc_mps = env_clip(o.vc_fps * 0.3048, 0.0, 1400.0);
}
// The same publication straight from the frame scratch, for the substeps in which only the other lanes' missiles read
// it: the geodetic sines and cosines are already there (Fukushima), so no inverse trigonometry and no sincos.
ENV_DEV void publish_from_frame(const Frame& f, const GeoOrigin& org, PubAc& p) {
  p.h = env_clip(f.h_asl * 0.3048, -500.0, 26000.0);
  const double ac = WGS84_A * f.cosLatGd, bs = WGS84_B * f.sinLatGd;
  const double n = (WGS84_A * WGS84_A) * em_rsqrt(ac * ac + bs * bs);      // a^2 / hypot(a cos, b sin); no overflow at 6e6 m
  const double x = (n + p.h) * f.cosLatGd * f.cosLon, y = (n + p.h) * f.cosLatGd * f.sinLon;
  const double z = (n * ((WGS84_B / WGS84_A) * (WGS84_B / WGS84_A)) + p.h) * f.sinLatGd;
  const double u = x - org.x0, v = y - org.y0, w = z - org.z0;
  const double t = org.clo * u + org.slo * v;
  p.f.e = -org.slo * u + org.clo * v;
  p.f.u = org.cla * t + org.sla * w;
  p.f.n = -org.sla * t + org.cla * w;
  p.f.vn = env_clip(f.vel.x * 0.3048, -700.0, 700.0);
  p.f.ve = env_clip(f.vel.y * 0.3048, -700.0, 700.0);
  p.f.vd = env_clip(f.vel.z * 0.3048, -700.0, 700.0);
  p.u_mps = env_clip(f.uvw.x * 0.3048, -700.0, 700.0);
}
ENV_DEV void store_derived(const EnvView& v, int row, const PubAc& p, double v_mps, double w_mps, double vc_mps) {
  AD(v, AD_POS_N, row) = p.f.n; AD(v, AD_POS_E, row) = p.f.e; AD(v, AD_POS_U, row) = p.f.u;
  AD(v, AD_VEL_N, row) = p.f.vn; AD(v, AD_VEL_E, row) = p.f.ve; AD(v, AD_VEL_D, row) = p.f.vd;
  AD(v, AD_H_SL_M, row) = p.h; AD(v, AD_U_MPS, row) = p.u_mps; AD(v, AD_V_MPS, row) = v_mps; AD(v, AD_W_MPS, row) = w_mps;
  AD(v, AD_VC_MPS, row) = vc_mps;
}
ENV_DEV void load_pub(const EnvView& v, int row, PubAc& p) {
  p.f.n = AD(v, AD_POS_N, row); p.f.e = AD(v, AD_POS_E, row); p.f.u = AD(v, AD_POS_U, row);
  p.f.vn = AD(v, AD_VEL_N, row); p.f.ve = AD(v, AD_VEL_E, row); p.f.vd = AD(v, AD_VEL_D, row);
  p.h = AD(v, AD_H_SL_M, row); p.u_mps = AD(v, AD_U_MPS, row); p.bloods = AD(v, AD_BLOODS, row);
  p.status = AI(v, AI_STATUS, row);
}

// ============================================================================================== substep kernel
// Which launched missiles still need run(): every one that can still move or change status.  A missile is INERT -- run()
// can no longer change anything observable -- once it is MISS and (its target is dead | t > t_max | |v| < v_min): all three
// are permanent within an episode and send every later run() into the MISS branch without a state transition
// (E/core/simulatior.py:528-531; only t, distance_pre and the increment window keep changing, and nothing reads them).
// A HIT missile is NOT inert: its next run() turns it MISS (target no longer alive), which the step's rewards observe.
ENV_DEV bool missile_inert(const EnvView& v, const int mid, const int env) {
  if (MI(v, MI_STATUS, mid) != MS_MISS) return false;
  const MissileParams pr = missile_params(MI(v, MI_KIND, mid));
  if (AI(v, AI_STATUS, env * v.A + MI(v, MI_TARGET, mid)) != ST_ALIVE) return true;
  if (MD(v, MD_T, mid) > pr.t_max) return true;
  const double vn = MD(v, MD_VEL_N, mid), ve = MD(v, MD_VEL_E, mid), vu = MD(v, MD_VEL_U, mid);
  return em_sqrt(vn * vn + ve * ve + vu * vu) < pr.v_min;
}
// bit s = slot s of this aircraft holds a missile that still runs (slots >= 64 cannot exist: acs_env_create caps S)
ENV_DEV unsigned long long live_missiles(const EnvView& v, const Lane& L) {
  unsigned long long live = 0;
  const int nl = AI(v, AI_N_LAUNCHED, L.row);
  for (int s = 0; s < nl; s++) {
    const int mid = L.row * v.S + s;
    if (MI(v, MI_DETACHED, mid) || missile_inert(v, mid, L.env)) continue;
    live |= 1ull << s;
  }
  return live;
}

// The missile phase of one substep for the lanes of one env (E/envs/env_base.py:141-154).  Reference order inside a
// substep: every aircraft run(), then every missile run() in dict order, then every chaff run(), then the
// chaff x missile test.  Missiles are processed in parallel by their shooter's lane, each lane walking its own list of
// live slots (so a warp iterates max-over-lanes of the LIVE count, not of the launched count); the only order-dependent
// outcome -- two missiles inside the fuze radius of one target in the same substep, where the first in dict order scores
// the HIT and later ones see a dead target and go MISS -- is resolved with a shared-memory atomicMin on the dict order.
// The chaff x missile test of a missile depends on nothing but that missile's position after its own run() and the
// chaff clouds (which no missile influences), so chaff run() is moved in front of the missiles' and the test is folded
// into the same traversal as run(): one pass fewer, and none at all while the env has no effective chaff.
__device__ __noinline__ void missile_phase(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const GeoOrigin& org,
                                           PubAc* sP, int* sWin, int* sShot, PubChaff* sCh, int substep_count,
                                           unsigned long long& live_io) {
  const double dt = cfg.sim_dt;
  const int maxlen = (int)(5.0 / dt);
  unsigned long long live = live_io;
  // ---- phase 1: who is inside its fuze radius (uses the state before this substep's missile moves)
  for (unsigned long long m = live; m; m &= m - 1) {
    const int mid = L.row * v.S + (__ffsll((long long)m) - 1);
    const int target = MI(v, MI_TARGET, mid);
    const Feat& tg = sP[L.gbase + target].f;
    const double ax = MD(v, MD_POS_N, mid) - tg.n, ay = MD(v, MD_POS_E, mid) - tg.e, az = tg.u - MD(v, MD_POS_U, mid);
    const double d = em_sqrt0(ax * ax + ay * ay + az * az);
    if (d < missile_params(MI(v, MI_KIND, mid)).Rc && MI(v, MI_STATUS, mid) != MS_MISS) atomicMin(&sWin[L.gbase + target], MI(v, MI_ORDER, mid));
  }
  // ---- chaff run() (:377-381) and publication
  {
    PubChaff c;
    c.state = CH_NONE; c.count = 0; c.n = c.e = c.u = 0.0;
    if (L.valid) {
      c.state = AI(v, AI_CH_STATE, L.row);
      if (c.state != CH_NONE) {
        const double t = AD(v, AD_CH_T, L.row) + dt;
        AD(v, AD_CH_T, L.row) = t;
        if (t > 20.0) c.state = CH_DONE;
        AI(v, AI_CH_STATE, L.row) = c.state;
        c.count = AI(v, AI_CH_COUNT, L.row);
        c.n = AD(v, AD_CH_N, L.row); c.e = AD(v, AD_CH_E, L.row); c.u = AD(v, AD_CH_U, L.row);
      }
    }
    sCh[L.tid] = c;
  }
  __syncwarp(L.gmask);                              // sWin and sCh of the whole env are complete
  const bool any_chaff = (__ballot_sync(L.gmask, sCh[L.tid].state == CH_ACTIVE) & L.gmask) != 0;
  // ---- phase 2: MissileSimulator.run() (:520-533), then this missile against every effective chaff (E/envs/env_base.py:146-154)
  for (unsigned long long mm = live; mm; mm &= mm - 1) {
    const int slot = __ffsll((long long)mm) - 1;
    const int mid = L.row * v.S + slot;
    Missile m;
    missile_load(v, mid, m);
    const MissileParams pr = missile_params(m.kind);
    const PubAc& tg = sP[L.gbase + m.target];
    m.t += dt;
    double ny, nz, dist;
    missile_guidance(m, pr, tg.f, ny, nz, dist);
    m.consec = (dist > m.d_prev) ? m.consec + 1 : 0;
    m.d_prev = dist;
    const int order = MI(v, MI_ORDER, mid);
    const bool target_alive = (tg.status == ST_ALIVE) && !(sWin[L.gbase + m.target] < order);
    if (dist < pr.Rc && target_alive && m.status != MS_MISS) {
      m.status = MS_HIT;
      sShot[L.gbase + m.target] = 1;
    } else if (m.t > pr.t_max || em_sqrt(m.vn * m.vn + m.ve * m.ve + m.vu * m.vu) < pr.v_min || m.consec >= maxlen || !target_alive) {
      // inert from here on (see missile_inert); `consec >= maxlen` alone is not permanent
      if (m.status == MS_MISS && (m.t > pr.t_max || !target_alive || em_sqrt(m.vn * m.vn + m.ve * m.ve + m.vu * m.vu) < pr.v_min)) live &= ~(1ull << slot);
      m.status = MS_MISS;
    } else {
      missile_state_trans(m, pr, org, ny, nz, dt);
    }
    if (any_chaff && m.status == MS_LAUNCHED) {
      const int keyn = MI(v, MI_KEYN, mid);
      // the episode is part of the key: substep counters restart at every reset
      const int64_t when = ((int64_t)EI(v, EI_EPISODE, L.env) << 20) + substep_count;
      bool missed = false;
      for (int j = 0; j < v.A; j++) {
        const PubChaff& c = sCh[L.gbase + j];
        if (c.state != CH_ACTIVE) continue;
        const double dx = c.n - m.pn, dy = c.e - m.pe, dz = c.u - m.pu;
        if (em_sqrt0(dx * dx + dy * dy + dz * dz) <= 300.0) {
          for (int q = 0; q < c.count; q++)
            if (env_u01(cfg.seed, cfg.env_offset + L.env, RNG_CHAFF, when, L.lane * 64 + keyn, j * 64 + q) < 0.85) missed = true;
        }
      }
      if (missed) m.status = MS_MISS;
    }
    missile_store(v, mid, m);
  }
  live_io = live;
}

// OR over the lanes of an env.  Not __reduce_or_sync(L.gmask, x): with a partial mask that compiles to WARPSYNC.EXCLUSIVE +
// REDUX, which splits the warp into its 32 / G groups, and the groups reconverge only much later (ncu of the three-warp
// frame: the state load executed 16 times per warp with 2 active lanes, Propagate 2.25 times with 14: +23 % warp
// instructions).  A butterfly of full-warp shuffles leaves the warp converged; every call site is warp-uniform.
ENV_DEV unsigned group_or(const Lane& L, unsigned x) {
  const int G = __popc(L.gmask);
  for (int o = 1; o < G; o <<= 1) x |= __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// ---------------------------------------------------------------------------------------------- shared by the substep kernels
// lane of an aircraft slot (the multi-warp frames give several threads the same slot)
ENV_DEV Lane lane_of_slot(const EnvView& v, const int lg, const int slot, const int gid) {
  Lane L;
  const int G = 1 << lg;
  L.tid = slot; L.env = gid >> lg; L.lane = gid & (G - 1); L.gbase = slot - L.lane;
  L.valid = (L.env < v.B) && (L.lane < v.A);
  L.row = L.env * v.A + L.lane;
  L.gmask = ((1u << G) - 1u) << ((unsigned)(slot & 31) & ~(unsigned)(G - 1));
  return L;
}

// normalize_action + set_property_values(action_var) with the catalog clip (E/tasks/heading_task.py:102-110,
// E/tasks/singlecombat_task.py:141-153, E/core/catalog.py:192-197); the shoot bits go to the aircraft arena
struct Controls { double u0, u1, u2, u3; };
ENV_DEV Controls decode_action(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const int32_t* __restrict__ actions) {
  const int adim = 4 + cfg.shoot_dim;
  const int32_t* act = actions + (size_t)L.row * adim;
  Controls c;
  if (cfg.act_kind == ACS_ACT_HEADING) {
    c.u0 = act[0] * 2. / (41 - 1.) - 1.; c.u1 = act[1] * 2. / (41 - 1.) - 1.; c.u2 = act[2] * 2. / (41 - 1.) - 1.; c.u3 = act[3] * 0.5 / (30 - 1.) + 0.4;
  } else {
    c.u0 = act[0] / 20. - 1.; c.u1 = act[1] / 20. - 1.; c.u2 = act[2] / 20. - 1.; c.u3 = act[3] / 58. + 0.4;
  }
  c.u0 = env_clip(c.u0, -1.0, 1.0); c.u1 = env_clip(c.u1, -1.0, 1.0); c.u2 = env_clip(c.u2, -1.0, 1.0); c.u3 = env_clip(c.u3, 0.0, 0.9);
  int shoot = 0;
  for (int k = 0; k < cfg.shoot_dim; k++) shoot |= (act[4 + k] != 0) << k;
  if (cfg.shoot_dim > 0) AI(v, AI_SHOOT, L.row) = shoot;
  return c;
}
// The command side of an aircraft at the start of a step: the actions are applied to dead aircraft too (they only stop
// running); returns true when the carried state was loaded (aircraft alive), with the commands set.
ENV_DEV bool load_commanded(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const int32_t* __restrict__ actions, const bool alive,
                            AcCore& a, Props& p, FcsState& s) {
  const int N = v.rows;
  const Controls c = decode_action(v, cfg, L, actions);
  if (alive) {
    f16_props_init(p, s);
    load_state(v.fdm, N, L.row, a, p, s);
    p.fcs_aileron_cmd_norm = c.u0; p.fcs_elevator_cmd_norm = c.u1; p.fcs_rudder_cmd_norm = c.u2; p.fcs_throttle_cmd_norm = c.u3;
    return true;
  }
  v.fdm[(size_t)(F_CMD0 + 0) * N + L.row] = c.u0; v.fdm[(size_t)(F_CMD0 + 1) * N + L.row] = c.u1;
  v.fdm[(size_t)(F_CMD0 + 2) * N + L.row] = c.u2; v.fdm[(size_t)(F_CMD0 + 3) * N + L.row] = c.u3;
  return false;
}

// What the thread that integrates an aircraft carries through a step besides the FDM state.
struct EomLane {
  PubAc me;
  double v_mps, w_mps, vc_mps;
  int status, sc0;
  bool has_ms, was_alive;
  bool deferred;               // the env's missiles / chaff run after the K loop in k_env_missiles (this kernel only records every
                               // aircraft's position / velocity per substep in v.traj)
  bool snap;                   // a missile may reach this aircraft within the step: record its state after every substep
  unsigned long long live;     // slots of this aircraft whose missile still runs (live_missiles)
  AcOut o;
};
// Which aircraft of the env can a missile of this lane reach (fuze radius) within this interaction step?  Conservative
// bound on the separation over the T = K dt seconds of the step from the radial closing rate r' = d . v_rel / |d| now:
// |d(t)| >= d(0)/|d(0)| . d(t) >= |d0| + r' t - a_max t^2 / 2, concave in t, so its minimum over [0, T] is at an end point.
// a_max = 1500 m/s^2 covers the missile (753 m/s^2 of thrust for the AIM-120B numbers of simulatior.py:700-712, at most
// 50 g = 490 m/s^2 of commanded lateral acceleration, drag, gravity) plus 15 g for the target; 20 m of margin.  The
// fuze is only tested at the substep instants, so the continuous-time bound covers it.  Returns a bit per target lane.
ENV_DEV unsigned missile_threatens(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const unsigned long long live) {
  const double T = cfg.substeps * cfg.sim_dt;
  unsigned mask = 0;
  for (unsigned long long m = live; m; m &= m - 1) {
    const int mid = L.row * v.S + (__ffsll((long long)m) - 1);
    if (MI(v, MI_STATUS, mid) != MS_LAUNCHED) continue;          // HIT / MISS missiles never score again
    const int target = MI(v, MI_TARGET, mid);
    const int trow = L.env * v.A + target;
    const double dx = MD(v, MD_POS_N, mid) - AD(v, AD_POS_N, trow), dy = MD(v, MD_POS_E, mid) - AD(v, AD_POS_E, trow),
                 dz = MD(v, MD_POS_U, mid) - AD(v, AD_POS_U, trow);
    const double wx = MD(v, MD_VEL_N, mid) - AD(v, AD_VEL_N, trow), wy = MD(v, MD_VEL_E, mid) - AD(v, AD_VEL_E, trow),
                 wz = MD(v, MD_VEL_U, mid) + AD(v, AD_VEL_D, trow);         // the aircraft publishes v_down
    const double d0 = sqrt(dx * dx + dy * dy + dz * dz);
    const double rate = d0 > 0.0 ? (dx * wx + dy * wy + dz * wz) / d0 : -sqrt(wx * wx + wy * wy + wz * wz);
    const double at_end = d0 + rate * T - 0.5 * 1500.0 * T * T;
    if (fmin(d0, at_end) < missile_params(MI(v, MI_KIND, mid)).Rc + 20.0) mask |= 1u << target;
  }
  return mask;
}
// allow_defer (the one-thread kernel): the launch is followed by k_env_missiles, which then does ALL missile and chaff work
// of the step (allocated only for tasks that fly missiles: v.traj; otherwise there is no such work).  The substep
// kernel integrates the aircraft without stopping and records, for the envs that have missiles or chaff in the air, every
// aircraft's position / velocity per substep (v.traj).  The one feedback from missiles to aircraft -- a hit freezes the
// target from the next substep on (simulatior.py:520-533, :210-229) -- is handled by recording the FDM state after every
// substep for the aircraft a missile can reach within the step (missile_threatens; v.snap): k_env_missiles restores the
// state of the substep in which the hit happened.  EI_DEFERRED tells k_env_missiles what the env needs: 0 nothing,
// 1 no missile can score this step (statuses are constant, the missiles are independent: each is integrated over the K
// substeps in registers), 2 the per-substep lockstep of the env's lanes (fuze arbitration in dict order).
ENV_DEV void eom_begin(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, EomLane& E, const bool allow_defer) {
  E.v_mps = E.w_mps = E.vc_mps = 0;
  E.status = ST_CRASH; E.has_ms = false; E.live = 0; E.deferred = false; E.snap = false;
  E.me.status = ST_CRASH; E.me.bloods = 0; E.me.h = 0; E.me.u_mps = 0;
  E.me.f.n = E.me.f.e = E.me.f.u = E.me.f.vn = E.me.f.ve = E.me.f.vd = 0;
  bool chaff = false;
  unsigned threat = 0;
  const bool ext = allow_defer && v.traj != nullptr;
  if (L.valid) {
    load_pub(v, L.row, E.me);
    E.status = E.me.status;
    E.live = live_missiles(v, L);
    chaff = AI(v, AI_CH_STATE, L.row) == CH_ACTIVE;
    if (ext) threat = missile_threatens(v, cfg, L, E.live);
  }
  // full-warp votes (every call site of eom_begin is warp-uniform): a vote on the group's own mask can compile to
  // WARPSYNC.EXCLUSIVE, which leaves the warp split into its groups (see group_or)
  const bool any_live = (__ballot_sync(0xffffffffu, E.live != 0) & L.gmask) != 0;
  const bool any_chaff = (__ballot_sync(0xffffffffu, chaff) & L.gmask) != 0;
  threat = group_or(L, threat);
  E.deferred = ext && (any_live || any_chaff);
  E.snap = E.deferred && L.valid && ((threat >> L.lane) & 1u);
  // in-kernel missile phase (no k_env_missiles): per-substep exchange while missiles run or a chaff cloud is effective
  E.has_ms = !ext && (any_live || any_chaff);
  E.was_alive = L.valid && E.status == ST_ALIVE;
  E.sc0 = (L.env < v.B) ? EI(v, EI_SUBSTEP_COUNT, L.env) : 0;
  if (L.valid && L.lane == 0) EI(v, EI_DEFERRED, L.env) = E.deferred ? (threat ? 2 : 1) : 0;
}
// the record k_env_missiles reads: what the aircraft of `row` published after substep k
ENV_DEV void traj_store(const EnvView& v, const int k, const int row, const Feat& f) {
  double* t = v.traj + (size_t)k * 6 * v.rows + row;
  t[0] = f.n; t[(size_t)v.rows] = f.e; t[(size_t)2 * v.rows] = f.u;
  t[(size_t)3 * v.rows] = f.vn; t[(size_t)4 * v.rows] = f.ve; t[(size_t)5 * v.rows] = f.vd;
}
ENV_DEV void traj_load(const EnvView& v, const int k, const int row, Feat& f) {
  const double* t = v.traj + (size_t)k * 6 * v.rows + row;
  f.n = t[0]; f.e = t[(size_t)v.rows]; f.u = t[(size_t)2 * v.rows];
  f.vn = t[(size_t)3 * v.rows]; f.ve = t[(size_t)4 * v.rows]; f.vd = t[(size_t)5 * v.rows];
}
// AircraftSimulator.run's gate (simulatior.py:210-229): an aircraft whose bloods ran out turns SHOTDOWN and still
// integrates this frame
ENV_DEV bool eom_runs(const Lane& L, EomLane& E) {
  if (!(L.valid && E.status == ST_ALIVE)) return false;
  if (E.me.bloods <= 0) E.status = ST_SHOTDOWN;
  return true;
}
// After the aircraft's frame of substep k: missiles / chaff of the env, then the property read-back.
// _update_properties (simulatior.py:238-257) is only materialised when somebody reads it: the other lanes' missiles need
// position / velocity every substep (published straight from the frame, no inverse trig); the full property set is
// extracted once, after the aircraft's last frame of the step (or of its life).
ENV_DEV void eom_after_frame(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const GeoOrigin& org, EomLane& E, const AcCore& a,
                             const Frame& f, const bool ran, const int k, const int K, PubAc* sP, int* sWin, int* sShot, PubChaff* sCh) {
  if (E.has_ms) {
    if (ran) publish_from_frame(f, org, E.me);
    E.me.status = E.status;
    sP[L.tid] = E.me;
    sWin[L.tid] = 0x7fffffff;
    sShot[L.tid] = 0;
    __syncwarp(L.gmask);
    missile_phase(v, cfg, L, org, sP, sWin, sShot, sCh, E.sc0 + k, E.live);
    __syncwarp(L.gmask);
    if (sShot[L.tid] && E.status == ST_ALIVE) E.status = ST_SHOTDOWN;   // target_aircraft.shotdown() (simulatior.py:527)
    __syncwarp(L.gmask);
  }
  if (ran && (k == K - 1 || E.status != ST_ALIVE)) {
    fdm_outputs(a, f, E.o);
    derive_aircraft(E.o, org, E.me, E.v_mps, E.w_mps, E.vc_mps);
  }
}
// end of the step; OWN_ALL: this thread also owns the flight-control side of the carried state (one-thread frame)
template <bool OWN_ALL>
ENV_DEV void eom_end(const EnvView& v, const Lane& L, const EomLane& E, const AcCore& a, const Props& p, const FcsState& s, const int K) {
  if (!L.valid) return;
  if (E.was_alive) {
    if (OWN_ALL) store_state(v.fdm, v.rows, L.row, a, p, s);
    else store_state_role<false>(v.fdm, v.rows, L.row, a, p, s);
    store_out(v.out, v.rows, L.row, E.o);
    store_derived(v, L.row, E.me, E.v_mps, E.w_mps, E.vc_mps);
  }
  AI(v, AI_STATUS, L.row) = E.status;
  if (L.lane == 0) EI(v, EI_SUBSTEP_COUNT, L.env) = E.sc0 + K;
}
// Propulsion of one frame (engine + fuel), as fdm_frame runs it
ENV_DEV void eom_propulsion(AcCore& a, const Props& p, Frame& f, const double* __restrict__ T, const double dt) {
  const int flags = (int)a.engflags;
  bool augmentation = flags & 2;
  f.thrust = fdm_stage_engine(a, p, f.atm, f.qbar, T, g_atmo, dt, flags & 1, augmentation);
  const bool starved_next = fdm_stage_consume_fuel(a, dt, flags & 1, false);
  a.engflags = (double)((starved_next ? 1 : 0) | (augmentation ? 2 : 0));
}

#ifndef ACS_FRAME_SYNC
#define ACS_FRAME_SYNC 0      // tuning builds: block barrier at the top of every frame of k_env_substeps (measured: no effect once the spills were gone)
#endif
constexpr int N_SNAP = FDM_N_CORE + F16_N_CARRIED + 4;     // the FDM state + FrameKeep
// state of the aircraft after substep k, for k_env_missiles (a hit in substep k leaves the aircraft in exactly this state)
ENV_DEV void snap_store(const EnvView& v, const int k, const int row, const AcCore& a, const Props& p, const FcsState& s, const FrameKeep& keep) {
  double* b = v.snap + (size_t)k * N_SNAP * v.rows;
  store_state(b, v.rows, row, a, p, s);
  double* q = b + (size_t)(FDM_N_CORE + F16_N_CARRIED) * v.rows + row;
  q[0] = keep.pilot_nx; q[(size_t)v.rows] = keep.vcas; q[(size_t)2 * v.rows] = keep.beta; q[(size_t)3 * v.rows] = keep.thrust;
}
ENV_DEV void snap_load(const EnvView& v, const int k, const int row, AcCore& a, Props& p, FcsState& s, FrameKeep& keep) {
  const double* b = v.snap + (size_t)k * N_SNAP * v.rows;
  f16_props_init(p, s);
  load_state(b, v.rows, row, a, p, s);
  const double* q = b + (size_t)(FDM_N_CORE + F16_N_CARRIED) * v.rows + row;
  keep.pilot_nx = q[0]; keep.vcas = q[(size_t)v.rows]; keep.beta = q[(size_t)2 * v.rows]; keep.thrust = q[(size_t)3 * v.rows];
}

// The throughput kernel: one thread per aircraft, lean frame (fdm_core.cuh): nothing of a frame's scratch outlives it; the
// property read-back (_update_properties, simulatior.py:238-257) is materialised once, from the state the aircraft's
// last frame left, after the K loop.  No missile or chaff code in here at all: k_env_missiles follows (eom_begin);
// aircraft of envs with missiles or chaff in the air publish their trajectory and, when a missile can reach them, their
// state per substep to global memory.
__global__ void __launch_bounds__(FDM_BLOCK, ACS_FDM_MIN_BLOCKS) k_env_substeps(const EnvView v, const __grid_constant__ AcsTaskConfig cfg, const int lg,
                                                           const int32_t* __restrict__ actions) {
  __shared__ double sT[F16_NTAB];
#ifdef ACS_PAD_NOPS      // tuning builds: shift the code that follows by 16 bytes per NOP (code-placement sensitivity)
#pragma unroll
  for (int i = 0; i < ACS_PAD_NOPS; i++) asm volatile("nanosleep.u32 0;");
#endif
  stage_tables(sT);
  const Lane L = lane_setup(v, lg);
  const int K = cfg.substeps;
  const double dt = cfg.sim_dt, fcs_dt = cfg.fcs_dt;
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  AcCore a; Props p; FcsState s;
  EomLane E;
  eom_begin(v, cfg, L, E, true);
  if (L.valid) load_commanded(v, cfg, L, actions, E.status == ST_ALIVE, a, p, s);
  FrameKeep keep;
  keep.pilot_nx = keep.vcas = keep.beta = keep.thrust = 0.0;
  const bool deferred = E.deferred, snap = E.snap;
  const double bloods = E.me.bloods;
  int status = E.status;
  for (int k = 0; k < K; k++) {
    if (ACS_FRAME_SYNC) __syncthreads();
    const bool ran = L.valid && status == ST_ALIVE;
    if (ran) {
      if (bloods <= 0) status = ST_SHOTDOWN;     // AircraftSimulator.run's gate (simulatior.py:220-226): still integrates this frame
      fdm_frame_lean(a, p, s, keep, sT, g_atmo, dt, fcs_dt, [&](const Frame& f) {
        if (deferred) { PubAc pub; publish_from_frame(f, org, pub); traj_store(v, k, L.row, pub.f); }
      });
      if (snap) snap_store(v, k, L.row, a, p, s, keep);
    } else if (deferred && L.valid) {
      // an aircraft that does not run this substep stays where it stopped: the arena's position before its first frame of
      // the step, its own last record afterwards (only a gun kill -- bloods <= 0 -- stops an aircraft in this kernel)
      Feat f = E.me.f;
      if (k > 0) traj_load(v, k - 1, L.row, f);
      traj_store(v, k, L.row, f);
    }
  }
  E.status = status;
  if (E.was_alive) {
    Frame f;
    fdm_refresh(a, p, keep, f);
    fdm_outputs(a, f, E.o);
    derive_aircraft(E.o, org, E.me, E.v_mps, E.w_mps, E.vc_mps);
  }
  eom_end<true>(v, L, E, a, p, s, K);
}

// ---------------------------------------------------------------------------------------------- missiles after the K loop
// All missile and chaff work of a step whose aircraft were integrated by k_env_substeps<true>, one thread per shooter, the
// lanes of an env adjacent as everywhere.  Per env (EI_DEFERRED, decided in eom_begin from the state at the start of
// the step):
//   1  no missile can score within this step.  Target statuses are constant, nothing feeds back into the aircraft, the
//      missiles do not interact: each missile is loaded once, run() K times against the target's recorded trajectory
//      with its state in registers, and stored once -- instead of K x (load, run, store) between the frames of a
//      255-register thread.  Chaff (E/core/simulatior.py:377-381, E/envs/env_base.py:146-154): a cloud is a fixed position
//      with a timer, so its state at substep k follows from its state at the start of the step; every lane publishes
//      its cloud once, each missile tests the clouds still effective at that substep (same keyed draws), the owner
//      lane advances its timer by the same K additions.  A fuze condition met here would mean the reach bound of
//      missile_threatens was wrong: it is counted in the env's fault counter (asserted zero by every parity test).
//   2  some missile may score: the env's lanes run the K missile phases in lockstep (missile_phase: dict-order fuze
//      arbitration, chaff), reading the aircraft's recorded positions; a HIT marks the target SHOTDOWN from that substep
//      on and -- because k_env_substeps integrated it to the end of the step -- its lane restores the state it had after
//      the substep of the hit (v.snap) and re-derives the outputs from it, which is exactly where the reference leaves
//      an aircraft that stops running (simulatior.py:210-229,520-533).
// Same expressions as the in-kernel missile phase of the multi-warp frames.
#ifdef ACS_MISSILE_PROFILE
// tuning builds only: cycles of the lockstep section (warps that ran it) and of the rest of the kernel, per warp: sum, max, count
__device__ unsigned long long g_missile_prof[8];
#define MPROF_T0 const long long mp0_ = clock64(); long long mp1_ = mp0_; const bool mp_m2_ = warp_m2;
#define MPROF_T1 mp1_ = clock64();
#define MPROF_OUT if ((threadIdx.x & 31) == 0) { const long long e_ = clock64(); \
    if (mp_m2_) { atomicAdd(&g_missile_prof[0], (unsigned long long)(mp1_ - mp0_)); atomicMax(&g_missile_prof[1], (unsigned long long)(mp1_ - mp0_)); atomicAdd(&g_missile_prof[2], 1ull); } \
    atomicAdd(&g_missile_prof[3], (unsigned long long)(e_ - mp1_)); atomicMax(&g_missile_prof[4], (unsigned long long)(e_ - mp1_)); atomicAdd(&g_missile_prof[5], 1ull); \
    atomicMax(&g_missile_prof[6], (unsigned long long)(e_ - mp0_)); }
#else
#define MPROF_T0
#define MPROF_T1
#define MPROF_OUT
#endif
// One independent missile (mode 1) for the K substeps of the step, state in registers: `stid` = block-local thread of the
// shooter's lane, `mslot` = its slot.  Everything it touches is addressed by that pair.
ENV_DEV void missile_run_independent(const EnvView& v, const AcsTaskConfig& cfg, const GeoOrigin& org, const int lg, const int stid,
                                     const int mslot, const PubChaff* sCh, const int* sChEnd) {
  const int K = cfg.substeps, G = 1 << lg;
  const double dt = cfg.sim_dt;
  const int maxlen = (int)(5.0 / dt);
  const int gid = blockIdx.x * blockDim.x + stid;
  const int env = gid >> lg, lane = gid & (G - 1), row = env * v.A + lane, gbase = stid - lane;
  const int mid = row * v.S + mslot;
  bool any_chaff = false;
  for (int j = 0; j < v.A; j++) any_chaff = any_chaff || sChEnd[gbase + j] > 0;
  const int64_t when0 = ((int64_t)EI(v, EI_EPISODE, env) << 20) + (EI(v, EI_SUBSTEP_COUNT, env) - K);   // k_env_substeps advanced the counter
  Missile m;
  missile_load(v, mid, m);
  const MissileParams pr = missile_params(m.kind);
  const int trow = env * v.A + m.target;
  const bool target_alive = AI(v, AI_STATUS, trow) == ST_ALIVE;     // constant over the step in this mode
  const int keyn = MI(v, MI_KEYN, mid);
  Feat tg;
  traj_load(v, 0, trow, tg);
  for (int k = 0; k < K; k++) {
    Feat nxt = tg;
    if (k + 1 < K) traj_load(v, k + 1, trow, nxt);                  // in flight while this substep computes
    m.t += dt;
    double ny, nz, dist;
    missile_guidance(m, pr, tg, ny, nz, dist);
    m.consec = (dist > m.d_prev) ? m.consec + 1 : 0;
    m.d_prev = dist;
    const double speed = em_sqrt(m.vn * m.vn + m.ve * m.ve + m.vu * m.vu);
    if (dist < pr.Rc && target_alive && m.status != MS_MISS) {
      atomicAdd(&EI(v, EI_FAULTS, env), 1);                         // cannot happen (missile_threatens)
      m.status = MS_HIT;
    } else if (m.t > pr.t_max || speed < pr.v_min || m.consec >= maxlen || !target_alive) {
      const bool inert = m.status == MS_MISS && (m.t > pr.t_max || !target_alive || speed < pr.v_min);
      m.status = MS_MISS;
      if (inert) break;             // every later run() is the same no-op (missile_inert)
    } else {
      missile_state_trans(m, pr, org, ny, nz, dt);
    }
    if (any_chaff && m.status == MS_LAUNCHED) {
      bool missed = false;
      for (int j = 0; j < v.A; j++) {
        if (k >= sChEnd[gbase + j]) continue;                       // no cloud, or done by this substep
        const PubChaff& c = sCh[gbase + j];
        const double dx = c.n - m.pn, dy = c.e - m.pe, dz = c.u - m.pu;
        if (em_sqrt0(dx * dx + dy * dy + dz * dz) <= 300.0) {
          for (int q = 0; q < c.count; q++)
            if (env_u01(cfg.seed, cfg.env_offset + env, RNG_CHAFF, when0 + k, lane * 64 + keyn, j * 64 + q) < 0.85) missed = true;
        }
      }
      if (missed) m.status = MS_MISS;
    }
    tg = nxt;
  }
  missile_store(v, mid, m);
}

constexpr int MISSILE_LIST_CAP = 512;      // independent missiles of one block handed out through the list; a lane whose slots
                                           // do not fit (more than 4 live missiles per lane on average) walks the rest itself
__global__ void __launch_bounds__(128, 2) k_env_missiles(const EnvView v, const __grid_constant__ AcsTaskConfig cfg, const int lg) {
  __shared__ PubAc sP[128];
  __shared__ int sWin[128];
  __shared__ int sShot[128];
  __shared__ PubChaff sCh[128];
  __shared__ int sChEnd[128];          // mode 1: first substep at which the lane's cloud is no longer effective (0: none)
  __shared__ int sList[MISSILE_LIST_CAP];   // mode 1: the block's independent missiles, compacted: (shooter tid << 8) | slot
  __shared__ int sWarpN[4], sWarpM2[4];
  const Lane L = lane_setup(v, lg);
  const int K = cfg.substeps;
  const double dt = cfg.sim_dt;
  const int mode = L.valid ? EI(v, EI_DEFERRED, L.env) : 0;
  if (!__syncthreads_or(mode != 0)) return;
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  const int sc0 = L.valid ? EI(v, EI_SUBSTEP_COUNT, L.env) - K : 0;          // k_env_substeps advanced the counter
  const int wid = threadIdx.x >> 5, wl = threadIdx.x & 31;
  const bool warp_m2 = __any_sync(0xffffffffu, mode == 2);
  MPROF_T0

  // ================================================================ mode 1, part 1: chaff of the step, the block's missile list
  // ---- chaff run() of the whole step for this lane's cloud, and its publication
  const bool on1 = mode == 1;
  {
    PubChaff c;
    c.state = CH_NONE; c.count = 0; c.n = c.e = c.u = 0.0;
    int end = 0;
    if (on1) {
      c.state = AI(v, AI_CH_STATE, L.row);
      if (c.state != CH_NONE) {
        double t = AD(v, AD_CH_T, L.row);
        int st = c.state;
        end = st == CH_ACTIVE ? K : 0;
        for (int k = 0; k < K; k++) {
          t += dt;
          if (t > 20.0) { if (st == CH_ACTIVE) end = k; st = CH_DONE; }      // run() precedes the test of the same substep
        }
        AD(v, AD_CH_T, L.row) = t;
        AI(v, AI_CH_STATE, L.row) = st;
        c.count = AI(v, AI_CH_COUNT, L.row);
        c.n = AD(v, AD_CH_N, L.row); c.e = AD(v, AD_CH_E, L.row); c.u = AD(v, AD_CH_U, L.row);
      }
    }
    sCh[L.tid] = c;
    sChEnd[L.tid] = end;
  }
  // ---- the independent missiles, compacted over the block.  A lane owns the missiles its aircraft launched, but most lanes
  // have none in the air and a few have two (ncu, 4v4: 12 of 32 lanes active when every lane walked its own slots).  The
  // live (shooter, slot) pairs of the whole block go into one dense list; a missile's K substeps run on whichever thread
  // draws its entry.  The list is consumed by the warps that have NO threatened env: the lockstep section below is the
  // longest chain of the kernel (~120-200 k cycles per warp against ~100 k for an independent missile), and with the two
  // running side by side a block costs the longer of them instead of their sum.
  const int nl = on1 ? AI(v, AI_N_LAUNCHED, L.row) : 0;
  unsigned long long mine = 0;                       // live slots of this lane (slots >= 64 cannot exist: acs_env_create caps S)
  for (int sl = 0; sl < nl; sl++) {
    const int mid = L.row * v.S + sl;
    if (MI(v, MI_DETACHED, mid) || missile_inert(v, mid, L.env)) continue;
    mine |= 1ull << sl;
  }
  const int cnt = __popcll(mine);
  int incl = cnt;                                    // inclusive warp scan
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (wl >= o) incl += y; }
  if (wl == 31) { sWarpN[wid] = incl; sWarpM2[wid] = warp_m2 ? 1 : 0; }
  __syncthreads();
  int base = incl - cnt, total = 0, nfree = 0, frank = 0;
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const int n = sWarpN[w];
    if (w < wid) { base += n; frank += sWarpM2[w] ? 0 : 1; }
    total += n; nfree += sWarpM2[w] ? 0 : 1;
  }
  unsigned long long rest = mine;                    // what does not fit into the list stays with its lane
  for (int i = base; rest && i < MISSILE_LIST_CAP; i++) {
    sList[i] = (L.tid << 8) | (__ffsll((long long)rest) - 1);
    rest &= rest - 1;
  }
  __syncthreads();                                   // the list, sCh and sChEnd are complete: the last block-wide barrier
  const int listed = min(total, MISSILE_LIST_CAP);

  // ================================================================ mode 2: lockstep phases, hits, state restore
  if (__any_sync(0xffffffffu, mode == 2)) {
    const bool on = mode == 2;
    unsigned long long live = 0;
    PubAc me;
    me.status = ST_CRASH; me.bloods = 0; me.h = 0; me.u_mps = 0; me.f.n = me.f.e = me.f.u = me.f.vn = me.f.ve = me.f.vd = 0;
    if (on) {
      live = live_missiles(v, L);
      me.status = AI(v, AI_STATUS, L.row);       // as k_env_substeps left it: constant over the step unless a missile scores
    }
    int hit_k = K;
    for (int k = 0; k < K; k++) {
      if (on && hit_k == K) traj_load(v, k, L.row, me.f);       // a shot-down aircraft stays where it was hit
      sP[L.tid] = me;
      sWin[L.tid] = 0x7fffffff;
      sShot[L.tid] = 0;
      __syncwarp(L.gmask);
      if ((__ballot_sync(0xffffffffu, on) & L.gmask) != 0) missile_phase(v, cfg, L, org, sP, sWin, sShot, sCh, sc0 + k, live);
      __syncwarp(L.gmask);
      if (on && sShot[L.tid] && me.status == ST_ALIVE) { me.status = ST_SHOTDOWN; hit_k = k; }   // target_aircraft.shotdown() (simulatior.py:527)
      __syncwarp(L.gmask);
    }
    if (on && hit_k < K) {
      AI(v, AI_STATUS, L.row) = ST_SHOTDOWN;
      if (hit_k < K - 1) {
        // k_env_substeps ran K frames; the aircraft stopped after frame hit_k
        AcCore a; Props p; FcsState s; FrameKeep keep;
        snap_load(v, hit_k, L.row, a, p, s, keep);
        // the record must be this step's: its clock is (K - 1 - hit_k) frames behind the one k_env_substeps stored
        const double t_end = v.fdm[(size_t)F_SIM_TIME * v.rows + L.row];
        if (fabs(a.sim_time - (t_end - (K - 1 - hit_k) * dt)) > 0.25 * dt) atomicAdd(&EI(v, EI_FAULTS, L.env), 1);   // hit on an unrecorded aircraft
        store_state(v.fdm, v.rows, L.row, a, p, s);
        Frame f; AcOut o; PubAc pub; double v_mps, w_mps, vc_mps;
        fdm_refresh(a, p, keep, f);
        fdm_outputs(a, f, o);
        derive_aircraft(o, org, pub, v_mps, w_mps, vc_mps);
        store_out(v.out, v.rows, L.row, o);
        store_derived(v, L.row, pub, v_mps, w_mps, vc_mps);
      }
    }
  }
  MPROF_T1

  // ================================================================ mode 1, part 2: the list, then what did not fit
  if (!warp_m2 || nfree == 0) {
    const int rank = nfree ? frank : wid, nw = nfree ? nfree : 4;
    for (int i = rank * 32 + wl; i < listed; i += nw * 32) {
      const int ent = sList[i];
      missile_run_independent(v, cfg, org, lg, ent >> 8, ent & 255, sCh, sChEnd);
    }
  }
  for (; rest; rest &= rest - 1) missile_run_independent(v, cfg, org, lg, L.tid, __ffsll((long long)rest) - 1, sCh, sChEnd);
  MPROF_OUT
}

// ---------------------------------------------------------------------------------------------- two-warp frame
// Latency variant of k_env_substeps for batches too small to fill the machine (the headline 4096-env config puts one
// warp on fewer than half of the 592 SM sub-partitions, so a step costs one warp's dependent-issue latency).  Each
// aircraft is integrated by TWO threads in two different warps of the block, running the stage functions of fdm_frame
// concurrently:
//
//   role A (warps 0 .. S/32-1)                           role B (warps S/32 .. 2S/32-1)
//   Propagate                    -- EARLY, run flag -->
//   ------------------------------------------------ pair barrier 1
//   Inertial, Atmosphere, MassBalance, Auxiliary         FCS (inputs: last frame's auxiliary values, as in JSBSim)
//                                -- AUX, 2 Vt      -->   <-- SURF --
//   ------------------------------------------------ pair barrier 2
//   Propulsion, aero axes SPLIT_AXES_A                   aero axes SPLIT_AXES_B
//                                                        <-- axis sums --
//   ------------------------------------------------ pair barrier 3
//   Aircraft + Accelerations, missiles / chaff
//
// Every value is computed by the same expression as in fdm_frame, only on another thread; the exchange lists are
// generated from the model's dataflow (F16_X_* in gen/f16_gen.cuh).  A pair synchronises on its own named barrier
// (64 threads), so pairs never wait for each other.  Buffer reuse: EARLY aliases SURF and the axis sums alias AUX --
// in both cases the reader of the first finishes (same thread) before it writes the second, and the next writer of
// the first is behind a later barrier.
#ifndef ACS_S2_MIN_BLOCKS
#define ACS_S2_MIN_BLOCKS 1
#endif
constexpr int SPLIT_N_BUF1 = F16_N_X_SURF > F16_N_X_EARLY ? F16_N_X_SURF : F16_N_X_EARLY;
constexpr int SPLIT_N_BUF2 = (F16_N_X_AUX + 1) > 6 ? (F16_N_X_AUX + 1) : 6;

#ifdef ACS_SPLIT_PROFILE
// tuning builds only: absolute clock64 stamps of one frame of the first pair of block 0, taken inside the same asm
// statement as the barrier (arrival / release), so the compiler cannot move them relative to it
__device__ long long g_split_prof[2][8];
#define PROF_DECL long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int pi_ = 0; const bool prof_ = blockIdx.x == 0 && slot == 0;
#define PROF_FRAME(k) pi_ = ((k) == 6) ? 0 : 8;
#define PROF_TOP(k) { if ((k) == 6) pt_[6] = clock64(); if ((k) == 7) pt_[7] = clock64(); }
#define PROF_OUT(r) if (prof_) { for (int i_ = 0; i_ < 8; i_++) g_split_prof[r][i_] = pt_[i_]; }
#define pair_barrier(id) { long long t0_, t1_; __syncwarp(); \
  asm volatile("mov.u64 %0, %%clock64;\n\tbar.sync %2, 64;\n\tmov.u64 %1, %%clock64;" : "=l"(t0_), "=l"(t1_) : "r"(id) : "memory"); \
  if (pi_ < 8) { pt_[pi_] = t0_; pt_[pi_ + 1] = t1_; pi_ += 2; } }
#else
#define PROF_DECL
#define PROF_FRAME(k)
#define PROF_TOP(k)
#define PROF_OUT(r)
ENV_DEV void pair_barrier(const int id) { __syncwarp(); asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
#endif

__global__ void __launch_bounds__(2 * SPLIT_SLOTS, ACS_S2_MIN_BLOCKS) k_env_substeps_split(const EnvView v, const __grid_constant__ AcsTaskConfig cfg,
                                                                          const int lg, const int32_t* __restrict__ actions) {
  __shared__ double sT[F16_NTAB];
  __shared__ PubAc sP[SPLIT_SLOTS];
  __shared__ int sWin[SPLIT_SLOTS];
  __shared__ int sShot[SPLIT_SLOTS];
  __shared__ PubChaff sCh[SPLIT_SLOTS];
  __shared__ double sX1[SPLIT_N_BUF1][SPLIT_SLOTS];
  __shared__ double sX2[SPLIT_N_BUF2][SPLIT_SLOTS];
  __shared__ int sRun[SPLIT_SLOTS];
  stage_tables(sT);
  const bool role_b = threadIdx.x >= SPLIT_SLOTS;            // warp-uniform
  const int slot = threadIdx.x - (role_b ? SPLIT_SLOTS : 0);
  const int bar = 1 + (slot >> 5);                           // barrier 0 is __syncthreads
  const Lane L = lane_of_slot(v, lg, slot, blockIdx.x * SPLIT_SLOTS + slot);
  const int K = cfg.substeps;
  const double dt = cfg.sim_dt, fcs_dt = cfg.fcs_dt;
#define XW1(e) sX1[xi++][slot] = e;
#define XR1(e) e = sX1[xi++][slot];
#define XW2(e) sX2[xi++][slot] = e;
#define XR2(e) e = sX2[xi++][slot];

  if (role_b) {
    // ================================================================ role B: flight controls + its aero axes
    Props p; FcsState s;
    AcCore unused;      // only the carried properties are live on this side; the core loads / stores are dead code
    const bool loaded = L.valid && load_commanded(v, cfg, L, actions, AI(v, AI_STATUS, L.row) == ST_ALIVE, unused, p, s);
    PROF_DECL
    for (int k = 0; k < K; k++) {
      PROF_FRAME(k)
      pair_barrier(bar);                                       // 1: EARLY + run flag are there
      const bool ran = sRun[slot] != 0;
      if (ran) {
        { int xi = 0; F16_X_EARLY(XR1) }
        f16_fcs(p, s, sT, fcs_dt);
        { int xi = 0; F16_X_SURF(XW1) }
      }
      pair_barrier(bar);                                       // 2: AUX is there, SURF is out
      if (ran) {
        double twovel, c[6];
        { int xi = 0; F16_X_AUX(XR2) twovel = sX2[xi][slot]; }
        f16_aero<SPLIT_AXES_B>(p, sT, twovel, c);
#pragma unroll
        for (int i = 0; i < 6; i++) if (SPLIT_AXES_B & (1 << i)) sX2[i][slot] = c[i];
      }
      pair_barrier(bar);                                       // 3: axis sums are out
    }
    PROF_OUT(1)
    if (loaded) store_state_role<true>(v.fdm, v.rows, L.row, unused, p, s);
    return;
  }

  // ================================================================== role A: everything else (as k_env_substeps)
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  AcCore a; Props p; FcsState s; Frame f;
  EomLane E;
  eom_begin(v, cfg, L, E, false);
  if (E.was_alive) { f16_props_init(p, s); load_state(v.fdm, v.rows, L.row, a, p, s); }
  PROF_DECL
  for (int k = 0; k < K; k++) {
    PROF_FRAME(k)
    PROF_TOP(k)
    const bool ran = eom_runs(L, E);
    if (ran) {
      fdm_stage_propagate(a, p, f, dt);
      { int xi = 0; F16_X_EARLY(XW1) }
    }
    sRun[slot] = ran;
    pair_barrier(bar);                                         // 1
    WindAxes w;
    if (ran) {
      fdm_stage_gravity(f);
      fdm_stage_atmosphere(p, f, g_atmo);
      fdm_stage_massbalance(a, f);
      fdm_stage_auxiliary(a, p, f, g_atmo, w);
      { int xi = 0; F16_X_AUX(XW2) sX2[xi][slot] = 2 * f.Vt; }
    }
    pair_barrier(bar);                                         // 2
    double c[6];
    if (ran) {
      { int xi = 0; F16_X_SURF(XR1) }
      eom_propulsion(a, p, f, sT, dt);
      f16_aero<SPLIT_AXES_A>(p, sT, 2 * f.Vt, c);
    }
    pair_barrier(bar);                                         // 3
    if (ran) {
#pragma unroll
      for (int i = 0; i < 6; i++) if (SPLIT_AXES_B & (1 << i)) c[i] = sX2[i][slot];
      fdm_stage_accelerations(a, f, w, c);
    }
    eom_after_frame(v, cfg, L, org, E, a, f, ran, k, K, sP, sWin, sShot, sCh);
  }
  PROF_OUT(0)
  eom_end<false>(v, L, E, a, p, s, K);
#undef XW1
#undef XR1
#undef XW2
#undef XR2
}

// exchange-list accessors of the three- / four-warp frames (row = position in the generated list, column = aircraft slot)
#define XW(buf, e) buf[xi++][slot] = e;
#define XR(buf, e) e = buf[xi++][slot];
#define XW_K(e) XW(sK, e)
#define XR_K(e) XR(sK, e)
#define XW_S(e) XW(sS, e)
#define XR_S(e) XR(sS, e)
#define XW_R(e) XW(sR, e)
#define XR_R(e) XR(sR, e)
#define XW_L(e) XW(sL, e)
#define XR_L(e) XR(sL, e)

// Barrier sets: the three-warp frame synchronises its three roles at every barrier; in the four-warp frame the look-ahead
// role stays out of barrier 2 (it publishes nothing the others wait for there), and a barrier 0 ends its prologue.
struct TripleSync {
  int id;
  ENV_DEV void b0() const {}
  ENV_DEV void b1() const { __syncwarp(); asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }
  ENV_DEV void b2() const { b1(); }
  ENV_DEV void b3() const { b1(); }
};
struct QuadSync {
  int quad;
  ENV_DEV void b0() const { __syncwarp(); asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * quad) : "memory"); }
  ENV_DEV void b1() const { b0(); }
  ENV_DEV void b2() const { __syncwarp(); asm volatile("bar.sync %0, 96;" ::"r"(2 + 2 * quad) : "memory"); }
  ENV_DEV void b3() const { b0(); }
};

constexpr int S3_AXES_B = (1 << 1) | (1 << 3) | (1 << 5), S3_AXES_C = 63 & ~S3_AXES_B;

// Role B of the three- and four-warp frames: the flight controls between barriers 1 and 2, the axes SIDE, ROLL, YAW between
// 2 and 3; owns the commands, FCS outputs and PID states of the carried state.
template <int NS, class Sync>
ENV_DEV void role_fcs_axes(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const int32_t* __restrict__ actions,
                           const double* __restrict__ sT, const int slot, const Sync sync, const int* sRun, double (*sE)[NS],
                           double (*sK)[NS], double (*sS)[NS], double (*sR)[NS], double (*sSum)[NS]) {
  const int K = cfg.substeps;
  const double fcs_dt = cfg.fcs_dt;
  Props p; FcsState s;
  AcCore unused;      // only the carried properties are live on this side; the core loads / stores are dead code
  const bool loaded = L.valid && load_commanded(v, cfg, L, actions, AI(v, AI_STATUS, L.row) == ST_ALIVE, unused, p, s);
  sync.b0();
  for (int k = 0; k < K; k++) {
    sync.b1();
    const bool ran = sRun[slot] != 0;
    if (ran) {
      p.attitude_cos_pitch_cos_roll = sE[0][slot]; p.velocities_u_fps = sE[1][slot]; p.velocities_v_fps = sE[2][slot];
      f16_fcs(p, s, sT, fcs_dt);
      { int xi = 0; F16_X_SURF(XW_S) }
    }
    sync.b2();
    if (ran) {
      double twovel, c[6];
      { int xi = 0; F16_X_KIN(XR_K) twovel = sK[xi][slot]; }
      { int xi = 0; F16_X_AIR(XR_R) }
      f16_aero<S3_AXES_B>(p, sT, twovel, c);
#pragma unroll
      for (int i = 0; i < 6; i++) if (S3_AXES_B & (1 << i)) sSum[i][slot] = c[i];
    }
    sync.b3();
  }
  if (loaded) store_state_role<true>(v.fdm, v.rows, L.row, unused, p, s);
}

// ---------------------------------------------------------------------------------------------- three-warp frame
// One more warp per 32 aircraft.  ncu of the two-warp frame shows role A never waiting and role B waiting for half of its
// time: A's chain Propagate -> Atmosphere -> Auxiliary -> Propulsion -> Accelerations is the critical path.  Here the
// atmosphere and the air-data half of Auxiliary (dynamic pressure, Mach, calibrated airspeed) move to a third role, which
// then shares the aerodynamic axes with the flight-control role:
//
//   role A (equations of motion)        role B (flight controls)          role C (air data)
//   Propagate        -- E: cos(pitch)cos(roll), u, v, w, |r|, cos(lat_gc), run flag -->
//   ------------------------------------------------------------------------------------ triple barrier 1
//   Inertial, MassBalance,              FCS                                Atmosphere, air-data half
//   kinematic half of Auxiliary                                            of Auxiliary
//        -- K: alpha, beta, rates, pilot g, ... -->    -- S: surfaces -->    -- R: qbar, Mach, Vc, T, rho, h -->
//   ------------------------------------------------------------------------------------ triple barrier 2
//   Propulsion                          axes SIDE, ROLL, YAW               axes DRAG, LIFT, PITCH
//                                            -- axis sums -->                   -- axis sums -->
//   ------------------------------------------------------------------------------------ triple barrier 3
//   Aircraft + Accelerations, missiles / chaff
//
// Same stage functions, same expressions, same ownership of the carried state as the two-warp frame (role C owns none:
// what it computes is re-published every frame).  Every exchange buffer has one writer and is rewritten only behind a
// barrier all of its readers have passed.
#ifndef ACS_SPLIT3_SLOTS
#define ACS_SPLIT3_SLOTS 64     // measured: 64 slots (two warps per role share an instruction stream) 0.087 ms, 32 slots 0.093 ms
#endif
constexpr int S3 = ACS_SPLIT3_SLOTS;
#ifndef ACS_S3_MIN_BLOCKS
#define ACS_S3_MIN_BLOCKS 1
#endif
static_assert(S3 % 32 == 0 && S3 >= 32 && 3 * S3 <= 384, "whole warps, at most 4 triples per block");

#define TRIPLE_BARRIER(id) { __syncwarp(); asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(3 * S3, ACS_S3_MIN_BLOCKS) k_env_substeps_split3(const EnvView v, const __grid_constant__ AcsTaskConfig cfg,
                                                                  const int lg, const int32_t* __restrict__ actions) {
  __shared__ double sT[F16_NTAB];
  __shared__ PubAc sP[S3];
  __shared__ int sWin[S3];
  __shared__ int sShot[S3];
  __shared__ PubChaff sCh[S3];
  __shared__ double sE[6][S3];                  // A -> B, C
  __shared__ double sK[F16_N_X_KIN + 1][S3];    // A -> B, C  (+ 2 Vt)
  __shared__ double sS[F16_N_X_SURF][S3];       // B -> A, C
  __shared__ double sR[F16_N_X_AIR + 5][S3];    // C -> A, B  (+ T, rho, h, Vc in ft/s, density altitude)
  __shared__ double sSum[6][S3];                // B, C -> A
  __shared__ int sRun[S3];
  stage_tables(sT);
  // warp-uniform role: 0 = A, 1 = B, 2 = C (which thread group takes which role makes no measurable difference)
#ifndef ACS_S3_ORDER
#define ACS_S3_ORDER 12       // decimal digits: role of thread group 0, 1, 2
#endif
  const int group = threadIdx.x / S3;
  const int role = group == 0 ? (ACS_S3_ORDER / 100) : (group == 1 ? (ACS_S3_ORDER / 10) % 10 : ACS_S3_ORDER % 10);
  const int slot = threadIdx.x - group * S3;
  const int bar = 1 + (slot >> 5);
  const Lane L = lane_of_slot(v, lg, slot, blockIdx.x * S3 + slot);
  const int K = cfg.substeps;
  const double dt = cfg.sim_dt, fcs_dt = cfg.fcs_dt;

  if (role == 1) {
    role_fcs_axes<S3>(v, cfg, L, actions, sT, slot, TripleSync{bar}, sRun, sE, sK, sS, sR, sSum);
    return;
  }
  if (role == 2) {
    // ================================================================ role C: air data + axes DRAG, LIFT, PITCH (no state)
    Props p; FcsState s;
    f16_props_init(p, s);
    Frame f;
    for (int k = 0; k < K; k++) {
      TRIPLE_BARRIER(bar)                                      // 1
      const bool ran = sRun[slot] != 0;
      if (ran) {
        f.uvw.x = sE[1][slot]; f.uvw.y = sE[2][slot]; f.uvw.z = sE[3][slot]; f.radius = sE[4][slot]; f.cosLatGc = sE[5][slot];
        fdm_stage_atmosphere(p, f, g_atmo);
        fdm_airspeed(f);
        fdm_stage_aux_air(p, f, g_atmo);
        { int xi = 0; F16_X_AIR(XW_R) sR[xi][slot] = f.atm.T; sR[xi + 1][slot] = f.atm.rho; sR[xi + 2][slot] = f.h_asl;
          sR[xi + 3][slot] = f.vcas; sR[xi + 4][slot] = p.atmosphere_density_altitude; }
      }
      TRIPLE_BARRIER(bar)                                      // 2
      if (ran) {
        double c[6];
        { int xi = 0; F16_X_KIN(XR_K) }
        { int xi = 0; F16_X_SURF(XR_S) }
        f16_aero<S3_AXES_C>(p, sT, 2 * f.Vt, c);
#pragma unroll
        for (int i = 0; i < 6; i++) if (S3_AXES_C & (1 << i)) sSum[i][slot] = c[i];
      }
      TRIPLE_BARRIER(bar)                                      // 3
    }
    return;
  }

  // ================================================================== role A: equations of motion, propulsion, missiles
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  AcCore a; Props p; FcsState s; Frame f;
  EomLane E;
  eom_begin(v, cfg, L, E, false);
  if (E.was_alive) { f16_props_init(p, s); load_state(v.fdm, v.rows, L.row, a, p, s); }
  for (int k = 0; k < K; k++) {
    const bool ran = eom_runs(L, E);
    if (ran) {
      fdm_stage_propagate(a, p, f, dt);
      sE[0][slot] = p.attitude_cos_pitch_cos_roll; sE[1][slot] = f.uvw.x; sE[2][slot] = f.uvw.y; sE[3][slot] = f.uvw.z;
      sE[4][slot] = f.radius; sE[5][slot] = f.cosLatGc;
    }
    sRun[slot] = ran;
    TRIPLE_BARRIER(bar)                                        // 1
    WindAxes w;
    if (ran) {
      fdm_stage_gravity(f);
      fdm_stage_massbalance(a, f);
      fdm_stage_aux_kin(a, p, f, w);
      { int xi = 0; F16_X_KIN(XW_K) sK[xi][slot] = 2 * f.Vt; }
    }
    TRIPLE_BARRIER(bar)                                        // 2
    if (ran) {
      // what the frame keeps of role C's work: the engine reads Mach / density altitude / qbar / T / rho, the outputs h, Mach, Vc
      { int xi = 0; F16_X_AIR(XR_R) f.atm.T = sR[xi][slot]; f.atm.rho = sR[xi + 1][slot]; f.h_asl = sR[xi + 2][slot];
        f.vcas = sR[xi + 3][slot]; p.atmosphere_density_altitude = sR[xi + 4][slot]; }
      f.qbar = p.aero_qbar_psf; f.mach = p.velocities_mach;
      { int xi = 0; F16_X_SURF(XR_S) }
      eom_propulsion(a, p, f, sT, dt);
    }
    TRIPLE_BARRIER(bar)                                        // 3
    if (ran) {
      double c[6];
#pragma unroll
      for (int i = 0; i < 6; i++) c[i] = sSum[i][slot];
      fdm_stage_accelerations(a, f, w, c);
    }
    eom_after_frame(v, cfg, L, org, E, a, f, ran, k, K, sP, sWin, sShot, sCh);
  }
  eom_end<false>(v, L, E, a, p, s, K);
}

// ---------------------------------------------------------------------------------------------- four-warp frame
// The position chain of Propagate -- inertial position (Adams-Bashforth 3 on PAST velocities), earth rotation angle,
// ECEF location, geodetic latitude / altitude, local frame -- and the gravity and atmosphere that hang off it depend only
// on the inertial velocity the previous frame ended with.  A fourth role computes them ONE FRAME AHEAD, so they leave
// the critical path altogether:
//
//   role A (equations of motion)      role B (flight controls)   role C (air data)           role D (position, look-ahead)
//   Accelerations of frame k-1, missiles;                        reads atmosphere(k) from L
//   attitude / rate / velocity update, reads L(k), body matrices
//        -- E: cos(pitch)cos(roll), u, v, w, v_inertial(k), run flag -->
//   --------------------------------------------------------------------------------------------- barrier 1 (all four)
//   MassBalance, kinematic Auxiliary   FCS                        air-data Auxiliary          position(k+1), location,
//        -- K -->                          -- S -->                   -- R: qbar, Mach, Vc -->    gravity, atmosphere
//   ----------------------------------------------------------------------- barrier 2 (A, B, C)  ... -> L(k+1)
//   Propulsion                         axes SIDE ROLL YAW         axes DRAG LIFT PITCH
//   --------------------------------------------------------------------------------------------- barrier 3 (all four)
//
// L is read by A and C only between barrier 3 of frame k-1 and barrier 1 of frame k, and written by D only between
// barrier 1 and barrier 3 of frame k: one buffer.  D runs ahead of the aircraft's status: when the aircraft does not run
// frame k (shot down in frame k-1), D rolls its four state vectors back to the frame that did run.  The last frame of a
// step computes no look-ahead (the next step's first frame is computed by D's prologue from the stored state).
#ifndef ACS_SPLIT4_SLOTS
#define ACS_SPLIT4_SLOTS 64
#endif
constexpr int S4 = ACS_SPLIT4_SLOTS;
static_assert(S4 % 32 == 0 && S4 >= 32 && S4 <= 128, "whole warps, at most 4 quads per block (two named barriers each)");
#define S4_L_FIELDS(X) X(ri_x) X(ri_y) X(f.sin_epa) X(f.cos_epa) X(f.ecef.x) X(f.ecef.y) X(f.ecef.z) X(f.radius) X(f.rxy) X(f.geodAlt) \
  X(f.sinLatGd) X(f.cosLatGd) X(f.sinLon) X(f.cosLon) X(f.cosLatGc) X(f.gd_s1) X(f.gd_cc) \
  X(f.Ti2l.m[0][0]) X(f.Ti2l.m[0][1]) X(f.Ti2l.m[0][2]) X(f.Ti2l.m[1][0]) X(f.Ti2l.m[1][1]) X(f.Ti2l.m[1][2]) \
  X(f.Ti2l.m[2][0]) X(f.Ti2l.m[2][1]) X(f.Ti2l.m[2][2]) X(f.grav.x) X(f.grav.y) X(f.grav.z) X(f.h_asl) X(f.atm.T) X(f.atm.rho) \
  X(p.atmosphere_density_altitude)
constexpr int S4_NL = 33 + 2;                        // the list above + atm.P, atm.a (role C only)
constexpr int S4_NE = 7, S4_NK = F16_N_X_KIN + 1, S4_NS = F16_N_X_SURF, S4_NR = F16_N_X_AIR + 1, S4_NSUM = 6;
constexpr int S4_ROWS = S4_NL + S4_NE + S4_NK + S4_NS + S4_NR + S4_NSUM;
constexpr size_t S4_DYN_SMEM = sizeof(double) * S4_ROWS * S4;

#define QUAD_BARRIER_ALL(q) { __syncwarp(); asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * (q)) : "memory"); }
#define QUAD_BARRIER_ABC(q) { __syncwarp(); asm volatile("bar.sync %0, 96;" ::"r"(2 + 2 * (q)) : "memory"); }

// core state fields owned by role D of the four-warp frame (position chain); names of FDM_CORE_FIELDS
#define S4_CORE_ROLE_D(name) (f16_streq(name, "ri_x") || f16_streq(name, "ri_y") || f16_streq(name, "ri_z") || f16_streq(name, "epa") || \
  f16_streq(name, "dqv0_x") || f16_streq(name, "dqv0_y") || f16_streq(name, "dqv0_z") || f16_streq(name, "dqv1_x") || \
  f16_streq(name, "dqv1_y") || f16_streq(name, "dqv1_z"))
// ROLE: 0 = A (core minus D's fields + the carried properties the core publishes), 3 = D (position chain)
template <int ROLE>
ENV_DEV void store_state_role4(double* __restrict__ st, int N, int i, const AcCore& a, const Props& p, const FcsState& s) {
  int k = 0;
#define ST(name, expr) { constexpr bool d_ = S4_CORE_ROLE_D(name); if (d_ == (ROLE == 3)) st[(size_t)k * N + i] = expr; k++; }
  FDM_CORE_FIELDS(ST)
#undef ST
#define ST(name, expr) { constexpr bool b_ = F16_CARRIED_ROLE_B(name); if (!b_ && ROLE == 0) st[(size_t)k * N + i] = expr; k++; }
  F16_CARRIED_FIELDS(ST)
#undef ST
}

__global__ void __launch_bounds__(4 * S4, 1) k_env_substeps_split4(const EnvView v, const __grid_constant__ AcsTaskConfig cfg,
                                                                  const int lg, const int32_t* __restrict__ actions) {
  __shared__ double sT[F16_NTAB];
  __shared__ PubAc sP[S4];
  __shared__ int sWin[S4];
  __shared__ int sShot[S4];
  __shared__ PubChaff sCh[S4];
  __shared__ int sRun[S4];
  extern __shared__ double sDyn[];
  double (*sL)[S4] = reinterpret_cast<double (*)[S4]>(sDyn);
  double (*sE)[S4] = sL + S4_NL;
  double (*sK)[S4] = sE + S4_NE;
  double (*sS)[S4] = sK + S4_NK;
  double (*sR)[S4] = sS + S4_NS;
  double (*sSum)[S4] = sR + S4_NR;
  stage_tables(sT);
  const int role = threadIdx.x / S4;            // warp-uniform: 0 = A, 1 = B, 2 = C, 3 = D
  const int slot = threadIdx.x - role * S4;
  const int quad = slot >> 5;
  const Lane L = lane_of_slot(v, lg, slot, blockIdx.x * S4 + slot);
  const int K = cfg.substeps;
  const double dt = cfg.sim_dt, fcs_dt = cfg.fcs_dt;

  if (role == 3) {
    // ================================================================ role D: position chain, one frame ahead
    AcCore a; Props p; FcsState s; Frame f;
    bool loaded = false;
    if (L.valid && AI(v, AI_STATUS, L.row) == ST_ALIVE) {
      f16_props_init(p, s);
      load_state(v.fdm, v.rows, L.row, a, p, s);     // ri, dqv0, dqv1, epa (+ the velocity the last step ended with) are live
      loaded = true;
    }
    V3 pri = a.ri, pq0 = a.dqv0, pq1 = a.dqv1;        // the state of the last frame that is known to have run
    double pepa = a.epa;
    bool ahead = false;
    auto look_ahead = [&](const V3& v0) {
      pri = a.ri; pq0 = a.dqv0; pq1 = a.dqv1; pepa = a.epa;
      fdm_propagate_pos(a, f, v0, dt);
      fdm_stage_gravity(f);
      fdm_stage_atmosphere(p, f, g_atmo);
      const double ri_x = a.ri.x, ri_y = a.ri.y;
      { int xi = 0; S4_L_FIELDS(XW_L) sL[xi][slot] = f.atm.P; sL[xi + 1][slot] = f.atm.a; }
      ahead = true;
    };
    if (loaded) look_ahead(a.vi);                        // frame 0 of this step
    QUAD_BARRIER_ALL(quad)                               // 0: L(0) is there
    for (int k = 0; k < K; k++) {
      QUAD_BARRIER_ALL(quad)                             // 1
      const bool ran = sRun[slot] != 0;
      if (!ran) {
        if (ahead) { a.ri = pri; a.dqv0 = pq0; a.dqv1 = pq1; a.epa = pepa; ahead = false; }   // that frame never ran
      } else if (k < K - 1) {
        look_ahead(v3(sE[4][slot], sE[5][slot], sE[6][slot]));
      } else ahead = false;
      QUAD_BARRIER_ALL(quad)                             // 3: L(k+1) is there
    }
    if (loaded) store_state_role4<3>(v.fdm, v.rows, L.row, a, p, s);
    return;
  }
  if (role == 1) {
    role_fcs_axes<S4>(v, cfg, L, actions, sT, slot, QuadSync{quad}, sRun, sE, sK, sS, sR, sSum);
    return;
  }
  if (role == 2) {
    // ================================================================ role C: air data + axes DRAG, LIFT, PITCH (no state)
    Props p; FcsState s;
    f16_props_init(p, s);
    Frame f;
    QUAD_BARRIER_ALL(quad)                               // 0
    for (int k = 0; k < K; k++) {
      // the atmosphere of this frame, computed ahead by D (must be read before D overwrites L after barrier 1)
      f.atm.T = sL[30][slot]; f.atm.rho = sL[31][slot]; f.atm.P = sL[33][slot]; f.atm.a = sL[34][slot];
      QUAD_BARRIER_ALL(quad)                             // 1
      const bool ran = sRun[slot] != 0;
      if (ran) {
        f.uvw.x = sE[1][slot]; f.uvw.y = sE[2][slot]; f.uvw.z = sE[3][slot];
        fdm_airspeed(f);
        fdm_stage_aux_air(p, f, g_atmo);
        { int xi = 0; F16_X_AIR(XW_R) sR[xi][slot] = f.vcas; }
      }
      QUAD_BARRIER_ABC(quad)                             // 2
      if (ran) {
        double c[6];
        { int xi = 0; F16_X_KIN(XR_K) }
        { int xi = 0; F16_X_SURF(XR_S) }
        f16_aero<S3_AXES_C>(p, sT, 2 * f.Vt, c);
#pragma unroll
        for (int i = 0; i < 6; i++) if (S3_AXES_C & (1 << i)) sSum[i][slot] = c[i];
      }
      QUAD_BARRIER_ALL(quad)                             // 3
    }
    return;
  }

  // ================================================================== role A: equations of motion, propulsion, missiles
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  AcCore a; Props p; FcsState s; Frame f;
  EomLane E;
  eom_begin(v, cfg, L, E, false);
  if (E.was_alive) { f16_props_init(p, s); load_state(v.fdm, v.rows, L.row, a, p, s); }
  QUAD_BARRIER_ALL(quad)                                 // 0
  for (int k = 0; k < K; k++) {
    const bool ran = eom_runs(L, E);
    if (ran) {
      V3 v0;
      fdm_propagate_rot(a, dt, v0);
      double ri_x, ri_y;
      { int xi = 0; S4_L_FIELDS(XR_L) }
      fdm_propagate_combine(a, p, f, ri_x, ri_y);
      sE[0][slot] = p.attitude_cos_pitch_cos_roll; sE[1][slot] = f.uvw.x; sE[2][slot] = f.uvw.y; sE[3][slot] = f.uvw.z;
      sE[4][slot] = a.vi.x; sE[5][slot] = a.vi.y; sE[6][slot] = a.vi.z;
    }
    sRun[slot] = ran;
    QUAD_BARRIER_ALL(quad)                               // 1
    WindAxes w;
    if (ran) {
      fdm_stage_massbalance(a, f);
      fdm_stage_aux_kin(a, p, f, w);
      { int xi = 0; F16_X_KIN(XW_K) sK[xi][slot] = 2 * f.Vt; }
    }
    QUAD_BARRIER_ABC(quad)                               // 2
    if (ran) {
      { int xi = 0; F16_X_AIR(XR_R) f.vcas = sR[xi][slot]; }
      f.qbar = p.aero_qbar_psf; f.mach = p.velocities_mach;
      { int xi = 0; F16_X_SURF(XR_S) }
      eom_propulsion(a, p, f, sT, dt);
    }
    QUAD_BARRIER_ALL(quad)                               // 3
    if (ran) {
      double c[6];
#pragma unroll
      for (int i = 0; i < 6; i++) c[i] = sSum[i][slot];
      fdm_stage_accelerations(a, f, w, c);
    }
    eom_after_frame(v, cfg, L, org, E, a, f, ran, k, K, sP, sWin, sShot, sCh);
  }
  // as eom_end<false>, minus the position chain role D stores
  if (L.valid) {
    if (E.was_alive) {
      store_state_role4<0>(v.fdm, v.rows, L.row, a, p, s);
      store_out(v.out, v.rows, L.row, E.o);
      store_derived(v, L.row, E.me, E.v_mps, E.w_mps, E.vc_mps);
    }
    AI(v, AI_STATUS, L.row) = E.status;
    if (L.lane == 0) EI(v, EI_SUBSTEP_COUNT, L.env) = E.sc0 + K;
  }
}

#undef XW
#undef XR
#undef XW_K
#undef XR_K
#undef XW_S
#undef XR_S
#undef XW_R
#undef XR_R
#undef XW_L
#undef XR_L

// ============================================================================================== per-step logic
// k_env_post executes every instruction once per launch, so its time is the fetch of its own instruction stream -- and a
// kernel that carries all 8 observation packers x 6 launch rules x 12 reward classes streams 57 k SASS instructions to
// execute ~3 k of them (ncu: stall_no_inst on top).  The per-step logic is therefore compiled per TASK FAMILY: the
// observation packer, the launch rule and the set of reward classes are template constants (PostSpec), acs_env_create
// picks the instantiation that covers the task (same expressions, dead branches removed) and falls back to the generic
// one (all three read from AcsTaskConfig at run time) for a combination that has none.
template <int OBS, int LAUNCH, unsigned RMASK>
struct PostSpec {
  static constexpr int obs = OBS, launch = LAUNCH;
  static constexpr unsigned rmask = RMASK;
  static ENV_DEV int obs_kind(const AcsTaskConfig& c) { return OBS >= 0 ? OBS : c.obs_kind; }
  static ENV_DEV int launch_kind(const AcsTaskConfig& c) { return LAUNCH >= 0 ? LAUNCH : c.launch_kind; }
  static ENV_DEV bool has(const int kind) { return (RMASK >> kind) & 1u; }      // folds: `kind` is a literal at every use
};
using PostGeneric = PostSpec<-1, -1, 0xfffu>;

struct StepCtx {
  const EnvView& v;
  const AcsTaskConfig& cfg;
  const Lane& L;
  PubAc* sP;       // block-wide, index gbase + agent
  int cs;          // env.current_step (already incremented)
};

// first live incoming missile in under_missiles (append order): AircraftSimulator.check_missile_warning (simulatior.py:321-325)
ENV_DEV int missile_warning(const StepCtx& c, int agent) {
  int best = -1, best_born = 0x7fffffff;
  for (int j = 0; j < c.v.A; j++) {
    if (same_team(c.cfg, agent, j)) continue;
    const int row = c.L.env * c.v.A + j;
    const int nl = AI(c.v, AI_N_LAUNCHED, row);
    for (int s = 0; s < nl; s++) {
      const int mid = row * c.v.S + s;
      if (MI(c.v, MI_TARGET, mid) == agent && MI(c.v, MI_STATUS, mid) == MS_LAUNCHED) {
        const int b = MI(c.v, MI_BORN, mid);
        if (b < best_born) { best_born = b; best = mid; }
      }
    }
  }
  return best;
}
ENV_DEV void attack_geometry(const PubAc& ag, const PubAc& en, double& distance, double& ang_deg) {
  const double tx = en.f.n - ag.f.n, ty = en.f.e - ag.f.e, tz = en.f.u - ag.f.u;
  distance = em_sqrt0(tx * tx + ty * ty + tz * tz);
  const double hv = em_sqrt0(ag.f.vn * ag.f.vn + ag.f.ve * ag.f.ve + ag.f.vd * ag.f.vd);
  const double sum = tx * ag.f.vn + ty * ag.f.ve + tz * ag.f.vd;
  ang_deg = em_acos(env_clip(sum / (distance * hv + 1e-8), -1.0, 1.0)) * (180.0 / 3.14159265358979323846);
}
// MissileSimulator.create -> launch + target (simulatior.py:497-518), env.add_temp_simulator (env_base.py:90-92)
ENV_DEV int launch_missile(const StepCtx& c, int a, int target, int kind, int keyn) {
  const EnvView& v = c.v;
  const int row = c.L.env * v.A + a;
  const int slot = AI(v, AI_N_LAUNCHED, row);
  if (slot >= v.S) { EI(v, EI_FAULTS, c.L.env) += 1; return -1; }   // cannot happen: S = max launches per episode (taskspec.py)
  AI(v, AI_N_LAUNCHED, row) = slot + 1;
  const int mid = row * v.S + slot;
  const PubAc& pa = c.sP[c.L.gbase + a];
  const MissileParams pr = missile_params(kind);
  MD(v, MD_POS_N, mid) = pa.f.n; MD(v, MD_POS_E, mid) = pa.f.e; MD(v, MD_POS_U, mid) = pa.f.u;
  MD(v, MD_VEL_N, mid) = pa.f.vn; MD(v, MD_VEL_E, mid) = pa.f.ve; MD(v, MD_VEL_U, mid) = pa.f.vd;
  MD(v, MD_THETA, mid) = OUTF(v, O_PITCH, row); MD(v, MD_PHI, mid) = OUTF(v, O_HEADING, row);
  { double st, ct; em_sincos(OUTF(v, O_PITCH, row), &st, &ct); MD(v, MD_SIN_THETA, mid) = st; MD(v, MD_COS_THETA, mid) = ct; }
  MD(v, MD_ALT, mid) = pa.h; MD(v, MD_T, mid) = 0.0; MD(v, MD_M, mid) = pr.m0; MD(v, MD_DTHETA, mid) = 0.0; MD(v, MD_DPHI, mid) = 0.0;
  MD(v, MD_D_PREV, mid) = INFINITY;
  MI(v, MI_STATUS, mid) = MS_LAUNCHED; MI(v, MI_KIND, mid) = kind; MI(v, MI_TARGET, mid) = target; MI(v, MI_CONSEC, mid) = 0;
  MI(v, MI_KEYN, mid) = keyn; MI(v, MI_DETACHED, mid) = 0;
  // dict semantics of env._tempsims[uid] = sim: a re-used uid keeps the old entry's position and drops the old object
  int order = -1;
  for (int s = 0; s < slot; s++) {
    const int om = row * v.S + s;
    if (!MI(v, MI_DETACHED, om) && MI(v, MI_KEYN, om) == keyn) { MI(v, MI_DETACHED, om) = 1; order = MI(v, MI_ORDER, om); }
  }
  if (order < 0) { order = EI(v, EI_ORDER_SEQ, c.L.env); EI(v, EI_ORDER_SEQ, c.L.env) = order + 1; }
  MI(v, MI_ORDER, mid) = order;
  const int born = EI(v, EI_BORN_SEQ, c.L.env);
  EI(v, EI_BORN_SEQ, c.L.env) = born + 1;
  MI(v, MI_BORN, mid) = born;
  return slot;
}
ENV_DEV bool last_shot_free(const StepCtx& c, int row) {
  const int ls = AI(c.v, AI_LAST_SHOT_SLOT, row);
  if (ls < 0) return true;
  const int st = MI(c.v, MI_STATUS, row * c.v.S + ls);
  return st == MS_HIT || st == MS_MISS;
}
// scenario tasks: get_target = argmax distance (E/tasks/scenario2_task.py:150-156); a2a_launch_available (:116-148)
ENV_DEV int scenario_target(const StepCtx& c, int a) {
  const PubAc& ag = c.sP[c.L.gbase + a];
  int best = -1; double bd = -1.0;
  for (int j = 0; j < c.v.A; j++) {
    if (same_team(c.cfg, a, j)) continue;
    const PubAc& en = c.sP[c.L.gbase + j];
    const double tx = en.f.n - ag.f.n, ty = en.f.e - ag.f.e, tz = en.f.u - ag.f.u;
    const double d = em_sqrt0(tx * tx + ty * ty + tz * tz);
    if (d > bd) { bd = d; best = j; }   // np.argmax: first maximum
  }
  return best;
}
ENV_DEV void a2a_available(const StepCtx& c, int a, bool ret[3], int& enemy) {
  ret[0] = ret[1] = ret[2] = false;
  enemy = scenario_target(c, a);
  const PubAc& en = c.sP[c.L.gbase + enemy];
  if (en.status != ST_ALIVE) return;
  double distance, ang;
  attack_geometry(c.sP[c.L.gbase + a], en, distance, ang);
  if (distance / 1000 < 3 && ang < 5) ret[0] = true;
  if (distance / 1000 < 37 && ang < 90) ret[1] = true;
  if (distance / 1000 < 7 && ang < 90) ret[2] = true;
  if (c.cfg.use_baseline && a >= c.cfg.n_ego) {
    ret[1] = false;
    if (distance / 1000 < 37 && ang < 90 / 2) ret[1] = true;
  }
}

// task.step for agent a (executed by lane a while the other lanes of the env wait)
template <class S>
ENV_DEV void task_step_agent(const StepCtx& c, int a) {
  const EnvView& v = c.v;
  const AcsTaskConfig& cfg = c.cfg;
  const int row = c.L.env * v.A + a;
  PubAc* sP = c.sP + c.L.gbase;
  const bool alive = sP[a].status == ST_ALIVE;
  const int shoot = AI(v, AI_SHOOT, row);
  if (S::launch_kind(cfg) == ACS_L_RULE_LOCK) {          // E/tasks/singlecombat_with_missile_task.py:109-127
    int e0 = -1;
    for (int j = 0; j < v.A && e0 < 0; j++) if (!same_team(cfg, a, j)) e0 = j;
    double distance, ang;
    attack_geometry(sP[a], sP[e0], distance, ang);
    // deque(maxlen = lock_len) of booleans, as a bit window
    unsigned long long bits = ((unsigned long long)(unsigned)AI(v, AI_LOCK_HI, row) << 32) | (unsigned)AI(v, AI_LOCK_LO, row);
    int n = AI(v, AI_LOCK_N, row);
    bits = (bits << 1) | (ang < cfg.max_attack_angle ? 1ull : 0ull);
    n = min(n + 1, cfg.lock_len);
    const unsigned long long wmask = cfg.lock_len >= 64 ? ~0ull : ((1ull << cfg.lock_len) - 1ull);
    bits &= wmask;
    AI(v, AI_LOCK_LO, row) = (int)(unsigned)(bits & 0xffffffffull); AI(v, AI_LOCK_HI, row) = (int)(unsigned)(bits >> 32); AI(v, AI_LOCK_N, row) = n;
    const int locked = __popcll(bits);
    const int interval = c.cs - AI(v, AI_LAST_SHOOT_TIME, row);
    const int rem = AI(v, AI_REM_MISSILES, row);
    if (alive && locked >= cfg.lock_len && distance <= cfg.max_attack_distance && rem > 0 && interval >= cfg.min_attack_interval) {
      launch_missile(c, a, e0, 0, rem);
      AI(v, AI_REM_MISSILES, row) = rem - 1;
      AI(v, AI_LAST_SHOOT_TIME, row) = c.cs;
    }
  } else if (S::launch_kind(cfg) == ACS_L_RL_SINGLE) {   // E/tasks/singlecombat_with_missile_task.py:194-204
    int e0 = -1;
    for (int j = 0; j < v.A && e0 < 0; j++) if (!same_team(cfg, a, j)) e0 = j;
    const int rem = AI(v, AI_REM_MISSILES, row);
    if (alive && (shoot & 1) && rem > 0 && last_shot_free(c, row)) {
      AI(v, AI_LAST_SHOT_SLOT, row) = launch_missile(c, a, e0, 0, rem);
      AI(v, AI_REM_MISSILES, row) = rem - 1;
    }
  } else if (S::launch_kind(cfg) == ACS_L_RL_NEAREST) {  // E/tasks/multiplecombat_task.py:278-299
    int ti = -1; double bd = INFINITY;
    for (int j = 0; j < v.A; j++) {
      if (same_team(cfg, a, j)) continue;
      const double tx = sP[j].f.n - sP[a].f.n, ty = sP[j].f.e - sP[a].f.e, tz = sP[j].f.u - sP[a].f.u;
      const double d = em_sqrt0(tx * tx + ty * ty + tz * tz);
      if (d < bd) { bd = d; ti = j; }              // np.argmin: first minimum
    }
    double distance, ang;
    attack_geometry(sP[a], sP[ti], distance, ang);
    const int interval = c.cs - AI(v, AI_LAST_SHOOT_TIME, row);
    const int rem = AI(v, AI_REM_MISSILES, row);
    if (alive && (shoot & 1) && rem > 0 && ang <= cfg.max_attack_angle && distance <= cfg.max_attack_distance &&
        interval >= cfg.min_attack_interval) {
      launch_missile(c, a, ti, 0, rem);
      AI(v, AI_REM_MISSILES, row) = rem - 1;
      AI(v, AI_LAST_SHOOT_TIME, row) = c.cs;
    }
  } else if (S::launch_kind(cfg) == ACS_L_AUTO_GUN) {    // E/tasks/WVR_task.py:67-81, singlecombat_task.py:290-297
    const int enemy = scenario_target(c, a);          // farthest enemy; no ammunition, no alive checks on either side
    double distance, ang;
    attack_geometry(sP[a], sP[enemy], distance, ang);
    if (distance / 1000 < 3 && ang < 5) sP[enemy].bloods -= 5;
  } else if (S::launch_kind(cfg) == ACS_L_SCENARIO) {    // E/tasks/scenario2_task.py:73-114
    const bool f_gun = alive && (shoot & 1) && AI(v, AI_REM_GUN, row) > 0;
    const bool f_9m = alive && (shoot & 2) && AI(v, AI_REM_9M, row) > 0;
    const bool f_120 = alive && (shoot & 4) && AI(v, AI_REM_120B, row) > 0;
    const bool f_chaff = alive && (shoot & 8) && AI(v, AI_REM_CHAFF, row) > 0;
    bool av[3]; int enemy;
    if (f_gun && last_shot_free(c, row)) {
      a2a_available(c, a, av, enemy);
      if (av[0]) { sP[enemy].bloods -= 5; AI(v, AI_REM_GUN, row) -= 1; }
    }
    if (f_120 && last_shot_free(c, row)) {
      a2a_available(c, a, av, enemy);
      if (av[1]) {
        const int rem = AI(v, AI_REM_120B, row);
        AI(v, AI_LAST_SHOT_SLOT, row) = launch_missile(c, a, scenario_target(c, a), 1, rem);
        AI(v, AI_REM_120B, row) = rem - 1;
      }
    }
    if (f_9m && last_shot_free(c, row)) {
      a2a_available(c, a, av, enemy);
      if (av[2]) {
        const int rem = AI(v, AI_REM_9M, row);
        AI(v, AI_LAST_SHOT_SLOT, row) = launch_missile(c, a, scenario_target(c, a), 1, rem);
        AI(v, AI_REM_9M, row) = rem - 1;
      }
    }
    if (f_chaff && AI(v, AI_CH_STATE, row) != CH_ACTIVE) {
      // one chaff per missile (done ones included) within 1000 m that targets this aircraft (:105-111)
      int cnt = 0;
      for (int j = 0; j < v.A; j++) {
        if (same_team(cfg, a, j)) continue;
        const int jr = c.L.env * v.A + j;
        const int nl = AI(v, AI_N_LAUNCHED, jr);
        for (int s = 0; s < nl; s++) {
          const int mid = jr * v.S + s;
          if (MI(v, MI_DETACHED, mid) || MI(v, MI_TARGET, mid) != a) continue;
          const double dx = sP[a].f.n - MD(v, MD_POS_N, mid), dy = sP[a].f.e - MD(v, MD_POS_E, mid), dz = sP[a].f.u - MD(v, MD_POS_U, mid);
          if (em_sqrt0(dx * dx + dy * dy + dz * dz) < 1000) cnt++;
        }
      }
      if (cnt > 0) {
        AD(v, AD_CH_N, row) = sP[a].f.n; AD(v, AD_CH_E, row) = sP[a].f.e; AD(v, AD_CH_U, row) = sP[a].f.u; AD(v, AD_CH_T, row) = 0.0;
        AI(v, AI_CH_STATE, row) = CH_ACTIVE; AI(v, AI_CH_COUNT, row) = cnt;
        AI(v, AI_REM_CHAFF, row) -= cnt;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- observations
ENV_DEV void obs_ego9(const EnvView& v, int row, const PubAc& s, double* o) {
  double sr, cr, sp, cp;
  em_sincos(OUTF(v, O_ROLL, row), &sr, &cr);
  em_sincos(OUTF(v, O_PITCH, row), &sp, &cp);
  o[0] = s.h / 5000; o[1] = sr; o[2] = cr; o[3] = sp; o[4] = cp;
  o[5] = s.u_mps / 340; o[6] = AD(v, AD_V_MPS, row) / 340; o[7] = AD(v, AD_W_MPS, row) / 340; o[8] = AD(v, AD_VC_MPS, row) / 340;
}
ENV_DEV void obs_rel6(const PubAc& ego, const PubAc& other, bool two_d, double* o) {
  const AoTaR g = get_ao_ta_r(ego.f, other.f, two_d);
  o[0] = (other.u_mps - ego.u_mps) / 340; o[1] = (other.h - ego.h) / 1000; o[2] = g.AO; o[3] = g.TA; o[4] = g.R / 10000; o[5] = g.side;
}
ENV_DEV bool obs_missile6(const StepCtx& c, int a, double* o) {
  const int mid = missile_warning(c, a);
  if (mid < 0) return false;
  const EnvView& v = c.v;
  const PubAc& ego = c.sP[c.L.gbase + a];
  Feat mf;
  mf.n = MD(v, MD_POS_N, mid); mf.e = MD(v, MD_POS_E, mid); mf.u = MD(v, MD_POS_U, mid);
  mf.vn = MD(v, MD_VEL_N, mid); mf.ve = MD(v, MD_VEL_E, mid); mf.vd = MD(v, MD_VEL_U, mid);
  const AoTaR g = get_ao_ta_r(ego.f, mf, false);
  o[0] = (em_sqrt0(mf.vn * mf.vn + mf.ve * mf.ve + mf.vd * mf.vd) - ego.u_mps) / 340; o[1] = (mf.u - ego.h) / 1000;
  o[2] = g.AO; o[3] = g.TA; o[4] = g.R / 10000; o[5] = g.side;
  return true;
}
// get_obs for agent a into o[obs_dim] (global memory); citations per branch in taskspec.py
template <class S>
ENV_DEV void write_obs(const StepCtx& c, int a, double* __restrict__ o) {
  const EnvView& v = c.v;
  const AcsTaskConfig& cfg = c.cfg;
  const int row = c.L.env * v.A + a;
  const PubAc* sP = c.sP + c.L.gbase;
  const PubAc& s = sP[a];
  const int D = cfg.obs_dim, k = S::obs_kind(cfg);
  double t[9];
  for (int i = 0; i < D; i++) o[i] = 0.0;
  if (k == ACS_OBS_HEADING) {                      // E/tasks/heading_task.py:67-100
    const double psi_deg = OUTF(v, O_HEADING, row) * RADTODEG;
    const double d_alt = env_clip((ED(v, ED_TGT_ALT, c.L.env) - OUTF(v, O_H_SL_FT, row)) * 0.3048, -40000.0, 40000.0);
    const double d_head = delta_heading_deg(ED(v, ED_TGT_HEADING, c.L.env), psi_deg);
    const double d_vel = env_clip(ED(v, ED_TGT_VEL, c.L.env) - s.u_mps, -1400.0, 1400.0);
    o[0] = d_alt / 1000; o[1] = d_head / 180 * 3.14159265358979323846; o[2] = d_vel / 340;
    obs_ego9(v, row, s, t);
    for (int i = 0; i < 9; i++) o[3 + i] = t[i];
    for (int i = 0; i < 12; i++) o[i] = env_clip(o[i], -10.0, 10.0);
    return;
  }
  obs_ego9(v, row, s, t);
  for (int i = 0; i < 9; i++) o[i] = t[i];
  double r6[6];
  if (k == ACS_OBS_1V1) {                          // E/tasks/singlecombat_task.py:88-139
    int e0 = -1;
    for (int j = 0; j < v.A && e0 < 0; j++) if (!same_team(cfg, a, j)) e0 = j;
    obs_rel6(s, sP[e0], true, r6);
    for (int i = 0; i < 6; i++) o[9 + i] = r6[i];
    for (int i = 0; i < 15; i++) o[i] = env_clip(o[i], -10.0, 10.0);
    return;
  }
  if (k == ACS_OBS_1V1_RWR) {                      // E/tasks/scenario1_task.py:222-314: nearest LIVE enemy, else enemies[0]
    int e = -1, first = -1;
    double bd = INFINITY;
    for (int j = 0; j < v.A; j++) {
      if (same_team(cfg, a, j)) continue;
      if (first < 0) first = j;
      const double tx = sP[j].f.n - s.f.n, ty = sP[j].f.e - s.f.e, tz = sP[j].f.u - s.f.u;
      const double d = em_sqrt0(tx * tx + ty * ty + tz * tz);
      if (sP[j].status == ST_ALIVE && d < bd) { bd = d; e = j; }   // stable sort by distance: the first minimum wins
    }
    if (e < 0) e = first;
    obs_rel6(s, sP[e], false, r6);
    for (int i = 0; i < 6; i++) o[9 + i] = r6[i];
    return;                                          // missile_sim is hard-wired to None; [15..22] stay 0
  }
  if (k == ACS_OBS_1V1_MISSILE || k == ACS_OBS_NV_MISSILE) {   // E/tasks/singlecombat_with_missile_task.py:31-99; multiplecombat_with_missile_task.py:32-117
    const int ti = (k == ACS_OBS_1V1_MISSILE) ? 0 : (a < cfg.n_ego ? a : a - cfg.n_ego);
    int cnt = 0, e = -1;
    for (int j = 0; j < v.A; j++) if (!same_team(cfg, a, j)) { if (cnt == ti) e = j; cnt++; }
    obs_rel6(s, sP[e], false, r6);
    for (int i = 0; i < 6; i++) o[9 + i] = r6[i];
    if (obs_missile6(c, a, r6)) for (int i = 0; i < 6; i++) o[15 + i] = r6[i];
    return;
  }
  int off = 9;
  if (k == ACS_OBS_MULTI || k == ACS_OBS_MULTI_MISSILE || k == ACS_OBS_NVN) {   // E/tasks/multiplecombat_task.py:105-135,232-267; scenario2_task.py:256-316
    for (int pass = 0; pass < 2; pass++)
      for (int j = 0; j < v.A; j++) {
        if (j == a || same_team(cfg, a, j) != (pass == 0)) continue;
        obs_rel6(s, sP[j], false, r6);
        for (int i = 0; i < 6; i++) o[off + i] = r6[i];
        off += 6;
      }
    if (k != ACS_OBS_NVN) for (int i = 0; i < off; i++) o[i] = env_clip(o[i], -10.0, 10.0);
    if (k != ACS_OBS_MULTI && obs_missile6(c, a, r6)) for (int i = 0; i < 6; i++) o[off + i] = r6[i];
  }
}

// ---------------------------------------------------------------------------------------------- rewards
ENV_DEV double reward_process(const StepCtx& c, int ri, int row, double new_reward) {   // BaseRewardFunction._process
  const AcsRewardSpec& r = c.cfg.rewards[ri];
  double reward = new_reward * r.scale;
  if (r.potential) {
    const double pre = AD(c.v, AD_PRE_REWARD0 + ri, row);
    AD(c.v, AD_PRE_REWARD0 + ri, row) = reward;
    reward = reward - pre;
  }
  return reward;
}
// geo[0..n) = get_AO_TA_R of agent a against each enemy in enemy order, computed once per agent per step (the reference
// recomputes it inside every reward class)
template <class S>
ENV_DEV double reward_one(const StepCtx& c, int ri, int a, const AoTaR* geo, int n) {
  const EnvView& v = c.v;
  const AcsTaskConfig& cfg = c.cfg;
  const AcsRewardSpec& r = cfg.rewards[ri];
  const int env = c.L.env, row = env * v.A + a;
  const PubAc* sP = c.sP + c.L.gbase;
  const PubAc& s = sP[a];
  const double FT = 1 / 3.28084, PI = 3.14159265358979323846;
  const int kind = r.kind;
  if (S::has(ACS_R_ALTITUDE) && kind == ACS_R_ALTITUDE) {            // altitude_reward.py:20-40
      const double ego_z = s.f.u / 1000, ego_vz = s.f.vd / 340;
      double Pv = 0., PH = 0.;
      if (ego_z <= r.p0) Pv = -env_clip(ego_vz / r.p2 * (r.p0 - ego_z) / r.p0, 0., 1.);
      if (ego_z <= r.p1) PH = env_clip(ego_z / r.p1, 0., 1.) - 1. - 1.;
      return reward_process(c, ri, row, Pv + PH);
    }
  if (S::has(ACS_R_POSTURE) && kind == ACS_R_POSTURE) {             // posture_reward.py:26-75
      double nr = 0;
      for (int i = 0; i < n; i++) nr += posture_orientation((int)r.p0, geo[i].AO, geo[i].TA) * posture_range((int)r.p1, geo[i].R / 1000, r.p2);
      return reward_process(c, ri, row, nr);
    }
  if (S::has(ACS_R_EVENT) && kind == ACS_R_EVENT) {               // event_driven_reward.py:15-34
      double rew = 0;
      if (s.status == ST_SHOTDOWN) rew -= 200; else if (s.status == ST_CRASH) rew -= 200;
      const int nl = AI(v, AI_N_LAUNCHED, row);
      for (int q = 0; q < nl; q++) if (MI(v, MI_STATUS, row * v.S + q) == MS_HIT) rew += 200;
      return reward_process(c, ri, row, rew);
    }
  if (S::has(ACS_R_MISSILE_POSTURE) && kind == ACS_R_MISSILE_POSTURE) {     // missile_posture_reward.py:18-46 (bypasses _process; previous_missile_v aliases a live array)
      double rew = 0;
      const int mid = missile_warning(c, a);
      if (mid >= 0) {
        int ref = EI(v, EI_PMV_REF, env);
        if (ref < 0) { ref = mid; EI(v, EI_PMV_REF, env) = ref; }
        const double mvx = MD(v, MD_VEL_N, mid), mvy = MD(v, MD_VEL_E, mid), mvz = MD(v, MD_VEL_U, mid);
        const double pvx = MD(v, MD_VEL_N, ref), pvy = MD(v, MD_VEL_E, ref), pvz = MD(v, MD_VEL_U, ref);
        const double nm = em_sqrt0(mvx * mvx + mvy * mvy + mvz * mvz), np_ = em_sqrt0(pvx * pvx + pvy * pvy + pvz * pvz);
        const double na = em_sqrt0(s.f.vn * s.f.vn + s.f.ve * s.f.ve + s.f.vd * s.f.vd);
        const double v_dec = (np_ - nm) / 340 * r.scale;
        const double ang = (mvx * s.f.vn + mvy * s.f.ve + mvz * s.f.vd) / (nm * na);
        rew = ang < 0 ? ang / (env_max(v_dec, 0.0) + 1) : ang * env_max(v_dec, 0.0);
      } else {
        EI(v, EI_PMV_REF, env) = -1;
      }
      return rew;
    }
  if (S::has(ACS_R_SHOOT_PENALTY) && kind == ACS_R_SHOOT_PENALTY) {       // shoot_penalty_reward.py:17-32
      double rew = 0;
      const int rem = AI(v, AI_REM_MISSILES, row);
      if (rem == AI(v, AI_PRE_REMAINING, row) - 1) rew -= 30;
      AI(v, AI_PRE_REMAINING, row) = rem;
      return reward_process(c, ri, row, rew);
    }
  if (S::has(ACS_R_HEADING) && kind == ACS_R_HEADING) {             // heading_reward.py:18-71
      const double roll = OUTF(v, O_ROLL, row), p = OUTF(v, O_P, row), q = OUTF(v, O_Q, row);
      const double psi_deg = OUTF(v, O_HEADING, row) * RADTODEG;
      const double d_head = delta_heading_deg(ED(v, ED_TGT_HEADING, env), psi_deg);
      const double d_alt = env_clip((ED(v, ED_TGT_ALT, env) - OUTF(v, O_H_SL_FT, row)) * 0.3048, -40000.0, 40000.0);
      const double d_vel = env_clip(ED(v, ED_TGT_VEL, env) - s.u_mps, -1400.0, 1400.0);
      const double heading_r = em_exp(-((d_head / 5.0) * (d_head / 5.0)));
      const double alt_r = em_exp(-((d_alt / 15.24) * (d_alt / 15.24)));
      const double roll_r = em_exp(-((roll / 0.35) * (roll / 0.35)));
      const double speed_r = em_exp(-((d_vel / 24) * (d_vel / 24)));
      double rew = em_sqrt0(em_sqrt0(heading_r * alt_r * roll_r * speed_r));   // x^(1/4)
      if (c.cs > 1) rew = rew + (-fabs(p - AD(v, AD_HR_P, row)) * 1.0) + (-fabs(q - AD(v, AD_HR_Q, row)) * 1.0);
      AD(v, AD_HR_ROLL, row) = roll; AD(v, AD_HR_P, row) = p; AD(v, AD_HR_Q, row) = q;
      return reward_process(c, ri, row, rew);
    }
  if (S::has(ACS_R_RELATIVE_ALTITUDE) && kind == ACS_R_RELATIVE_ALTITUDE) {   // relative_altitude_reward.py:18-32
      int e0 = -1;
      for (int j = 0; j < v.A && e0 < 0; j++) if (!same_team(cfg, a, j)) e0 = j;
      const double ego_z = s.f.u / 1000, enm_z = sP[e0].f.u / 1000;
      return reward_process(c, ri, row, env_min(r.p0 - fabs(ego_z - enm_z), 0.0));
    }
  // per-enemy geometry rewards share the (AO, TA, R) list
  double nr = 0;
  if (S::has(ACS_R_COMBAT_GEOMETRY) && kind == ACS_R_COMBAT_GEOMETRY) {          // combat_geometry_reward.py:28-68 (index never advances; prev list only grows)
    if (!EI(v, EI_CG_VALID, env)) { ED(v, ED_CG_PREV0, env) = geo[0].AO; ED(v, ED_CG_PREV1, env) = geo[0].TA; EI(v, EI_CG_VALID, env) = 1; }
    const double p0 = ED(v, ED_CG_PREV0, env), p1 = ED(v, ED_CG_PREV1, env);
    for (int i = 0; i < n; i++) nr += -(geo[0].AO - p0) - (geo[0].TA - p1);
  } else if (S::has(ACS_R_GUN_BEHIT) && kind == ACS_R_GUN_BEHIT) {         // gun_behit_reward.py:27-54
    for (int i = 0; i < n; i++) if (geo[i].R >= 500 * FT && geo[i].R <= 3000 * FT && geo[i].AO >= 179 * PI / 180) nr += -5;
  } else if (S::has(ACS_R_GUN_WEZ) && kind == ACS_R_GUN_WEZ) {           // gun_WEZ_reward.py:28-55
    for (int i = 0; i < n; i++)
      if (geo[i].R >= 500 * FT && geo[i].R <= 3000 * FT && geo[i].AO <= 1 * PI / 180) nr += 5 + 5 * (3000 * FT - geo[i].R) / (2500 * FT);
  } else if ((S::has(ACS_R_GUN_TARGETTAIL) && kind == ACS_R_GUN_TARGETTAIL) || (S::has(ACS_R_GUN_WEZDOT) && kind == ACS_R_GUN_WEZDOT)) {   // gun_targettail_reward.py:28-78; gun_WEZDOT_reward.py:29-77
    const bool tt = kind == ACS_R_GUN_TARGETTAIL;
    double d[ACS_MAX_AGENTS];
    for (int i = 0; i < n; i++) {
      const double R = geo[i].R;
      if (tt) {
        if (R >= 3000 * FT && R <= 5000 * FT) d[i] = R * em_sin(geo[i].TA);
        else if (R <= 3000 * FT) d[i] = em_sqrt0(R * R + (3000 * FT) * (3000 * FT) - 2 * R * (3000 * FT) * em_cos(geo[i].TA));
        else d[i] = em_sqrt0(R * R + (5000 * FT) * (5000 * FT) - 2 * R * (5000 * FT) * em_cos(geo[i].TA));
      } else {
        if (R >= 500 * FT && R <= 3000 * FT) d[i] = R * em_sin(geo[i].AO);
        else d[i] = em_sqrt0(R * R + (3000 * FT) * (3000 * FT) - 2 * R * (3000 * FT) * em_cos(geo[i].AO));
      }
    }
    const int fvalid = tt ? EI_TT_VALID : EI_WD_VALID, fprev = tt ? ED_TT_PREV0 : ED_WD_PREV0;
    if (!EI(v, fvalid, env)) {     // the first evaluation after reset records prev = [d0, d0, d1, ...]
      ED(v, fprev, env) = d[0];
      for (int i = 1; i < n; i++) ED(v, fprev + i, env) = d[i - 1];
      EI(v, fvalid, env) = 1;
    }
    for (int i = 0; i < n; i++) nr += -1 / 60.0 * em_tanh((d[i] - ED(v, fprev + i, env)) / em_sqrt0(geo[i].R));
  }
  return reward_process(c, ri, row, nr);
}
// Rewards whose evaluation order across the agents of an env matters (they share env-level state): the
// first evaluation after a reset of the three "prev list" rewards, and MissilePostureReward's shared reference always.
ENV_DEV bool reward_is_order_dependent(const StepCtx& c, int kind) {
  const int env = c.L.env;
  if (kind == ACS_R_MISSILE_POSTURE) return true;
  if (kind == ACS_R_COMBAT_GEOMETRY) return !EI(c.v, EI_CG_VALID, env);
  if (kind == ACS_R_GUN_TARGETTAIL) return !EI(c.v, EI_TT_VALID, env);
  if (kind == ACS_R_GUN_WEZDOT) return !EI(c.v, EI_WD_VALID, env);
  return false;
}
ENV_DEV int enemy_geometry(const StepCtx& c, int a, AoTaR* geo) {
  int n = 0;
  const PubAc* sP = c.sP + c.L.gbase;
  for (int j = 0; j < c.v.A; j++) if (!same_team(c.cfg, a, j)) geo[n++] = get_ao_ta_r(sP[a].f, sP[j].f, false);
  return n;
}
// task.get_reward gating for agent a (E/tasks/singlecombat_task.py:190-195, multiplecombat_task.py:147-151): true = the
// agent gets no reward this step (and no reward function is evaluated for it)
ENV_DEV bool reward_gated(const StepCtx& c, int a) {
  const int row = c.L.env * c.v.A + a;
  const PubAc& s = c.sP[c.L.gbase + a];
  if (c.cfg.reward_gate == ACS_G_DIE_FLAG) {
    if (AI(c.v, AI_DIE_FLAG, row)) return true;
    AI(c.v, AI_DIE_FLAG, row) = (s.status != ST_ALIVE);
  } else if (c.cfg.reward_gate == ACS_G_ALIVE) {
    if (s.status != ST_ALIVE) return true;
  }
  return false;
}
// All agents' rewards of one env.  The reference evaluates agent by agent, reward class by reward class; the classes
// without cross-agent state are evaluated by all lanes at once, the order-dependent ones lane after lane, and the
// per-agent total is summed in the reference's class order.
template <class S>
ENV_DEV double env_rewards(const StepCtx& c, bool valid) {
  const Lane& L = c.L;
  const int nr = c.cfg.n_rewards;
  double vals[ACS_MAX_REWARDS];
  AoTaR geo[ACS_MAX_AGENTS];
  int n = 0;
  bool gated = true;
  unsigned serial = 0;
  if (valid) {
    gated = reward_gated(c, L.lane);
    for (int ri = 0; ri < nr; ri++) if (reward_is_order_dependent(c, c.cfg.rewards[ri].kind)) serial |= 1u << ri;
    if (!gated) n = enemy_geometry(c, L.lane, geo);
  }
  serial = group_or(L, serial);
  // pass 0: every lane evaluates the classes without cross-agent state; passes 1 .. A (only when some class is
  // order-dependent this step): lane pass-1 evaluates those -- one call site of the reward code
  const int passes = serial ? c.v.A + 1 : 1;
  for (int pass = 0; pass < passes; pass++) {
    if (valid && !gated && (pass == 0 || L.lane == pass - 1)) {
      const unsigned sel = pass == 0 ? ~serial : serial;
      for (int ri = 0; ri < nr; ri++) if (sel >> ri & 1) vals[ri] = reward_one<S>(c, ri, L.lane, geo, n);
    }
    if (pass > 0) __syncwarp(L.gmask);
  }
  double tot = 0.0;
  if (valid && !gated) for (int ri = 0; ri < nr; ri++) tot += vals[ri];
  return tot;
}

// ---------------------------------------------------------------------------------------------- terminations
// task.get_termination for agent a: conditions in order, first done short-circuits (E/tasks/task_base.py:90-112).
// Returns the cause (ACS_T_*) or -1.  crash() side effects go to sP[a].status.
template <class S>
ENV_DEV int agent_termination(const StepCtx& c, int a) {
  const EnvView& v = c.v;
  const AcsTaskConfig& cfg = c.cfg;
  const int env = c.L.env, row = env * v.A + a;
  PubAc* sP = c.sP + c.L.gbase;
  for (int ti = 0; ti < cfg.n_terms; ti++) {
    const int t = cfg.terms[ti];
    bool done = false;
    if ((S::obs < 0 || S::obs == ACS_OBS_HEADING) && t == ACS_T_UNREACH_HEADING) {              // unreach_heading.py:22-65 (heading task only)
      const double sim_time = v.fdm[(size_t)F_SIM_TIME * v.rows + row];
      const double check_time = ED(v, ED_CHECK_TIME, env);
      if (sim_time >= check_time) {
        const double d_head = delta_heading_deg(ED(v, ED_TGT_HEADING, env), OUTF(v, O_HEADING, row) * RADTODEG);
        if (fabs(d_head) > 10) done = true;
        else {
          const int tc = EI(v, EI_TURN_COUNTS, env);
          const double inc_tab[5] = {0.2, 0.4, 0.6, 0.8, 1.0};
          const double inc = inc_tab[tc < 4 ? tc : 4];
          const int ep = EI(v, EI_EPISODE, env);
          const double d0 = env_u01(cfg.seed, cfg.env_offset + env, RNG_HEADING, ep, tc, 0);
          const double d1 = env_u01(cfg.seed, cfg.env_offset + env, RNG_HEADING, ep, tc, 1);
          const double d2 = env_u01(cfg.seed, cfg.env_offset + env, RNG_HEADING, ep, tc, 2);
          const double dh = (-inc + 2 * inc * d0) * cfg.heading_increments[0];
          const double da = (-inc + 2 * inc * d1) * cfg.heading_increments[1];
          const double dv = (-inc + 2 * inc * d2) * cfg.heading_increments[2];
          double nh = fmod(ED(v, ED_TGT_HEADING, env) + dh + 360, 360.0);
          if (nh < 0) nh += 360.0;
          ED(v, ED_TGT_HEADING, env) = env_clip(nh, 0.0, 360.0);
          ED(v, ED_TGT_ALT, env) = env_clip(ED(v, ED_TGT_ALT, env) + da, -1400.0, 85000.0);
          ED(v, ED_TGT_VEL, env) = env_clip(ED(v, ED_TGT_VEL, env) + dv, -700.0, 700.0);
          ED(v, ED_CHECK_TIME, env) = env_clip(check_time + cfg.check_interval, 0.0, 1000000.0);
          EI(v, EI_TURN_COUNTS, env) = tc + 1;
        }
      }
    } else if (t == ACS_T_EXTREME_STATE) {         // extreme_state.py:14-33 + catalog.py:386-416
      const double pp = OUTF(v, O_P, row), qq = OUTF(v, O_Q, row), rr = OUTF(v, O_R, row);
      const bool ev = OUTF(v, O_ECI_VMAG, row) >= 1e10;
      const bool er = em_sqrt0(pp * pp + qq * qq + rr * rr) >= 1000;
      const bool ea = OUTF(v, O_H_SL_FT, row) >= 1e10;
      const bool eacc = env_max(env_max(fabs(OUTF(v, O_NPX, row)), fabs(OUTF(v, O_NPY, row))), fabs(OUTF(v, O_NPZ, row))) > 1e1;
      done = ea || er || ev || eacc;
      if (done) sP[a].status = ST_CRASH;
    } else if (t == ACS_T_OVERLOAD) {              // overload.py:18-46
      const double sim_time = v.fdm[(size_t)F_SIM_TIME * v.rows + row];
      if (sim_time > 10)
        if (fabs(OUTF(v, O_NPX, row)) > cfg.acc_limit[0] || fabs(OUTF(v, O_NPY, row)) > cfg.acc_limit[1] ||
            fabs(OUTF(v, O_NPZ, row) + 1) > cfg.acc_limit[2]) done = true;
      if (done) sP[a].status = ST_CRASH;
    } else if (t == ACS_T_LOW_ALTITUDE) {          // low_altitude.py:15-34
      done = sP[a].h <= cfg.altitude_limit;
      if (done) sP[a].status = ST_CRASH;
    } else if (t == ACS_T_TIMEOUT) {               // timeout.py:14-32
      done = c.cs >= cfg.max_steps;
    } else if (t == ACS_T_SAFE_RETURN) {           // safe_return.py:15-50
      if (sP[a].status == ST_SHOTDOWN || sP[a].status == ST_CRASH) done = true;
      else {
        bool all_dead = true;
        for (int j = 0; j < v.A; j++) if (!same_team(cfg, a, j) && sP[j].status == ST_ALIVE) all_dead = false;
        if (all_dead && missile_warning(c, a) < 0) done = true;
      }
    }
    if (done) return t;
  }
  return -1;
}

ENV_DEV void reset_copy_warp(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const bool on, double* __restrict__ obs,
                             double* __restrict__ share_obs, const ResetTpl& tp);

// fuse_reset != 0 (auto-reset with a reset template, i.e. every task but the heading task): an env whose agents are all
// done is reset right here -- rewards / dones / info of the terminal step are already written, the reset observation
// replaces the terminal one (R/envs/env_wrappers.py:191-204) -- instead of by two more kernel launches whose code
// would be fetched cold.
//
// obs_split != 0 (blockDim = 256): get_obs runs on a second set of four warps next to the terminations and rewards.  The
// kernel executes every instruction once, so its time is the fetch of its own instruction stream; two streams fetch in
// parallel.  Legal when nothing the step logic writes is read by get_obs: no weapons (task.step is empty) and not the
// heading task (UnreachHeading re-targets what the observation shows); get_obs precedes terminations and rewards in the
// reference (E/envs/env_base.py:155-171), and here it reads its own copy of the pre-termination aircraft state.
template <class S>
__global__ void __launch_bounds__(256, 2) k_env_post(const __grid_constant__ EnvView v, const __grid_constant__ AcsTaskConfig cfg, const int lg, double* __restrict__ obs,
                                                  double* __restrict__ share_obs, double* __restrict__ rewards,
                                                  uint8_t* __restrict__ dones, int32_t* __restrict__ info, uint8_t* __restrict__ env_done,
                                                  const int fuse_reset, const __grid_constant__ ResetTpl tpl, const int obs_split) {
  __shared__ PubAc sP2[2][128];
  __shared__ double sRew[128];
  __shared__ int sDone[128];
  const int obs_role = obs_split ? (threadIdx.x >> 7) : 0;      // warp-uniform
  PubAc* sP = sP2[obs_role];
  const Lane L = lane_of_slot(v, lg, threadIdx.x & 127, blockIdx.x * 128 + (threadIdx.x & 127));
  const int A = v.A;
  if (fuse_reset && tpl.full && !obs_role) tpl_prefetch(tpl);     // long before the first env of this warp can need it
  PubAc me;
  me.status = ST_CRASH; me.bloods = 0; me.h = 0; me.u_mps = 0; me.f.n = me.f.e = me.f.u = me.f.vn = me.f.ve = me.f.vd = 0;
  if (L.valid) load_pub(v, L.row, me);
  sP[L.tid] = me;
  const int cs = (L.env < v.B) ? EI(v, EI_CURRENT_STEP, L.env) + 1 : 0;
  __syncwarp(L.gmask);
  const StepCtx c{v, cfg, L, sP, cs};
  if (!obs_role) {
    sRew[L.tid] = 0.0;
    sDone[L.tid] = 1;
    // ---- task.step: artillery (E/tasks/singlecombat_task.py:163-188), then the launch rules, agents in dict order
    if (cfg.use_artillery) {
      for (int a = 0; a < A; a++) {
        if (L.valid && L.lane == a) {
          for (int j = 0; j < A; j++) {
            if (same_team(cfg, a, j) || sP[L.gbase + j].status != ST_ALIVE) continue;
            const AoTaR g = get_ao_ta_r(sP[L.gbase + a].f, sP[L.gbase + j].f, false);
            double of = 0.0;
            if (g.AO >= 0 && g.AO <= 0.5236) of = 1 - g.AO / 0.5236; else if (g.AO >= -0.5236 && g.AO <= 0) of = 1 + g.AO / 0.5236;
            const double Rk = g.R / 1000;
            const double df = Rk <= 1 ? 1.0 : (Rk <= 3 ? (3 - Rk) / 2. : 0.0);
            sP[L.gbase + j].bloods -= of * df;
          }
        }
        __syncwarp(L.gmask);
      }
    }
    if (S::launch_kind(cfg) != ACS_L_NONE) {
      for (int a = 0; a < A; a++) {
        if (L.valid && L.lane == a) task_step_agent<S>(c, a);
        __syncwarp(L.gmask);
      }
    }
  }
  // ---- get_obs (every agent, before terminations / rewards): by the observation warps when there are any (the one call
  // site of the packer in this kernel; task.step above is empty whenever obs_split is legal)
  if ((obs_role || !obs_split) && L.valid) write_obs<S>(c, L.lane, obs + ((size_t)L.env * A + L.lane) * cfg.obs_dim);
  if (obs_role) {
    __syncthreads();        // pairs with the one below: the observations are complete
    return;
  }
  __syncwarp(L.gmask);
  // ---- dones and rewards in the reference's order
  int cause = -1;
  if (cfg.dones_before_rewards) {
    for (int a = 0; a < A; a++) {
      if (L.valid && L.lane == a) cause = agent_termination<S>(c, a);
      __syncwarp(L.gmask);
    }
    sRew[L.tid] = env_rewards<S>(c, L.valid);
    __syncwarp(L.gmask);
  } else {
    sRew[L.tid] = env_rewards<S>(c, L.valid);
    __syncwarp(L.gmask);
    if (cfg.team_mean) {                            // E/envs/multiplecombat_env.py:170-175
      double ego = 0.0, enm = 0.0;
      for (int j = 0; j < cfg.n_ego; j++) ego += sRew[L.gbase + j];
      for (int j = cfg.n_ego; j < A; j++) enm += sRew[L.gbase + j];
      ego /= cfg.n_ego; if (cfg.n_enm > 0) enm /= cfg.n_enm;
      __syncwarp(L.gmask);
      sRew[L.tid] = L.lane < cfg.n_ego ? ego : enm;
    }
    for (int a = 0; a < A; a++) {
      if (L.valid && L.lane == a) cause = agent_termination<S>(c, a);
      __syncwarp(L.gmask);
    }
  }
  if (L.valid) sDone[L.tid] = cause >= 0;
  // win of this agent as the reference's `success` flag survives the termination loop: ended ALIVE through SafeReturn
  const bool won = L.valid && L.lane < cfg.n_ego && cause == ACS_T_SAFE_RETURN && sP[L.tid].status == ST_ALIVE;
  const bool env_won = (__ballot_sync(0xffffffffu, won) & L.gmask) != 0;
  __syncwarp(L.gmask);
  if (obs_split) __syncthreads();     // the observation warps are done (share_obs and the reset below read / replace obs)
  if (L.valid) {
    const size_t oa = (size_t)L.env * A + L.lane;
    rewards[oa] = sRew[L.tid];
    dones[oa] = cause >= 0;
    AI(v, AI_STATUS, L.row) = sP[L.tid].status;
    AD(v, AD_BLOODS, L.row) = sP[L.tid].bloods;
    if (info) {
      info[oa * ACS_INFO_DIM + 0] = cause; info[oa * ACS_INFO_DIM + 1] = sP[L.tid].status;
      info[oa * ACS_INFO_DIM + 2] = cs; info[oa * ACS_INFO_DIM + 3] = EI(v, EI_TURN_COUNTS, L.env);
    }
    if (share_obs) {                                 // get_state: hstack of every agent's obs (E/envs/env_base.py:183-189)
      const int D = cfg.obs_dim;
      double* so = share_obs + oa * (size_t)(A * D);
      const double* src = obs + (size_t)L.env * A * D;
      for (int i = 0; i < A * D; i++) so[i] = src[i];
    }
    if (L.lane == 0) {
      bool all = true;
      for (int j = 0; j < A; j++) all = all && sDone[L.gbase + j];
      if (env_done) env_done[L.env] = all;
      EI(v, EI_CURRENT_STEP, L.env) = cs;
      if (all && cfg.curriculum_window > 0) {
        // the episode's outcome joins the env's record (last `window` episodes); the stage rule runs here, i.e. before the
        // reset below, as task.reset applies it before reset_simulators_curriculum (E/tasks/scenario2_task.py:183-187)
        const int W = min(cfg.curriculum_window, 31);
        unsigned bits = ((unsigned)EI(v, EI_CUR_BITS, L.env) << 1 | (env_won ? 1u : 0u)) & ((1u << W) - 1u);
        int cnt = min(EI(v, EI_CUR_COUNT, L.env) + 1, W);
        const double rate = (double)__popc(bits) / (double)cnt;
        // rule 1 is the reference's: len(record) > window can never hold (the record is capped at `window`)
        const bool adv = rate >= cfg.curriculum_threshold && (cfg.curriculum_rule == 2 ? cnt >= W : (cfg.curriculum_rule == 1 ? cnt > W : false));
        if (adv) { EI(v, EI_STAGE, L.env) = EI(v, EI_STAGE, L.env) + 1; bits = 0; cnt = 0; }
        EI(v, EI_CUR_BITS, L.env) = (int)bits; EI(v, EI_CUR_COUNT, L.env) = cnt;
      }
    }
  }
  __syncwarp();      // the stage written by lane 0 is read by the whole warp in the reset below
  if (fuse_reset) {
    bool all = L.valid;
    for (int j = 0; j < A; j++) all = all && sDone[L.gbase + j];
    // the reset synchronises the lanes of each env on their mask: the whole warp takes the call together
    if (__any_sync(0xffffffffu, all)) reset_copy_warp(v, cfg, L, all, obs, share_obs, tpl);     // fuse_reset implies tpl.full (acs_env_step)
  }
}

// ============================================================================================== reset
// Stage 1: per aircraft FDM reload (AircraftSimulator.reload, simulatior.py:152-190) for masked envs.
__global__ void __launch_bounds__(FDM_BLOCK) k_env_reset_fdm(const EnvView v, const __grid_constant__ AcsTaskConfig cfg, const int lg,
                                                            const uint8_t* __restrict__ env_mask) {
  __shared__ double sT[F16_NTAB];
  const Lane L = lane_setup(v, lg);
  const bool on = L.valid && (env_mask == nullptr || env_mask[L.env]);
  // the auto-reset launch follows every step; a block none of whose environments finished leaves before it touches the
  // tables or the bulk of the code (cold after an L2 flush)
  if (!__syncthreads_or(on)) return;
  stage_tables(sT);
  if (!on) return;
  const int N = v.rows;
  const GeoOrigin org = geo_origin(cfg.center[0], cfg.center[1], cfg.center[2]);
  IcParams c;
  const double* r = cfg.init_state[L.lane];
  c.lon_deg = r[0]; c.lat_geod_deg = r[1]; c.h_sl_ft = r[2]; c.psi_deg = r[3]; c.u = r[4]; c.v = r[5]; c.w = r[6];
  c.p = r[7]; c.q = r[8]; c.r = r[9]; c.phi_deg = r[10]; c.theta_deg = r[11];
  const int episode = EI(v, EI_EPISODE, L.env) + 1;    // the task stage stores it
  if (cfg.obs_kind == ACS_OBS_HEADING) {               // E/envs/singlecontrol_env.py:32-49
    const int64_t e = cfg.env_offset + L.env;
    c.psi_deg = 0.0 + 180.0 * env_u01(cfg.seed, e, RNG_RESET, episode, 0, 0);
    c.h_sl_ft = 14000.0 + 16000.0 * env_u01(cfg.seed, e, RNG_RESET, episode, 1, 0);
    c.u = 400.0 + 800.0 * env_u01(cfg.seed, e, RNG_RESET, episode, 2, 0);
    ED(v, ED_TGT_HEADING, L.env) = env_clip(c.psi_deg, 0.0, 360.0);
    ED(v, ED_TGT_ALT, L.env) = env_clip(c.h_sl_ft, -1400.0, 85000.0);
    ED(v, ED_TGT_VEL, L.env) = env_clip(c.u * 0.3048, -700.0, 700.0);
    ED(v, ED_CHECK_TIME, L.env) = 0.0;
    EI(v, EI_TURN_COUNTS, L.env) = 0;
  }
  AcCore a; Props p; FcsState s; Frame f;
  fdm_reset(a, p, s, f, sT, g_atmo, c, cfg.fcs_dt);
  store_state(v.fdm, N, L.row, a, p, s);
  AcOut o;
  fdm_outputs(a, f, o);
  store_out(v.out, N, L.row, o);
  PubAc me;
  double v_mps, w_mps, vc_mps;
  derive_aircraft(o, org, me, v_mps, w_mps, vc_mps);
  store_derived(v, L.row, me, v_mps, w_mps, vc_mps);
  AD(v, AD_BLOODS, L.row) = 100.0;
  AI(v, AI_STATUS, L.row) = ST_ALIVE;
}
// Stage 2: task.reset + reward-function resets + get_obs for masked envs.
// With a reset template (tpl.fdm != nullptr) stage 1 is folded in: every task but the heading task reloads each aircraft
// from fixed per-lane initial conditions, so sim.reload() always produces the same state.  It is computed once per
// handle by k_env_reset_fdm on a one-env arena (acs_env_create / acs_env_set_init_states) and copied here -- bit-identical
// to recomputing it, and the auto-reset that follows every step no longer launches the 2-frame FDM reload.
// The lanes of the environments being reset (`on`): used by k_env_reset_task and, fused, by k_env_post's auto-reset.
// sP is the block's PubAc exchange array; every lane of the warp calls this (warp-level syncs on the env's lane mask).
// The whole reset as a copy of the template (ResetTpl::full), by the whole warp: for every env of the warp that is being
// reset, the 32 lanes scatter the packed template words (coalesced reads, 32 independent words in flight per iteration,
// a dozen instructions of code -- the per-lane, per-arena unrolled copy this replaces spent its time fetching its own
// cold instructions).  Every lane of the warp must call this; `on` lanes belong to envs being reset.
ENV_DEV void reset_copy_warp(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const bool on, double* __restrict__ obs,
                             double* __restrict__ share_obs, const ResetTpl& tp) {
  const unsigned wl = threadIdx.x & 31;
  __syncwarp();   // the step logic's own stores to these words (terminal obs, status, bloods) are ordered before the scatter
  unsigned leaders = __ballot_sync(0xffffffffu, on && L.lane == 0);
  const int A = v.A, S = v.S, AD_ = A * cfg.obs_dim;
  while (leaders) {
    const int src = __ffs(leaders) - 1;
    leaders &= leaders - 1;
    const int env = __shfl_sync(0xffffffffu, L.env, src);
    const size_t row0 = (size_t)env * A;
    // the env's curriculum stage picks the value set (same words, other initial conditions)
    const double* v64 = tp.v64; const int* v32 = tp.v32; const double* tobs = tp.obs;
    if (tp.n_stages > 1) {
      const int stage = min(max(EI(v, EI_STAGE, env), 0), tp.n_stages - 1);
      if (stage > 0) {
        const char* sb = tp.stage_base + (size_t)stage * tp.stage_stride;
        tobs = (const double*)sb; v64 = (const double*)(sb + tp.stage_off64); v32 = (const int*)(sb + tp.stage_off32);
      }
    }
#pragma unroll 2
    for (int w = wl; w < tp.n64; w += 32) {
      const int d = __ldg(tp.d64 + w);
      const double x = __ldg(v64 + w);
      const int k = d >> 24, f = (d >> 12) & 0xfff, j = d & 0xfff;
      double* b = k == 0 ? v.fdm : (k == 1 ? v.out : (k == 2 ? v.ad : (k == 6 ? v.md : v.ed)));
      const size_t stride = k == 6 ? (size_t)v.rows * S : (k == 4 ? (size_t)v.B : (size_t)v.rows);
      const size_t i0 = k == 6 ? row0 * S + j : (k == 4 ? (size_t)env : row0 + j);
      b[f * stride + i0] = x;
    }
#pragma unroll 2
    for (int w = wl; w < tp.n32; w += 32) {
      const int d = __ldg(tp.d32 + w);
      const int x = __ldg(v32 + w);
      const int k = d >> 24, f = (d >> 12) & 0xfff, j = d & 0xfff;
      int* b = k == 3 ? v.ai : (k == 7 ? v.mi : v.ei);
      const size_t stride = k == 7 ? (size_t)v.rows * S : (k == 5 ? (size_t)v.B : (size_t)v.rows);
      const size_t i0 = k == 7 ? row0 * S + j : (k == 5 ? (size_t)env : row0 + j);
      b[f * stride + i0] = x;
    }
    if (wl == 0) EI(v, EI_EPISODE, env) = EI(v, EI_EPISODE, env) + 1;     // the one read-modify-write of reset()
    for (int w = wl; w < AD_; w += 32) {
      const double x = __ldg(tobs + w);
      obs[row0 * cfg.obs_dim + w] = x;
      if (share_obs) for (int a = 0; a < A; a++) share_obs[(row0 + a) * (size_t)AD_ + w] = x;
    }
  }
}

template <class S>
__device__ __noinline__ void reset_task_lanes(const EnvView& v, const AcsTaskConfig& cfg, const Lane& L, const bool on, PubAc* sP,
                                              double* __restrict__ obs, double* __restrict__ share_obs, const ResetTpl& tp) {
  if (tp.full) { reset_copy_warp(v, cfg, L, on, obs, share_obs, tp); return; }
  const EnvView& tpl = tp.t;
  const int A = v.A;
  PubAc me;
  me.status = ST_CRASH; me.bloods = 0; me.h = 0; me.u_mps = 0; me.f.n = me.f.e = me.f.u = me.f.vn = me.f.ve = me.f.vd = 0;
  if (on && tpl.fdm != nullptr) {
    const int row = L.row, N = v.rows, l = L.lane;
    // batches of independent loads first, then the stores: the two arenas may alias as far as the compiler knows, and a
    // load -> store -> load chain would pay one (cold) memory round trip per field
    const double* __restrict__ tf = tpl.fdm; const double* __restrict__ to = tpl.out; const double* __restrict__ ta = tpl.ad;
    constexpr int NS = FDM_N_CORE + F16_N_CARRIED, CH = 25;
    static_assert(NS % CH == 0 && FDM_N_OUT == CH, "template copy is written for 25-field batches");
#pragma unroll 1
    for (int k0 = 0; k0 < NS; k0 += CH) {
      double t[CH];
#pragma unroll
      for (int j = 0; j < CH; j++) t[j] = __ldg(tf + (k0 + j) * A + l);
#pragma unroll
      for (int j = 0; j < CH; j++) v.fdm[(size_t)(k0 + j) * N + row] = t[j];
    }
    {
      double t[CH];
#pragma unroll
      for (int j = 0; j < CH; j++) t[j] = __ldg(to + j * A + l);
#pragma unroll
      for (int j = 0; j < CH; j++) v.out[(size_t)j * N + row] = t[j];
    }
    // what store_derived and the tail of k_env_reset_fdm write
    const int adf[12] = {AD_POS_N, AD_POS_E, AD_POS_U, AD_VEL_N, AD_VEL_E, AD_VEL_D, AD_H_SL_M, AD_U_MPS, AD_V_MPS, AD_W_MPS, AD_VC_MPS, AD_BLOODS};
    double t[12];
#pragma unroll
    for (int k = 0; k < 12; k++) t[k] = __ldg(ta + adf[k] * A + l);
#pragma unroll
    for (int k = 0; k < 12; k++) AD(v, adf[k], row) = t[k];
    AI(v, AI_STATUS, row) = ST_ALIVE;
  }
  if (on) {
    const int row = L.row;
    const int nm = cfg.num_missiles[L.lane];
    AI(v, AI_DIE_FLAG, row) = 0;
    AI(v, AI_REM_MISSILES, row) = nm; AI(v, AI_REM_9M, row) = nm; AI(v, AI_REM_120B, row) = nm; AI(v, AI_REM_GUN, row) = nm;
    AI(v, AI_REM_CHAFF, row) = nm;
    AI(v, AI_LAST_SHOOT_TIME, row) = -cfg.min_attack_interval;
    AI(v, AI_LAST_SHOT_SLOT, row) = -1; AI(v, AI_N_LAUNCHED, row) = 0;
    AI(v, AI_LOCK_LO, row) = 0; AI(v, AI_LOCK_HI, row) = 0; AI(v, AI_LOCK_N, row) = 0;
    AI(v, AI_PRE_REMAINING, row) = nm; AI(v, AI_SHOOT, row) = 0; AI(v, AI_CH_STATE, row) = CH_NONE; AI(v, AI_CH_COUNT, row) = 0;
    AD(v, AD_HR_ROLL, row) = 0; AD(v, AD_HR_P, row) = 0; AD(v, AD_HR_Q, row) = 0;
    AD(v, AD_CH_N, row) = 0; AD(v, AD_CH_E, row) = 0; AD(v, AD_CH_U, row) = 0; AD(v, AD_CH_T, row) = 0;
    for (int ri = 0; ri < ACS_MAX_REWARDS; ri++) AD(v, AD_PRE_REWARD0 + ri, row) = 0.0;
    for (int s = 0; s < v.S; s++) {
      const int mid = row * v.S + s;
      MI(v, MI_STATUS, mid) = MS_INACTIVE; MI(v, MI_DETACHED, mid) = 0; MI(v, MI_KEYN, mid) = -1; MI(v, MI_TARGET, mid) = 0;
      MI(v, MI_KIND, mid) = 0; MI(v, MI_CONSEC, mid) = 0; MI(v, MI_ORDER, mid) = 0; MI(v, MI_BORN, mid) = 0;
    }
    if (L.lane == 0) {
      const int env = L.env;
      EI(v, EI_CURRENT_STEP, env) = 0; EI(v, EI_EPISODE, env) = EI(v, EI_EPISODE, env) + 1; EI(v, EI_SUBSTEP_COUNT, env) = 0;
      EI(v, EI_PMV_REF, env) = -1; EI(v, EI_CG_VALID, env) = 0; EI(v, EI_TT_VALID, env) = 0; EI(v, EI_WD_VALID, env) = 0;
      EI(v, EI_ORDER_SEQ, env) = 0; EI(v, EI_BORN_SEQ, env) = 0;
      if (cfg.obs_kind != ACS_OBS_HEADING) EI(v, EI_TURN_COUNTS, env) = 0;
    }
    load_pub(v, row, me);
  }
  sP[L.tid] = me;
  __syncwarp(L.gmask);
  const StepCtx c{v, cfg, L, sP, 0};
  // reward_function.reset (E/reward_functions/reward_function_base.py:20-32): potential rewards seed pre_rewards
  AoTaR geo[ACS_MAX_AGENTS];
  const int n_geo = on ? enemy_geometry(c, L.lane, geo) : 0;
  for (int ri = 0; ri < cfg.n_rewards; ri++) {
    if (!cfg.rewards[ri].potential) continue;
    for (int a = 0; a < A; a++) {
      if (on && L.lane == a) {
        const double r0 = reward_one<S>(c, ri, a, geo, n_geo);
        AD(v, AD_PRE_REWARD0 + ri, L.row) = r0;
      }
      __syncwarp(L.gmask);
    }
  }
  if (on) write_obs<S>(c, L.lane, obs + ((size_t)L.env * A + L.lane) * cfg.obs_dim);
  __syncwarp(L.gmask);
  if (on && share_obs) {
    const int D = cfg.obs_dim;
    double* so = share_obs + ((size_t)L.env * A + L.lane) * (size_t)(A * D);
    const double* src = obs + (size_t)L.env * A * D;
    for (int i = 0; i < A * D; i++) so[i] = src[i];
  }
}

__global__ void __launch_bounds__(128) k_env_reset_task(const __grid_constant__ EnvView v, const __grid_constant__ AcsTaskConfig cfg, const int lg,
                                                        const uint8_t* __restrict__ env_mask, double* __restrict__ obs,
                                                        double* __restrict__ share_obs, const __grid_constant__ ResetTpl tpl) {
  __shared__ PubAc sP[128];
  const Lane L = lane_setup(v, lg);
  const bool on = L.valid && (env_mask == nullptr || env_mask[L.env]);
  if (!__syncthreads_or(on)) return;
  if (tpl.full) tpl_prefetch(tpl);
  reset_task_lanes<PostGeneric>(v, cfg, L, on, sP, obs, share_obs, tpl);
}

// ---------------------------------------------------------------------------------------------- k_env_post instantiations
// (observation packer, launch rule, reward classes) of the task families the shipped yamls resolve to (tasks.py);
// post_kernel_for() returns the first entry that covers a config, the generic kernel otherwise.
#define RM(x) (1u << (x))
constexpr unsigned RM_BASE = RM(ACS_R_ALTITUDE) | RM(ACS_R_POSTURE) | RM(ACS_R_EVENT);
constexpr unsigned RM_SCENARIO = RM_BASE | RM(ACS_R_COMBAT_GEOMETRY) | RM(ACS_R_GUN_BEHIT) | RM(ACS_R_GUN_TARGETTAIL) | RM(ACS_R_GUN_WEZDOT) |
                                 RM(ACS_R_GUN_WEZ) | RM(ACS_R_RELATIVE_ALTITUDE) | RM(ACS_R_MISSILE_POSTURE) | RM(ACS_R_SHOOT_PENALTY);
constexpr unsigned RM_GUNS = RM_BASE | RM(ACS_R_COMBAT_GEOMETRY) | RM(ACS_R_GUN_BEHIT) | RM(ACS_R_GUN_TARGETTAIL) | RM(ACS_R_GUN_WEZDOT) |
                             RM(ACS_R_GUN_WEZ) | RM(ACS_R_RELATIVE_ALTITUDE);
#define ACS_POST_FAMILIES(X) \
  X(ACS_OBS_HEADING, ACS_L_NONE, RM(ACS_R_HEADING) | RM(ACS_R_ALTITUDE))            /* heading, approach */ \
  X(ACS_OBS_1V1, ACS_L_NONE, RM_BASE)                                               /* 1v1 NoWeapon */ \
  X(ACS_OBS_1V1_MISSILE, ACS_L_RL_SINGLE, RM_BASE | RM(ACS_R_SHOOT_PENALTY))        /* 1v1 ShootMissile */ \
  X(ACS_OBS_1V1_MISSILE, ACS_L_RULE_LOCK, RM_BASE | RM(ACS_R_MISSILE_POSTURE))      /* 1v1 DodgeMissile */ \
  X(ACS_OBS_MULTI, ACS_L_NONE, RM_BASE)                                             /* 2v2 NoWeapon */ \
  X(ACS_OBS_MULTI_MISSILE, ACS_L_RL_NEAREST, RM_BASE | RM(ACS_R_MISSILE_POSTURE))   /* 2v2 ShootMissile (shoot nearest) */ \
  X(ACS_OBS_NV_MISSILE, ACS_L_SCENARIO, RM_SCENARIO)                                /* scenario2, scenario3 */ \
  X(ACS_OBS_1V1_MISSILE, ACS_L_SCENARIO, RM_SCENARIO)                               /* scenario1 */ \
  X(ACS_OBS_NVN, ACS_L_SCENARIO, RM_SCENARIO)                                       /* scenario2/3 _nvn, _rwr */ \
  X(ACS_OBS_1V1, ACS_L_AUTO_GUN, RM_GUNS)                                           /* wvr, maneuver_curriculum */
typedef void (*PostKernel)(const EnvView, const AcsTaskConfig, const int, double*, double*, double*, uint8_t*, int32_t*, uint8_t*, const int,
                           const ResetTpl, const int);
static PostKernel post_kernel_for(const AcsTaskConfig& cfg, int* family) {
  unsigned need = 0;
  for (int i = 0; i < cfg.n_rewards; i++) need |= 1u << cfg.rewards[i].kind;
  int k = 0;
#define X(OBS, LAUNCH, MASK) \
  if (cfg.obs_kind == (OBS) && cfg.launch_kind == (LAUNCH) && (need & ~(unsigned)(MASK)) == 0) { *family = k; return k_env_post<PostSpec<OBS, LAUNCH, (MASK)>>; } \
  k++;
  ACS_POST_FAMILIES(X)
#undef X
  *family = -1;
  return k_env_post<PostGeneric>;
}
#undef RM
