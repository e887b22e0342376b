/* acs.h -- C ABI of the B200-native batched air-combat simulator ("acs").
 *
 * Drop-in boundary for the env-step hot path of junghoseong/aircombat-selfplay.  Every entry point
 * below replaces a call (or a Python loop of calls) the reference makes through the pip `jsbsim`
 * Cython binding; the replaced interface is cited as reference file:line ("R/" = the reference repo
 * root, "E/" = R/envs/JSBSim/).
 *
 * Conventions: plain pointers and sizes only; every `*_dev` pointer is DEVICE memory owned by the
 * caller (PyTorch tensors in the shipped host layer); the library owns only its structure-of-arrays
 * state arena.  Calls are stream-ordered on `stream` (a cudaStream_t passed as void*) and never
 * synchronise.  Return 0 on success, non-zero on error; acs_last_error() gives the message.  One
 * handle per GPU; a handle is not thread-safe.  There is no CPU fallback: acs_create fails if no
 * CUDA device is usable.
 */
#ifndef ACS_H_
#define ACS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct AcsHandle AcsHandle;

/* Number of aircraft state rows etc. are fixed at creation. */
typedef struct AcsConfig {
  int32_t n_envs;        /* B: environments owned by this handle (this GPU's contiguous env slice)        */
  int32_t n_agents;      /* A: aircraft per environment (ego team first, then enemy team)                 */
  double  sim_dt;        /* 1 / sim_freq (E/envs/env_base.py:26)                                          */
  double  fcs_dt;        /* dt latched by the FCS components at load_model time; the reference calls      */
                         /* set_dt() after load_model, so this is 1/120 (E/core/simulatior.py:167-169)    */
} AcsConfig;

const char* acs_last_error(void);
int acs_version(void);

/* replaces: jsbsim.FGFDMExec(root); load_model('f16') per aircraft (E/core/simulatior.py:165-168) */
int acs_create(const AcsConfig* cfg, int device, AcsHandle** out);
int acs_destroy(AcsHandle* h);

/* Layout of the SoA arenas: field-major, `n_rows = n_envs*n_agents` doubles per field. */
int acs_n_rows(const AcsHandle* h);
int acs_n_state_fields(void);
const char* acs_state_field_name(int i);
int acs_n_output_fields(void);
const char* acs_output_field_name(int i);

/* replaces: AircraftSimulator.reload(): IC property writes, run_ic(), engine.init_running(),
 * propulsion.get_steady_state() (E/core/simulatior.py:152-190).
 * ic_dev: [n_rows][12] doubles = lon_deg, lat_geod_deg, h_sl_ft, psi_deg, u, v, w (fps), p, q, r (rad/s),
 * phi_deg, theta_deg.  mask_dev: [n_rows] bytes, rows with 0 are left untouched; NULL = all rows. */
int acs_fdm_reset(AcsHandle* h, const uint8_t* mask_dev, const double* ic_dev, void* stream);

/* replaces: set_property_values(action_var, a) on fcs/{aileron,elevator,rudder,throttle}-cmd-norm
 * (E/envs/env_base.py:135-137).  controls_dev: [n_rows][4], already clipped by the caller or not --
 * the catalog clip ranges [-1,1]x3, [0,0.9] (E/core/catalog.py:192-197) are applied here. */
int acs_fdm_set_controls(AcsHandle* h, const double* controls_dev, void* stream);

/* replaces: `for _ in range(n): jsbsim_exec.run()` for every aircraft (E/core/simulatior.py:210-229).
 * alive_dev: [n_rows] bytes (NULL = all alive); rows with 0 do not integrate (dead aircraft are frozen). */
int acs_fdm_run(AcsHandle* h, int n_frames, const uint8_t* alive_dev, void* stream);

/* replaces: get_property_value reads / parity-test state injection.  Whole-arena device copies. */
int acs_get_state(const AcsHandle* h, double* dst_dev, void* stream);    /* [n_state_fields][n_rows]  */
int acs_set_state(AcsHandle* h, const double* src_dev, void* stream);
int acs_get_outputs(const AcsHandle* h, double* dst_dev, void* stream);  /* [n_output_fields][n_rows] */


/* ======================================================================================================
 * Environment layer: the gym-style reset/step of SingleControlEnv / SingleCombatEnv / MultipleCombatEnv
 * (E/envs/env_base.py:98-173, E/envs/multiplecombat_env.py:66-182) for n_envs environments at once.
 * AcsTaskConfig is what a reference Task class + yaml config boils down to once Python class composition
 * is resolved (E/tasks/*.py, E/reward_functions/*.py, E/termination_conditions/*.py); the host layer
 * (aircombat_selfplay_b200/taskspec.py) builds it from the same yaml files the reference parses.
 * ====================================================================================================== */
#define ACS_MAX_AGENTS 8
#define ACS_MAX_REWARDS 12
#define ACS_MAX_TERMS 6
#define ACS_INFO_DIM 4   /* per agent: done cause (-1 none, else ACS_T_*), status (0 alive,1 crash,2 shotdown), spare, spare */

enum { ACS_OBS_HEADING = 0, ACS_OBS_1V1 = 1, ACS_OBS_1V1_MISSILE = 2, ACS_OBS_NV_MISSILE = 3, ACS_OBS_MULTI = 4,
       ACS_OBS_MULTI_MISSILE = 5, ACS_OBS_NVN = 6, ACS_OBS_1V1_RWR = 7 };
enum { ACS_ACT_HEADING = 0, ACS_ACT_COMBAT = 1 };
enum { ACS_R_ALTITUDE = 0, ACS_R_POSTURE, ACS_R_EVENT, ACS_R_MISSILE_POSTURE, ACS_R_SHOOT_PENALTY, ACS_R_HEADING,
       ACS_R_RELATIVE_ALTITUDE, ACS_R_COMBAT_GEOMETRY, ACS_R_GUN_BEHIT, ACS_R_GUN_TARGETTAIL, ACS_R_GUN_WEZ, ACS_R_GUN_WEZDOT };
enum { ACS_T_UNREACH_HEADING = 0, ACS_T_EXTREME_STATE, ACS_T_OVERLOAD, ACS_T_LOW_ALTITUDE, ACS_T_TIMEOUT, ACS_T_SAFE_RETURN };
enum { ACS_L_NONE = 0, ACS_L_RULE_LOCK, ACS_L_RL_SINGLE, ACS_L_RL_NEAREST, ACS_L_SCENARIO, ACS_L_AUTO_GUN };
enum { ACS_G_NONE = 0, ACS_G_DIE_FLAG, ACS_G_ALIVE };

typedef struct AcsRewardSpec {
  int32_t kind;       /* ACS_R_* */
  int32_t potential;  /* BaseRewardFunction.is_potential (E/reward_functions/reward_function_base.py:15) */
  double scale;       /* reward_scale */
  double p0, p1, p2;  /* Altitude: safe_altitude, danger_altitude, Kv | Posture: orientation version, range version, target_dist | RelativeAltitude: KH */
} AcsRewardSpec;

typedef struct AcsTaskConfig {
  int32_t n_envs, n_ego, n_enm;
  int32_t substeps;                 /* agent_interaction_steps (E/envs/env_base.py:27) */
  int32_t max_steps;
  double sim_dt, fcs_dt;
  double altitude_limit, acc_limit[3], center[3];   /* center = battle_field_center (lon, lat, alt) */
  int32_t obs_kind, obs_dim, act_kind, shoot_dim;   /* action row = 4 low-level ints + shoot_dim shoot bits */
  int32_t n_rewards;
  AcsRewardSpec rewards[ACS_MAX_REWARDS];
  int32_t n_terms;
  int32_t terms[ACS_MAX_TERMS];     /* evaluated in this order, first done short-circuits (E/tasks/task_base.py:90-112) */
  int32_t dones_before_rewards;     /* BaseEnv.step order (env_base.py:159-171) vs MultipleCombatEnv.step (:163-180) */
  int32_t team_mean, share_obs, reward_gate, launch_kind, use_artillery, use_baseline;
  double max_attack_angle, max_attack_distance;
  int32_t min_attack_interval, lock_len;
  int32_t num_missiles[ACS_MAX_AGENTS];
  int32_t n_missile_slots;          /* per aircraft */
  double init_state[ACS_MAX_AGENTS][12];  /* per aircraft, layout of acs_fdm_reset's ic rows */
  double heading_increments[3];     /* UnreachHeading increment sizes (heading deg, altitude ft, velocity m/s) */
  double check_interval;
  uint64_t seed;                    /* keyed counter RNG seed (replaces np.random / gymnasium np_random draws) */
  int32_t env_offset;               /* global index of this handle's first env (multi-GPU sharding keeps RNG streams per env) */
  int32_t reserved;
  /* Curriculum tasks (E/tasks/scenario2_task.py:172-223, scenario1_task.py:164-194, WVR_task.py:38-64): every env keeps a
   * record of its last `curriculum_window` episode outcomes (win = an ego aircraft ended alive through SafeReturn) and a
   * stage; when an episode ends the record is updated and the stage rule is applied BEFORE the auto-reset, so the reset
   * that follows already starts from the new stage (acs_env_set_stage_init_states).  curriculum_rule: 0 = stages only
   * change when the caller writes them (acs_env_set_arena on "stage"); 1 = the reference's rule verbatim -- advance when
   * rate >= threshold and len(record) > window, which can never hold because the record is capped at `window` entries
   * (the reference's stage therefore never advances by itself); 2 = advance when the record is full (len == window). */
  int32_t curriculum_rule, curriculum_window;
  double curriculum_threshold;
} AcsTaskConfig;

typedef struct AcsEnv AcsEnv;

/* replaces: constructing n_envs Env objects (BaseEnv.__init__ -> load_task, load_simulator; E/envs/env_base.py:24-87) */
int acs_env_create(const AcsTaskConfig* cfg, int device, AcsEnv** out);
int acs_env_destroy(AcsEnv* e);
/* replaces: env.reload-time changes of init_state (reset_simulators / curriculum resets, E/envs/singlecombat_env.py:45-122) */
int acs_env_set_init_states(AcsEnv* e, const double* init_host /* [n_agents][12], HOST memory */);

/* replaces: reset_simulators_curriculum(angle) per env process (E/envs/singlecombat_env.py:87-122, multiplecombat_env.py:185-248).
 * Registers the initial conditions of curriculum stage `stage` (0 <= stage < 256; stage 0 = the handle's own init_state /
 * acs_env_set_init_states).  Every env resets from the stage in its "stage" field (arena 5, per-env ints), which the
 * curriculum rule above advances on the device and acs_env_set_arena can write.  Needs the whole-reset template (every
 * task but the heading task).  init_host [n_agents][12], HOST memory. */
int acs_env_set_stage_init_states(AcsEnv* e, int stage, const double* init_host);

/* replaces: env.seed(seed) (E/envs/env_base.py:251-266): re-keys the counter RNG and restarts the per-env episode counters, so
 * seed(s); reset(); ... replays the same draws (the reference's determinism test, R/tests/test_jsbsim.py:55-64). */
int acs_env_set_seed(AcsEnv* e, uint64_t seed, void* stream);

/* replaces: BaseEnv.reset() for every env with env_mask_dev[env] != 0 (NULL = all): sim.reload() for every aircraft,
 * task.reset (reward-function resets included), get_obs (E/envs/env_base.py:98-113).
 * obs_dev [n_envs][n_agents][obs_dim] doubles; share_obs_dev [n_envs][n_agents][n_agents*obs_dim] or NULL.  Rows of
 * unmasked envs are left untouched. */
int acs_env_reset(AcsEnv* e, const uint8_t* env_mask_dev, double* obs_dev, double* share_obs_dev, void* stream);

/* replaces: BaseEnv.step(action) / MultipleCombatEnv.step(action) for all envs, plus -- when auto_reset != 0 -- the
 * VecEnv worker's `if all(done): obs = env.reset()` (R/envs/env_wrappers.py:191-204,380-393).
 * actions_dev [n_envs][n_agents][4+shoot_dim] int32 LOW-LEVEL discrete actions (hierarchical tasks run their GRU
 * controller above this boundary); rewards_dev [n_envs][n_agents] doubles; dones_dev [n_envs][n_agents] bytes;
 * info_dev [n_envs][n_agents][ACS_INFO_DIM] int32 or NULL; env_done_dev [n_envs] bytes (all agents done) or NULL. */
int acs_env_step(AcsEnv* e, const int32_t* actions_dev, double* obs_dev, double* share_obs_dev, double* rewards_dev,
                 uint8_t* dones_dev, int32_t* info_dev, uint8_t* env_done_dev, int auto_reset, void* stream);

/* Introspection for parity tests: named SoA arenas.  which: 0 FDM state [n_state_fields][rows] doubles, 1 FDM outputs,
 * 2 per-aircraft doubles, 3 per-aircraft ints, 4 per-env doubles, 5 per-env ints, 6 missile doubles [f][rows*slots],
 * 7 missile ints.  acs_env_arena_size returns the number of fields of an arena (elements per field via its own
 * second output).  Copies are device-to-device on `stream`. */
int acs_env_arena_info(const AcsEnv* e, int which, int* n_fields, int* n_per_field, int* is_int);
const char* acs_env_arena_field_name(int which, int field);
int acs_env_get_arena(const AcsEnv* e, int which, void* dst_dev, void* stream);
/* zero-copy: the device address of an arena (layout as acs_env_arena_info); valid for the life of the handle.  The host
 * layer reads aircraft state through it for the rule-based opponents (E/model/baseline.py) without a copy per step. */
int acs_env_arena_ptr(const AcsEnv* e, int which, void** dev_ptr);
int acs_env_set_arena(AcsEnv* e, int which, const void* src_dev, void* stream);
/* the FDM batch inside an env handle (for acs_fdm_* / acs_get_state on the same aircraft rows) */
AcsHandle* acs_env_fdm(AcsEnv* e);

/* Tuning knob (no reference counterpart).  "frame_split": which substep kernel acs_env_step launches -- 0 = one thread per
 * aircraft (throughput kernel), 1 / 2 / 3 = the two- / three- / four-warp frame (that many threads in different warps
 * per aircraft running the stages of one FDM frame concurrently; lower latency while the batch leaves SM
 * sub-partitions idle; 3 is opt-in only), -1 = choose by batch size
 * (default; the environment variable ACS_FRAME_SPLIT overrides the default at acs_env_create).  Both kernels evaluate
 * the same expressions.  Returns non-zero for an unknown option or value. */
int acs_env_set_option(AcsEnv* e, const char* name, int value);
/* reads an option back; "frame_split_effective" = 1 if the next acs_env_step launches the two-warp frame;
 * "launches_per_step" = kernels one auto-resetting acs_env_step launches (2 when the reset is fused into k_env_post) */
int acs_env_get_option(const AcsEnv* e, const char* name, int* value);

/* Measurement (no reference counterpart): with timing on, acs_env_step brackets each of its kernels with CUDA events on the
 * caller's stream (event-record nodes when the stream is being captured into a CUDA graph, so the intervals contain no
 * launch gaps); acs_env_get_timing synchronises those events and returns the accumulated milliseconds per kernel --
 * ms[0] k_env_substeps*, ms[1] k_env_post, ms[2] the two reset kernels, ms[3] k_env_missiles -- and the number of steps covered.  reset != 0
 * forgets the recorded steps; reset == 0 keeps them (a replayed graph re-records the same events every launch). */
int acs_env_set_timing(AcsEnv* e, int on);
int acs_env_get_timing(AcsEnv* e, double ms[4], int* n_steps, int reset);

/* Test helper (no reference counterpart): evaluates ONE function of csrc/fmath.cuh -- the guard-free fp64 division /
 * reciprocal / square-root sequences and the constant-memory polynomial kernels the FDM frame, the missile path and the
 * per-step logic use instead of `/`, sqrt() and libdevice -- over device arrays, so that the device build (real MUFU
 * seeds) can be compared with libm (tests/test_fmath_gpu.py).  a_dev, b_dev, out_dev, out2_dev: double[n] on the device;
 * b_dev may be NULL for one-operand functions, out2_dev may be NULL.  out2 = second result where there is one (SQRT:
 * fm_sqrt0; SINCOS / SINCOS_SMALL: cosine; SIN: fm_cos; TANH: fm_atanh; POW_RATIO: the second ISA exponent). */
enum {
  ACS_FMATH_DIV = 0, ACS_FMATH_RCP, ACS_FMATH_SQRT, ACS_FMATH_RSQRT, ACS_FMATH_SINCOS, ACS_FMATH_SIN, ACS_FMATH_EXP,
  ACS_FMATH_LOG, ACS_FMATH_ATAN2, ACS_FMATH_ACOS, ACS_FMATH_TANH, ACS_FMATH_POW_RATIO, ACS_FMATH_ANGLE_SC,
  ACS_FMATH_SINCOS_SMALL, ACS_FMATH_N_OPS
};
int acs_debug_fmath(int op, const double* a_dev, const double* b_dev, double* out_dev, double* out2_dev, int n, void* stream);

/* Measurement helper (no reference counterpart): runs a dependent-free fp64 FMA loop on every SM and returns the
 * achieved FLOP/s in *flops_out; used by bench.py to put an fp64-pipe roof beside the HBM roof. */
int acs_bench_fp64_peak(int device, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* ACS_H_ */
