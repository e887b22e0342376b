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

#ifdef __cplusplus
}
#endif
#endif /* ACS_H_ */
