// TEST INFRASTRUCTURE: the PRODUCT's flight-dynamics source -- csrc/fdm_core.cuh, csrc/gen/f16_gen.cuh, csrc/fmath.cuh, the
// very files libacs.so is built from -- compiled for the HOST, so that `pytest -m "not gpu"` checks the device arithmetic
// (stage functions, generated FCS / aero code, table brackets, the guard-free division / root / elementary-function
// sequences) against the CPU oracle without a GPU.  Nothing here ships: the simulator has no CPU path (acs_create fails
// without a CUDA device); this file is compiled by tests/test_fdm_host.py into a temporary directory.
//
// What differs from the device build: the MUFU seeds of fmath.cuh are emulated (deliberately poorly, 2^-12), g++ does not
// contract a*b+c into FMAs (-ffp-contract=off), and libm stands in for the few libdevice calls left.  Results therefore
// agree with the device build to rounding, not bit for bit -- the bound the test states is the parity bound of the GPU
// tests (1e-9 per step), which is what a regression in this source would break.
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static
#include "fdm_core.cuh"

static AtmoConst g_atmo;
static bool g_init = false;
static void init_once() {
  if (g_init) return;
  host_atmo(g_atmo);
  for (int k = 0; k < F16_NTAB; k++) g_f16_kc[k] = F16_TAB_HOST[k];     // what acs_create copies to constant memory
  g_init = true;
}

#define NAME_ONLY(name, expr) name,
static const char* const STATE_NAMES[] = {FDM_CORE_FIELDS(NAME_ONLY) F16_CARRIED_FIELDS(NAME_ONLY)};
static const char* const OUTPUT_NAMES[] = {FDM_OUT_FIELDS(NAME_ONLY)};
static constexpr int N_STATE = FDM_N_CORE + F16_N_CARRIED;

struct HostFdm {
  AcCore a; Props p; FcsState s; Frame f; FrameKeep keep;
  double sim_dt, fcs_dt;
};

static void store(const HostFdm& h, double* st, double* out) {
  const AcCore& a = h.a; const Props& p = h.p; const FcsState& s = h.s;
  int k = 0;
#define ST(name, expr) st[k++] = expr;
  FDM_CORE_FIELDS(ST)
  F16_CARRIED_FIELDS(ST)
#undef ST
  AcOut o;
  fdm_outputs(h.a, h.f, o);
  k = 0;
#define ST(name, expr) out[k++] = expr;
  FDM_OUT_FIELDS(ST)
#undef ST
}

extern "C" {
int fh_n_state(void) { return N_STATE; }
int fh_n_out(void) { return FDM_N_OUT; }
const char* fh_state_name(int i) { return STATE_NAMES[i]; }
const char* fh_out_name(int i) { return OUTPUT_NAMES[i]; }

void* fh_create(double sim_dt, double fcs_dt) {
  init_once();
  HostFdm* h = new HostFdm();
  std::memset(h, 0, sizeof(HostFdm));
  h->sim_dt = sim_dt; h->fcs_dt = fcs_dt;
  return h;
}
void fh_destroy(void* h) { delete (HostFdm*)h; }

// what k_fdm_reset does for one aircraft (csrc/acs.cu)
void fh_reset(void* hv, const double* r) {
  HostFdm* h = (HostFdm*)hv;
  IcParams c;
  c.lon_deg = r[0]; c.lat_geod_deg = r[1]; c.h_sl_ft = r[2]; c.psi_deg = r[3]; c.u = r[4]; c.v = r[5]; c.w = r[6];
  c.p = r[7]; c.q = r[8]; c.r = r[9]; c.phi_deg = r[10]; c.theta_deg = r[11];
  fdm_reset(h->a, h->p, h->s, h->f, F16_TAB_HOST, g_atmo, c, h->fcs_dt);
  h->keep.pilot_nx = h->keep.vcas = h->keep.beta = h->keep.thrust = 0.0;
}
// what k_set_controls does: the catalog clip of the four commands
void fh_set_controls(void* hv, const double* u) {
  HostFdm* h = (HostFdm*)hv;
  const double lo[4] = {-1.0, -1.0, -1.0, 0.0}, hi[4] = {1.0, 1.0, 1.0, 0.9};
  double v[4];
  for (int k = 0; k < 4; k++) v[k] = u[k] < lo[k] ? lo[k] : (u[k] > hi[k] ? hi[k] : u[k]);
  h->p.fcs_aileron_cmd_norm = v[0]; h->p.fcs_elevator_cmd_norm = v[1]; h->p.fcs_rudder_cmd_norm = v[2]; h->p.fcs_throttle_cmd_norm = v[3];
}
// lean = 0: fdm_frame, the frame of k_fdm_run and of the multi-warp kernels' stage functions;
// lean = 1: fdm_frame_lean + fdm_refresh, the frame of the throughput kernel k_env_substeps
void fh_run(void* hv, int n_frames, int lean) {
  HostFdm* h = (HostFdm*)hv;
  if (!lean) {
    for (int k = 0; k < n_frames; k++) fdm_frame(h->a, h->p, h->s, h->f, F16_TAB_HOST, g_atmo, h->sim_dt, h->fcs_dt, false);
  } else {
    for (int k = 0; k < n_frames; k++)
      fdm_frame_lean(h->a, h->p, h->s, h->keep, F16_TAB_HOST, g_atmo, h->sim_dt, h->fcs_dt, [](const Frame&) {});
    fdm_refresh(h->a, h->p, h->keep, h->f);
  }
}
void fh_get(void* hv, double* state, double* out) { store(*(HostFdm*)hv, state, out); }
// what load_state does (csrc/acs.cu): the state arena's words, in STATE_NAMES order, into the registers of one aircraft
void fh_set_state(void* hv, const double* st) {
  HostFdm* h = (HostFdm*)hv;
  AcCore& a = h->a; Props& p = h->p; FcsState& s = h->s;
  f16_props_init(p, s);
  int k = 0;
#define LD(name, expr) expr = st[k++];
  FDM_CORE_FIELDS(LD)
  F16_CARRIED_FIELDS(LD)
#undef LD
  f16_props_derive(p);
}
}
