// acs.cu -- kernels + C ABI (include/acs.h) of the batched air-combat simulator, sm_100a only.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <cmath>
#include "../../include/acs.h"
#include "fdm_core.cuh"
#include "env_core.cuh"

// ----------------------------------------------------------------------------- device constants
__constant__ AtmoConst g_atmo;
__device__ double g_f16_tab[F16_NTAB];

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// ----------------------------------------------------------------------------- state arena layout
#define NAME_ONLY(name, expr) name,
static const char* const STATE_NAMES[] = {FDM_CORE_FIELDS(NAME_ONLY) F16_CARRIED_FIELDS(NAME_ONLY)};
static const char* const OUTPUT_NAMES[] = {FDM_OUT_FIELDS(NAME_ONLY)};
static constexpr int N_STATE = FDM_N_CORE + F16_N_CARRIED;
static_assert(sizeof(STATE_NAMES) / sizeof(STATE_NAMES[0]) == N_STATE, "state field count");
static_assert(sizeof(OUTPUT_NAMES) / sizeof(OUTPUT_NAMES[0]) == FDM_N_OUT, "output field count");

struct AcsHandle {
  AcsConfig cfg;
  int device;
  int n_rows;
  double* state;   // [N_STATE][n_rows]
  double* out;     // [FDM_N_OUT][n_rows]
};

__device__ __forceinline__ void load_state(const double* __restrict__ st, int N, int i, AcCore& a, Props& p, FcsState& s) {
  int k = 0;
#define LD(name, expr) expr = st[(size_t)(k++) * N + i];
  FDM_CORE_FIELDS(LD)
  F16_CARRIED_FIELDS(LD)
#undef LD
  f16_props_derive(p);
}
__device__ __forceinline__ void store_state(double* __restrict__ st, int N, int i, const AcCore& a, const Props& p, const FcsState& s) {
  int k = 0;
#define ST(name, expr) st[(size_t)(k++) * N + i] = expr;
  FDM_CORE_FIELDS(ST)
  F16_CARRIED_FIELDS(ST)
#undef ST
}
// two-warp frame: role A (ROLE_B = false) owns the core fields and the carried properties the core models publish, role B
// the flight-control side (commands, FCS outputs, PID states); the ownership predicate folds at compile time
template <bool ROLE_B>
__device__ __forceinline__ void store_state_role(double* __restrict__ st, int N, int i, const AcCore& a, const Props& p, const FcsState& s) {
  int k = 0;
#define ST(name, expr) { if (!ROLE_B) st[(size_t)k * N + i] = expr; k++; }
  FDM_CORE_FIELDS(ST)
#undef ST
#define ST(name, expr) { constexpr bool b_ = F16_CARRIED_ROLE_B(name); if (b_ == ROLE_B) st[(size_t)k * N + i] = expr; k++; }
  F16_CARRIED_FIELDS(ST)
#undef ST
}
__device__ __forceinline__ void store_out(double* __restrict__ out, int N, int i, const AcOut& o) {
  int k = 0;
#define ST(name, expr) out[(size_t)(k++) * N + i] = expr;
  FDM_OUT_FIELDS(ST)
#undef ST
}

__device__ __forceinline__ void stage_tables(double* sT) {
  for (int k = threadIdx.x; k < F16_NTAB; k += blockDim.x) sT[k] = g_f16_tab[k];
  __syncthreads();
}

// ----------------------------------------------------------------------------- kernels
#ifndef ACS_FDM_BLOCK
#define ACS_FDM_BLOCK 128
#endif
#ifndef ACS_FDM_MIN_BLOCKS
#define ACS_FDM_MIN_BLOCKS 1
#endif
constexpr int FDM_BLOCK = ACS_FDM_BLOCK;
// two-warp frame: aircraft per block (each gets one thread of role A and one of role B), and the axes role A builds
#ifndef ACS_SPLIT_SLOTS
#define ACS_SPLIT_SLOTS 64
#endif
#ifndef ACS_SPLIT_AXES_A
#define ACS_SPLIT_AXES_A 0   // measured best on B200: role A keeps the engine, role B builds all six axes (one set of brackets)
#endif
#ifndef ACS_SPLIT_AUTO_LANES_PER_SM
#define ACS_SPLIT_AUTO_LANES_PER_SM 128
#endif
constexpr int SPLIT_SLOTS = ACS_SPLIT_SLOTS;
constexpr int SPLIT_AXES_A = ACS_SPLIT_AXES_A, SPLIT_AXES_B = 63 & ~SPLIT_AXES_A;
static_assert(SPLIT_SLOTS % 32 == 0 && SPLIT_SLOTS >= 32 && SPLIT_SLOTS <= 128, "pairs of whole warps, at most 4 pairs per block");

__global__ void __launch_bounds__(FDM_BLOCK, ACS_FDM_MIN_BLOCKS) k_fdm_run(double* __restrict__ state, double* __restrict__ out,
                                                      const uint8_t* __restrict__ alive, int N, int n_frames, double dt,
                                                      double fcs_dt) {
  __shared__ double sT[F16_NTAB];
  stage_tables(sT);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (alive != nullptr && !alive[i]) return;
  AcCore a; Props p; FcsState s; Frame f;
  f16_props_init(p, s);
  load_state(state, N, i, a, p, s);
  for (int k = 0; k < n_frames; k++) fdm_frame(a, p, s, f, sT, g_atmo, dt, fcs_dt, false);
  store_state(state, N, i, a, p, s);
  AcOut o;
  fdm_outputs(a, f, o);
  store_out(out, N, i, o);
}

__global__ void __launch_bounds__(FDM_BLOCK) k_fdm_reset(double* __restrict__ state, double* __restrict__ out,
                                                        const uint8_t* __restrict__ mask, const double* __restrict__ ic, int N,
                                                        double fcs_dt) {
  __shared__ double sT[F16_NTAB];
  stage_tables(sT);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (mask != nullptr && !mask[i]) return;
  IcParams c;
  const double* r = ic + (size_t)i * 12;
  c.lon_deg = r[0]; c.lat_geod_deg = r[1]; c.h_sl_ft = r[2]; c.psi_deg = r[3]; c.u = r[4]; c.v = r[5]; c.w = r[6];
  c.p = r[7]; c.q = r[8]; c.r = r[9]; c.phi_deg = r[10]; c.theta_deg = r[11];
  AcCore a; Props p; FcsState s; Frame f;
  fdm_reset(a, p, s, f, sT, g_atmo, c, fcs_dt);
  store_state(state, N, i, a, p, s);
  AcOut o;
  fdm_outputs(a, f, o);
  store_out(out, N, i, o);
}

__global__ void k_set_controls(double* __restrict__ state, const double* __restrict__ u, int N, int f_ail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  // catalog clip ranges (reference envs/JSBSim/core/catalog.py:192-197)
  const double lo[4] = {-1.0, -1.0, -1.0, 0.0}, hi[4] = {1.0, 1.0, 1.0, 0.9};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double v = u[(size_t)i * 4 + k];
    v = v < lo[k] ? lo[k] : (v > hi[k] ? hi[k] : v);
    state[(size_t)(f_ail + k) * N + i] = v;
  }
}

#include "env_kernels.cuh"

// ----------------------------------------------------------------------------- host side
// host_atmo(AtmoConst&): fdm_core.cuh (host side of the ISA layer constants)

static int find_state_field(const char* name) {
  for (int i = 0; i < N_STATE; i++) if (!std::strcmp(STATE_NAMES[i], name)) return i;
  return -1;
}

extern "C" {

const char* acs_last_error(void) { return g_err.c_str(); }
int acs_version(void) { return 1; }

int acs_create(const AcsConfig* cfg, int device, AcsHandle** out) {
  if (!cfg || !out) return fail("acs_create: null argument");
  if (cfg->n_envs <= 0 || cfg->n_agents <= 0) return fail("acs_create: n_envs and n_agents must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("acs_create: no CUDA device (there is no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  AcsHandle* h = new AcsHandle();
  h->cfg = *cfg; h->device = device; h->n_rows = cfg->n_envs * cfg->n_agents;
  h->state = nullptr; h->out = nullptr;
  CUDA_TRY(cudaMalloc(&h->state, sizeof(double) * (size_t)N_STATE * h->n_rows));
  CUDA_TRY(cudaMalloc(&h->out, sizeof(double) * (size_t)FDM_N_OUT * h->n_rows));
  CUDA_TRY(cudaMemset(h->state, 0, sizeof(double) * (size_t)N_STATE * h->n_rows));
  CUDA_TRY(cudaMemset(h->out, 0, sizeof(double) * (size_t)FDM_N_OUT * h->n_rows));
  AtmoConst ac;
  host_atmo(ac);
  CUDA_TRY(cudaMemcpyToSymbol(g_atmo, &ac, sizeof(ac)));
  CUDA_TRY(cudaMemcpyToSymbol(g_f16_tab, F16_TAB_HOST, sizeof(double) * F16_NTAB));
  CUDA_TRY(cudaMemcpyToSymbol(g_f16_kc, F16_TAB_HOST, sizeof(double) * F16_NTAB));
  *out = h;
  return 0;
}

int acs_destroy(AcsHandle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->state); cudaFree(h->out);
  delete h;
  return 0;
}

int acs_n_rows(const AcsHandle* h) { return h ? h->n_rows : 0; }
int acs_n_state_fields(void) { return N_STATE; }
const char* acs_state_field_name(int i) { return (i >= 0 && i < N_STATE) ? STATE_NAMES[i] : nullptr; }
int acs_n_output_fields(void) { return FDM_N_OUT; }
const char* acs_output_field_name(int i) { return (i >= 0 && i < FDM_N_OUT) ? OUTPUT_NAMES[i] : nullptr; }

int acs_fdm_reset(AcsHandle* h, const uint8_t* mask_dev, const double* ic_dev, void* stream) {
  if (!h || !ic_dev) return fail("acs_fdm_reset: null argument");
  const int N = h->n_rows;
  k_fdm_reset<<<(N + FDM_BLOCK - 1) / FDM_BLOCK, FDM_BLOCK, 0, (cudaStream_t)stream>>>(h->state, h->out, mask_dev, ic_dev, N, h->cfg.fcs_dt);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int acs_fdm_set_controls(AcsHandle* h, const double* controls_dev, void* stream) {
  if (!h || !controls_dev) return fail("acs_fdm_set_controls: null argument");
  const int N = h->n_rows;
  const int f_ail = find_state_field("fcs/aileron-cmd-norm");
  if (f_ail < 0 || find_state_field("fcs/throttle-cmd-norm") != f_ail + 3) return fail("acs_fdm_set_controls: control fields not contiguous");
  k_set_controls<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, controls_dev, N, f_ail);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int acs_fdm_run(AcsHandle* h, int n_frames, const uint8_t* alive_dev, void* stream) {
  if (!h) return fail("acs_fdm_run: null handle");
  if (n_frames <= 0) return fail("acs_fdm_run: n_frames must be positive");
  const int N = h->n_rows;
  k_fdm_run<<<(N + FDM_BLOCK - 1) / FDM_BLOCK, FDM_BLOCK, 0, (cudaStream_t)stream>>>(h->state, h->out, alive_dev, N, n_frames, h->cfg.sim_dt, h->cfg.fcs_dt);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int acs_get_state(const AcsHandle* h, double* dst_dev, void* stream) {
  if (!h || !dst_dev) return fail("acs_get_state: null argument");
  CUDA_TRY(cudaMemcpyAsync(dst_dev, h->state, sizeof(double) * (size_t)N_STATE * h->n_rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int acs_set_state(AcsHandle* h, const double* src_dev, void* stream) {
  if (!h || !src_dev) return fail("acs_set_state: null argument");
  CUDA_TRY(cudaMemcpyAsync(h->state, src_dev, sizeof(double) * (size_t)N_STATE * h->n_rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int acs_get_outputs(const AcsHandle* h, double* dst_dev, void* stream) {
  if (!h || !dst_dev) return fail("acs_get_outputs: null argument");
  CUDA_TRY(cudaMemcpyAsync(dst_dev, h->out, sizeof(double) * (size_t)FDM_N_OUT * h->n_rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"

// ============================================================================= env layer host side
struct AcsEnv {
  AcsTaskConfig cfg;
  AcsHandle* fdm;
  EnvView v;
  int G, lg;   // lanes per env (1, 2, 4 or 8) and its log2
  ResetTpl tpl;                  // reset template (tpl.t.fdm == nullptr: none, e.g. the heading task)
  char* tpl_block = nullptr;     // the one allocation behind the template
  char* stage_block = nullptr;   // values of curriculum stages 1 .. ACS_MAX_STAGES-1 (acs_env_set_stage_init_states), allocated on first use
  size_t n64max = 0, n32max = 0;
  double* tpl_obs = nullptr;     // [A][obs_dim]
  int tpl_mode = 2;              // 0 none, 1 FDM reload only, 2 whole reset (ACS_RESET_TEMPLATE)
  bool fused_reset = true;       // auto-reset inside k_env_post (needs the template)
  bool post_split = true;        // get_obs on its own warps in k_env_post where that is legal (ACS_POST_SPLIT)
  PostKernel post = nullptr;     // k_env_post instantiation of this task's family (post_kernel_for)
  int post_family = -1;          // its index in ACS_POST_FAMILIES, -1 = the generic kernel
  int frame_split = -1;          // substep kernel: 0 one thread per aircraft, 1 two-warp frame, -1 by batch size
  int split_max_threads = 0;     // auto: use the two-warp frame up to this many aircraft lanes
  int n_sms = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;   // 5 events per timed step: before substeps, after substeps, after missiles, after post, after reset
  size_t ev_used = 0;
};

static constexpr int ACS_MAX_STAGES = 256;
static int build_reset_template(AcsEnv* e, int stage = 0);
// which substep kernel acs_env_step launches: 0 one thread, 1 two warps, 2 three warps per aircraft
static int frame_split_effective(const AcsEnv* e) {
  if (e->frame_split >= 0) return e->frame_split;
  const int lanes = e->v.B * e->G;
  // three warps per aircraft while every block still gets an SM of its own (one wave), two warps up to 128 aircraft per
  // SM, one thread per aircraft beyond (measured crossovers on B200)
  if (lanes <= e->n_sms * S3) return 2;
  if (lanes <= e->split_max_threads) return 1;
  return 0;
}

static cudaEvent_t timing_event(AcsEnv* e, cudaStream_t st) {
  if (e->ev_used == e->ev.size()) { cudaEvent_t x; cudaEventCreate(&x); e->ev.push_back(x); }
  cudaEvent_t x = e->ev[e->ev_used++];
  cudaEventRecord(x, st);
  return x;
}

static int env_arena(const AcsEnv* e, int which, void** ptr, int* nf, int* per, int* is_int) {
  const EnvView& v = e->v;
  switch (which) {
    case 0: *ptr = v.fdm; *nf = N_STATE; *per = v.rows; *is_int = 0; return 0;
    case 1: *ptr = v.out; *nf = FDM_N_OUT; *per = v.rows; *is_int = 0; return 0;
    case 2: *ptr = v.ad; *nf = N_AD; *per = v.rows; *is_int = 0; return 0;
    case 3: *ptr = v.ai; *nf = N_AI; *per = v.rows; *is_int = 1; return 0;
    case 4: *ptr = v.ed; *nf = N_ED; *per = v.B; *is_int = 0; return 0;
    case 5: *ptr = v.ei; *nf = N_EI; *per = v.B; *is_int = 1; return 0;
    case 6: *ptr = v.md; *nf = N_MD; *per = v.rows * v.S; *is_int = 0; return 0;
    case 7: *ptr = v.mi; *nf = N_MI; *per = v.rows * v.S; *is_int = 1; return 0;
  }
  return fail("unknown arena id");
}

// one fmath.cuh function per launch over arrays (acs_debug_fmath: the DEVICE build of the guard-free sequences, with the real
// MUFU seeds, against libm in tests/test_fmath_gpu.py)
__global__ void k_fmath_probe(const int op, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                              double* __restrict__ out2, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b ? b[i] : 0.0;
  double r = 0.0, r2 = 0.0;
  switch (op) {
    case ACS_FMATH_DIV: r = fm_div(x, y); break;
    case ACS_FMATH_RCP: r = fm_rcp(x); break;
    case ACS_FMATH_SQRT: r = fm_sqrt(x); r2 = fm_sqrt0(x); break;
    case ACS_FMATH_RSQRT: r = fm_rsqrt(x); break;
    case ACS_FMATH_SINCOS: fm_sincos(x, &r, &r2); break;
    case ACS_FMATH_SIN: r = fm_sin(x); r2 = fm_cos(x); break;
    case ACS_FMATH_EXP: r = fm_exp(x); break;
    case ACS_FMATH_LOG: r = fm_log(x); break;
    case ACS_FMATH_ATAN2: r = fm_atan2(x, y); break;
    case ACS_FMATH_ACOS: r = fm_acos(x); break;
    case ACS_FMATH_TANH: r = fm_tanh(x); r2 = fm_atanh(x); break;
    case ACS_FMATH_POW_RATIO: r = fm_pow_ratio(x, y, -5.255876113278518); r2 = fm_pow_ratio(x, y, 34.16319474407325); break;
    case ACS_FMATH_ANGLE_SC: {   // (x, y) = (component across, component along): the sine / cosine pair as Auxiliary forms it
      const double ih = fm_rcp(fm_sqrt(x * x + y * y));
      r = fm_angle_sc(x * ih, y * ih, x, y);
      break;
    }
    case ACS_FMATH_SINCOS_SMALL: fm_sincos_small(x, &r, &r2); break;
  }
  out[i] = r;
  if (out2) out2[i] = r2;
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" {

int acs_env_create(const AcsTaskConfig* cfg, int device, AcsEnv** out) {
  if (!cfg || !out) return fail("acs_env_create: null argument");
  const int A = cfg->n_ego + cfg->n_enm;
  if (cfg->n_envs <= 0 || A <= 0 || A > ACS_MAX_AGENTS) return fail("acs_env_create: need 1..8 aircraft per env and n_envs > 0");
  if (cfg->n_rewards < 0 || cfg->n_rewards > ACS_MAX_REWARDS || cfg->n_terms < 0 || cfg->n_terms > ACS_MAX_TERMS)
    return fail("acs_env_create: too many rewards / terminations");
  if (cfg->substeps <= 0 || cfg->obs_dim <= 0) return fail("acs_env_create: substeps and obs_dim must be positive");
  if (cfg->shoot_dim != 0 && cfg->shoot_dim != 1 && cfg->shoot_dim != 4) return fail("acs_env_create: shoot_dim must be 0, 1 or 4");
  if (cfg->lock_len > 64) return fail("acs_env_create: lock_len > 64 is not supported");
  if (cfg->n_missile_slots > 64) return fail("acs_env_create: more than 64 missile slots per aircraft are not supported");
  if (std::strcmp(STATE_NAMES[F_SIM_TIME], "sim_time") || std::strcmp(STATE_NAMES[F_CMD0], "fcs/aileron-cmd-norm") ||
      std::strcmp(STATE_NAMES[F_CMD0 + 3], "fcs/throttle-cmd-norm"))
    return fail("acs_env_create: FDM state layout changed (sim_time / control fields)");
  AcsConfig fc;
  fc.n_envs = cfg->n_envs; fc.n_agents = A; fc.sim_dt = cfg->sim_dt; fc.fcs_dt = cfg->fcs_dt;
  AcsHandle* fh = nullptr;
  if (acs_create(&fc, device, &fh)) return 1;
  AcsEnv* e = new AcsEnv();
  e->cfg = *cfg; e->fdm = fh;
  e->lg = A <= 1 ? 0 : (A <= 2 ? 1 : (A <= 4 ? 2 : 3));
  e->G = 1 << e->lg;
  EnvView& v = e->v;
  v.B = cfg->n_envs; v.A = A; v.S = cfg->n_missile_slots > 0 ? cfg->n_missile_slots : 1; v.rows = v.B * A;
  v.fdm = fh->state; v.out = fh->out;
  const size_t rows = v.rows, B = v.B, ms = rows * v.S;
  CUDA_TRY(cudaMalloc(&v.ad, sizeof(double) * N_AD * rows)); CUDA_TRY(cudaMemset(v.ad, 0, sizeof(double) * N_AD * rows));
  CUDA_TRY(cudaMalloc(&v.ai, sizeof(int) * N_AI * rows));    CUDA_TRY(cudaMemset(v.ai, 0, sizeof(int) * N_AI * rows));
  CUDA_TRY(cudaMalloc(&v.ed, sizeof(double) * N_ED * B));    CUDA_TRY(cudaMemset(v.ed, 0, sizeof(double) * N_ED * B));
  CUDA_TRY(cudaMalloc(&v.ei, sizeof(int) * N_EI * B));       CUDA_TRY(cudaMemset(v.ei, 0, sizeof(int) * N_EI * B));
  CUDA_TRY(cudaMalloc(&v.md, sizeof(double) * N_MD * ms));   CUDA_TRY(cudaMemset(v.md, 0, sizeof(double) * N_MD * ms));
  CUDA_TRY(cudaMalloc(&v.mi, sizeof(int) * N_MI * ms));      CUDA_TRY(cudaMemset(v.mi, 0, sizeof(int) * N_MI * ms));
  // episode counters start at -1 so the first reset is episode 0
  CUDA_TRY(cudaMemset(v.ei + (size_t)EI_EPISODE * B, 0xff, sizeof(int) * B));
  // reset template: every task with fixed per-lane initial conditions
  std::memset(&e->tpl, 0, sizeof(ResetTpl));
  if (const char* s = std::getenv("ACS_RESET_TEMPLATE")) e->tpl_mode = std::atoi(s);
  if (const char* s = std::getenv("ACS_FUSED_RESET")) e->fused_reset = std::atoi(s) != 0;
  if (const char* s = std::getenv("ACS_POST_SPLIT")) e->post_split = std::atoi(s) != 0;
  e->post = post_kernel_for(e->cfg, &e->post_family);
  if (const char* s = std::getenv("ACS_POST_GENERIC")) if (std::atoi(s) != 0) { e->post = k_env_post<PostGeneric>; e->post_family = -1; }
  v.traj = nullptr; v.snap = nullptr;
  if (cfg->launch_kind != ACS_L_NONE && cfg->launch_kind != ACS_L_AUTO_GUN) {     // tasks that fly missiles
    CUDA_TRY(cudaMalloc(&v.traj, sizeof(double) * 6 * (size_t)cfg->substeps * rows));
    CUDA_TRY(cudaMemset(v.traj, 0, sizeof(double) * 6 * (size_t)cfg->substeps * rows));
    // state per substep of the aircraft a missile may hit within the step (7.6 KB per aircraft at K = 12; touched rarely)
    CUDA_TRY(cudaMalloc(&v.snap, sizeof(double) * N_SNAP * (size_t)cfg->substeps * rows));
    CUDA_TRY(cudaMemset(v.snap, 0, sizeof(double) * N_SNAP * (size_t)cfg->substeps * rows));
  }
  if (cfg->obs_kind != ACS_OBS_HEADING && e->tpl_mode > 0) {
    EnvView& t = e->tpl.t;
    t.B = 1; t.A = A; t.S = v.S; t.rows = A; t.traj = nullptr; t.snap = nullptr;
    // one contiguous block (256-byte aligned pieces): arenas of one env, reset observation, field lists
    const size_t n64max = (size_t)(N_STATE + FDM_N_OUT + N_AD) * A + (size_t)N_MD * A * v.S + N_ED;
    const size_t n32max = (size_t)N_AI * A + (size_t)N_MI * A * v.S + N_EI;
    e->n64max = n64max; e->n32max = n32max;
    const size_t sz[13] = {sizeof(double) * N_STATE * A, sizeof(double) * FDM_N_OUT * A, sizeof(double) * N_AD * A, sizeof(int) * N_AI * A,
                           sizeof(double) * N_ED, sizeof(int) * N_EI, sizeof(double) * N_MD * A * v.S, sizeof(int) * N_MI * A * v.S,
                           sizeof(double) * A * cfg->obs_dim, sizeof(double) * n64max, sizeof(int) * n64max, sizeof(int) * n32max,
                           sizeof(int) * n32max};
    size_t off[14] = {0};
    for (int k = 0; k < 13; k++) off[k + 1] = off[k] + ((sz[k] + 255) / 256) * 256;
    char* blk = nullptr;
    CUDA_TRY(cudaMalloc(&blk, off[13]));
    e->tpl_block = blk;
    t.fdm = (double*)(blk + off[0]); t.out = (double*)(blk + off[1]); t.ad = (double*)(blk + off[2]); t.ai = (int*)(blk + off[3]);
    t.ed = (double*)(blk + off[4]);  t.ei = (int*)(blk + off[5]);     t.md = (double*)(blk + off[6]); t.mi = (int*)(blk + off[7]);
    e->tpl_obs = (double*)(blk + off[8]);
    e->tpl.obs = e->tpl_obs;
    e->tpl.v64 = (double*)(blk + off[9]); e->tpl.d64 = (int*)(blk + off[10]);
    e->tpl.v32 = (int*)(blk + off[11]);   e->tpl.d32 = (int*)(blk + off[12]);
    // what a resetting warp reads: observation + packed words (the arenas in front are only read in FDM-only mode)
    e->tpl.base = blk + off[8]; e->tpl.bytes = (int)(off[13] - off[8]);
    if (build_reset_template(e)) return 1;
  }
  // substep-kernel choice: the two-warp frame wins while the batch leaves SM sub-partitions idle
  {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    e->split_max_threads = prop.multiProcessorCount * ACS_SPLIT_AUTO_LANES_PER_SM;
    e->n_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaFuncSetAttribute(k_env_substeps_split4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4_DYN_SMEM));
    if (const char* s = std::getenv("ACS_FRAME_SPLIT")) e->frame_split = std::atoi(s);
  }
  *out = e;
  return 0;
}

int acs_env_set_option(AcsEnv* e, const char* name, int value) {
  if (!e || !name) return fail("acs_env_set_option: null argument");
  if (!std::strcmp(name, "frame_split")) {
    if (value < -1 || value > 3) return fail("acs_env_set_option: frame_split must be -1 (auto), 0, 1, 2 or 3");
    e->frame_split = value;
    return 0;
  }
  return fail(std::string("acs_env_set_option: unknown option ") + name);
}

} // extern "C"

// reset() of one env on the template arenas (legacy stream, synchronous: called at create / when the initial conditions
// change, never on the step path).  Run twice over two different fill patterns: a word that comes out the same both
// times is one reset() writes (and its value does not depend on what was there); the others it leaves alone.
static int build_reset_template(AcsEnv* e, int stage) {
  ResetTpl& tp = e->tpl;
  EnvView& t = tp.t;
  if (t.fdm == nullptr) return 0;
  if (stage == 0) { tp.full = 0; tp.n64 = tp.n32 = 0; tp.n_stages = 1; }     // new stage-0 conditions drop the registered stages
  const int A = t.A, S = t.S, D = e->cfg.obs_dim;
  struct Arena { void* p; int nf, per; size_t elt; };
  const Arena ar[8] = {{t.fdm, N_STATE, A, 8}, {t.out, FDM_N_OUT, A, 8}, {t.ad, N_AD, A, 8}, {t.ai, N_AI, A, 4},
                       {t.ed, N_ED, 1, 8},     {t.ei, N_EI, 1, 4},       {t.md, N_MD, A * S, 8}, {t.mi, N_MI, A * S, 4}};
  std::vector<unsigned char> snap[2][8];
  ResetTpl none;
  std::memset(&none, 0, sizeof(none));
  // a stage's reset observation goes straight to its slot (the template arenas are scratch in whole-reset mode)
  double* obs_dst = stage == 0 ? e->tpl_obs : (double*)(e->stage_block + (size_t)stage * tp.stage_stride);
  CUDA_TRY(cudaDeviceSynchronize());
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < 8; k++) CUDA_TRY(cudaMemset(ar[k].p, pass ? 0x55 : 0x00, ar[k].elt * ar[k].nf * ar[k].per));
    CUDA_TRY(cudaMemset(obs_dst, 0, sizeof(double) * A * D));
    k_env_reset_fdm<<<1, FDM_BLOCK>>>(t, e->cfg, e->lg, nullptr);
    CUDA_TRY(cudaGetLastError());
    k_env_reset_task<<<1, 128>>>(t, e->cfg, e->lg, nullptr, obs_dst, nullptr, none);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    for (int k = 0; k < 8; k++) {
      snap[pass][k].resize(ar[k].elt * ar[k].nf * ar[k].per);
      CUDA_TRY(cudaMemcpy(snap[pass][k].data(), ar[k].p, snap[pass][k].size(), cudaMemcpyDeviceToHost));
    }
  }
  if (e->tpl_mode < 2) return stage == 0 ? 0 : fail("curriculum stages need the whole-reset template (ACS_RESET_TEMPLATE=2)");
  std::vector<double> v64; std::vector<int> d64, v32, d32;
  for (int k = 0; k < 8; k++) {
    for (int f = 0; f < ar[k].nf; f++) {
      int same = 0;
      for (int j = 0; j < ar[k].per; j++) {
        const size_t o = ((size_t)f * ar[k].per + j) * ar[k].elt;
        same += std::memcmp(&snap[0][k][o], &snap[1][k][o], ar[k].elt) == 0;
      }
      if (same != ar[k].per) {
        if (same != 0) return 0;     // written for some lanes / slots only: keep the computed reset
        // A word that differs between the passes must be one reset() leaves alone, i.e. still the fill pattern in both
        // (the episode counter, the one read-modify-write reset_copy_warp knows about, excepted).  Anything else is a
        // value that depends on what was there before: not a template, keep the computed reset.
        if (!(k == 5 && f == EI_EPISODE)) {
          for (int j = 0; j < ar[k].per; j++) {
            const size_t o = ((size_t)f * ar[k].per + j) * ar[k].elt;
            for (size_t b = 0; b < ar[k].elt; b++)
              if (snap[0][k][o + b] != 0x00 || snap[1][k][o + b] != 0x55) return 0;
          }
        }
        continue;
      }
      for (int j = 0; j < ar[k].per; j++) {
        const size_t o = ((size_t)f * ar[k].per + j) * ar[k].elt;
        const int d = (k << 24) | (f << 12) | j;
        if (ar[k].elt == 8) { double x; std::memcpy(&x, &snap[1][k][o], 8); v64.push_back(x); d64.push_back(d); }
        else { int x; std::memcpy(&x, &snap[1][k][o], 4); v32.push_back(x); d32.push_back(d); }
      }
    }
  }
  if (ar[6].per >= 4096 || N_STATE >= 4096) return 0;
  if (stage > 0) {
    // same words as stage 0 (which words reset() writes does not depend on the initial conditions), other values
    if (!tp.full) return fail("acs_env_set_stage_init_states: the task has no whole-reset template");
    std::vector<int> h64(tp.n64), h32(tp.n32);
    CUDA_TRY(cudaMemcpy(h64.data(), tp.d64, sizeof(int) * tp.n64, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(h32.data(), tp.d32, sizeof(int) * tp.n32, cudaMemcpyDeviceToHost));
    if ((int)d64.size() != tp.n64 || (int)d32.size() != tp.n32 || h64 != d64 || h32 != d32)
      return fail("acs_env_set_stage_init_states: reset() of this stage writes a different set of words than stage 0");
    char* sb = e->stage_block + (size_t)stage * tp.stage_stride;
    CUDA_TRY(cudaMemcpy(sb + tp.stage_off64, v64.data(), sizeof(double) * v64.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(sb + tp.stage_off32, v32.data(), sizeof(int) * v32.size(), cudaMemcpyHostToDevice));
    if (stage + 1 > tp.n_stages) tp.n_stages = stage + 1;
    return 0;
  }
  CUDA_TRY(cudaMemcpy((void*)tp.v64, v64.data(), sizeof(double) * v64.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((void*)tp.d64, d64.data(), sizeof(int) * d64.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((void*)tp.v32, v32.data(), sizeof(int) * v32.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((void*)tp.d32, d32.data(), sizeof(int) * d32.size(), cudaMemcpyHostToDevice));
  tp.n64 = (int)v64.size(); tp.n32 = (int)v32.size();
  tp.full = 1;
  return 0;
}

extern "C" {

int acs_env_get_option(const AcsEnv* e, const char* name, int* value) {
  if (!e || !name || !value) return fail("acs_env_get_option: null argument");
  if (!std::strcmp(name, "frame_split")) { *value = e->frame_split; return 0; }
  if (!std::strcmp(name, "frame_split_effective")) {   // what the next acs_env_step launches
    *value = frame_split_effective(e);
    return 0;
  }
  if (!std::strcmp(name, "post_family")) { *value = e->post_family; return 0; }
  if (!std::strcmp(name, "reset_template")) { *value = e->tpl.t.fdm == nullptr ? 0 : (e->tpl.full ? 2 : 1); return 0; }
  if (!std::strcmp(name, "launches_per_step")) {       // kernels one auto-resetting acs_env_step launches
    *value = (e->tpl.t.fdm != nullptr && e->tpl.full && e->fused_reset) ? 2 : ((e->tpl.t.fdm != nullptr) ? 3 : 4);
    if (frame_split_effective(e) == 0 && e->v.traj != nullptr) *value += 1;     // k_env_missiles
    return 0;
  }
  return fail(std::string("acs_env_get_option: unknown option ") + name);
}

int acs_env_destroy(AcsEnv* e) {
  if (!e) return 0;
  cudaSetDevice(e->fdm->device);
  cudaFree(e->v.ad); cudaFree(e->v.ai); cudaFree(e->v.ed); cudaFree(e->v.ei); cudaFree(e->v.md); cudaFree(e->v.mi);
  if (e->v.traj) cudaFree(e->v.traj);
  if (e->v.snap) cudaFree(e->v.snap);
  if (e->tpl_block) cudaFree(e->tpl_block);
  if (e->stage_block) cudaFree(e->stage_block);
  for (cudaEvent_t x : e->ev) cudaEventDestroy(x);
  acs_destroy(e->fdm);
  delete e;
  return 0;
}

int acs_env_set_init_states(AcsEnv* e, const double* init_host) {
  if (!e || !init_host) return fail("acs_env_set_init_states: null argument");
  std::memcpy(e->cfg.init_state, init_host, sizeof(double) * 12 * e->v.A);
  return build_reset_template(e);
}

int acs_env_set_stage_init_states(AcsEnv* e, int stage, const double* init_host) {
  if (!e || !init_host) return fail("acs_env_set_stage_init_states: null argument");
  if (stage < 0 || stage >= ACS_MAX_STAGES) return fail("acs_env_set_stage_init_states: stage must be in [0, 256)");
  if (stage == 0) return acs_env_set_init_states(e, init_host);
  if (e->tpl.t.fdm == nullptr || !e->tpl.full) return fail("acs_env_set_stage_init_states: the task has no whole-reset template");
  ResetTpl& tp = e->tpl;
  const int A = e->v.A, D = e->cfg.obs_dim;
  if (e->stage_block == nullptr) {
    const size_t o64 = ((sizeof(double) * A * D + 255) / 256) * 256;
    const size_t o32 = o64 + ((sizeof(double) * e->n64max + 255) / 256) * 256;
    const size_t stride = o32 + ((sizeof(int) * e->n32max + 255) / 256) * 256;
    CUDA_TRY(cudaMalloc(&e->stage_block, stride * ACS_MAX_STAGES));
    CUDA_TRY(cudaMemset(e->stage_block, 0, stride * ACS_MAX_STAGES));
    tp.stage_base = e->stage_block; tp.stage_stride = (int)stride; tp.stage_off64 = (int)o64; tp.stage_off32 = (int)o32;
  }
  // the stage's reset() is evaluated on the template arenas with its own initial conditions; the handle's stay as they are
  double saved[ACS_MAX_AGENTS][12];
  std::memcpy(saved, e->cfg.init_state, sizeof(saved));
  std::memcpy(e->cfg.init_state, init_host, sizeof(double) * 12 * A);
  const int rc = build_reset_template(e, stage);
  std::memcpy(e->cfg.init_state, saved, sizeof(saved));
  return rc;
}

AcsHandle* acs_env_fdm(AcsEnv* e) { return e ? e->fdm : nullptr; }

int acs_env_set_seed(AcsEnv* e, uint64_t seed, void* stream) {
  if (!e) return fail("acs_env_set_seed: null handle");
  e->cfg.seed = seed;
  CUDA_TRY(cudaMemsetAsync(e->v.ei + (size_t)EI_EPISODE * e->v.B, 0xff, sizeof(int) * e->v.B, (cudaStream_t)stream));
  return 0;
}

int acs_env_reset(AcsEnv* e, const uint8_t* env_mask_dev, double* obs_dev, double* share_obs_dev, void* stream) {
  if (!e || !obs_dev) return fail("acs_env_reset: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = e->v.B * e->G;
  if (e->tpl.t.fdm == nullptr) {
    k_env_reset_fdm<<<(threads + FDM_BLOCK - 1) / FDM_BLOCK, FDM_BLOCK, 0, st>>>(e->v, e->cfg, e->lg, env_mask_dev);
    CUDA_TRY(cudaGetLastError());
  }
  k_env_reset_task<<<(threads + 127) / 128, 128, 0, st>>>(e->v, e->cfg, e->lg, env_mask_dev, obs_dev, share_obs_dev, e->tpl);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int acs_env_step(AcsEnv* e, const int32_t* actions_dev, double* obs_dev, double* share_obs_dev, double* rewards_dev,
                 uint8_t* dones_dev, int32_t* info_dev, uint8_t* env_done_dev, int auto_reset, void* stream) {
  if (!e || !actions_dev || !obs_dev || !rewards_dev || !dones_dev) return fail("acs_env_step: null argument");
  if (auto_reset && !env_done_dev) return fail("acs_env_step: auto_reset needs env_done_dev");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = e->v.B * e->G;
  if (e->timing) timing_event(e, st);
  const int split = frame_split_effective(e);
  if (split == 3) k_env_substeps_split4<<<(threads + S4 - 1) / S4, 4 * S4, S4_DYN_SMEM, st>>>(e->v, e->cfg, e->lg, actions_dev);
  else if (split == 2) k_env_substeps_split3<<<(threads + S3 - 1) / S3, 3 * S3, 0, st>>>(e->v, e->cfg, e->lg, actions_dev);
  else if (split == 1) k_env_substeps_split<<<(threads + SPLIT_SLOTS - 1) / SPLIT_SLOTS, 2 * SPLIT_SLOTS, 0, st>>>(e->v, e->cfg, e->lg, actions_dev);
  else {
    // the one-thread frame carries no missile code: k_env_missiles does the missile / chaff work of the step
    k_env_substeps<<<(threads + FDM_BLOCK - 1) / FDM_BLOCK, FDM_BLOCK, 0, st>>>(e->v, e->cfg, e->lg, actions_dev);
    CUDA_TRY(cudaGetLastError());
    if (e->timing) timing_event(e, st);
    if (e->v.traj != nullptr) k_env_missiles<<<(threads + 127) / 128, 128, 0, st>>>(e->v, e->cfg, e->lg);
  }
  CUDA_TRY(cudaGetLastError());
  if (e->timing && split != 0) timing_event(e, st);     // (the multi-warp frames carry their missile phase: interval 0)
  if (e->timing) timing_event(e, st);
  const int fuse = (auto_reset && e->tpl.t.fdm != nullptr && e->tpl.full && e->fused_reset) ? 1 : 0;
  const int obs_split = (e->post_split && e->cfg.launch_kind == ACS_L_NONE && !e->cfg.use_artillery &&
                         e->cfg.obs_kind != ACS_OBS_HEADING) ? 1 : 0;
  e->post<<<(threads + 127) / 128, obs_split ? 256 : 128, 0, st>>>(e->v, e->cfg, e->lg, obs_dev, share_obs_dev, rewards_dev, dones_dev,
                                                                   info_dev, env_done_dev, fuse, e->tpl, obs_split);
  CUDA_TRY(cudaGetLastError());
  if (e->timing) timing_event(e, st);
  int rc = 0;
  if (auto_reset && !fuse) rc = acs_env_reset(e, env_done_dev, obs_dev, share_obs_dev, stream);
  if (e->timing) timing_event(e, st);
  return rc;
}

int acs_env_set_timing(AcsEnv* e, int on) {
  if (!e) return fail("acs_env_set_timing: null handle");
  e->timing = on != 0;
  e->ev_used = 0;
  return 0;
}

int acs_env_get_timing(AcsEnv* e, double ms[4], int* n_steps, int reset) {
  if (!e || !ms || !n_steps) return fail("acs_env_get_timing: null argument");
  ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
  const size_t n = e->ev_used / 5;
  // intervals in stream order: substeps, missiles, post, reset -> ms[0], ms[3], ms[1], ms[2]
  for (size_t k = 0; k < n; k++) {
    CUDA_TRY(cudaEventSynchronize(e->ev[5 * k + 4]));
    for (int j = 0; j < 4; j++) {
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, e->ev[5 * k + j], e->ev[5 * k + j + 1]));
      ms[j == 0 ? 0 : (j == 1 ? 3 : j - 1)] += t;
    }
  }
  *n_steps = (int)n;
  if (reset) e->ev_used = 0;
  return 0;
}

int acs_env_arena_info(const AcsEnv* e, int which, int* n_fields, int* n_per_field, int* is_int) {
  if (!e) return fail("acs_env_arena_info: null handle");
  void* p; int nf, per, ii;
  if (env_arena(e, which, &p, &nf, &per, &ii)) return 1;
  if (n_fields) *n_fields = nf; if (n_per_field) *n_per_field = per; if (is_int) *is_int = ii;
  return 0;
}
const char* acs_env_arena_field_name(int which, int f) {
  switch (which) {
    case 0: return acs_state_field_name(f);
    case 1: return acs_output_field_name(f);
    case 2: return (f >= 0 && f < N_AD) ? AD_NAMES[f] : nullptr;
    case 3: return (f >= 0 && f < N_AI) ? AI_NAMES[f] : nullptr;
    case 4: return (f >= 0 && f < N_ED) ? ED_NAMES[f] : nullptr;
    case 5: return (f >= 0 && f < N_EI) ? EI_NAMES[f] : nullptr;
    case 6: return (f >= 0 && f < N_MD) ? MD_NAMES[f] : nullptr;
    case 7: return (f >= 0 && f < N_MI) ? MI_NAMES[f] : nullptr;
  }
  return nullptr;
}
int acs_env_get_arena(const AcsEnv* e, int which, void* dst_dev, void* stream) {
  if (!e || !dst_dev) return fail("acs_env_get_arena: null argument");
  void* p; int nf, per, ii;
  if (env_arena(e, which, &p, &nf, &per, &ii)) return 1;
  CUDA_TRY(cudaMemcpyAsync(dst_dev, p, (size_t)nf * per * (ii ? sizeof(int) : sizeof(double)), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int acs_env_arena_ptr(const AcsEnv* e, int which, void** dev_ptr) {
  if (!e || !dev_ptr) return fail("acs_env_arena_ptr: null argument");
  int nf, per, ii;
  return env_arena(e, which, dev_ptr, &nf, &per, &ii);
}
int acs_env_set_arena(AcsEnv* e, int which, const void* src_dev, void* stream) {
  if (!e || !src_dev) return fail("acs_env_set_arena: null argument");
  void* p; int nf, per, ii;
  if (env_arena(e, which, &p, &nf, &per, &ii)) return 1;
  CUDA_TRY(cudaMemcpyAsync(p, src_dev, (size_t)nf * per * (ii ? sizeof(int) : sizeof(double)), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

#ifdef ACS_SPLIT_PROFILE
int acs_debug_split_profile(long long* out16) {
  CUDA_TRY(cudaMemcpyFromSymbol(out16, g_split_prof, sizeof(long long) * 16));
  return 0;
}
#endif

#ifdef ACS_MISSILE_PROFILE
int acs_debug_missile_profile(long long* out8, int reset) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out8, g_missile_prof, sizeof(long long) * 8));
  if (reset) { long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; CUDA_TRY(cudaMemcpyToSymbol(g_missile_prof, z, sizeof(z))); }
  return 0;
}
#endif

#ifdef ACS_FRAME_PROFILE
int acs_debug_frame_profile(long long* out16) {
  CUDA_TRY(cudaMemcpyFromSymbol(out16, g_frame_prof, sizeof(long long) * 16));
  return 0;
}
#endif

int acs_debug_fmath(int op, const double* a_dev, const double* b_dev, double* out_dev, double* out2_dev, int n, void* stream) {
  if (!a_dev || !out_dev || n < 0) return fail("acs_debug_fmath: null argument");
  if (op < 0 || op >= ACS_FMATH_N_OPS) return fail("acs_debug_fmath: unknown op");
  if ((op == ACS_FMATH_DIV || op == ACS_FMATH_ATAN2 || op == ACS_FMATH_POW_RATIO || op == ACS_FMATH_ANGLE_SC) && !b_dev)
    return fail("acs_debug_fmath: this op takes two operands");
  if (n == 0) return 0;
  k_fmath_probe<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, a_dev, b_dev, out_dev, out2_dev, n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int acs_bench_fp64_peak(int device, double* flops_out) {
  if (!flops_out) return fail("acs_bench_fp64_peak: null argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  k_fp64_peak<<<blocks, threads>>>(buf, 1024);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(buf, iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  *flops_out = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
  return 0;
}

}  // extern "C"
