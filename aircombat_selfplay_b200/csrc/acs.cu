// acs.cu -- kernels + C ABI (include/acs.h) of the batched air-combat simulator, sm_100a only.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <cmath>
#include "../../include/acs.h"
#include "fdm_core.cuh"

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
constexpr int FDM_BLOCK = 128;

__global__ void __launch_bounds__(FDM_BLOCK) k_fdm_run(double* __restrict__ state, double* __restrict__ out,
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

// ----------------------------------------------------------------------------- host side
static void host_atmo(AtmoConst& c) {
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
}

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
