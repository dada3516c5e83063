// C-ABI implementation (include/cmpc.h): handle, HBM layout, kernel launches.
// There is deliberately no CPU path in this file: every compute entry point launches
// sm_100a kernels and fails with CMPC_ERR_CUDA when that is not possible.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cmpc.h"
#include "handle.cuh"

using namespace cmpc;

namespace {
thread_local std::string g_err;
}

namespace cmpc {
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
extern const ShapeOps kOps_cent_par;
extern const ShapeOps kOps_coop_par;
extern const ShapeOps kOps_ncoop_par;
extern const ShapeOps kOps_cent_ser;
extern const ShapeOps kOps_coop_ser;
extern const ShapeOps kOps_ncoop_ser;
extern const ShapeOps kOps_ncoop_ser_old;
const ShapeOps* const kShapeOps[kNumShapes] = {&kOps_cent_par, &kOps_coop_par, &kOps_ncoop_par, &kOps_cent_ser, &kOps_coop_ser, &kOps_ncoop_ser, &kOps_ncoop_ser_old};
}  // namespace cmpc

namespace {

int find_shape(const cmpc_config& c) {
  const int ny = c.n_controlled_outputs[0];
  if (c.n_controllers == 2 && c.n_controlled_outputs[1] != ny) return -1;
  for (int i = 0; i < kNumShapes; ++i) {
    const ShapeOps& o = *kShapeOps[i];
    if (o.plant == c.plant && o.ny == ny && o.nu == c.n_sub_control_inputs && o.nctrl == c.n_controllers) return i;
  }
  return -1;
}

// Room for the four window events of n_steps control steps, starting at win_used = 0.
int reserve_window_events(cmpc_handle* h, size_t n_steps) {
  const size_t old = h->win_ev.size();
  if (old < 4 * n_steps) {
    h->win_ev.resize(4 * n_steps);
    for (size_t i = old; i < h->win_ev.size(); ++i) {
      cudaError_t e = cudaEventCreate(&h->win_ev[i]);
      if (e != cudaSuccess) {
        h->win_ev.resize(i);
        return fail(CMPC_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
      }
    }
  }
  h->win_used = 0;
  return CMPC_OK;
}

// Measured time of control step i of the last windowed run, in nanoseconds.
int window_ns(cmpc_handle* h, size_t i, int64_t* ns) {
  float a = 0.f, b = 0.f;
  cudaEvent_t* w = &h->win_ev[4 * i];
  cudaError_t e = cudaEventElapsedTime(&a, w[0], w[1]);
  if (e == cudaSuccess) e = cudaEventElapsedTime(&b, w[2], w[3]);
  if (e != cudaSuccess) return fail(CMPC_ERR_CUDA, std::string("cudaEventElapsedTime: ") + cudaGetErrorString(e));
  *ns = int64_t((double(a) + double(b)) * 1e6 + 0.5);
  return CMPC_OK;
}

// The general path keeps its parameter block in device memory: setters rewrite it (setup time only;
// nothing may be in flight).
int upload_generic_params(cmpc_handle* h) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_genp, &h->GP, sizeof(GenParams), cudaMemcpyHostToDevice));
  return CMPC_OK;
}

// StepParams::obs_states_free: do the observer gains leave the plant-state estimates alone?
void update_obs_states_free(cmpc_handle* h) {
  bool free_ = true;
  for (int c = 0; c < h->NCTRL; ++c)
    for (int i = 0; i < h->N * 4; ++i)
      if (h->P.c[c].M[i] != 0.0) free_ = false;
  h->P.obs_states_free = free_ ? 1 : 0;
}

// Device address of a host buffer that is page-locked and mapped into the device (cudaHostAlloc /
// cudaHostRegister, pinned torch tensors) and aligned as asked; null for anything else.
void* mapped_host(const void* host, size_t align) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return (reinterpret_cast<uintptr_t>(at.devicePointer) % align == 0) ? at.devicePointer : nullptr;
}

// First statement of every entry point that takes a handle: argument check, then the handle's
// device becomes current until the entry point returns (DeviceGuard puts the caller's back).
#define CMPC_ENTER(h)                                                                              \
  if (!(h)) return fail(CMPC_ERR_ARG, "null handle");                                              \
  DeviceGuard device_guard_((h)->device);                                                          \
  if (device_guard_.err != cudaSuccess)                                                            \
    return fail(CMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_guard_.err))
#define CMPC_ENTER_DEVICE(dev)                                                                     \
  DeviceGuard device_guard_(dev);                                                                  \
  if (device_guard_.err != cudaSuccess)                                                            \
    return fail(CMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_guard_.err))

}  // namespace

extern "C" {

const char* cmpc_last_error(void) { return g_err.c_str(); }

int cmpc_plant_dims(int plant, int* n_states, int* n_inputs) {
  if (plant != 0 && plant != 1) return fail(CMPC_ERR_ARG, "plant must be 0 or 1");
  if (n_states) *n_states = plant == 0 ? 11 : 10;
  if (n_inputs) *n_inputs = plant == 0 ? 9 : 8;
  return CMPC_OK;
}

int cmpc_plant_defaults(int plant, double* x, double* u) {
  if (plant != 0 && plant != 1) return fail(CMPC_ERR_ARG, "plant must be 0 or 1");
  const double xp[11] = {0.916, 1.145, 0.152, 440, 0, 0.916, 1.145, 0.152, 440, 0, 1.12};
  const double up[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  const double xs[10] = {0.867, 1.03, 0.176, 395, 0, 0.999, 1.19, 0.176, 395, 0};
  const double us[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  if (x) std::memcpy(x, plant == 0 ? xp : xs, sizeof(double) * (plant == 0 ? 11 : 10));
  if (u) std::memcpy(u, plant == 0 ? up : us, sizeof(double) * (plant == 0 ? 9 : 8));
  return CMPC_OK;
}

int cmpc_default_config(int plant, int mode, int batch, cmpc_config* cfg) {
  if (!cfg) return fail(CMPC_ERR_ARG, "null cfg");
  if ((plant != 0 && plant != 1) || mode < 0 || mode > 3 || (mode == 3 && plant != 1))
    return fail(CMPC_ERR_ARG, "bad plant/mode");
  std::memset(cfg, 0, sizeof *cfg);
  cfg->plant = plant;
  cfg->mode = mode;
  cfg->p = 100;
  cfg->m = 2;
  cfg->Ts = 0.05;
  cfg->batch = batch;
  const int delays[4] = {0, 40, 0, 40};
  std::memcpy(cfg->delays, delays, sizeof delays);
  cfg->n_disturbance_states = 4;
  const int id1[4] = {0, 1, 2, 3}, id2[4] = {2, 3, 0, 1};
  std::memcpy(cfg->control_input_indices[0], id1, sizeof id1);
  std::memcpy(cfg->control_input_indices[1], id2, sizeof id2);
  auto set_out = [&](int c, std::initializer_list<int> o) {
    cfg->n_controlled_outputs[c] = int(o.size());
    int i = 0;
    for (int v : o) cfg->controlled_output_indices[c][i++] = v;
  };
  if (mode == CMPC_MODE_CENTRALIZED) {
    cfg->n_controllers = 1;
    cfg->n_sub_control_inputs = 4;
    cfg->n_iterations = 1;
    if (plant == 0) set_out(0, {0, 1, 3}); else set_out(0, {0, 1, 2, 3});
  } else {
    cfg->n_controllers = 2;
    cfg->n_sub_control_inputs = 2;
    cfg->n_iterations = 9;
    if (plant == 0 && mode == 1) { set_out(0, {0, 1, 3}); set_out(1, {0, 1, 3}); }
    if (plant == 0 && mode == 2) { set_out(0, {0, 3}); set_out(1, {1, 3}); }
    if (plant == 1 && mode == 1) { set_out(0, {0, 1, 2, 3}); set_out(1, {0, 1, 2, 3}); }
    if (plant == 1 && mode == 2) { set_out(0, {0, 1}); set_out(1, {2, 3}); }
    if (plant == 1 && mode == 3) { set_out(0, {0, 1, 2}); set_out(1, {2, 3, 1}); }
  }
  return CMPC_OK;
}

// Own inputs of sub-controller c.
static int nu_of(const cmpc_config& c, int ctrl) {
  return c.n_sub_control_inputs_per[ctrl] > 0 ? c.n_sub_control_inputs_per[ctrl] : c.n_sub_control_inputs;
}

// The tuned kernels (step_kernel.cuh) cover the reference's own instantiations: m = 2,
// Delays = {0,40,0,40}, one of the seven controller shapes.  Everything else runs on the general
// path (generic_kernels.cuh).  CMPC_FORCE_GENERIC=1 sends the reference's shapes there too (tests).
static int fast_path_shape(const cmpc_config& c) {
  if (const char* e = getenv("CMPC_FORCE_GENERIC"))
    if (e[0] == '1') return -1;
  const int delays[4] = {0, kDelay, 0, kDelay};
  if (c.m != 2 || std::memcmp(c.delays, delays, sizeof delays) != 0 || c.n_controllers > 2 || c.p < 2) return -1;
  for (int k = 0; k < c.n_controllers; ++k) {
    if (nu_of(c, k) != c.n_sub_control_inputs) return -1;
    for (int i = 0; i < 4; ++i) {
      const int v = c.control_input_indices[k][i];
      // delays are attached to the local position (aug_lin_sys.cc:160-173): the tuned kernels need
      // the permutation to map delayed positions onto delayed plant inputs
      if (v < 0 || v > 3 || delays[i] != delays[v]) return -1;
    }
  }
  return find_shape(c);
}

static int fill_generic_params(const cmpc_config& cfg, int N, GenParams* out) {
  GenParams& P = *out;
  std::memset(&P, 0, sizeof P);
  P.plant = cfg.plant; P.p = cfg.p; P.m = cfg.m; P.n_iter = cfg.n_iterations; P.batch = cfg.batch;
  P.n_ctrl = cfg.n_controllers; P.n = N; P.n_obs = N + kNDist; P.Ts = cfg.Ts;
  const double Ts = cfg.Ts;
  P.rk[0] = Ts; P.rk[1] = Ts * Ts / 2.0; P.rk[2] = Ts * Ts * Ts / 6.0; P.rk[3] = Ts * Ts * Ts * Ts / 24.0;
  if (cfg.m < 1 || 4 * cfg.m > kGenMaxPred) return fail(CMPC_ERR_UNSUPPORTED, "move horizon m must be in [1, 4]");
  if (cfg.p < cfg.m || cfg.p > kGenMaxP) return fail(CMPC_ERR_UNSUPPORTED, "prediction horizon must be in [m, 256]");
  if (cfg.n_controllers < 1 || cfg.n_controllers > kGenMaxCtrl)
    return fail(CMPC_ERR_UNSUPPORTED, "1 to 4 sub-controllers");
  int sum_d = 0, dmax = 1, ring = 0;
  for (int i = 0; i < 4; ++i) {
    const int d = cfg.delays[i];
    // a delay of one sample has no chain state and the reference's BComposite index arithmetic
    // (aug_lin_sys.cc:182-199) then points at another input's state
    if (d < 0 || d == 1 || d > kGenMaxDelay) return fail(CMPC_ERR_UNSUPPORTED, "delays must be 0 or in [2, 128] samples");
    P.delays_sys[i] = d;
    P.ring_off[i] = ring;
    ring += d;
    sum_d += d;
    dmax = d > dmax ? d : dmax;
  }
  if (sum_d > 240) return fail(CMPC_ERR_UNSUPPORTED, "more than 240 delay states");
  P.ring_total = ring > 0 ? ring : 1;
  P.dmax = dmax;
  P.n_total = N + kNDist + sum_d;
  P.state_stride = (N + P.n_total + 8 + 1) & ~1;
  int in_off = 0, pred_off = 0, yref_off = 0;
  for (int c = 0; c < cfg.n_controllers; ++c) {
    GenCtrl& K = P.c[c];
    K.nu = nu_of(cfg, c);
    K.ny = cfg.n_controlled_outputs[c];
    if (K.nu < 1 || K.nu > 4 || K.ny < 1 || K.ny > 4) return fail(CMPC_ERR_ARG, "bad input / output count of a sub-controller");
    K.nv = cfg.m * K.nu;
    if (K.nv > kGenMaxNv)
      return fail(CMPC_ERR_UNSUPPORTED, "m * n_sub_control_inputs must be at most 8 (the active-set word has 32 bits)");
    K.no = 4 - K.nu;
    K.nvo = cfg.m * K.no;
    K.reduced = K.nu != 4;
    unsigned seen = 0;
    for (int i = 0; i < 4; ++i) {
      const int v = cfg.control_input_indices[c][i];
      if (v < 0 || v > 3) return fail(CMPC_ERR_ARG, "control_input_indices out of range");
      seen |= 1u << v;
      K.ctrl_idx[i] = v;
      K.out_idx[i] = i < K.ny ? cfg.controlled_output_indices[c][i] : 0;
      if (K.out_idx[i] < 0 || K.out_idx[i] > 3) return fail(CMPC_ERR_ARG, "controlled_output_indices out of range");
      // NerveCenter adds the first moves of sub-controller c to the system inputs that follow those
      // of the sub-controllers before it (nerve_center.h:313-319) and hands every sub-controller
      // du[ControlInputIndices] (:322-328): the two agree only for this layout
      if (i < K.nu && K.reduced && v != in_off + i)
        return fail(CMPC_ERR_ARG, "the own inputs of a sub-controller must be the system inputs that follow "
                                  "those of the sub-controllers before it (nerve_center.h:313-328)");
    }
    if (seen != 0xF) return fail(CMPC_ERR_ARG, "control_input_indices must be a permutation");
    // each sub-controller's AugmentedLinearizedSystem sees the delays in its own input order; a
    // controller that is not reduced keeps the system order (aug_lin_sys.cc:158)
    int n_del = 0;
    for (int i = 0; i < 4; ++i) {
      K.delay[i] = cfg.delays[K.reduced ? K.ctrl_idx[i] : i];
      if (K.delay[i] > 0) ++n_del;
    }
    int head = N + kNDist, chain = N + kNDist + n_del;
    for (int i = 0; i < 4; ++i) {
      K.head[i] = K.chain[i] = -1;
      if (K.delay[i] > 0) {
        K.head[i] = head++;
        K.chain[i] = chain;
        chain += K.delay[i] - 1;
      }
    }
    K.pred_off = pred_off; K.in_off = in_off; K.yref_off = yref_off;
    pred_off += K.nv; in_off += K.nu; yref_off += cfg.p * K.ny;
    for (int i = 0; i < 4; ++i) { K.lower[i] = -1e30; K.upper[i] = 1e30; K.rate_lower[i] = -1e30; K.rate_upper[i] = 1e30; }
    for (int i = 0; i < K.ny; ++i) K.Q[i * K.ny + i] = 1.0;
    for (int i = 0; i < K.nu; ++i) K.R[i * K.nu + i] = 1.0;
    for (int i = 0; i < 4; ++i) K.M[(N + i) * 4 + i] = 1.0;   // default gain [0; I]
  }
  if (in_off != 4) return fail(CMPC_ERR_ARG, "the sub-controllers' own inputs must add up to the four control inputs");
  P.n_pred = pred_off;
  return CMPC_OK;
}

int cmpc_create(const cmpc_config* cfg, int device, cmpc_handle** out) {
  if (!cfg || !out) return fail(CMPC_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->batch <= 0) return fail(CMPC_ERR_ARG, "batch must be positive");
  if (cfg->plant != 0 && cfg->plant != 1) return fail(CMPC_ERR_ARG, "plant must be 0 or 1");
  if (cfg->n_iterations < 1) return fail(CMPC_ERR_ARG, "n_iterations must be >= 1");
  if (cfg->n_disturbance_states != kNDist)
    return fail(CMPC_ERR_UNSUPPORTED, "4 disturbance states (one per plant output, as in both plants of the reference)");
  if (cfg->n_controllers < 1 || cfg->n_controllers > CMPC_MAX_CONTROLLERS) return fail(CMPC_ERR_ARG, "bad n_controllers");
  for (int c = 0; c < cfg->n_controllers; ++c)
    for (int i = 0; i < 4 && i < cfg->n_controlled_outputs[c]; ++i)
      if (cfg->controlled_output_indices[c][i] < 0 || cfg->controlled_output_indices[c][i] > 3)
        return fail(CMPC_ERR_ARG, "controlled_output_indices out of range");
  const int shape = fast_path_shape(*cfg);
  int N = 0, NIN = 0;
  cmpc_plant_dims(cfg->plant, &N, &NIN);
  GenParams gp;
  if (shape < 0) {
    if (int rc = fill_generic_params(*cfg, N, &gp)) return rc;
  } else if (cfg->p > 256) {
    return fail(CMPC_ERR_UNSUPPORTED, "prediction horizon must be in [2, 256]");
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(CMPC_ERR_CUDA, "no CUDA device: the control step has no CPU path");
  if (device < 0 || device >= n_dev) return fail(CMPC_ERR_ARG, "bad device index");
  CMPC_ENTER_DEVICE(device);

  cmpc_handle* h = new cmpc_handle;
  h->cfg = *cfg;
  h->device = device;
  h->shape = shape;
  h->generic = shape < 0;
  h->ops = h->generic ? (cfg->plant == 0 ? &kOps_generic_par : &kOps_generic_ser) : kShapeOps[shape];
  h->N = N;
  h->NIN = NIN;
  h->NCTRL = cfg->n_controllers;
  h->NV = cfg->m * cfg->n_sub_control_inputs;
  h->NVO = cfg->m * (4 - cfg->n_sub_control_inputs);
  std::memset(&h->P, 0, sizeof h->P);
  std::memset(&h->G, 0, sizeof h->G);
  std::memset(&h->GS, 0, sizeof h->GS);
  h->GP = gp;
  StepParams& P = h->P;
  P.p = cfg->p;
  P.b_max = (cfg->p + kBaby - 1) / kBaby;
  P.n_pow = ladder_stages(P.b_max);
  P.b_full = full_blocks(P.p, P.b_max);
  P.ldr = giant_stride(giant_cols(P.b_max, P.b_full));
  P.n_iter = cfg->n_iterations;
  P.batch = cfg->batch;
  P.Ts = cfg->Ts;
  {
    const double Ts = cfg->Ts;
    P.rk[0] = Ts; P.rk[1] = Ts * Ts / 2.0; P.rk[2] = Ts * Ts * Ts / 6.0; P.rk[3] = Ts * Ts * Ts * Ts / 24.0;
  }
  const int B = cfg->batch, NC = h->NCTRL;
  for (int c = 0; c < NC && c < 2 && !h->generic; ++c) {
    CtrlParams& cp = P.c[c];
    const int ny = cfg->n_controlled_outputs[c], nu = cfg->n_sub_control_inputs;
    for (int i = 0; i < 4; ++i) {
      cp.out_idx[i] = i < ny ? cfg->controlled_output_indices[c][i] : 0;
      cp.ctrl_idx[i] = cfg->control_input_indices[c][i];
      cp.lower[i] = -1e30; cp.upper[i] = 1e30; cp.rate_lower[i] = -1e30; cp.rate_upper[i] = 1e30;
    }
    for (int i = 0; i < ny; ++i) cp.Q[i * ny + i] = 1.0;
    for (int i = 0; i < nu; ++i) cp.R[i * nu + i] = 1.0;
    for (int i = 0; i < 4; ++i) cp.M[(N + i) * 4 + i] = 1.0;  // default gain [0; I]
  }
  if (!h->generic) update_obs_states_free(h);
  DeviceState& G = h->G;
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  if (h->generic) {
    GenState& S = h->GS;
    A(dalloc(&S.ctrl, size_t(B) * NC * gp.state_stride));
    A(dalloc(&S.guess, size_t(B) * NC));
    A(dalloc(&S.scen, size_t(B) * kGenScenStride));
    A(dalloc(&S.u_offset, size_t(B) * NIN));
    A(dalloc(&S.qpH, size_t(B) * NC * 64));
    A(dalloc(&S.qpf, size_t(B) * NC * 8));
    A(dalloc(&S.status, size_t(B) * NC));
    A(dalloc(&S.active, size_t(B) * NC));
    A(dalloc(&S.objective, size_t(B) * NC));
    A(dalloc(&h->d_yref, size_t(NC) * cfg->p * 4));
    A(dalloc(&h->d_genp, 1));
    A(dalloc(&h->d_ring, size_t(B) * gp.ring_total));
    S.yref = h->d_yref;
    // the read-back entry points find results where the tuned path keeps them
    G.status = S.status; G.active = S.active; G.objective = S.objective;
  } else {
    A(dalloc(&G.ctrl, size_t(B) * NC * kCtrlStateStride));
    A(dalloc(&G.guess, size_t(B) * NC));
    A(dalloc(&G.scen, size_t(B) * kScenStateStride));
    A(dalloc(&G.u_offset, size_t(B) * h->NIN));
    A(dalloc(&G.work, size_t(B) * NC * kWorkStride));
    A(dalloc(&G.qpH, size_t(B) * NC * h->NV * h->NV));
    A(dalloc(&G.qpf, size_t(B) * NC * h->NV));
    A(dalloc(&G.qpG, size_t(B) * NC * h->NV * (h->NVO > 0 ? h->NVO : 1)));
    A(dalloc(&G.status, size_t(B) * NC));
    A(dalloc(&G.active, size_t(B) * NC));
    A(dalloc(&G.objective, size_t(B) * NC));
    A(dalloc(&G.ticks, size_t(B) * 32));
    A(dalloc(&h->d_yref, size_t(NC) * cfg->p * 4));
    A(dalloc(&h->d_ring, size_t(B) * 2 * kDelay));
  }
  A(dalloc(&h->d_y, size_t(B) * 4));
  A(dalloc(&h->d_u, size_t(B) * 4));
  A(dalloc(&h->d_xinit, size_t(B) * N));
  A(dalloc(&h->d_uinit, size_t(B) * 4));
  A(dalloc(&h->d_uinitfull, size_t(B) * h->NIN));
  A(dalloc(&h->d_yinit, size_t(B) * 4));
  A(dalloc(&h->d_x, size_t(B) * N));
  A(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (e == cudaSuccess && h->generic) e = cudaMemcpy(h->d_genp, &h->GP, sizeof(GenParams), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cmpc_destroy(h);
    return fail(CMPC_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  P.yref = h->d_yref;
  int rc = h->ops->setup(h);
  if (rc) {
    cmpc_destroy(h);
    return rc;
  }
  *out = h;
  return CMPC_OK;
}

int cmpc_destroy(cmpc_handle* h) {
  if (!h) return CMPC_OK;
  DeviceGuard device_guard_(h->device);
  cudaDeviceSynchronize();
  DeviceState& G = h->G;
  if (h->generic) G.status = nullptr, G.active = nullptr, G.objective = nullptr;   // aliases of the GenState arrays
  GenState& S = h->GS;
  void* ptrs[] = {S.ctrl, S.guess, S.scen, S.u_offset, S.qpH, S.qpf, S.status, S.active, S.objective, h->d_genp,
                  G.ctrl, G.guess, G.scen, G.u_offset, G.work, G.qpH, G.qpf, G.qpG, G.lin, G.etab, G.status,
                  G.active, G.objective, G.ticks, h->d_yref, h->d_y, h->d_u, h->d_xinit, h->d_uinit,
                  h->d_uinitfull, h->d_yinit, h->d_x, h->d_ring, h->d_block_end, h->d_block_off,
                  h->d_step_end, h->d_step_off, h->d_rec};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->ev_plant) cudaEventDestroy(h->ev_plant);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->win_ev) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CMPC_OK;
}

int cmpc_set_weights(cmpc_handle* h, int ctrl, const double* uwt, const double* ywt) {
  CMPC_ENTER(h);
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  const int ny = h->cfg.n_controlled_outputs[ctrl], nu = nu_of(h->cfg, ctrl);
  if (ywt)
    for (int i = 0; i < ny; ++i)
      for (int j = 0; j < i; ++j)
        if (ywt[i * ny + j] != ywt[j * ny + i])
          return fail(CMPC_ERR_UNSUPPORTED, "ywt must be symmetric (H = Su' Q Su is kept as a symmetric matrix)");
  // the reference adds the full u_weight_ to H (mpc_qp_solver.cc:27); the Hessian is kept as a
  // symmetric matrix here, so an input weight that would make it non-symmetric is refused as well
  if (uwt)
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < i; ++j)
        if (uwt[i * nu + j] != uwt[j * nu + i])
          return fail(CMPC_ERR_UNSUPPORTED, "uwt must be symmetric (H = Su' Q Su + R is kept as a symmetric matrix)");
  for (int i = 0; uwt && i < nu * nu; ++i)
    if (!std::isfinite(uwt[i])) return fail(CMPC_ERR_ARG, "uwt must be finite");
  for (int i = 0; ywt && i < ny * ny; ++i)
    if (!std::isfinite(ywt[i])) return fail(CMPC_ERR_ARG, "ywt must be finite");
  if (h->generic) {
    GenCtrl& K = h->GP.c[ctrl];
    if (uwt) std::memcpy(K.R, uwt, sizeof(double) * nu * nu);
    if (ywt) std::memcpy(K.Q, ywt, sizeof(double) * ny * ny);
    return upload_generic_params(h);
  }
  CtrlParams& cp = h->P.c[ctrl];
  if (uwt) std::memcpy(cp.R, uwt, sizeof(double) * nu * nu);
  if (ywt) std::memcpy(cp.Q, ywt, sizeof(double) * ny * ny);
  return CMPC_OK;
}

int cmpc_set_output_reference(cmpc_handle* h, const double* yref) {
  CMPC_ENTER(h);
  if (!yref) return fail(CMPC_ERR_ARG, "null yref");
  // nerve_center.h:237-249: each controller keeps its controlled outputs of every prediction row
  const int p = h->cfg.p;
  std::vector<double> sub(size_t(h->NCTRL) * p * 4, 0.0);
  for (int c = 0; c < h->NCTRL; ++c) {
    const int ny = h->cfg.n_controlled_outputs[c];
    for (int r = 0; r < p; ++r)
      for (int i = 0; i < ny; ++i)
        sub[size_t(c) * p * 4 + size_t(r) * ny + i] = yref[r * 4 + h->cfg.controlled_output_indices[c][i]];
  }
  if (h->generic) {   // one p x n_y block per sub-controller, back to back (GenCtrl::yref_off)
    std::vector<double> packed;
    for (int c = 0; c < h->NCTRL; ++c) {
      const int nyc = h->cfg.n_controlled_outputs[c];
      packed.insert(packed.end(), sub.begin() + size_t(c) * p * 4, sub.begin() + size_t(c) * p * 4 + size_t(p) * nyc);
    }
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h->d_yref, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice));
    return CMPC_OK;
  }
  // device layout [NCTRL][NY][p] with NY common to both controllers
  const int ny = h->cfg.n_controlled_outputs[0];
  std::vector<double> packed(size_t(h->NCTRL) * p * ny);
  for (int c = 0; c < h->NCTRL; ++c)
    for (int r = 0; r < p; ++r)
      for (int i = 0; i < ny; ++i) packed[(size_t(c) * ny + i) * p + r] = sub[size_t(c) * p * 4 + size_t(r) * ny + i];
  CU(cudaMemcpy(h->d_yref, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice));
  return CMPC_OK;
}

int cmpc_set_constraints(cmpc_handle* h, int ctrl, const double* lower, const double* upper,
                         const double* rate_lower, const double* rate_upper) {
  CMPC_ENTER(h);
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  if (!lower || !upper || !rate_lower || !rate_upper) return fail(CMPC_ERR_ARG, "null constraint array");
  const int nu = nu_of(h->cfg, ctrl);
  // a bound is a finite number or +-infinity (kept as +-1e30, which no input ever reaches); NaN is
  // refused.  lower > upper is accepted like the reference accepts it: every QP is then infeasible
  // and every step applies the zero move (mpc_qp_solver.cc:66-69), with status != 0 in cmpc_get_step_info.
  auto clamp = [](double v) { return v > 1e30 ? 1e30 : (v < -1e30 ? -1e30 : v); };
  for (int i = 0; i < nu; ++i)
    if (std::isnan(lower[i]) || std::isnan(upper[i]) || std::isnan(rate_lower[i]) || std::isnan(rate_upper[i]))
      return fail(CMPC_ERR_ARG, "constraint bounds must not be NaN");
  if (h->generic) {
    GenCtrl& K = h->GP.c[ctrl];
    for (int i = 0; i < nu; ++i) {
      K.lower[i] = clamp(lower[i]); K.upper[i] = clamp(upper[i]);
      K.rate_lower[i] = clamp(rate_lower[i]); K.rate_upper[i] = clamp(rate_upper[i]);
    }
    return upload_generic_params(h);
  }
  CtrlParams& cp = h->P.c[ctrl];
  for (int i = 0; i < nu; ++i) {
    cp.lower[i] = clamp(lower[i]); cp.upper[i] = clamp(upper[i]);
    cp.rate_lower[i] = clamp(rate_lower[i]); cp.rate_upper[i] = clamp(rate_upper[i]);
  }
  return CMPC_OK;
}

int cmpc_set_observer_gain(cmpc_handle* h, int ctrl, const double* M) {
  CMPC_ENTER(h);
  if (ctrl < 0 || ctrl >= h->NCTRL || !M) return fail(CMPC_ERR_ARG, "bad argument");
  if (h->generic) {
    std::memcpy(h->GP.c[ctrl].M, M, sizeof(double) * (h->N + kNDist) * 4);
    return upload_generic_params(h);
  }
  std::memcpy(h->P.c[ctrl].M, M, sizeof(double) * (h->N + kNDist) * 4);
  update_obs_states_free(h);
  return CMPC_OK;
}

int cmpc_set_capture(cmpc_handle* h, int on) {
  CMPC_ENTER(h);
  if (h->generic && on)
    return fail(CMPC_ERR_UNSUPPORTED, "the linearisation / prediction read-back hooks exist on the tuned path only");
  const size_t B = h->cfg.batch, NC = h->NCTRL;
  const size_t ny = h->cfg.n_controlled_outputs[0];
  if (on && !h->G.lin) {
    CU(dalloc(&h->G.lin, B * NC * (h->N * h->N + h->N * 5)));
    CU(dalloc(&h->G.etab, B * NC * size_t(h->cfg.p) * ny * 5));
  } else if (!on && h->G.lin) {
    CU(cudaDeviceSynchronize());
    cudaFree(h->G.lin); cudaFree(h->G.etab);
    h->G.lin = nullptr; h->G.etab = nullptr;
  }
  h->capture = on != 0;
  return CMPC_OK;
}

int cmpc_initialize(cmpc_handle* h, const double* x_init, const double* u_init,
                    const double* u_init_full, const double* y_init) {
  CMPC_ENTER(h);
  if (!x_init || !u_init || !u_init_full || !y_init) return fail(CMPC_ERR_ARG, "null argument");
  const size_t B = h->cfg.batch;
  CU(cudaDeviceSynchronize());   // nothing launched earlier (on any stream) may still be using the state
  CU(cudaMemcpy(h->d_xinit, x_init, B * h->N * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_uinit, u_init, B * 4 * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_uinitfull, u_init_full, B * h->NIN * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_yinit, y_init, B * 4 * sizeof(double), cudaMemcpyHostToDevice));
  int rc = h->ops->init(h, h->d_xinit, h->d_uinit, h->d_uinitfull, h->d_yinit, h->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  h->initialized = true;
  h->loop_started = false;   // the controller was restarted on its own: the on-device plants no longer match it
  return CMPC_OK;
}

int cmpc_get_next_input_device(cmpc_handle* h, const double* y_dev, double* u_dev, void* stream) {
  CMPC_ENTER(h);
  if (!h->initialized) return fail(CMPC_ERR_STATE, "cmpc_initialize has not been called");
  if (h->lin_ahead) return fail(CMPC_ERR_STATE, "a closed-loop run owns the controller state: call cmpc_initialize first");
  if (!y_dev || !u_dev) return fail(CMPC_ERR_ARG, "null argument");
  return h->ops->step(h, y_dev, u_dev, static_cast<cudaStream_t>(stream));
}

int cmpc_get_next_input(cmpc_handle* h, const double* y, double* u) {
  CMPC_ENTER(h);
  if (!h->initialized) return fail(CMPC_ERR_STATE, "cmpc_initialize has not been called");
  if (h->lin_ahead) return fail(CMPC_ERR_STATE, "a closed-loop run owns the controller state: call cmpc_initialize first");
  if (!y || !u) return fail(CMPC_ERR_ARG, "null argument");
  const size_t bytes = size_t(h->cfg.batch) * 4 * sizeof(double);
  // Page-locked buffers that are mapped into the device are used directly (as in cmpc_closed_loop_step): the
  // linearisation kernel brings the measurements in, the solve kernel writes the inputs out, and no copy
  // stands in front of or behind the step.  Anything else goes through the handle's device buffers.
  const bool no_direct = h->generic || getenv("CMPC_NO_DIRECT_HOST_IO") != nullptr;
  const double* y_map = no_direct ? nullptr : static_cast<const double*>(mapped_host(y, 8));
  double* u_map = no_direct ? nullptr : static_cast<double*>(mapped_host(u, 16));
  if (!y_map) CU(cudaMemcpyAsync(h->d_y, y, bytes, cudaMemcpyHostToDevice, h->stream));
  h->y_mapped_src = y_map;
  int rc = h->ops->step(h, h->d_y, u_map ? u_map : h->d_u, h->stream);
  h->y_mapped_src = nullptr;
  if (rc) return rc;
  if (!u_map) CU(cudaMemcpyAsync(u, h->d_u, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return CMPC_OK;
}

int cmpc_get_next_input_timed(cmpc_handle* h, const double* y, double* u, int n_timing_iterations,
                              int64_t* time_ns) {
  CMPC_ENTER(h);
  if (!h->initialized) return fail(CMPC_ERR_STATE, "cmpc_initialize has not been called");
  if (h->lin_ahead) return fail(CMPC_ERR_STATE, "a closed-loop run owns the controller state: call cmpc_initialize first");
  if (!y || !u || !time_ns) return fail(CMPC_ERR_ARG, "null argument");
  if (int rc = reserve_window_events(h, 1)) return rc;
  const size_t bytes = size_t(h->cfg.batch) * 4 * sizeof(double);
  cudaEvent_t* w = h->win_ev.data();
  CU(cudaEventRecord(w[0], h->stream));
  CU(cudaMemcpyAsync(h->d_y, y, bytes, cudaMemcpyHostToDevice, h->stream));
  h->window_on = true;
  h->window_n = n_timing_iterations;
  int rc = h->ops->step(h, h->d_y, h->d_u, h->stream);
  h->window_on = false;
  if (rc) return rc;
  CU(cudaMemcpyAsync(u, h->d_u, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(w[3], h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return window_ns(h, 0, time_ns);
}

int cmpc_get_step_info(cmpc_handle* h, int32_t* status, uint32_t* active, double* objective) {
  CMPC_ENTER(h);
  const size_t n = size_t(h->cfg.batch) * h->NCTRL;
  CU(cudaDeviceSynchronize());
  if (status) CU(cudaMemcpy(status, h->G.status, n * sizeof(int), cudaMemcpyDeviceToHost));
  if (active) CU(cudaMemcpy(active, h->G.active, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (objective) CU(cudaMemcpy(objective, h->G.objective, n * sizeof(double), cudaMemcpyDeviceToHost));
  return CMPC_OK;
}

int cmpc_debug_phase_ticks(cmpc_handle* h, long long* out /* B x 32 */) {
  CMPC_ENTER(h);
  if (!out) return fail(CMPC_ERR_ARG, "null argument");
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, h->G.ticks, size_t(h->cfg.batch) * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  return CMPC_OK;
}

int cmpc_launch_count(cmpc_handle* h, int64_t* n) {
  if (!h || !n) return fail(CMPC_ERR_ARG, "null argument");
  *n = h->launches;
  return CMPC_OK;
}

int cmpc_run_closed_loop_device(cmpc_handle* h, int first_step, int n_steps, int total_steps,
                                const double* x0_dev, int n_blocks, const int32_t* block_end_dev,
                                const double* block_off_dev, double* traj_dev, uint32_t* qp_active_dev,
                                double* qp_objective_dev, int32_t* qp_status_dev, void* stream) {
  CMPC_ENTER(h);
  if (first_step < 0 || n_steps < 0 || first_step + n_steps > total_steps || n_blocks < 1 ||
      !block_end_dev || !block_off_dev)
    return fail(CMPC_ERR_ARG, "bad closed-loop arguments");
  const bool reinit = first_step == 0;
  if (reinit && !x0_dev) return fail(CMPC_ERR_ARG, "x0 required to start the scenarios");
  if (!reinit && !h->loop_started)
    return fail(CMPC_ERR_STATE, "no closed loop to continue: start one with first_step = 0");
  // The plant delay rings are indexed by the record number and the trajectory slots by
  // (record, total_steps): a continuation that does not pick up exactly where the previous call
  // stopped, with the same run description, would silently desynchronise them.
  if (!reinit && (first_step != h->loop_next || total_steps != h->loop_total || n_blocks != h->loop_blocks ||
                  block_end_dev != h->loop_block_end || block_off_dev != h->loop_block_off ||
                  traj_dev != h->loop_traj))
    return fail(CMPC_ERR_STATE, "continuation does not match the running closed loop (expected first_step = " +
                                    std::to_string(h->loop_next) + " of " + std::to_string(h->loop_total) +
                                    " records and the same block / trajectory arrays)");
  ClosedLoopArrays A;
  A.x = h->d_x; A.y = h->d_y; A.u = h->d_u; A.ring = h->d_ring;
  A.block_end = block_end_dev; A.block_off = block_off_dev; A.n_blocks = n_blocks;
  A.traj = traj_dev; A.qp_active = qp_active_dev; A.qp_objective = qp_objective_dev;
  A.qp_status = qp_status_dev; A.n_steps = total_steps; A.rec_base = 0;
  A.stream_io = 0;
  A.phases = 3;
  const int rc = h->ops->closed_loop(h, first_step, n_steps, x0_dev, A, reinit,
                                                   static_cast<cudaStream_t>(stream));
  h->stream_next = -1;   // a cmpc_closed_loop_start / _step sequence does not survive a run of this kind
  if (rc == CMPC_OK) {
    h->loop_next = first_step + n_steps;
    h->loop_total = total_steps;
    h->loop_blocks = n_blocks;
    h->loop_block_end = block_end_dev;
    h->loop_block_off = block_off_dev;
    h->loop_traj = traj_dev;
  } else {
    h->loop_started = false;   // part of the launches may have gone out: the run cannot be continued
  }
  return rc;
}

int cmpc_closed_loop_start(cmpc_handle* h, const double* x0) {
  CMPC_ENTER(h);
  if (!x0) return fail(CMPC_ERR_ARG, "null argument");
  const size_t B = h->cfg.batch;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_xinit, x0, B * h->N * sizeof(double), cudaMemcpyHostToDevice));
  if (!h->d_step_end) {
    CU(dalloc(&h->d_step_end, B));
    CU(dalloc(&h->d_step_off, B * h->NIN));
    CU(dalloc(&h->d_rec, B * (1 + h->N + 8)));
    std::vector<int> never(B, 0x7fffffff);   // a single block that never ends: the offsets of a sample are
    CU(cudaMemcpy(h->d_step_end, never.data(), B * sizeof(int), cudaMemcpyHostToDevice));   // whatever the call brings
  }
  h->stream_next = 0;
  h->loop_started = false;
  h->ctrl_ahead = false;
  return CMPC_OK;
}

int cmpc_closed_loop_pipeline(cmpc_handle* h, int on) {
  CMPC_ENTER(h);
  if (h->stream_next > 0) return fail(CMPC_ERR_STATE, "the pipelining of a closed loop is chosen before its first step");
  if (on && h->generic) return fail(CMPC_ERR_UNSUPPORTED, "general configurations run the closed loop unpipelined");
  if (on && !h->ev_plant) CU(cudaEventCreateWithFlags(&h->ev_plant, cudaEventDisableTiming));
  h->stream_pipeline = on != 0;
  return CMPC_OK;
}

int cmpc_closed_loop_step(cmpc_handle* h, const double* plant_offset, double* record) {
  CMPC_ENTER(h);
  if (!plant_offset || !record) return fail(CMPC_ERR_ARG, "null argument");
  if (h->stream_next < 0) return fail(CMPC_ERR_STATE, "cmpc_closed_loop_start has not been called");
  if (h->stream_next > 0 && !h->loop_started)
    return fail(CMPC_ERR_STATE, "the controller was restarted since cmpc_closed_loop_start");
  const size_t B = h->cfg.batch, REC = 1 + h->N + 8;
  const int k = h->stream_next;
  // Page-locked host buffers that are mapped into the device (cudaHostAlloc / cudaHostRegister, e.g. pinned
  // torch tensors) are read and written by the plant kernel itself, in contiguous chunks, while the solve
  // kernel runs / the plant is integrated: no copy sits in front of the control step or behind the plant
  // advance.  Anything else goes through the handle's device buffers and two copies on the stream.
  auto mapped = [&](const void* host, size_t align) -> void* { return mapped_host(host, align); };
  const bool no_direct = h->generic || getenv("CMPC_NO_DIRECT_HOST_IO") != nullptr;   // (the env variable: A/B measurements)
  const double* off_dev = no_direct ? nullptr : static_cast<const double*>(mapped(plant_offset, 8));
  double* rec_dev = no_direct ? nullptr : static_cast<double*>(mapped(record, 16));
  const bool rec_direct = rec_dev != nullptr;
  if (!off_dev) {
    CU(cudaMemcpyAsync(h->d_step_off, plant_offset, B * h->NIN * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    off_dev = h->d_step_off;
  }
  ClosedLoopArrays A;
  A.x = h->d_x; A.y = h->d_y; A.u = h->d_u; A.ring = h->d_ring;
  A.block_end = h->d_step_end; A.block_off = off_dev; A.n_blocks = 1;
  A.traj = rec_direct ? rec_dev : h->d_rec; A.qp_active = nullptr; A.qp_objective = nullptr; A.qp_status = nullptr;
  A.n_steps = 1; A.rec_base = k;
  A.stream_io = h->generic ? 0 : 1;
  A.phases = 3;
  auto give_up = [&](int rc) {
    h->stream_next = -1;
    h->loop_started = false;
    h->ctrl_ahead = false;
    return rc;
  };
  if (!h->stream_pipeline) {
    if (int rc = h->ops->closed_loop(h, k, 1, h->d_xinit, A, k == 0, h->stream)) return give_up(rc);
    if (!rec_direct) CU(cudaMemcpyAsync(record, h->d_rec, B * REC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  } else {
    // Pipelined (cmpc_closed_loop_pipeline): a control step needs the measurement, not the plant-input offsets
    // of its record, so the control step of record k + 1 is launched right behind the plant advance of record k
    // and runs while the caller looks at record k and prepares the next offsets; this call then only adds the
    // plant advance and waits for it.
    if (!h->ctrl_ahead) {
      A.phases = 1;
      if (int rc = h->ops->closed_loop(h, k, 1, h->d_xinit, A, k == 0, h->stream)) return give_up(rc);
    }
    A.phases = 2;
    if (int rc = h->ops->closed_loop(h, k, 1, h->d_xinit, A, false, h->stream)) return give_up(rc);
    if (!rec_direct) CU(cudaMemcpyAsync(record, h->d_rec, B * REC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev_plant, h->stream));
    A.phases = 1;
    A.rec_base = k + 1;
    if (int rc = h->ops->closed_loop(h, k + 1, 1, h->d_xinit, A, false, h->stream)) return give_up(rc);
    h->ctrl_ahead = true;
    CU(cudaEventSynchronize(h->ev_plant));
  }
  h->stream_next = k + 1;
  h->loop_next = k + 1;
  h->loop_total = -1;   // not a run the device-resident variant could continue
  return CMPC_OK;
}

int cmpc_set_timing(cmpc_handle* h, int on) {
  CMPC_ENTER(h);
  h->timing = on != 0;
  h->ev_used = 0;
  return CMPC_OK;
}

int cmpc_get_timing(cmpc_handle* h, int64_t* n_steps, double* step_ms, double* assemble_ms) {
  CMPC_ENTER(h);
  CU(cudaDeviceSynchronize());
  double total = 0.0, asm_ms = 0.0;
  for (size_t i = 0; i + 3 < h->ev_used; i += 4) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 3]));
    total += ms;
    CU(cudaEventElapsedTime(&ms, h->ev[i + 1], h->ev[i + 2]));
    asm_ms += ms;
  }
  if (n_steps) *n_steps = int64_t(h->ev_used / 4);
  if (step_ms) *step_ms = total;
  if (assemble_ms) *assemble_ms = asm_ms;
  h->ev_used = 0;
  return CMPC_OK;
}

int cmpc_run_closed_loop(cmpc_handle* h, int n_steps, const double* x0, int n_blocks,
                         const int32_t* block_end, const double* block_off, double* traj,
                         uint32_t* qp_active, double* qp_objective, int32_t* qp_status) {
  return cmpc_run_closed_loop_timed(h, n_steps, x0, n_blocks, block_end, block_off, traj, qp_active, qp_objective,
                                    qp_status, -1, nullptr);
}

int cmpc_run_closed_loop_timed(cmpc_handle* h, int n_steps, const double* x0, int n_blocks,
                               const int32_t* block_end, const double* block_off, double* traj,
                               uint32_t* qp_active, double* qp_objective, int32_t* qp_status,
                               int n_timing_iterations, int64_t* step_ns) {
  CMPC_ENTER(h);
  if (!x0 || !block_end || !block_off || n_blocks < 1 || n_steps < 0)
    return fail(CMPC_ERR_ARG, "bad closed-loop arguments");
  if (step_ns)
    if (int rc = reserve_window_events(h, size_t(n_steps))) return rc;
  const size_t B = h->cfg.batch, NC = h->NCTRL, REC = 1 + h->N + 8;
  for (size_t b = 0; b < B; ++b)
    for (int i = 0; i < n_blocks; ++i) {
      const int32_t e = block_end[b * n_blocks + i];
      if (e < 0 || (i > 0 && e < block_end[b * n_blocks + i - 1]))
        return fail(CMPC_ERR_ARG, "block_end must be non-negative and non-decreasing per scenario");
    }
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_xinit, x0, B * h->N * sizeof(double), cudaMemcpyHostToDevice));
  if (h->block_cap < B * n_blocks) {
    if (h->d_block_end) cudaFree(h->d_block_end);
    if (h->d_block_off) cudaFree(h->d_block_off);
    h->d_block_end = nullptr;
    h->d_block_off = nullptr;
    h->block_cap = 0;
    CU(dalloc(&h->d_block_end, B * n_blocks));
    CU(dalloc(&h->d_block_off, B * n_blocks * h->NIN));
    h->block_cap = B * n_blocks;
  }
  CU(cudaMemcpy(h->d_block_end, block_end, B * n_blocks * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_block_off, block_off, B * n_blocks * h->NIN * sizeof(double), cudaMemcpyHostToDevice));
  DevBuf<double> d_traj, d_obj;
  DevBuf<unsigned> d_act;
  DevBuf<int> d_st;
  const size_t nrec = B * size_t(n_steps);
  if (traj) CU(d_traj.alloc(nrec * REC));
  if (qp_active) CU(d_act.alloc(nrec * NC));
  if (qp_objective) CU(d_obj.alloc(nrec * NC));
  if (qp_status) CU(d_st.alloc(nrec * NC));
  h->window_on = step_ns != nullptr;
  h->window_n = n_timing_iterations;
  const int rc_run = cmpc_run_closed_loop_device(h, 0, n_steps, n_steps, h->d_xinit, n_blocks, h->d_block_end,
                                                 h->d_block_off, d_traj.p, d_act.p, d_obj.p, d_st.p, h->stream);
  h->window_on = false;
  if (rc_run) return rc_run;
  CU(cudaStreamSynchronize(h->stream));
  for (int k = 0; step_ns && k < n_steps; ++k)
    if (int rc = window_ns(h, size_t(k), &step_ns[k])) return rc;
  if (traj) CU(d_traj.download(traj, nrec * REC));
  if (qp_active) CU(d_act.download(qp_active, nrec * NC));
  if (qp_objective) CU(d_obj.download(qp_objective, nrec * NC));
  if (qp_status) CU(d_st.download(qp_status, nrec * NC));
  // the temporaries go away with this call: the run cannot be continued through the device variant
  h->loop_traj = nullptr;
  h->loop_started = false;
  return CMPC_OK;
}

// ---- parity hooks ---------------------------------------------------------------------
int cmpc_get_linearization(cmpc_handle* h, int ctrl, double* Aorig, double* Bd, double* f) {
  CMPC_ENTER(h);
  if (!h->G.lin) return fail(CMPC_ERR_STATE, "enable cmpc_set_capture before the step");
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  const int N = h->N, rec = N * N + N * 5;
  const size_t B = h->cfg.batch;
  std::vector<double> buf(B * h->NCTRL * rec);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(buf.data(), h->G.lin, buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t b = 0; b < B; ++b) {
    const double* r = buf.data() + (b * h->NCTRL + ctrl) * rec;
    if (Aorig) std::memcpy(Aorig + b * N * N, r, sizeof(double) * N * N);
    for (int i = 0; i < N; ++i) {
      if (Bd) for (int c = 0; c < 4; ++c) Bd[(b * N + i) * 4 + c] = r[N * N + i * 5 + c];
      if (f) f[b * N + i] = r[N * N + i * 5 + 4];
    }
  }
  return CMPC_OK;
}

int cmpc_get_qp(cmpc_handle* h, int ctrl, double* H, double* f, double* Gx) {
  CMPC_ENTER(h);
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  if (h->generic) {   // H nv x nv and f nv of this sub-controller; the cross term is not kept in HBM
    if (Gx) return fail(CMPC_ERR_UNSUPPORTED, "the cross term is read back on the tuned path only");
    const size_t B = h->cfg.batch, NC = h->NCTRL, nv = h->GP.c[ctrl].nv;
    CU(cudaDeviceSynchronize());
    if (H) CU(cudaMemcpy2D(H, nv * nv * sizeof(double), h->GS.qpH + ctrl * 64, NC * 64 * sizeof(double),
                           nv * nv * sizeof(double), B, cudaMemcpyDeviceToHost));
    if (f) CU(cudaMemcpy2D(f, nv * sizeof(double), h->GS.qpf + ctrl * 8, NC * 8 * sizeof(double), nv * sizeof(double), B,
                           cudaMemcpyDeviceToHost));
    return CMPC_OK;
  }
  const size_t B = h->cfg.batch, NC = h->NCTRL, NV = h->NV, NVO = h->NVO;
  CU(cudaDeviceSynchronize());
  if (H) CU(cudaMemcpy2D(H, NV * NV * sizeof(double), h->G.qpH + ctrl * NV * NV, NC * NV * NV * sizeof(double),
                         NV * NV * sizeof(double), B, cudaMemcpyDeviceToHost));
  if (f) CU(cudaMemcpy2D(f, NV * sizeof(double), h->G.qpf + ctrl * NV, NC * NV * sizeof(double),
                         NV * sizeof(double), B, cudaMemcpyDeviceToHost));
  if (Gx && NVO > 0)
    CU(cudaMemcpy2D(Gx, NV * NVO * sizeof(double), h->G.qpG + ctrl * NV * NVO, NC * NV * NVO * sizeof(double),
                    NV * NVO * sizeof(double), B, cudaMemcpyDeviceToHost));
  return CMPC_OK;
}

int cmpc_generate_prediction(cmpc_handle* h, int ctrl, double* Su, double* Su_other) {
  CMPC_ENTER(h);
  if (!h->G.etab) return fail(CMPC_ERR_STATE, "enable cmpc_set_capture before the step");
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  // Rebuild Su / Su_other from the impulse-response table exactly as GeneratePrediction
  // accumulates them (aug_lin_sys.cc:311-328): Su[r, move0] = G_r, Su[r, move1] = sum_{k<r} G_k.
  const size_t B = h->cfg.batch;
  const int p = h->cfg.p, ny = h->cfg.n_controlled_outputs[0], nu = h->cfg.n_sub_control_inputs;
  const int no = 4 - nu, NV = h->NV, NVO = h->NVO;
  const size_t rec = size_t(p) * ny * 5;
  std::vector<double> E(B * h->NCTRL * rec);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(E.data(), h->G.etab, E.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t b = 0; b < B; ++b) {
    const double* e = E.data() + (b * h->NCTRL + ctrl) * rec;
    std::vector<double> run(size_t(ny) * 4, 0.0);
    for (int r = 0; r < p; ++r)
      for (int y = 0; y < ny; ++y)
        for (int i = 0; i < 4; ++i) {
          const bool delayed = i & 1;
          const int k = delayed ? r - kDelay : r;
          const double gval = k >= 0 ? e[(size_t(k) * ny + y) * 5 + i] : 0.0;
          const size_t row = b * size_t(p) * ny + size_t(r) * ny + y;
          if (i < nu) {
            if (Su) { Su[row * NV + i] = gval; Su[row * NV + nu + i] = run[y * 4 + i]; }
          } else if (Su_other && NVO > 0) {
            Su_other[row * NVO + (i - nu)] = gval;
            Su_other[row * NVO + no + (i - nu)] = run[y * 4 + i];
          }
          run[y * 4 + i] += gval;
        }
  }
  return CMPC_OK;
}

int cmpc_get_controller_state(cmpc_handle* h, int ctrl, double* x_hat, double* dx_aug, double* y_old,
                              double* u_old) {
  CMPC_ENTER(h);
  if (ctrl < 0 || ctrl >= h->NCTRL) return fail(CMPC_ERR_ARG, "bad controller index");
  const size_t B = h->cfg.batch;
  if (h->generic) {   // x_hat[n] | dx_aug[n_total] (the reference's state order) | y_old[4] | u_old[4]
    const size_t st = h->GP.state_stride, nt = h->GP.n_total, n = h->N;
    std::vector<double> buf(B * h->NCTRL * st);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(buf.data(), h->GS.ctrl, buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; ++b) {
      const double* r = buf.data() + (b * h->NCTRL + ctrl) * st;
      if (x_hat) std::memcpy(x_hat + b * n, r, sizeof(double) * n);
      if (dx_aug) std::memcpy(dx_aug + b * nt, r + n, sizeof(double) * nt);
      if (y_old) std::memcpy(y_old + b * 4, r + n + nt, sizeof(double) * 4);
      if (u_old) std::memcpy(u_old + b * 4, r + n + nt + 4, sizeof(double) * 4);
    }
    return CMPC_OK;
  }
  std::vector<double> buf(B * h->NCTRL * kCtrlStateStride);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(buf.data(), h->G.ctrl, buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const int N = h->N, NT = N + kNAug;
  for (size_t b = 0; b < B; ++b) {
    const double* r = buf.data() + (b * h->NCTRL + ctrl) * kCtrlStateStride;
    if (x_hat) std::memcpy(x_hat + b * N, r + kOffXhat, sizeof(double) * N);
    if (dx_aug) {
      // logical order: [x | d | heads | chain 0 | chain 1]; the chains are rings starting at ring_pos
      std::memcpy(dx_aug + b * NT, r + kOffDx, sizeof(double) * (N + kNDist + 2));
      for (int d = 0; d < 2; ++d)
        for (int j = 0; j < kRing; ++j)
          dx_aug[b * NT + N + kNDist + 2 + d * kRing + j] =
              r[kOffDx + N + kNDist + 2 + d * kRing + (h->P.ring_pos + j) % kRing];
    }
    if (y_old) std::memcpy(y_old + b * 4, r + kOffYold, sizeof(double) * 4);
    if (u_old) std::memcpy(u_old + b * 4, r + kOffUold, sizeof(double) * 4);
  }
  return CMPC_OK;
}

int cmpc_solve_qp(int device, int nq, int nv, const double* H, const double* f, const double* lb,
                  const double* ub, const double* lbA, const double* ubA, uint32_t* guess_io, double* z,
                  uint32_t* active, double* objective, int32_t* status) {
  if (nq <= 0 || (nv != 4 && nv != 8)) return fail(CMPC_ERR_ARG, "nv must be 4 or 8");
  if (!H || !f || !lb || !ub || !lbA || !ubA || !guess_io || !z || !active || !objective || !status)
    return fail(CMPC_ERR_ARG, "null argument");
  CMPC_ENTER_DEVICE(device);
  DevBuf<double> dH, df, dlb, dub, dlbA, dubA, dz, dobj;
  DevBuf<unsigned> dg, dact;
  DevBuf<int> dst;
  const size_t n = nq;
  CU(dH.upload(H, n * nv * nv)); CU(df.upload(f, n * nv)); CU(dlb.upload(lb, n * nv)); CU(dub.upload(ub, n * nv));
  CU(dlbA.upload(lbA, n * nv)); CU(dubA.upload(ubA, n * nv)); CU(dg.upload(guess_io, n));
  CU(dz.alloc(n * nv)); CU(dobj.alloc(n)); CU(dact.alloc(n)); CU(dst.alloc(n));
  if (nv == 4)
    qp_kernel<4><<<(nq + 63) / 64, 64>>>(nq, dH.p, df.p, dlb.p, dub.p, dlbA.p, dubA.p, dg.p, dz.p, dact.p, dobj.p, dst.p);
  else
    qp_kernel<8><<<(nq + 63) / 64, 64>>>(nq, dH.p, df.p, dlb.p, dub.p, dlbA.p, dubA.p, dg.p, dz.p, dact.p, dobj.p, dst.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  CU(dz.download(z, n * nv));
  CU(dg.download(guess_io, n));
  CU(dact.download(active, n));
  CU(dobj.download(objective, n));
  CU(dst.download(status, n));
  return CMPC_OK;
}

int cmpc_plant_eval(int device, int plant, int nq, const double* x, const double* u, double* dxdt,
                    double* y, double* A, double* Bc, double* C) {
  if ((plant != 0 && plant != 1) || nq <= 0 || !x || !u) return fail(CMPC_ERR_ARG, "bad argument");
  CMPC_ENTER_DEVICE(device);
  const size_t n = nq, N = plant == 0 ? 11 : 10, NIN = plant == 0 ? 9 : 8;
  DevBuf<double> dx, du, dd, dy, dA, dB, dC;
  CU(dx.upload(x, n * N)); CU(du.upload(u, n * NIN));
  CU(dd.alloc(n * N)); CU(dy.alloc(n * 4)); CU(dA.alloc(n * N * N)); CU(dB.alloc(n * N * 4)); CU(dC.alloc(n * 4 * N));
  if (plant == 0)
    plant_eval_kernel<0><<<(nq + 63) / 64, 64>>>(nq, dx.p, du.p, dd.p, dy.p, dA.p, dB.p, dC.p);
  else
    plant_eval_kernel<1><<<(nq + 63) / 64, 64>>>(nq, dx.p, du.p, dd.p, dy.p, dA.p, dB.p, dC.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  if (dxdt) CU(dd.download(dxdt, n * N));
  if (y) CU(dy.download(y, n * 4));
  if (A) CU(dA.download(A, n * N * N));
  if (Bc) CU(dB.download(Bc, n * N * 4));
  if (C) CU(dC.download(C, n * 4 * N));
  return CMPC_OK;
}

int cmpc_plant_integrate(int device, int plant, int nq, double* x, const double* u, double Ts,
                         int32_t* n_substeps) {
  if ((plant != 0 && plant != 1) || nq <= 0 || !x || !u) return fail(CMPC_ERR_ARG, "bad argument");
  CMPC_ENTER_DEVICE(device);
  const size_t n = nq, N = plant == 0 ? 11 : 10, NIN = plant == 0 ? 9 : 8;
  DevBuf<double> dx, du;
  DevBuf<int> ds;
  CU(dx.upload(x, n * N)); CU(du.upload(u, n * NIN)); CU(ds.alloc(n));
  if (plant == 0)
    plant_integrate_kernel<0><<<(nq + 3) / 4, 128>>>(nq, dx.p, du.p, Ts, ds.p);
  else
    plant_integrate_kernel<1><<<(nq + 3) / 4, 128>>>(nq, dx.p, du.p, Ts, ds.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  CU(dx.download(x, n * N));
  if (n_substeps) CU(ds.download(n_substeps, n));
  return CMPC_OK;
}

int cmpc_inrange_math(int device, int n, const double* a, const double* b, double* sqrt_fast, double* sqrt_std,
                      double* div_fast, double* div_std, int32_t* flagged) {
  if (n <= 0 || !a || !b || !sqrt_fast || !sqrt_std || !div_fast || !div_std || !flagged)
    return fail(CMPC_ERR_ARG, "bad argument");
  CMPC_ENTER_DEVICE(device);
  DevBuf<double> da, db, d1, d2, d3, d4;
  DevBuf<int> df;
  CU(da.upload(a, n)); CU(db.upload(b, n));
  CU(d1.alloc(n)); CU(d2.alloc(n)); CU(d3.alloc(n)); CU(d4.alloc(n)); CU(df.alloc(n));
  inrange_math_kernel<<<(n + 127) / 128, 128>>>(n, da.p, db.p, d1.p, d2.p, d3.p, d4.p, df.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  CU(d1.download(sqrt_fast, n)); CU(d2.download(sqrt_std, n)); CU(d3.download(div_fast, n)); CU(d4.download(div_std, n));
  CU(df.download(flagged, n));
  return CMPC_OK;
}

}  // extern "C"
