// Plant-side kernels of the closed loop: one GPU thread per scenario.
//   SimulationSystem::{SetOffset,SetInput,operator(),Integrate}  include/simulation_system.h:66-116
//   (Dormand-Prince 5(4) with odeint's controlled stepper, rel/abs 1e-6, error measured in the
//    2-norm the reference installs at simulation_system.h:119-133)
//   TimeDelay::GetDelayedInput                                     include/time_delay.h:41-58
#pragma once
#include <cuda_runtime.h>

#include "plant_dev.cuh"
#include "step_kernel.cuh"

namespace cmpc {

// One try_step of the controlled Dormand-Prince stepper (FSAL).  true: accepted.
// odeint's integrate_adaptive puts no bound on the number of accepted steps: a plant state that runs
// away (unphysical inputs) makes it take ever smaller steps for ever.  A kernel must not do that:
// an interval stops after this many accepted steps (the nominal count is 1-4) and leaves the state
// where it is; the records of such a scenario are no longer meaningful.
constexpr int kMaxStepsPerInterval = 4000;

template <int PLANT>
__device__ bool dopri5_try_step(const double* u, double* x, double* dxdt, double* t, double* dt) {
  constexpr int N = PlantDims<PLANT>::N;
  constexpr double b21 = 1.0 / 5;
  constexpr double b31 = 3.0 / 40, b32 = 9.0 / 40;
  constexpr double b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9;
  constexpr double b51 = 19372.0 / 6561, b52 = -25360.0 / 2187, b53 = 64448.0 / 6561, b54 = -212.0 / 729;
  constexpr double b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247, b64 = 49.0 / 176,
                   b65 = -5103.0 / 18656;
  constexpr double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  constexpr double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                   dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  const double h = *dt;
  double k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], xt[N], xn[N];
  const double* k1 = dxdt;
#pragma unroll
  for (int i = 0; i < N; ++i) xt[i] = x[i] + h * b21 * k1[i];
  plant_derivative<PLANT>(xt, u, k2);
#pragma unroll
  for (int i = 0; i < N; ++i) xt[i] = x[i] + h * (b31 * k1[i] + b32 * k2[i]);
  plant_derivative<PLANT>(xt, u, k3);
#pragma unroll
  for (int i = 0; i < N; ++i) xt[i] = x[i] + h * (b41 * k1[i] + b42 * k2[i] + b43 * k3[i]);
  plant_derivative<PLANT>(xt, u, k4);
#pragma unroll
  for (int i = 0; i < N; ++i) xt[i] = x[i] + h * (b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i]);
  plant_derivative<PLANT>(xt, u, k5);
#pragma unroll
  for (int i = 0; i < N; ++i)
    xt[i] = x[i] + h * (b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i]);
  plant_derivative<PLANT>(xt, u, k6);
#pragma unroll
  for (int i = 0; i < N; ++i)
    xn[i] = x[i] + h * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
  plant_derivative<PLANT>(xn, u, k7);
  double sumsq = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xerr = h * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
    const double e = fabs(xerr) / (1e-6 + 1e-6 * (fabs(x[i]) + fabs(h) * fabs(k1[i])));
    sumsq += e * e;
  }
  double err = sqrt(sumsq);
  if (err > 1.0) {
    *dt = h * fmax(0.9 * pow(err, -1.0 / 3.0), 0.2);
    return false;
  }
  *t += h;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i] = xn[i];
    dxdt[i] = k7[i];
  }
  if (err < 0.5) {
    err = fmax(1.0 / 3125.0, err);
    *dt = h * 9.0 / 10.0 * pow(err, -1.0 / 5.0);
  }
  return true;
}

// integrate_adaptive over [0, Ts] starting with dt = Ts; returns accepted steps (-1: stuck).
template <int PLANT>
__device__ int integrate_interval(const double* u, double* x, double Ts) {
  constexpr int N = PlantDims<PLANT>::N;
  double dxdt[N];
  plant_derivative<PLANT>(x, u, dxdt);
  double t = 0.0, dt = Ts;
  int steps = 0, fails = 0;
  const double eps = 2.220446049250313e-16;
  while (Ts - t > eps && steps < kMaxStepsPerInterval) {
    if ((t + dt) - Ts > eps) dt = Ts - t;
    while (!dopri5_try_step<PLANT>(u, x, dxdt, &t, &dt)) {
      if (++fails > 500) return -1;
    }
    fails = 0;
    ++steps;
  }
  return steps;   // kMaxStepsPerInterval: the interval was cut short (runaway plant state)
}

// ---- lane-pair Dormand-Prince (same arithmetic per state as dopri5_try_step, same order of the
// error-norm sum, so step acceptance is bit-identical to the one-thread form) --------------------
template <int PLANT>
__device__ bool dopri5_try_step_pair(unsigned full, int c, const double uc[4], double u_tank, double xs[6],
                                     double k1[6], double* t, double* dt) {
  constexpr int NS = PLANT == 0 ? 6 : 5;
  constexpr double b21 = 1.0 / 5;
  constexpr double b31 = 3.0 / 40, b32 = 9.0 / 40;
  constexpr double b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9;
  constexpr double b51 = 19372.0 / 6561, b52 = -25360.0 / 2187, b53 = 64448.0 / 6561, b54 = -212.0 / 729;
  constexpr double b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247, b64 = 49.0 / 176,
                   b65 = -5103.0 / 18656;
  constexpr double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  constexpr double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                   dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  const double h = *dt;
  double k2[6], k3[6], k4[6], k5[6], k6[6], k7[6], xt[6], xn[6];
  xt[5] = xn[5] = 0.0;
#pragma unroll
  for (int i = 0; i < NS; ++i) xt[i] = xs[i] + h * b21 * k1[i];
  pair_derivative<PLANT>(full, c, xt, uc, u_tank, k2);
#pragma unroll
  for (int i = 0; i < NS; ++i) xt[i] = xs[i] + h * (b31 * k1[i] + b32 * k2[i]);
  pair_derivative<PLANT>(full, c, xt, uc, u_tank, k3);
#pragma unroll
  for (int i = 0; i < NS; ++i) xt[i] = xs[i] + h * (b41 * k1[i] + b42 * k2[i] + b43 * k3[i]);
  pair_derivative<PLANT>(full, c, xt, uc, u_tank, k4);
#pragma unroll
  for (int i = 0; i < NS; ++i) xt[i] = xs[i] + h * (b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i]);
  pair_derivative<PLANT>(full, c, xt, uc, u_tank, k5);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    xt[i] = xs[i] + h * (b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i]);
  pair_derivative<PLANT>(full, c, xt, uc, u_tank, k6);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    xn[i] = xs[i] + h * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
  pair_derivative<PLANT>(full, c, xn, uc, u_tank, k7);
  // squared scaled errors of this lane's states; summed in plant state order 0..N-1
  double e2[6];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const double xerr = h * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
    const double e = fabs(xerr) / (1e-6 + 1e-6 * (fabs(xs[i]) + fabs(h) * fabs(k1[i])));
    e2[i] = e * e;
  }
  double sumsq = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {   // compressor 0's states
    const double o = __shfl_xor_sync(full, e2[i], 1);
    sumsq += (c == 0) ? e2[i] : o;
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {   // compressor 1's states
    const double o = __shfl_xor_sync(full, e2[i], 1);
    sumsq += (c == 1) ? e2[i] : o;
  }
  if (PLANT == 0) sumsq += e2[5];
  double err = sqrt(sumsq);
  if (err > 1.0) {
    *dt = h * fmax(0.9 * pow(err, -1.0 / 3.0), 0.2);
    return false;
  }
  *t += h;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    xs[i] = xn[i];
    k1[i] = k7[i];
  }
  if (err < 0.5) {
    err = fmax(1.0 / 3125.0, err);
    *dt = h * 9.0 / 10.0 * pow(err, -1.0 / 5.0);
  }
  return true;
}

template <int PLANT>
__device__ int integrate_interval_pair(unsigned full, int c, const double uc[4], double u_tank, double xs[6],
                                       double Ts) {
  double k1[6];
  pair_derivative<PLANT>(full, c, xs, uc, u_tank, k1);
  double t = 0.0, dt = Ts;
  int steps = 0, fails = 0;
  const double eps = 2.220446049250313e-16;
  while (Ts - t > eps && steps < kMaxStepsPerInterval) {
    if ((t + dt) - Ts > eps) dt = Ts - t;
    while (!dopri5_try_step_pair<PLANT>(full, c, uc, u_tank, xs, k1, &t, &dt)) {
      if (++fails > 500) return -1;
    }
    fails = 0;
    ++steps;
  }
  return steps;   // kMaxStepsPerInterval: the interval was cut short (runaway plant state)
}

struct ClosedLoopArrays {
  double* x;           // [B][N] plant state
  double* y;           // [B][4] measurement handed to the controller
  double* u;           // [B][4] controller output
  double* ring;        // [B][2][kDelay] actuator delay rings (plant side)
  const int* block_end;     // [B][n_blocks]
  const double* block_off;  // [B][n_blocks][NIN]
  int n_blocks;
  double* traj;        // [B][T][1+N+8] or null
  unsigned* qp_active; // [B][T][NCTRL] or null
  double* qp_objective;
  int* qp_status;
  int n_steps;         // records per scenario in traj / qp_* ...
  int rec_base;        // ... whose slot 0 is record rec_base of the run
};

// Start of a closed-loop run: x = x0, y = GetOutput(x0), rings = 0.
template <int PLANT>
__global__ void cl_start_kernel(int B, const double* __restrict__ x0, ClosedLoopArrays A,
                                double* u_init, double* u_init_full) {
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[N], y[4];
  for (int i = 0; i < N; ++i) {
    x[i] = x0[size_t(b) * N + i];
    A.x[size_t(b) * N + i] = x[i];
  }
  plant_output<PLANT>(x, y);
  for (int i = 0; i < 4; ++i) {
    A.y[size_t(b) * 4 + i] = y[i];
    u_init[size_t(b) * 4 + i] = 0.0;
  }
  for (int i = 0; i < 2 * kDelay; ++i) A.ring[size_t(b) * 2 * kDelay + i] = 0.0;
  constexpr double udef_par[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  constexpr double udef_ser[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  for (int i = 0; i < NIN; ++i) u_init_full[size_t(b) * NIN + i] = PLANT == 0 ? udef_par[i] : udef_ser[i];
}

// ---- plant side of the closed loop: one lane pair per scenario (16 scenarios per warp) ---------
// After the control step of record k: write the record, push u through the
// actuator delay rings, integrate the plant over one sampling interval (one compressor per lane
// of the pair) and produce the next measurement.
template <int PLANT, int NCTRL>
__device__ __forceinline__ void advance_pair(int b, int c, unsigned pair_mask, int k, double t_k, double Ts,
                                             const ClosedLoopArrays& A, const int* __restrict__ status,
                                             const unsigned* __restrict__ active,
                                             const double* __restrict__ objective, double (&y_next)[4]) {
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN, REC = 1 + N + 8;
  double u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = A.u[size_t(b) * 4 + i];
  double xs[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) xs[i] = A.x[size_t(b) * N + 5 * c + i];
  xs[5] = PLANT == 0 ? A.x[size_t(b) * N + (PLANT == 0 ? 10 : 0)] : 0.0;
  if (A.traj) {
    double* r = A.traj + (size_t(b) * A.n_steps + (k - A.rec_base)) * REC;
#pragma unroll
    for (int i = 0; i < 5; ++i) r[1 + 5 * c + i] = xs[i];
    if (c == 0) {
      r[0] = t_k;
      if (PLANT == 0) r[1 + (PLANT == 0 ? 10 : 0)] = xs[5];
#pragma unroll
      for (int i = 0; i < 4; ++i) r[1 + N + i] = u[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[1 + N + 4 + i] = A.y[size_t(b) * 4 + i];
    }
  }
  if (c < NCTRL) {
    const size_t o = (size_t(b) * A.n_steps + (k - A.rec_base)) * NCTRL + c;
    if (A.qp_active) A.qp_active[o] = active[b * NCTRL + c];
    if (A.qp_objective) A.qp_objective[o] = objective[b * NCTRL + c];
    if (A.qp_status) A.qp_status[o] = status[b * NCTRL + c];
  }
  // plant-input offsets of the block this record belongs to (SetOffset); TimeDelay: this
  // compressor's recycle valve command (control input 2c+1) comes out 40 samples late
  int blk = 0;
  while (blk + 1 < A.n_blocks && k >= A.block_end[b * A.n_blocks + blk]) ++blk;
  const int pos = k % kDelay;
  double* ring = A.ring + (size_t(b) * 2 + c) * kDelay;
  const double ud = ring[pos];
  ring[pos] = u[2 * c + 1];
  constexpr double udef_par[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  constexpr double udef_ser[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  const double* off = A.block_off + (size_t(b) * A.n_blocks + blk) * NIN;
  double uc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) uc[i] = (PLANT == 0 ? udef_par[4 * c + i] : udef_ser[4 * c + i]) + off[4 * c + i];
  uc[0] += u[2 * c];
  uc[3] += ud;
  const double u_tank = PLANT == 0 ? udef_par[8] + off[PLANT == 0 ? 8 : 0] : 0.0;
  integrate_interval_pair<PLANT>(pair_mask, c, uc, u_tank, xs, Ts);
  // next measurement: the other compressor's pressures and flow come by shuffle
  double p2, sd;
  compressor_output(xs, &p2, &sd);
  const double p2_o = __shfl_xor_sync(pair_mask, p2, 1), sd_o = __shfl_xor_sync(pair_mask, sd, 1);
#pragma unroll
  for (int i = 0; i < 5; ++i) A.x[size_t(b) * N + 5 * c + i] = xs[i];
  double y[4] = {0.0, 0.0, 0.0, 0.0};
  if (c == 0) {
    if (PLANT == 0) {
      y[0] = sd; y[1] = sd_o; y[2] = p2 - p2_o; y[3] = xs[5];
      A.x[size_t(b) * N + (PLANT == 0 ? 10 : 0)] = xs[5];
    } else {
      y[0] = p2; y[1] = sd; y[2] = p2_o; y[3] = sd_o;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) A.y[size_t(b) * 4 + i] = y[i];
  }
  // both lanes leave with the measurement (lane 0 of the pair formed it)
#pragma unroll
  for (int i = 0; i < 4; ++i) y_next[i] = __shfl_sync(pair_mask, y[i], threadIdx.x & 30);
}

// Plant side of closed-loop record k, 16 scenarios per 128-thread block.  Warp 0: one lane pair per
// scenario, lane = compressor (record, delay rings, Dormand-Prince over [t_k, t_k + Ts], next
// measurement).  With lin_next the block also does lin_kernel's work for record k + 1, so that the
// next launch of the loop is assemble_kernel again: warps 1-3 fill the linearisation parts, one
// (scenario, sub-controller, part) per lane -- at the same time as the integration when the
// observer gain leaves the state estimate alone (P.obs_states_free: the linearisation point then
// does not depend on the measurement being produced), after it otherwise -- and finally warp 0
// stores the observer state with the new measurement.
// MINB: resident blocks per SM the register allocation aims at.  1 (195 registers, 2 blocks per SM)
// is the fastest single wave, which is what a batch of a few thousand scenarios is; batches of many
// waves (the 65 536-scenario sweep: 14 waves) are bound by how many integrations an SM holds at once.
template <class S, int MINB>
__global__ void __launch_bounds__(128, MINB)
cl_advance_kernel(StepParams P, DeviceState G, int k, double t_k, double Ts, ClosedLoopArrays A, bool lin_next) {
  pdl_wait();
  pdl_trigger();   // single wave
  constexpr int kScen = 16, kItems = kScen * S::NCTRL * 3;
  static_assert(kItems <= 96, "one linearisation item per lane of warps 1-3");
  __shared__ double y_sh[kScen][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * kScen;
  const bool early = lin_next && P.obs_states_free;
  // this lane's linearisation item (warps 1-3): scenario slot, sub-controller, part
  const int idx = (warp - 1) * 32 + lane;
  const int sl = idx / (3 * S::NCTRL), g = (idx % (3 * S::NCTRL)) / 3, part = idx % 3;
  const bool item = warp != 0 && idx < kItems && b0 + sl < P.batch;
  ObsPending<S> pend;
  if (warp == 0) {
    const int b = b0 + (lane >> 1), c = lane & 1;
    if (b < P.batch) {
      double y[4];
      advance_pair<S::PLANT, S::NCTRL>(b, c, 3u << (lane & 30), k, t_k, Ts, A, G.status, G.active, G.objective, y);
      if (c == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) y_sh[lane >> 1][i] = y[i];
      }
    }
  } else if (early && item) {
    lin_part_early<S>(P, G, b0 + sl, g, part, pend);
  }
  if (!lin_next) return;
  __syncthreads();   // the measurements are in y_sh; with `early`, every part has read the old observer state
  double y[4] = {0.0, 0.0, 0.0, 0.0};
  if (item) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = y_sh[sl][i];
  }
  if (early) {
    if (item && part == 0) lin_finish<S>(P, G, b0 + sl, g, pend, y);
    return;
  }
  if (item) lin_part<S>(P, G, b0 + sl, g, part, y, 0u);
  __syncthreads();   // every part has read the old observer state
  if (item && part == 0) lin_part<S>(P, G, b0 + sl, g, 3, y, 0u);
}

// NerveCenter::Initialize + DistributedController::Initialize (nerve_center.h:98-104,186-203,
// distributed_controller.cc:27-67): x_hat = x_init, dx_aug = 0, y_old = y_init, u_old = permuted
// u_init, no QP warm start; NerveCenter's own u_old_/du_old_ start at zero (nerve_center.h:91-93).
template <class S>
__global__ void init_kernel(int B, DeviceState G, const double* __restrict__ x_init,
                            const double* __restrict__ u_init, const double* __restrict__ u_init_full,
                            const double* __restrict__ y_init, StepParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int c = 0; c < S::NCTRL; ++c) {
    double* gs = G.ctrl + (size_t(b) * S::NCTRL + c) * kCtrlStateStride;
    for (int i = 0; i < kCtrlStateStride; ++i) gs[i] = 0.0;
    for (int i = 0; i < S::N; ++i) gs[kOffXhat + i] = x_init[size_t(b) * S::N + i];
    for (int i = 0; i < 4; ++i) {
      gs[kOffYold + i] = y_init[size_t(b) * 4 + i];
      gs[kOffUold + i] = u_init[size_t(b) * 4 + P.c[c].ctrl_idx[i]];
    }
    G.guess[size_t(b) * S::NCTRL + c] = kQpNoGuess;
    G.status[size_t(b) * S::NCTRL + c] = 0;
    G.active[size_t(b) * S::NCTRL + c] = 0;
    G.objective[size_t(b) * S::NCTRL + c] = 0.0;
  }
  for (int i = 0; i < kScenStateStride; ++i) G.scen[size_t(b) * kScenStateStride + i] = 0.0;
  for (int i = 0; i < S::NIN; ++i) G.u_offset[size_t(b) * S::NIN + i] = u_init_full[size_t(b) * S::NIN + i];
}

// Stand-alone plant evaluation / integration / QP kernels behind the parity hooks.
template <int PLANT>
__global__ void plant_eval_kernel(int nq, const double* __restrict__ x, const double* __restrict__ u,
                                  double* dxdt, double* y, double* A, double* Bc, double* C) {
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nq) return;
  double xl[N], ul[NIN], Al[N * N], Bl[N * 4], Cl[4 * N], fl[N], yl[4];
  for (int i = 0; i < N; ++i) xl[i] = x[size_t(b) * N + i];
  for (int i = 0; i < NIN; ++i) ul[i] = u[size_t(b) * NIN + i];
  plant_linearize<PLANT>(xl, ul, Al, Bl, Cl, fl);
  plant_output<PLANT>(xl, yl);
  if (dxdt) for (int i = 0; i < N; ++i) dxdt[size_t(b) * N + i] = fl[i];
  if (y) for (int i = 0; i < 4; ++i) y[size_t(b) * 4 + i] = yl[i];
  if (A) for (int i = 0; i < N * N; ++i) A[size_t(b) * N * N + i] = Al[i];
  if (Bc) for (int i = 0; i < N * 4; ++i) Bc[size_t(b) * N * 4 + i] = Bl[i];
  if (C) for (int i = 0; i < 4 * N; ++i) C[size_t(b) * 4 * N + i] = Cl[i];
}

// One warp per entry; the lane pairs of the warp all integrate the same scenario.
template <int PLANT>
__global__ void plant_integrate_kernel(int nq, double* x, const double* __restrict__ u, double Ts,
                                       int* n_substeps) {
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, ln = threadIdx.x & 31, c = ln & 1;
  if (b >= nq) return;
  double xs[6], uc[4];
  for (int i = 0; i < 5; ++i) xs[i] = x[size_t(b) * N + 5 * c + i];
  xs[5] = PLANT == 0 ? x[size_t(b) * N + 10] : 0.0;
  for (int i = 0; i < 4; ++i) uc[i] = u[size_t(b) * NIN + 4 * c + i];
  const double u_tank = PLANT == 0 ? u[size_t(b) * NIN + 8] : 0.0;
  const int s = integrate_interval_pair<PLANT>(0xffffffffu, c, uc, u_tank, xs, Ts);
  if (ln < 2) {
    for (int i = 0; i < 5; ++i) x[size_t(b) * N + 5 * c + i] = xs[i];
    if (PLANT == 0 && ln == 0) x[size_t(b) * N + 10] = xs[5];
  }
  if (n_substeps && ln == 0) n_substeps[b] = s;
}

template <int NV>
__global__ void qp_kernel(int nq, const double* __restrict__ H, const double* __restrict__ f,
                          const double* __restrict__ lb, const double* __restrict__ ub,
                          const double* __restrict__ lbA, const double* __restrict__ ubA,
                          unsigned* guess_io, double* z, unsigned* active, double* objective,
                          int* status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nq) return;
  QpData<NV> qd;
  double Hl[NV * NV], fl[NV], zl[NV];
  for (int i = 0; i < NV * NV; ++i) Hl[i] = H[size_t(b) * NV * NV + i];
  for (int i = 0; i < NV; ++i) {
    fl[i] = f[size_t(b) * NV + i];
    qd.lb[i] = lb[size_t(b) * NV + i];
    qd.ub[i] = ub[size_t(b) * NV + i];
    qd.lbA[i] = lbA[size_t(b) * NV + i];
    qd.ubA[i] = ubA[size_t(b) * NV + i];
  }
  unsigned g = guess_io[b], act = 0;
  double obj = 0.0;
  int st = 3;
  if (qp_invert_spd<NV>(Hl, qd.J)) {
    st = qp_solve<NV, NV / 2>(qd, Hl, fl, &g, zl, &act, &obj);
  } else {
    for (int i = 0; i < NV; ++i) zl[i] = 0.0;
  }
  for (int i = 0; i < NV; ++i) z[size_t(b) * NV + i] = zl[i];
  guess_io[b] = g;
  active[b] = act;
  objective[b] = obj;
  status[b] = st;
}

}  // namespace cmpc
