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

// ---- lane-pair Dormand-Prince (same arithmetic per state as dopri5_try_step up to the order of the
// stage sums, same order of the error-norm sum) --------------------------------------------------
// The plant kernel runs this code once per record on one warp per block, so it is bound by
// instruction fetch, not by arithmetic: the seven derivative evaluations of a step are ONE copy of
// the derivative inside a loop over the stages (the first pass pays the instruction-cache misses,
// the other six run from the cache), with the stage derivatives k_1..k_7 in shared memory, where a
// run-time stage index costs nothing (in registers it would need a select per stage and state).
// Butcher tableau by rows: row s (1..5) forms the argument of stage s + 1, row 6 is the 5th-order
// solution, whose derivative is k_7 (FSAL); kDopriE = 5th-order weights minus the embedded 4th-order ones.
static __constant__ double kDopriA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
static __constant__ double kDopriE[7] = {35.0 / 384 - 5179.0 / 57600, 0, 500.0 / 1113 - 7571.0 / 16695,
                                         125.0 / 192 - 393.0 / 640, -2187.0 / 6784 - (-92097.0 / 339200),
                                         11.0 / 84 - 187.0 / 2100, -1.0 / 40};
constexpr int kDopriSlots = 7 * 6;   // doubles per lane in the stage buffer
constexpr int kDopriLanes = 32;      // lane stride of the stage buffer: ks[(stage * 6 + state) * 32 + lane]

// Stages s_begin..6 of one step of size h from xs: k_s = f(xs + h sum_j A[s][j] k_j) into the stage
// buffer (stage 0 is k_1 = f(xs) itself).  Leaves the argument of the last stage -- the 5th-order
// solution -- in xn.
template <int PLANT>
__device__ __forceinline__ void dopri5_stages(int c, const double uc[4], double u_tank, const double xs[6],
                                              double h, double* ks, int s_begin, double xn[6],
                                              long long* tk = nullptr, long long tk0 = 0) {
  constexpr int NS = PLANT == 0 ? 6 : 5;
#pragma unroll 1
  for (int s = s_begin; s < 7; ++s) {
#ifdef CMPC_PHASE_TIMING
    if (tk && c == 0 && s <= 2) tk[25 + s] = clock64() - tk0;
#endif
    // all six columns with the zeros of the tableau (a stage not yet evaluated holds a finite leftover of
    // the previous step, or the zeros planted by the caller): no trip count, every load in flight at once
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double a = kDopriA[s][j];
#pragma unroll
      for (int i = 0; i < NS; ++i) acc[i] = fma(a, ks[(j * 6 + i) * kDopriLanes], acc[i]);
    }
    xn[5] = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) xn[i] = fma(h, acc[i], xs[i]);
    double d[6];
    pair_derivative<PLANT>(c, xn, uc, u_tank, d);
#pragma unroll
    for (int i = 0; i < NS; ++i) ks[(s * 6 + i) * kDopriLanes] = d[i];
  }
}

// One try_step of the controlled stepper (FSAL: stage 0 of the buffer holds f(xs) on entry and on exit).
// Every lane of the warp runs it (the shuffles name the whole warp: a shuffle whose mask differs from
// lane to lane -- one mask per lane pair -- is executed group by group through WARPSYNC.COLLECTIVE and
// costs more than the arithmetic between two of them); a lane with go == false computes along and
// leaves its state alone.
template <int PLANT>
__device__ __forceinline__ bool dopri5_try_step_pair(bool go, int c, const double uc[4], double u_tank, double xs[6],
                                                     double* ks, int s_begin, double* t, double* dt, double t_end,
                                                     long long* tk = nullptr, long long tk0 = 0) {
  constexpr int NS = PLANT == 0 ? 6 : 5;
  const double h = *dt;
  double xn[6];
  dopri5_stages<PLANT>(c, uc, u_tank, xs, h, ks, s_begin, xn, tk, tk0);
#ifdef CMPC_PHASE_TIMING
  if (tk && c == 0) tk[28] = clock64() - tk0 + (xn[0] != xn[0] ? 1 : 0);
#endif
  // squared scaled errors of this lane's states; summed in plant state order 0..N-1
  // (straight-line divisions and root, redone with the standard ones if an operand was out of range)
  double e2[6], num[6], den[6];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double es = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j != 1) es = fma(kDopriE[j], ks[(j * 6 + i) * kDopriLanes], es);
    num[i] = fabs(h * es);
    den[i] = 1e-6 + 1e-6 * (fabs(xs[i]) + fabs(h) * fabs(ks[i * kDopriLanes]));
#ifdef CMPC_NO_FAST_ERR
    const double e = num[i] / den[i];
#else
    const double e = div_inrange(num[i], den[i], bad);
#endif
    e2[i] = e * e;
  }
#ifndef CMPC_NO_FAST_ERR
  if (bad) {   // this lane's own six quotients: no other lane is involved
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const double e = num[i] / den[i];
      e2[i] = e * e;
    }
  }
#endif
  // sum in plant state order (compressor 0's states, compressor 1's, tank), as the one-thread form does:
  // lane 0 of the pair sums its five, lane 1 continues from there and hands the total back
  double part = e2[0];
#pragma unroll
  for (int i = 1; i < 5; ++i) part += e2[i];
  const double first = __shfl_xor_sync(0xffffffffu, part, 1);   // (lane 1 receives compressor 0's sum)
  double tot = first;
#pragma unroll
  for (int i = 0; i < 5; ++i) tot += e2[i];
  const double both = __shfl_xor_sync(0xffffffffu, tot, 1);      // (lane 0 receives the sum of all ten)
  double sumsq = (c == 0) ? both : tot;
  if (PLANT == 0) sumsq += e2[5];
#ifdef CMPC_NO_FAST_ERR
  double err = sqrt(sumsq);
#else
  bool bad_s = false;
  double err = sqrt_inrange(sumsq, bad_s);
  if (bad_s) err = sqrt(sumsq);
#endif
#ifdef CMPC_PHASE_TIMING
  if (tk && c == 0) tk[29] = clock64() - tk0 + (err != err ? 1 : 0);
#endif
  if (!go) return false;
  if (err > 1.0) {
    *dt = h * fmax(0.9 * pow(err, -1.0 / 3.0), 0.2);
    return false;
  }
  *t += h;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    xs[i] = xn[i];
    ks[i * kDopriLanes] = ks[(6 * 6 + i) * kDopriLanes];
  }
  // the step size that odeint proposes next is only worked out when another step follows (the pow is a
  // long dependent chain at the end of the interval's critical path)
  if (err < 0.5 && t_end - *t > 2.220446049250313e-16) {
    err = fmax(1.0 / 3125.0, err);
    *dt = h * 9.0 / 10.0 * pow(err, -1.0 / 5.0);
  }
  return true;
}

// ks: this lane's column of a kDopriSlots x kDopriLanes stage buffer in shared memory.  Called by all
// 32 lanes of a warp (live == false: a lane without a scenario); the loop runs until no lane has a
// step left, so that the warp stays converged for the shuffles inside.
template <int PLANT>
__device__ __forceinline__ int integrate_interval_pair(bool live, int c, const double uc[4], double u_tank, double xs[6],
                                                       double Ts, double* ks, long long* tk = nullptr, long long tk0 = 0) {
  double t = 0.0, dt = Ts;
  int steps = 0, fails = 0, s_begin = 0;   // the first try also evaluates k_1 = f(xs)
#pragma unroll
  for (int i = 0; i < 6 * 6; ++i) ks[i * kDopriLanes] = 0.0;   // 0 * leftover must not be 0 * NaN
  const double eps = 2.220446049250313e-16;
  for (;;) {
    const bool go = live && Ts - t > eps && steps < kMaxStepsPerInterval;
    if (!__any_sync(0xffffffffu, go)) break;
    if (go && (t + dt) - Ts > eps) dt = Ts - t;
    const bool ok = dopri5_try_step_pair<PLANT>(go, c, uc, u_tank, xs, ks, s_begin, &t, &dt, Ts, tk, tk0);
    s_begin = 1;
    if (go) {
      if (ok) {
        fails = 0;
        ++steps;
      } else if (++fails > 500) {
        live = false;
        steps = -1;
      }
    }
  }
  return steps;   // kMaxStepsPerInterval: the interval was cut short (runaway plant state); -1: stuck
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
  int phases;          // launch_closed_loop: bit 0 the control step of each record, bit 1 its plant advance (3: both)
  int stream_io;       // 1: one record per launch with n_blocks == 1 (cmpc_closed_loop_step): block_off and traj may be
                       // page-locked HOST memory mapped into the device, so a block moves its scenarios' offsets and
                       // record rows through shared memory as whole 16-byte pieces of contiguous chunks
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
// The work is split at the programmatic-launch wait.  advance_pre touches only what the solve kernel
// of this record does not write (plant state, delay rings, input offsets, the measurement already
// taken: all left by the previous record's plant kernel, which had completed before this record's
// assemble kernel started) and runs while the solve kernel is still busy; advance_post needs the inputs.
template <int PLANT>
struct AdvancePre {
  double xs[6], uc[4], u_tank, ud;
  double* ring_slot;
  double* rec;   // this record's trajectory row, or null
};

// (valid == false: a lane past the end of the batch; it works on the batch's last scenario so that
// the warp stays whole, and stores nothing)
template <int PLANT>
__device__ __forceinline__ void advance_pre(int b, int c, bool valid, int k, double t_k, const ClosedLoopArrays& A,
                                            AdvancePre<PLANT>& S, const double* off_sh, double* rec_sh) {
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN, REC = 1 + N + 8;
#pragma unroll
  for (int i = 0; i < 5; ++i) S.xs[i] = A.x[size_t(b) * N + 5 * c + i];
  S.xs[5] = PLANT == 0 ? A.x[size_t(b) * N + (PLANT == 0 ? 10 : 0)] : 0.0;
  // plant-input offsets of the block this record belongs to (SetOffset); TimeDelay: this
  // compressor's recycle valve command (control input 2c+1) comes out 40 samples late
  int blk = 0;
  if (!off_sh)
    while (blk + 1 < A.n_blocks && k >= A.block_end[b * A.n_blocks + blk]) ++blk;
  S.ring_slot = A.ring + (size_t(b) * 2 + c) * kDelay + k % kDelay;
  S.ud = *S.ring_slot;
  constexpr double udef_par[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  constexpr double udef_ser[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  const double* off = off_sh ? off_sh : A.block_off + (size_t(b) * A.n_blocks + blk) * NIN;   // (off_sh: this scenario's row)
#pragma unroll
  for (int i = 0; i < 4; ++i) S.uc[i] = (PLANT == 0 ? udef_par[4 * c + i] : udef_ser[4 * c + i]) + off[4 * c + i];
  S.uc[3] += S.ud;
  S.u_tank = PLANT == 0 ? udef_par[8] + off[PLANT == 0 ? 8 : 0] : 0.0;
  S.rec = nullptr;
  if (A.traj && valid) {
    double* r = rec_sh ? rec_sh : A.traj + (size_t(b) * A.n_steps + (k - A.rec_base)) * REC;   // (rec_sh: this scenario's row)
    S.rec = r;
#pragma unroll
    for (int i = 0; i < 5; ++i) r[1 + 5 * c + i] = S.xs[i];
    if (c == 0) {
      r[0] = t_k;
      if (PLANT == 0) r[1 + (PLANT == 0 ? 10 : 0)] = S.xs[5];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[1 + N + 4 + i] = A.y[size_t(b) * 4 + i];
    }
  }
}

template <int PLANT, int NCTRL>
__device__ __forceinline__ void advance_post(int b, int c, bool valid, int k, double Ts, AdvancePre<PLANT>& S,
                                             const ClosedLoopArrays& A, const int* __restrict__ status,
                                             const unsigned* __restrict__ active,
                                             const double* __restrict__ objective, double (&y_next)[4], double* ks,
                                             const double* rec_sh, double* rec_out, int rec_len,
                                             long long* tick_out = nullptr, long long tick_t0 = 0) {
  constexpr int N = PlantDims<PLANT>::N;
  double u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = A.u[size_t(b) * 4 + i];
  double (&xs)[6] = S.xs;
  double (&uc)[4] = S.uc;
  uc[0] += u[2 * c];
  if (valid) *S.ring_slot = u[2 * c + 1];
  if (S.rec && c == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) S.rec[1 + N + i] = u[i];
  }
  if (rec_sh && rec_out) {
    // the block's record rows are complete in shared memory: out they go as one contiguous chunk, 16 bytes
    // per lane and store (whole sectors, so the chunk may be mapped host memory), while the plant is integrated
    __syncwarp();
    const double2* src = reinterpret_cast<const double2*>(rec_sh);
    double2* dst = reinterpret_cast<double2*>(rec_out);
    for (int i = threadIdx.x & 31; i < rec_len / 2; i += 32) dst[i] = src[i];
    if ((rec_len & 1) && (threadIdx.x & 31) == 0) rec_out[rec_len - 1] = rec_sh[rec_len - 1];   // (ragged last block of an odd-length record)
  }
  if (c < NCTRL && valid) {
    const size_t o = (size_t(b) * A.n_steps + (k - A.rec_base)) * NCTRL + c;
    if (A.qp_active) A.qp_active[o] = active[b * NCTRL + c];
    if (A.qp_objective) A.qp_objective[o] = objective[b * NCTRL + c];
    if (A.qp_status) A.qp_status[o] = status[b * NCTRL + c];
  }
  const double u_tank = S.u_tank;
#ifdef CMPC_PHASE_TIMING
  if (c == 0) tick_out[13] = clock64() - tick_t0 + (uc[3] != uc[3] ? 1 : 0);   // loads have arrived
  const int n_acc = integrate_interval_pair<PLANT>(true, c, uc, u_tank, xs, Ts, ks, tick_out, tick_t0);
  if (c == 0) { tick_out[14] = (clock64() - tick_t0 + (xs[0] != xs[0] ? 1 : 0)) | ((long long)n_acc << 40); }
#else
  integrate_interval_pair<PLANT>(true, c, uc, u_tank, xs, Ts, ks);
#endif
  // next measurement: the other compressor's pressures and flow come by shuffle
  double p2, sd;
  compressor_output(xs, &p2, &sd);
  const double p2_o = __shfl_xor_sync(0xffffffffu, p2, 1), sd_o = __shfl_xor_sync(0xffffffffu, sd, 1);
  if (valid) {
#pragma unroll
    for (int i = 0; i < 5; ++i) A.x[size_t(b) * N + 5 * c + i] = xs[i];
  }
  double y[4] = {0.0, 0.0, 0.0, 0.0};
  if (c == 0) {
    if (PLANT == 0) {
      y[0] = sd; y[1] = sd_o; y[2] = p2 - p2_o; y[3] = xs[5];
      if (valid) A.x[size_t(b) * N + (PLANT == 0 ? 10 : 0)] = xs[5];
    } else {
      y[0] = p2; y[1] = sd; y[2] = p2_o; y[3] = sd_o;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) A.y[size_t(b) * 4 + i] = y[i];
    }
  }
  // both lanes leave with the measurement (lane 0 of the pair formed it)
#pragma unroll
  for (int i = 0; i < 4; ++i) y_next[i] = __shfl_sync(0xffffffffu, y[i], threadIdx.x & 30);
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
cl_advance_kernel(StepParams P, DeviceState G, int k, double t_k, double Ts, ClosedLoopArrays A, bool lin_next,
                  bool apriori_here) {
  constexpr int kScen = 16, kItems = kScen * S::NCTRL * 3;
  static_assert(kItems <= 96, "one linearisation item per lane of warps 1-3");
  __shared__ double y_sh[kScen][4];
  __shared__ double ks_sh[kDopriSlots * kDopriLanes];   // stage derivatives of the integrator (warp 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * kScen;
#ifdef CMPC_PHASE_TIMING
  const bool st_ = threadIdx.x == 0 && blockIdx.x * 16 < P.batch;
  if (st_) CMPC_GTIME_AT(blockIdx.x * 16, 22);
#endif
  // before the wait: the part of the plant side that does not depend on this record's solve kernel
  AdvancePre<S::PLANT> pre;
  const bool valid = b0 + (lane >> 1) < P.batch;            // warp 0: this lane pair has a scenario
  const int b = valid ? b0 + (lane >> 1) : P.batch - 1, c = lane & 1;
  constexpr int NIN = S::NIN, REC = 1 + S::N + 8;
  __shared__ __align__(16) double off_sh[kScen * NIN];
  __shared__ __align__(16) double rec_sh[kScen * REC];
  const int n_here = P.batch - b0 < kScen ? P.batch - b0 : kScen;   // scenarios of this block
  if (A.stream_io) {
    // this block's plant-input offsets: one contiguous chunk (possibly of mapped host memory), read once
    for (int i = threadIdx.x; i < n_here * NIN; i += blockDim.x) off_sh[i] = A.block_off[size_t(b0) * NIN + i];
    __syncthreads();
  }
  if (warp == 0)
    advance_pre<S::PLANT>(b, c, valid, k, t_k, A, pre, A.stream_io ? off_sh + (b - b0) * NIN : nullptr,
                          A.stream_io ? rec_sh + (b - b0) * REC : nullptr);
  pdl_wait();
  pdl_trigger();   // single wave
#ifdef CMPC_PHASE_TIMING
  if (st_) CMPC_GTIME_AT(blockIdx.x * 16, 23);
  const long long tk0_ = clock64();
#endif
  const bool early = lin_next && P.obs_states_free;
  // this lane's linearisation item (warps 1-3): scenario slot, sub-controller, part
  const int idx = (warp - 1) * 32 + lane;
  const int sl = idx / (3 * S::NCTRL), g = (idx % (3 * S::NCTRL)) / 3, part = idx % 3;
  const bool item = warp != 0 && idx < kItems && b0 + sl < P.batch;
  ObsPending<S> pend;
  if (warp == 0) {
    double y[4];
#ifdef CMPC_PHASE_TIMING
    advance_post<S::PLANT, S::NCTRL>(b, c, valid, k, Ts, pre, A, G.status, G.active, G.objective, y, ks_sh + lane,
                                     A.stream_io ? rec_sh : nullptr, A.traj ? A.traj + size_t(b0) * REC : nullptr, n_here * REC,
                                     G.ticks + size_t(b) * 32, tk0_);
#else
    advance_post<S::PLANT, S::NCTRL>(b, c, valid, k, Ts, pre, A, G.status, G.active, G.objective, y, ks_sh + lane,
                                     A.stream_io ? rec_sh : nullptr, A.traj ? A.traj + size_t(b0) * REC : nullptr, n_here * REC);
#endif
    if (c == 0 && valid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y_sh[lane >> 1][i] = y[i];
    }
  } else if (early && item) {
    lin_part_early<S>(P, G, b0 + sl, g, part, pend, apriori_here);
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
    if (item && part == 0 && g == 0) CMPC_TICK_AT(b0 + sl, 15, tk0_);
    if (item && part == 0 && g == 0 && sl == 0) CMPC_GTIME_AT(b0, 24);
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
  __shared__ double ks_sh[4][kDopriSlots * kDopriLanes];
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, ln = threadIdx.x & 31, c = ln & 1;
  if (b >= nq) return;
  double xs[6], uc[4];
  for (int i = 0; i < 5; ++i) xs[i] = x[size_t(b) * N + 5 * c + i];
  xs[5] = PLANT == 0 ? x[size_t(b) * N + 10] : 0.0;
  for (int i = 0; i < 4; ++i) uc[i] = u[size_t(b) * NIN + 4 * c + i];
  const double u_tank = PLANT == 0 ? u[size_t(b) * NIN + 8] : 0.0;
  const int s = integrate_interval_pair<PLANT>(true, c, uc, u_tank, xs, Ts, ks_sh[threadIdx.x >> 5] + ln);
  if (ln < 2) {
    for (int i = 0; i < 5; ++i) x[size_t(b) * N + 5 * c + i] = xs[i];
    if (PLANT == 0 && ln == 0) x[size_t(b) * N + 10] = xs[5];
  }
  if (n_substeps && ln == 0) n_substeps[b] = s;
}

// The straight-line square root and division of the plant integrator next to the standard operations
// (parity hook: cmpc_inrange_math).
static __global__ void inrange_math_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* sqrt_fast,
                                    double* sqrt_std, double* div_fast, double* div_std, int* flagged) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool bad_s = false, bad_d = false;
  sqrt_fast[i] = sqrt_inrange(a[i], bad_s);
  sqrt_std[i] = sqrt(a[i]);
  div_fast[i] = div_inrange(a[i], b[i], bad_d);
  div_std[i] = a[i] / b[i];
  flagged[i] = (bad_s ? 1 : 0) | (bad_d ? 2 : 0);
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
