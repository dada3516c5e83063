// The batched control step: one CTA per plant scenario, one thread group per
// sub-controller, everything between "new measurement y" and "next input u" on chip.
//
// Reference path replaced: NerveCenter::GetNextInputWithTiming
// (include/nerve_center.h:134-182) -> DistributedController::GenerateInitialQP
// (libs/distributed_controller.cc:72-108) -> Observer::ObserveAPosteriori
// (libs/observer.cc:24-40), AugmentedLinearizedSystem::Update / DiscretizeRK4 /
// GeneratePrediction (libs/aug_lin_sys.cc:145-177,232-255,260-334),
// DistributedSolver::GenerateDistributedQP (include/distributed_solver.h:83-94 ->
// libs/mpc_qp_solver.cc:19-40), then n_iterations Jacobi sweeps of
// DistributedController::GetInput (include/distributed_controller.h:206-226,
// distributed_solver.h:98-121, mpc_qp_solver.cc:45-75) and UpdateU / ObserveAPriori
// (distributed_controller.h:146-152, observer.cc:6-19).
//
// The prediction matrices Su/Sx/Sf/Su_other are never materialised.  Everything
// the QP needs follows from the impulse-response table
//     E[k][y][c] = C~ Ad^k [Bd | fd],   k = 0..p-1
// (C~ = controlled rows of C, Bd columns in this controller's input order):
//     G_k = E_k for undelayed inputs, E_{k-40} for delayed ones  (C~ A_aug^k B_aug)
//     Su[r] = [G_r | sum_{k<r} G_k],  Sf[r] fd = sum_{k<=r} E_k[fd],
//     Sx[r] x_aug = d + sum_t E_{r-t}[delayed] q[t]   (q = delay-line contents)
// and E itself is built as a product L R of baby steps L_a = C~ Ad^a (a < 8) and giant
// steps R_b = Ad^(8b) [Bd | fd], both obtained by repeated squaring/doubling, so the
// sequential depth is ~log2(p) small matrix products instead of p.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "plant_dev.cuh"
#include "qp_dev.cuh"

namespace cmpc {

constexpr int kDelay = 40;      // Delays = {0,40,0,40} (parallel/serial_compressors_constants.h)
constexpr int kNDist = 4;       // n_disturbance_states
constexpr int kNAug = kNDist + 2 * kDelay;  // 84
constexpr int kBaby = 8;        // baby steps a = 0..7
constexpr int kNC = 5;          // columns of [Bd | fd]
constexpr int kCtrlStateStride = 128;  // doubles per (scenario, controller) in global memory
constexpr int kScenStateStride = 16;   // doubles per scenario
constexpr int kMaxRows = 4;     // prediction rows owned by one thread (p <= 4 * TPC)

// offsets inside one controller's global state record
constexpr int kOffXhat = 0, kOffDx = 16, kOffYold = 112, kOffUold = 116;

template <int PLANT_, int NY_, int NU_, int NCTRL_>
struct Shape {
  static constexpr int PLANT = PLANT_, NY = NY_, NU = NU_, NCTRL = NCTRL_;
  static constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  static constexpr int NO = 4 - NU, NV = 2 * NU, NVO = 2 * NO;
  static constexpr int NOBS = N + kNDist, NTOT = N + kNAug;
  static constexpr int NH = NV * (NV + 1) / 2;     // upper triangle of H
  static constexpr int NACC = NH + NV * NVO + NV;  // H | Gx | f
  static constexpr int NCH = NY * kNC;             // scan channels
  static constexpr int TPC = 64;                   // threads per controller group
};

struct CtrlParams {
  int out_idx[4];     // ControlledOutputIndices
  int ctrl_idx[4];    // ControlInputIndices (local -> system control input)
  double Q[16];       // ywt, NY x NY row-major (symmetric)
  double R[16];       // uwt sub-matrix, NU x NU row-major
  double lower[4], upper[4], rate_lower[4], rate_upper[4];
  double M[15 * 4];   // observer gain, NOBS x 4 row-major
};

struct StepParams {
  int p, b_max, n_pow, n_iter, batch;
  double Ts;
  const double* yref;   // [NCTRL][p][NY]
  CtrlParams c[2];
};

// Global (HBM) arrays of one handle.
struct DeviceState {
  double* ctrl;        // [B][NCTRL][kCtrlStateStride]
  unsigned* guess;     // [B][NCTRL]
  double* scen;        // [B][kScenStateStride]: u_old (4, system order), du_old (8)
  double* u_offset;    // [B][NIN]
  // results / parity hooks of the last step
  double* qpH;         // [B][NCTRL][NV*NV]
  double* qpf;         // [B][NCTRL][NV]
  double* qpG;         // [B][NCTRL][NV*NVO]
  double* lin;         // [B][NCTRL][N*N + N*kNC]  (Ad | [Bd fd])   (capture only)
  double* etab;        // [B][NCTRL][p*NY*kNC]                       (capture only)
  int* status;         // [B][NCTRL]
  unsigned* active;    // [B][NCTRL]
  double* objective;   // [B][NCTRL]
};

__device__ __forceinline__ void group_sync(int g, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}

// Shared-memory footprint of one controller group, in doubles.  The big region is used three
// times: RK4 scratch -> powers Ad^(2^j) + L + R -> impulse-response table E -> reduction buffer.
template <class S>
struct SmemLayout {
  static constexpr int NN = S::N * S::N;
  int xh, dx, yv, yold, uold, ufull, ev, q, Cc, Ad, BF, carry, qp, region, L, R, E, total;
  bool e_alias;
  __host__ __device__ SmemLayout(int p, int b_max, int n_pow) {
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    xh = take(S::N);
    dx = take(S::NTOT);
    yv = take(4);
    yold = take(4);
    uold = take(4);
    ufull = take(S::NIN);
    ev = take(4);
    q = take(2 * kDelay);
    Cc = take(4 * S::N);
    Ad = take(NN);
    BF = take(S::N * kNC);
    carry = take((S::TPC / 32) * S::NCH);
    qp = take(S::NV * S::NV + S::NV + S::NV * (S::NVO > 0 ? S::NVO : 1) + QpFastLayout<S::NV>::size);
    region = o;
    const int n_scr = (n_pow > 5 ? n_pow : 5) * NN;
    L = region + n_scr;
    R = L + kBaby * S::NY * S::N;
    const int lr_end = R + b_max * S::N * kNC;
    const int e_size = kBaby * b_max * S::NCH;
    const int red_size = S::TPC * (S::NACC | 1);
    // E can overwrite its own inputs when every thread can hold its tiles in registers
    e_alias = (kBaby * b_max + S::TPC - 1) / S::TPC <= 2;
    E = e_alias ? region : lr_end;
    int end = e_alias ? (lr_end > region + e_size ? lr_end : region + e_size) : lr_end + e_size;
    if (end < E + red_size) end = E + red_size;
    total = (end + 1) & ~1;
  }
};

// Z = X * Y for N x N row-major matrices in shared memory, outputs strided over the group.
template <int N>
__device__ __forceinline__ void matmul_nn(const double* X, const double* Y, double* Z, int t, int nt) {
  for (int idx = t; idx < N * N; idx += nt) {
    const int i = idx / N, j = idx % N;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) s = fma(X[i * N + k], Y[k * N + j], s);
    Z[idx] = s;
  }
}

// C[r] . dx[0..N) for plant output row r at state x (the non-zeros of C, compressor.cc:169-170).
template <int PLANT>
__device__ __forceinline__ double plant_c_row_dot(const double* x, int r, const double* v) {
  if (PLANT == 0) {
    if (r < 2) {
      const double* xc = x + 5 * r;
      const double* vc = v + 5 * r;
      return 100 * xc[1] / (kSDc0 * xc[0] * xc[0]) * vc[0] - 100. / (kSDc0 * xc[0]) * vc[1] + 100 * vc[2];
    }
    return r == 2 ? v[1] - v[6] : v[10];
  } else {
    const int c = r >> 1;
    const double* xc = x + 5 * c;
    const double* vc = v + 5 * c;
    if ((r & 1) == 0) return vc[1];
    return 100 * xc[1] / (kSDc0 * xc[0] * xc[0]) * vc[0] - 100. / (kSDc0 * xc[0]) * vc[1] + 100 * vc[2];
  }
}

// One control step for the scenario owned by this CTA.  y4: the new measurement (4 doubles).
// u_out: 4 doubles.  All threads of the CTA must call it.
template <class S>
__device__ void control_step(const StepParams& P, const DeviceState& G, int scen, const double* y4,
                             double* u_out, double* smem) {
  constexpr int N = S::N, NY = S::NY, NU = S::NU, NV = S::NV, NVO = S::NVO, NO = S::NO;
  constexpr int NN = N * N, TPC = S::TPC, NTOT = S::NTOT, NOBS = S::NOBS, NCH = S::NCH, NH = S::NH;
  const int g = threadIdx.x / TPC, t = threadIdx.x % TPC;
  const int p = P.p, b_max = P.b_max;
  const SmemLayout<S> lay(p, b_max, P.n_pow);
  double* sm = smem + g * lay.total;
  double* zbuf = smem + S::NCTRL * lay.total;  // [NCTRL][NV] plans exchanged between sub-controllers
  const CtrlParams& cp = P.c[g];
  double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  double* ss = G.scen + size_t(scen) * kScenStateStride;

  double* xh = sm + lay.xh; double* dx = sm + lay.dx; double* yv = sm + lay.yv;
  double* yold = sm + lay.yold; double* uold = sm + lay.uold; double* ufull = sm + lay.ufull;
  double* ev = sm + lay.ev; double* q = sm + lay.q; double* Cc = sm + lay.Cc;
  double* Ad = sm + lay.Ad; double* BF = sm + lay.BF; double* scr = sm + lay.region;
  double* L = sm + lay.L; double* R = sm + lay.R; double* E = sm + lay.E;
  double* carry = sm + lay.carry; double* qpm = sm + lay.qp;

  // ---- phase 0: load state -------------------------------------------------------------
  for (int i = t; i < NTOT; i += TPC) dx[i] = gs[kOffDx + i];
  if (t < N) xh[t] = gs[kOffXhat + t];
  if (t < 4) {
    yold[t] = gs[kOffYold + t];
    uold[t] = gs[kOffUold + t];
    yv[t] = y4[t];
  }
  if (t >= 32 && t < 32 + S::NIN) {
    // u_full_old = GetPlantInput(u_old_, u_offset_)  (nerve_center.h:140)
    const int i = t - 32;
    double v = G.u_offset[size_t(scen) * S::NIN + i];
    if (i == 0) v += ss[0];
    if (i == 3) v += ss[1];
    if (i == 4) v += ss[2];
    if (i == 7) v += ss[3];
    ufull[i] = v;
  }
  group_sync(g, TPC);

  // ---- phase 1: Observer::ObserveAPosteriori (observer.cc:24-40), with the C of the
  //      previous linearisation (same x_hat) ------------------------------------------------
  // delay-line contents relative to u_old (AdjustAllDelayedStates, aug_lin_sys.h:141-154)
  for (int i = t; i < 2 * kDelay; i += TPC) {
    const int d = i / kDelay, tt = i % kDelay;
    const int slot = (tt == 0) ? (NOBS + d) : (NOBS + 2 + d * (kDelay - 1) + tt - 1);
    q[i] = dx[slot] - uold[1 + 2 * d];
  }
  if (t < 4) ev[t] = yv[t] - yold[t] - (plant_c_row_dot<S::PLANT>(xh, t, dx) + dx[N + t]);
  group_sync(g, TPC);
  double xh_new = 0.0;
  if (t < NOBS) {
    double acc = dx[t];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc = fma(cp.M[t * 4 + r], ev[r], acc);
    if (t < N) xh_new = xh[t] + acc;  // x_ += ObserveAPosteriori(y)  (distributed_controller.cc:80)
    dx[t] = acc;
  }
  group_sync(g, TPC);  // everyone has read the old xh/dx
  if (t < N) xh[t] = xh_new;
  // zero the continuous-time matrices before the sparse fill
  double* Ac = scr;            // continuous A
  double* A2 = scr + NN;
  double* A3 = scr + 2 * NN;
  double* Acom = scr + 3 * NN;
  double* Bc = scr + 4 * NN;   // N x 4
  double* fc = Bc + 4 * N;     // N
  for (int i = t; i < NN; i += TPC) Ac[i] = 0.0;
  for (int i = t; i < 4 * N; i += TPC) { Bc[i] = 0.0; Cc[i] = 0.0; }
  group_sync(g, TPC);

  // ---- phase 2: linearise at (x_hat, u_full_old)  (aug_lin_sys.cc:147); three threads ------
  if (t < 3) plant_linearize_part<S::PLANT>(t, xh, ufull, Ac, Bc, Cc, fc);
  group_sync(g, TPC);

  // ---- phase 3: DiscretizeRK4 (aug_lin_sys.cc:232-255) ------------------------------------
  matmul_nn<N>(Ac, Ac, A2, t, TPC);
  group_sync(g, TPC);
  matmul_nn<N>(A2, Ac, A3, t, TPC);
  group_sync(g, TPC);
  {
    const double Ts = P.Ts;
    const double c1 = Ts, c2 = Ts * Ts / 2.0, c3 = Ts * Ts * Ts / 6.0, c4 = Ts * Ts * Ts * Ts / 24.0;
    for (int idx = t; idx < NN; idx += TPC) {
      const int i = idx / N, j = idx % N;
      Acom[idx] = c1 * (i == j ? 1.0 : 0.0) + c2 * Ac[idx] + c3 * A2[idx] + c4 * A3[idx];
    }
  }
  group_sync(g, TPC);
  for (int idx = t; idx < NN + N * kNC; idx += TPC) {
    if (idx < NN) {
      const int i = idx / N, j = idx % N;
      double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], Ac[k * N + j], s);
      Ad[idx] = s;
    } else {
      // [Bd | fd] with Bd's columns permuted into this controller's input order
      // (aug_lin_sys.cc:156-173)
      const int r = idx - NN, i = r / kNC, c = r % kNC;
      double s = 0.0;
      if (c < 4) {
        const int col = cp.ctrl_idx[c];
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], Bc[k * 4 + col], s);
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], fc[k], s);
      }
      BF[r] = s;
    }
  }
  group_sync(g, TPC);
  if (G.lin) {
    double* gl = G.lin + (size_t(scen) * S::NCTRL + g) * (NN + N * kNC);
    for (int idx = t; idx < NN + N * kNC; idx += TPC) gl[idx] = (idx < NN) ? Ad[idx] : BF[idx - NN];
  }

  // ---- phase 4: powers Ad^(2^j) with baby (L) and giant (R) steps by doubling -------------
  // L_a = C~ Ad^a (a < 8): rows [2^j, 2^(j+1)) = rows [0, 2^j) * Ad^(2^j), j = 0..2
  // R_b = Ad^(8b) [Bd fd]: blocks [2^j, 2^(j+1)) = Ad^(8*2^j) * blocks [0, 2^j), j = 0..
  double* Pw = scr;  // Pw[j] = Ad^(2^j) at scr + j*NN; j = 0 is a copy of Ad
  for (int idx = t; idx < NN; idx += TPC) Pw[idx] = Ad[idx];
  for (int idx = t; idx < NY * N; idx += TPC) L[idx] = Cc[cp.out_idx[idx / N] * N + idx % N];
  for (int idx = t; idx < N * kNC; idx += TPC) R[idx] = BF[idx];
  group_sync(g, TPC);
  for (int s = 1; s <= P.n_pow; ++s) {
    // (a) Pw[s] = Pw[s-1]^2   (b) L doubling with Pw[s-1], s-1 < 3   (c) R doubling with Pw[s-1], s-1 >= 3
    const double* Pm = Pw + (s - 1) * NN;
    const int n_sq = (s < P.n_pow) ? NN : 0;
    const int j = s - 1;
    int n_l = 0, n_r = 0, r_base = 0;
    if (j < 3) {
      n_l = (1 << j) * NY * N;
    } else {
      r_base = 1 << (j - 3);
      int cnt = r_base;
      if (r_base + cnt > b_max) cnt = b_max - r_base;
      n_r = cnt > 0 ? cnt * N * kNC : 0;
    }
    for (int idx = t; idx < n_sq + n_l + n_r; idx += TPC) {
      if (idx < n_sq) {
        const int i = idx / N, jj = idx % N;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(Pm[i * N + k], Pm[k * N + jj], acc);
        Pw[s * NN + idx] = acc;
      } else if (idx < n_sq + n_l) {
        const int r = idx - n_sq, row = r / N, col = r % N;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(L[row * N + k], Pm[k * N + col], acc);
        L[((1 << j) * NY + row) * N + col] = acc;
      } else {
        const int r = idx - n_sq - n_l, blk = r / (N * kNC), rr = r % (N * kNC);
        const int i = rr / kNC, c = rr % kNC;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(Pm[i * N + k], R[(blk * N + k) * kNC + c], acc);
        R[((r_base + blk) * N + i) * kNC + c] = acc;
      }
    }
    group_sync(g, TPC);
  }

  // ---- phase 5: E[a + 8b] = L_a R_b --------------------------------------------------------
  const int K = kBaby * b_max;
  if (lay.e_alias) {
    // every thread keeps its (at most two) tiles in registers, then E overwrites L, R and the powers
    double acc[2][NY][kNC];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) {
      const int tile = t + ti * TPC;
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int c = 0; c < kNC; ++c) acc[ti][y][c] = 0.0;
      if (tile < K) {
        const double* Lr = L + (tile % kBaby) * NY * N;
        const double* Rb = R + (tile / kBaby) * N * kNC;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double rv[kNC];
#pragma unroll
          for (int c = 0; c < kNC; ++c) rv[c] = Rb[k * kNC + c];
#pragma unroll
          for (int y = 0; y < NY; ++y) {
            const double lv = Lr[y * N + k];
#pragma unroll
            for (int c = 0; c < kNC; ++c) acc[ti][y][c] = fma(lv, rv[c], acc[ti][y][c]);
          }
        }
      }
    }
    group_sync(g, TPC);
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) {
      const int tile = t + ti * TPC;
      if (tile < K) {
#pragma unroll
        for (int y = 0; y < NY; ++y)
#pragma unroll
          for (int c = 0; c < kNC; ++c) E[tile * NCH + y * kNC + c] = acc[ti][y][c];
      }
    }
  } else {
    for (int tile = t; tile < K; tile += TPC) {
      double acc[NY][kNC];
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int c = 0; c < kNC; ++c) acc[y][c] = 0.0;
      const double* Lr = L + (tile % kBaby) * NY * N;
      const double* Rb = R + (tile / kBaby) * N * kNC;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double rv[kNC];
#pragma unroll
        for (int c = 0; c < kNC; ++c) rv[c] = Rb[k * kNC + c];
#pragma unroll
        for (int y = 0; y < NY; ++y) {
          const double lv = Lr[y * N + k];
#pragma unroll
          for (int c = 0; c < kNC; ++c) acc[y][c] = fma(lv, rv[c], acc[y][c]);
        }
      }
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int c = 0; c < kNC; ++c) E[tile * NCH + y * kNC + c] = acc[y][c];
    }
  }
  group_sync(g, TPC);
  if (G.etab) {
    double* ge = G.etab + (size_t(scen) * S::NCTRL + g) * (size_t(p) * NCH);
    for (int idx = t; idx < p * NCH; idx += TPC) ge[idx] = E[idx];
  }

  // ---- phase 6: QP assembly.  Thread t owns prediction rows [t*rpt, (t+1)*rpt). ---------------
  //   G_r[y][c]: c < 4 input columns (delayed ones read 40 rows back), c = 4 the fd column
  //   prefix sums over r by a register scan: local totals -> warp scan -> carry across warps
  //   w_r = Sf fd + Sx x_aug - (y_ref - y)   (mpc_qp_solver.cc:31-37)
  //   H = Su' Q Su + R, Gx = Su' Q Su_other, f = Su' Q w accumulated per thread, then reduced.
  const int rpt = (p + TPC - 1) / TPC;
  const int r0 = t * rpt;
  const int r1 = (r0 + rpt < p) ? r0 + rpt : p;
  auto load_g = [&](int r, double* gv) {
    const double* Er = E + r * NCH;
    const bool del = r >= kDelay;
    const double* Ed = E + (del ? r - kDelay : 0) * NCH;
#pragma unroll
    for (int y = 0; y < NY; ++y) {
      gv[y * kNC + 0] = Er[y * kNC + 0];
      gv[y * kNC + 1] = del ? Ed[y * kNC + 1] : 0.0;
      gv[y * kNC + 2] = Er[y * kNC + 2];
      gv[y * kNC + 3] = del ? Ed[y * kNC + 3] : 0.0;
      gv[y * kNC + 4] = Er[y * kNC + 4];
    }
  };
  double off[NCH];
  {
    double tot[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tot[c] = 0.0;
    for (int r = r0; r < r1; ++r) {
      double gv[NCH];
      load_g(r, gv);
#pragma unroll
      for (int c = 0; c < NCH; ++c) tot[c] += gv[c];
    }
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      double v = tot[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += up;
      }
      off[c] = v - tot[c];                       // exclusive prefix inside the warp
      if (lane == 31) carry[warp * NCH + c] = v;  // warp total
    }
    group_sync(g, TPC);
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      for (int w = 0; w < warp; ++w) off[c] += carry[w * NCH + c];
  }
  // delay-line convolution for the owned rows, E rows shared between neighbouring rows
  double conv[kMaxRows][NY];
#pragma unroll
  for (int j = 0; j < kMaxRows; ++j)
#pragma unroll
    for (int y = 0; y < NY; ++y) conv[j][y] = 0.0;
  if (r0 < p) {
    const int k_hi = r1 - 1;
    const int k_lo = (r0 - (kDelay - 1) > 0) ? r0 - (kDelay - 1) : 0;
    for (int k = k_hi; k >= k_lo; --k) {
      double e1[NY], e3[NY];
      const double* Ek = E + k * NCH;
#pragma unroll
      for (int y = 0; y < NY; ++y) { e1[y] = Ek[y * kNC + 1]; e3[y] = Ek[y * kNC + 3]; }
#pragma unroll
      for (int j = 0; j < kMaxRows; ++j) {
        const int tt = r0 + j - k;
        if (j < rpt && tt >= 0 && tt < kDelay && r0 + j < p) {
          const double q0 = q[tt], q1 = q[kDelay + tt];
#pragma unroll
          for (int y = 0; y < NY; ++y) conv[j][y] = fma(e1[y], q0, fma(e3[y], q1, conv[j][y]));
        }
      }
    }
  }
  double acc[S::NACC];
#pragma unroll
  for (int i = 0; i < S::NACC; ++i) acc[i] = 0.0;
#pragma unroll
  for (int j = 0; j < kMaxRows; ++j) {
    const int r = r0 + j;
    if (j < rpt && r < p) {
      double gv[NCH], su[NY][NV], so[NY][NVO > 0 ? NVO : 1], qs[NY][NV], wv[NY];
      load_g(r, gv);
#pragma unroll
      for (int y = 0; y < NY; ++y) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < NU) {
            su[y][i] = gv[y * kNC + i];
            su[y][NU + i] = off[y * kNC + i];
          } else if (NVO > 0) {
            so[y][i - NU] = gv[y * kNC + i];
            so[y][NO + i - NU] = off[y * kNC + i];
          }
        }
        const int oy = cp.out_idx[y];
        const double yref = P.yref[(size_t(g) * p + r) * NY + y];
        wv[y] = (off[y * kNC + 4] + gv[y * kNC + 4]) + dx[N + oy] + conv[j][y] - (yref - yv[oy]);
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) off[c] += gv[c];
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double s = 0.0;
#pragma unroll
          for (int y2 = 0; y2 < NY; ++y2) s = fma(cp.Q[y * NY + y2], su[y2][v], s);
          qs[y][v] = s;
        }
      int hi = 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int v2 = v; v2 < NV; ++v2, ++hi)
#pragma unroll
          for (int y = 0; y < NY; ++y) acc[hi] = fma(su[y][v], qs[y][v2], acc[hi]);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int vo = 0; vo < NVO; ++vo)
#pragma unroll
          for (int y = 0; y < NY; ++y)
            acc[NH + v * NVO + vo] = fma(so[y][vo], qs[y][v], acc[NH + v * NVO + vo]);
#pragma unroll
        for (int y = 0; y < NY; ++y) acc[NH + NV * NVO + v] = fma(wv[y], qs[y][v], acc[NH + NV * NVO + v]);
      }
    }
  }
  group_sync(g, TPC);  // E is dead: reuse it as the reduction buffer
  {
    constexpr int RS = S::NACC | 1;  // odd stride: conflict-free 64-bit stores
    double* red = E;
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) red[t * RS + i] = acc[i];
    group_sync(g, TPC);
    // qpm: H (NV*NV) | f (NV) | Gx (NV*NVO)
    if (t < S::NACC) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < TPC; k += 2) {
        s0 += red[k * RS + t];
        s1 += red[(k + 1) * RS + t];
      }
      const double v = s0 + s1;
      if (t < NH) {
        int a = 0, rem = t;
        while (rem >= NV - a) { rem -= NV - a; ++a; }
        const int b = a + rem;
        double hv = v;
        if (a / NU == b / NU) hv += cp.R[(a % NU) * NU + (b % NU)];  // u_weight_ = I_m (x) uwt
        qpm[a * NV + b] = hv;
        qpm[b * NV + a] = hv;
        double* gH = G.qpH + (size_t(scen) * S::NCTRL + g) * NV * NV;
        gH[a * NV + b] = hv;
        gH[b * NV + a] = hv;
      } else if (t < NH + NV * NVO) {
        qpm[NV * NV + NV + (t - NH)] = v;
        G.qpG[(size_t(scen) * S::NCTRL + g) * NV * (NVO > 0 ? NVO : 1) + (t - NH)] = v;
      } else {
        const int v_i = t - NH - NV * NVO;
        qpm[NV * NV + v_i] = v;
        G.qpf[(size_t(scen) * S::NCTRL + g) * NV + v_i] = v;
      }
    }
  }
  __syncthreads();

  // ---- phase 8: n_iter Jacobi sweeps (nerve_center.h:146-158,275-296) ----------------------
  // lane c of warp 0 owns sub-controller c; plans are exchanged through zbuf.
  if (threadIdx.x < 32) {
    const int c = threadIdx.x;
    const bool on = c < S::NCTRL;
    QpData<NV> qd;
    double f0[NV], z[NV], lam[NV], fi[NV];
    unsigned wset = kQpNoGuess, act = 0;
    double obj = 0.0;
    int status = 0, nq = 0;
    bool pd = true, fast_ok = false;
    const double* qm = smem + (on ? c : 0) * lay.total + lay.qp;
    double* fast = smem + (on ? c : 0) * lay.total + lay.qp + NV * NV + NV + NV * (NVO > 0 ? NVO : 1);
    if (on) {
      const double* uo = smem + c * lay.total + lay.uold;
      const CtrlParams& cq = P.c[c];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        f0[i] = qm[NV * NV + i];
        z[i] = 0.0;
        qd.lb[i] = cq.lower[i % NU] - uo[i % NU];
        qd.ub[i] = cq.upper[i % NU] - uo[i % NU];
        qd.lbA[i] = cq.rate_lower[i % NU];
        qd.ubA[i] = cq.rate_upper[i % NU];
      }
      pd = qp_invert_spd<NV>(qm, qd.J);
      wset = G.guess[size_t(scen) * S::NCTRL + c];
      if (pd && wset != kQpNoGuess) fast_ok = qp_prepare<NV, NU>(qd, wset, fast, &nq);
#pragma unroll
      for (int i = 0; i < NV; ++i) zbuf[c * NV + i] = ss[4 + c * NV + i];  // du_prev = du_old_
    }
    __syncwarp();
    for (int it = 0; it < P.n_iter; ++it) {
      if (on) {
#pragma unroll
        for (int i = 0; i < NV; ++i) fi[i] = f0[i];
        if (NVO > 0) {
          const double* zo = zbuf + (1 - c) * NV;  // the other controller's previous plan
          const double* Gx = qm + NV * NV + NV;
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int k = 0; k < NVO; ++k) fi[i] = fma(Gx[i * NVO + k], zo[k], fi[i]);
        }
        if (!pd) {
          status = 3;
        } else if (fast_ok && qp_eval_fast<NV, NU>(qd, fast, nq, wset, fi, z, lam)) {
          status = 0;
        } else {
          // the working set changes (rare): general dual active-set solve, then re-prepare
          unsigned gset = wset;
          status = qp_solve<NV, NU>(qd, qm, fi, &gset, z, &act, &obj);
          if (status == 0) {
            wset = gset;
            fast_ok = qp_prepare<NV, NU>(qd, wset, fast, &nq);
            // multipliers of the new working set for the report below
            if (fast_ok) qp_eval_fast<NV, NU>(qd, fast, nq, wset, fi, z, lam);
          } else {
            fast_ok = false;
          }
        }
        if (status != 0) {
#pragma unroll
          for (int i = 0; i < NV; ++i) z[i] = 0.0;   // mpc_qp_solver.cc:66-69
        }
      }
      __syncwarp();
      if (on) {
#pragma unroll
        for (int i = 0; i < NV; ++i) zbuf[c * NV + i] = z[i];
      }
      __syncwarp();
    }
    if (on) {
      if (status == 0) {
        // report of the last sweep: active constraints (strictly positive multiplier), objective
        double fmax = 1.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) fmax = fmax > fabs(fi[i]) ? fmax : fabs(fi[i]);
        act = 0;
        int w = 0;
        for (int j = 0; j < 4 * NV; ++j)
          if ((wset >> j) & 1u) {
            if (wset != kQpNoGuess && lam[w] > 1e-9 * fmax) act |= 1u << j;
            ++w;
          }
        obj = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < NV; ++k) s = fma(qm[i * NV + k], z[k], s);
          obj += z[i] * (0.5 * s + fi[i]);
        }
        G.guess[size_t(scen) * S::NCTRL + c] = wset;
      } else {
        act = 0;
        obj = 0.0;
      }
      G.status[size_t(scen) * S::NCTRL + c] = status;
      G.active[size_t(scen) * S::NCTRL + c] = act;
      G.objective[size_t(scen) * S::NCTRL + c] = obj;
    }
  }
  __syncthreads();

  // ---- phase 9: apply first move, UpdateU / ObserveAPriori (observer.cc:6-19) --------------
  {
    double du[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < NU; ++i) du[i] = zbuf[g * NV + i];
    const double h0 = dx[NOBS + 0] - uold[1], h1 = dx[NOBS + 1] - uold[3];
    for (int i = t; i < NTOT; i += TPC) {
      double v;
      if (i < N) {
        v = BF[i * kNC + 0] * du[0] + BF[i * kNC + 2] * du[2] + BF[i * kNC + 1] * h0 +
            BF[i * kNC + 3] * h1 + BF[i * kNC + 4];
      } else if (i < NOBS) {
        v = dx[i];
      } else if (i < NOBS + 2) {
        v = dx[NOBS + 2 + (i - NOBS) * (kDelay - 1)];  // head <- first chain slot
      } else {
        const int cidx = i - NOBS - 2, d = cidx / (kDelay - 1), jj = cidx % (kDelay - 1);
        v = (jj == kDelay - 2) ? uold[1 + 2 * d] + du[1 + 2 * d] : dx[i + 1];
      }
      gs[kOffDx + i] = v;
    }
    if (t < N) gs[kOffXhat + t] = xh[t];
    if (t < 4) {
      gs[kOffYold + t] = yv[t];
      gs[kOffUold + t] = uold[t] + du[t];
    }
  }
  if (threadIdx.x < 4) {
    // nerve_center.h:162-167,313-319: u_old_ += first move of each controller's plan
    const int c = threadIdx.x / NU, i = threadIdx.x % NU;
    const double un = ss[threadIdx.x] + zbuf[c * NV + i];
    u_out[threadIdx.x] = un;
    ss[threadIdx.x] = un;
  }
  if (threadIdx.x < S::NCTRL * NV) ss[4 + threadIdx.x] = zbuf[threadIdx.x];  // du_old_ = du_prev
}

#ifndef CMPC_MIN_BLOCKS
#define CMPC_MIN_BLOCKS 3
#endif

template <class S>
__global__ void __launch_bounds__(S::NCTRL * S::TPC, CMPC_MIN_BLOCKS)
step_kernel(StepParams P, DeviceState G, const double* __restrict__ y, double* __restrict__ u) {
  extern __shared__ __align__(16) double smem[];
  const int scen = blockIdx.x;
  if (scen >= P.batch) return;
  control_step<S>(P, G, scen, y + size_t(scen) * 4, u + size_t(scen) * 4, smem);
}

}  // namespace cmpc
