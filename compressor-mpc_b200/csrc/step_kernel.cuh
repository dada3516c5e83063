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
constexpr int kMaxPow = 8;      // Ad^(2^j), j = 0..7

// offsets inside one controller's global state record
constexpr int kOffXhat = 0, kOffDx = 16, kOffYold = 112, kOffUold = 116;

template <int PLANT_, int NY_, int NU_, int NCTRL_>
struct Shape {
  static constexpr int PLANT = PLANT_, NY = NY_, NU = NU_, NCTRL = NCTRL_;
  static constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  static constexpr int NO = 4 - NU, NV = 2 * NU, NVO = 2 * NO;
  static constexpr int NOBS = N + kNDist, NTOT = N + kNAug;
  static constexpr int NACC = NV * (NV + NVO + 1);  // H | Gx | f
  static constexpr int TPC = 64;                    // threads per controller group
};

struct CtrlParams {
  int out_idx[4];     // ControlledOutputIndices
  int ctrl_idx[4];    // ControlInputIndices (local -> system control input)
  double Q[16];       // ywt, NY x NY row-major
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
  double* lin;         // [B][NCTRL][N*N + N*kNC]  (Ad | [Bd fd])
  double* etab;        // [B][NCTRL][p*NY*kNC] (optional, may be null)
  int* status;         // [B][NCTRL]
  unsigned* active;    // [B][NCTRL]
  double* objective;   // [B][NCTRL]
};

__device__ __forceinline__ void group_sync(int g, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}

// Shared-memory footprint of one controller group, in doubles.
template <class S>
struct SmemLayout {
  static constexpr int NN = S::N * S::N;
  int xh, dx, yv, yold, uold, ufull, ev, q, Cc, Ad, BF, scratch, L, R, E, CE, W, red, qp, total;
  __host__ __device__ SmemLayout(int p, int b_max) {
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
    scratch = take(kMaxPow * NN);            // RK4 scratch, then powers Ad^(2^j)
    L = take(kBaby * S::NY * S::N);
    R = take(b_max * S::N * kNC);
    E = take(kBaby * b_max * S::NY * kNC);
    CE = take((kBaby * b_max + 1) * S::NY * kNC);
    W = take(p * S::NY);
    red = take(2 * S::NACC);
    qp = take(S::NV * S::NV + S::NV + S::NV * S::NVO + 2 * S::NV + 8);
    total = o;
  }
};

// C = X * Y for N x N row-major matrices in shared memory, outputs strided over the group.
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

// One control step for the scenario owned by this CTA.  y4: the new measurement (global or
// shared pointer, 4 doubles).  u_out: 4 doubles.  All threads of the CTA must call it.
template <class S>
__device__ void control_step(const StepParams& P, const DeviceState& G, int scen, const double* y4,
                             double* u_out, double* smem) {
  constexpr int N = S::N, NY = S::NY, NU = S::NU, NV = S::NV, NVO = S::NVO, NO = S::NO;
  constexpr int NN = N * N, TPC = S::TPC, NTOT = S::NTOT, NOBS = S::NOBS;
  const int g = threadIdx.x / TPC, t = threadIdx.x % TPC;
  const int p = P.p, b_max = P.b_max;
  const SmemLayout<S> lay(p, b_max);
  double* sm = smem + g * lay.total;
  double* zbuf = smem + S::NCTRL * lay.total;  // [2][NCTRL][NV] Jacobi exchange + first moves
  const CtrlParams& cp = P.c[g];
  double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  double* ss = G.scen + size_t(scen) * kScenStateStride;

  double* xh = sm + lay.xh; double* dx = sm + lay.dx; double* yv = sm + lay.yv;
  double* yold = sm + lay.yold; double* uold = sm + lay.uold; double* ufull = sm + lay.ufull;
  double* ev = sm + lay.ev; double* q = sm + lay.q; double* Cc = sm + lay.Cc;
  double* Ad = sm + lay.Ad; double* BF = sm + lay.BF; double* scr = sm + lay.scratch;
  double* L = sm + lay.L; double* R = sm + lay.R; double* E = sm + lay.E; double* CE = sm + lay.CE;
  double* W = sm + lay.W; double* red = sm + lay.red; double* qpm = sm + lay.qp;

  // ---- phase 0: load state -------------------------------------------------------------
  for (int i = t; i < NTOT; i += TPC) dx[i] = gs[kOffDx + i];
  if (t < N) xh[t] = gs[kOffXhat + t];
  if (t < 4) {
    yold[t] = gs[kOffYold + t];
    uold[t] = gs[kOffUold + t];
    yv[t] = y4[t];
  }
  if (t < S::NIN) ufull[t] = G.u_offset[size_t(scen) * S::NIN + t];
  group_sync(g, TPC);
  // u_full_old = GetPlantInput(u_old_, u_offset_)  (nerve_center.h:140)
  if (t < 4) {
    const int plant_idx = (t == 0) ? 0 : (t == 1) ? 3 : (t == 2) ? 4 : 7;
    ufull[plant_idx] += ss[t];
  }
  // delay-line contents relative to u_old (AdjustAllDelayedStates, aug_lin_sys.h:141-154)
  for (int i = t; i < 2 * kDelay; i += TPC) {
    const int d = i / kDelay, tt = i % kDelay;
    const int slot = (tt == 0) ? (NOBS + d) : (NOBS + 2 + d * (kDelay - 1) + tt - 1);
    q[i] = dx[slot] - uold[1 + 2 * d];
  }
  if (t == 0) plant_c_entry<S::PLANT>(xh, Cc);  // C of the previous linearisation (same x_hat)
  group_sync(g, TPC);

  // ---- phase 1: Observer::ObserveAPosteriori (observer.cc:24-40) -------------------------
  if (t < 4) {
    double cy = dx[N + t];
    for (int k = 0; k < N; ++k) cy += Cc[t * N + k] * dx[k];
    ev[t] = yv[t] - yold[t] - cy;
  }
  group_sync(g, TPC);
  if (t < NOBS) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) acc += cp.M[t * 4 + r] * ev[r];
    dx[t] += acc;
    if (t < N) xh[t] += dx[t];  // x_ += ObserveAPosteriori(y)  (distributed_controller.cc:80)
  }
  group_sync(g, TPC);

  // ---- phase 2: linearise at (x_hat, u_full_old)  (aug_lin_sys.cc:147) --------------------
  double* Ac = scr;            // continuous A
  double* A2 = scr + NN;
  double* A3 = scr + 2 * NN;
  double* Acom = scr + 3 * NN;
  double* Bc = scr + 4 * NN;   // N x 4
  double* fc = Bc + 4 * N;     // N
  if (t == 0) plant_linearize<S::PLANT>(xh, ufull, Ac, Bc, Cc, fc);
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
  for (int tile = t; tile < K; tile += TPC) {
    const int a = tile % kBaby, b = tile / kBaby;
    double acc[NY][kNC];
#pragma unroll
    for (int y = 0; y < NY; ++y)
#pragma unroll
      for (int c = 0; c < kNC; ++c) acc[y][c] = 0.0;
    const double* Lr = L + a * NY * N;
    const double* Rb = R + b * N * kNC;
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
    double* Ek = E + tile * NY * kNC;
#pragma unroll
    for (int y = 0; y < NY; ++y)
#pragma unroll
      for (int c = 0; c < kNC; ++c) Ek[y * kNC + c] = acc[y][c];
  }
  group_sync(g, TPC);
  if (G.etab) {
    double* ge = G.etab + (size_t(scen) * S::NCTRL + g) * (size_t(p) * NY * kNC);
    for (int idx = t; idx < p * NY * kNC; idx += TPC) ge[idx] = E[idx];
  }

  // ---- phase 6: exclusive prefix sums over k, delay-line convolution, w -------------------
  if (t < NY * kNC) {
    double run = 0.0;
    for (int k = 0; k < K; ++k) {
      CE[k * NY * kNC + t] = run;
      run += E[k * NY * kNC + t];
    }
    CE[K * NY * kNC + t] = run;
  }
  group_sync(g, TPC);
  for (int r = t; r < p; r += TPC) {
    double conv[NY];
#pragma unroll
    for (int y = 0; y < NY; ++y) conv[y] = 0.0;
    const int tmax = r < kDelay - 1 ? r : kDelay - 1;
    for (int tt = 0; tt <= tmax; ++tt) {
      const double q0 = q[tt], q1 = q[kDelay + tt];
      const double* Er = E + (r - tt) * NY * kNC;
#pragma unroll
      for (int y = 0; y < NY; ++y) conv[y] = fma(Er[y * kNC + 1], q0, fma(Er[y * kNC + 3], q1, conv[y]));
    }
#pragma unroll
    for (int y = 0; y < NY; ++y) {
      const int oy = cp.out_idx[y];
      const double yref = P.yref[(size_t(g) * p + r) * NY + y];
      // Sf fd + Sx x_aug - (y_ref - y)   (mpc_qp_solver.cc:31-37)
      W[r * NY + y] = CE[(r + 1) * NY * kNC + y * kNC + 4] + dx[N + oy] + conv[y] - (yref - yv[oy]);
    }
  }
  group_sync(g, TPC);

  // ---- phase 7: H = Su' Q Su + R, Gx = Su' Q Su_other, f = Su' Q w -----------------------
  {
    double acc[S::NACC];
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) acc[i] = 0.0;
    for (int r = t; r < p; r += TPC) {
      double su[NY][NV], so[NY][NVO > 0 ? NVO : 1], qs[NY][NV], wv[NY];
      const double* Er = E + r * NY * kNC;
      const double* Cr = CE + r * NY * kNC;
      const bool del = r >= kDelay;
      const double* Ed = E + (del ? r - kDelay : 0) * NY * kNC;
      const double* Cd = CE + (del ? r - kDelay : 0) * NY * kNC;
#pragma unroll
      for (int y = 0; y < NY; ++y) {
        wv[y] = W[r * NY + y];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool delayed = (i & 1);
          const double gval = delayed ? (del ? Ed[y * kNC + i] : 0.0) : Er[y * kNC + i];
          const double pval = delayed ? (del ? Cd[y * kNC + i] : 0.0) : Cr[y * kNC + i];
          if (i < NU) {
            su[y][i] = gval;
            su[y][NU + i] = pval;
          } else if (NVO > 0) {
            so[y][i - NU] = gval;
            so[y][NO + i - NU] = pval;
          }
        }
      }
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double s = 0.0;
#pragma unroll
          for (int y2 = 0; y2 < NY; ++y2) s = fma(cp.Q[y * NY + y2], su[y2][v], s);
          qs[y][v] = s;
        }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int v2 = 0; v2 < NV; ++v2)
#pragma unroll
          for (int y = 0; y < NY; ++y) acc[v * NV + v2] = fma(su[y][v], qs[y][v2], acc[v * NV + v2]);
#pragma unroll
        for (int vo = 0; vo < NVO; ++vo)
#pragma unroll
          for (int y = 0; y < NY; ++y)
            acc[NV * NV + v * NVO + vo] = fma(so[y][vo], qs[y][v], acc[NV * NV + v * NVO + vo]);
#pragma unroll
        for (int y = 0; y < NY; ++y)
          acc[NV * NV + NV * NVO + v] = fma(wv[y], qs[y][v], acc[NV * NV + NV * NVO + v]);
      }
    }
    // reduce over the group: butterfly inside each warp, then across the two warps
#pragma unroll
    for (int i = 0; i < S::NACC; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((t & 31) == 0) red[(t >> 5) * S::NACC + i] = v;
    }
  }
  group_sync(g, TPC);
  // qpm: H (NV*NV) | f (NV) | Gx (NV*NVO)
  for (int i = t; i < S::NACC; i += TPC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < TPC / 32; ++w) v += red[w * S::NACC + i];
    if (i < NV * NV) {
      const int a = i / NV, b = i % NV;
      if (a / NU == b / NU) v += cp.R[(a % NU) * NU + (b % NU)];  // u_weight_ = I_m (x) uwt
      qpm[i] = v;
      G.qpH[(size_t(scen) * S::NCTRL + g) * NV * NV + i] = v;
    } else if (i < NV * NV + NV * NVO) {
      qpm[NV * NV + NV + (i - NV * NV)] = v;
      G.qpG[(size_t(scen) * S::NCTRL + g) * NV * (NVO > 0 ? NVO : 1) + (i - NV * NV)] = v;
    } else {
      const int v_i = i - NV * NV - NV * NVO;
      qpm[NV * NV + v_i] = v;
      G.qpf[(size_t(scen) * S::NCTRL + g) * NV + v_i] = v;
    }
  }
  __syncthreads();

  // ---- phase 8: n_iter Jacobi sweeps (nerve_center.h:146-158,275-296) ----------------------
  // lane c of warp 0 owns sub-controller c; plans are exchanged through zbuf.
  if (threadIdx.x < 32) {
    const int c = threadIdx.x;
    const bool on = c < S::NCTRL;
    QpData<NV> qd;
    double Hm[NV * NV], f0[NV], Gx[NV * (NVO > 0 ? NVO : 1)], z[NV];
    unsigned guess = kQpNoGuess, act = 0;
    double obj = 0.0;
    int status = 0;
    bool pd = true;
    if (on) {
      const double* qm = smem + c * lay.total + lay.qp;
      const double* uo = smem + c * lay.total + lay.uold;
      const CtrlParams& cq = P.c[c];
#pragma unroll
      for (int i = 0; i < NV * NV; ++i) Hm[i] = qm[i];
#pragma unroll
      for (int i = 0; i < NV; ++i) f0[i] = qm[NV * NV + i];
#pragma unroll
      for (int i = 0; i < NV * NVO; ++i) Gx[i] = qm[NV * NV + NV + i];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        qd.lb[i] = cq.lower[i % NU] - uo[i % NU];
        qd.ub[i] = cq.upper[i % NU] - uo[i % NU];
        qd.lbA[i] = cq.rate_lower[i % NU];
        qd.ubA[i] = cq.rate_upper[i % NU];
      }
      pd = qp_invert_spd<NV>(Hm, qd.J);
      guess = G.guess[size_t(scen) * S::NCTRL + c];
#pragma unroll
      for (int i = 0; i < NV; ++i) zbuf[c * NV + i] = ss[4 + c * NV + i];  // du_prev = du_old_
    }
    __syncwarp();
    for (int it = 0; it < P.n_iter; ++it) {
      if (on) {
        double fi[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) fi[i] = f0[i];
        if (NVO > 0) {
          const double* zo = zbuf + (1 - c) * NV;  // the other controller's previous plan
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int k = 0; k < NVO; ++k) fi[i] = fma(Gx[i * NVO + k], zo[k], fi[i]);
        }
        if (pd) {
          status = qp_solve<NV, NU>(qd, Hm, fi, &guess, z, &act, &obj);
        } else {
          status = 3;
#pragma unroll
          for (int i = 0; i < NV; ++i) z[i] = 0.0;
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
      G.guess[size_t(scen) * S::NCTRL + c] = guess;
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
    double newv[(NTOT + TPC - 1) / TPC];
    int cnt = 0;
    for (int i = t; i < NTOT; i += TPC, ++cnt) {
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
      newv[cnt] = v;
    }
    group_sync(g, TPC);
    cnt = 0;
    for (int i = t; i < NTOT; i += TPC, ++cnt) gs[kOffDx + i] = newv[cnt];
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

template <class S>
__global__ void __launch_bounds__(S::NCTRL * S::TPC)
step_kernel(StepParams P, DeviceState G, const double* __restrict__ y, double* __restrict__ u) {
  extern __shared__ __align__(16) double smem[];
  const int scen = blockIdx.x;
  if (scen >= P.batch) return;
  control_step<S>(P, G, scen, y + size_t(scen) * 4, u + size_t(scen) * 4, smem);
}

}  // namespace cmpc
