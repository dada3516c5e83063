// The control step for configurations outside the reference's own instantiations: the run-time
// replacement of its compile-time configuration (include/constexpr_array.h:10-142,
// include/{parallel,serial}_compressors_constants.h, include/nerve_center.h:19-38 -- a parameter
// pack of any number of sub-controllers).  Any per-input delay (0 or >= 2 samples), any move
// horizon m (m * n_u <= 8 per sub-controller), any controlled-output partition, one to four
// sub-controllers with their own input counts.
//
// One CTA per scenario does the whole step of every sub-controller in turn (the tuned kernels of
// step_kernel.cuh cover the reference's seven shapes; this path is about reach, not speed):
//   Observer::ObserveAPosteriori                      libs/observer.cc:24-40
//   AugmentedLinearizedSystem::Update / DiscretizeRK4  libs/aug_lin_sys.cc:145-177,232-255
//   GeneratePrediction                                 libs/aug_lin_sys.cc:260-334
//   DistributedSolver::GenerateDistributedQP           include/distributed_solver.h:83-94, libs/mpc_qp_solver.cc:19-40
//   NerveCenter sweeps / ApplyOtherInput / SolveQP     include/nerve_center.h:146-158,275-296,
//                                                      include/distributed_solver.h:98-121, libs/mpc_qp_solver.cc:45-75
//   UpdateUOld / SendU / UpdateU / ObserveAPriori      include/nerve_center.h:313-328, include/distributed_controller.h:146-152,
//                                                      libs/observer.cc:6-19
// As in step_kernel.cuh the prediction matrices are not materialised: with the delay-free impulse
// response E_k = C~ Ad^k Bd (k < p) and its running sum PE_k, the literal accumulation of
// aug_lin_sys.cc:311-328 gives  Su[(i,y)][(mv,q)] = E_{i-mv-d_q}  for mv < m-1  and  PE_{i-m+1-d_q}
// for the last move (d_q the delay of local input q, negative index = 0); Sf f_d + Sx x_aug is the
// free response of the augmented model, whose delay chains only shift.
#pragma once
#include <cuda_runtime.h>

#include "generic_params.cuh"
#include "plant_dev.cuh"
#include "plant_kernels.cuh"
#include "qp_dev.cuh"
#include "step_kernel.cuh"

namespace cmpc {

template <int NU>
__device__ __forceinline__ int gen_qp(const QpData<8>& qd, const double* H, const double* f, unsigned* wset,
                                      double* z, unsigned* act, double* obj) {
  return qp_solve<8, NU>(qd, H, f, wset, z, act, obj);
}

template <int PLANT>
__global__ void __launch_bounds__(kGenThreads)
gen_step_kernel(const GenParams* __restrict__ Pp, GenState G, const double* __restrict__ y, double* __restrict__ u) {
  extern __shared__ __align__(16) double sm[];
  const GenParams& P = *Pp;
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  const int scen = blockIdx.x;
  if (scen >= P.batch) return;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int p = P.p, m = P.m, n_ctrl = P.n_ctrl, n_total = P.n_total, dmax = P.dmax;
  int o = 0;
  auto take = [&](int k) { double* r = sm + o; o += k; return r; };
  double* A = take(N * N); double* A2 = take(N * N); double* A3 = take(N * N); double* Acom = take(N * N);
  double* Ad = take(N * N); take(2 * N * N);
  double* B4 = take(N * 4); double* Bd4 = take(N * 4); double* Bloc = take(N * 4); double* Cm = take(4 * N);
  double* fc = take(N); double* fd = take(N); take(2 * N);
  double* q = take(4 * dmax);
  double* vbuf = take(5 * 2 * 16);
  double* E = take(p * 16); double* PE = take(p * 16);
  double* yfree = take(p * 4); double* wv = take(p * 4);
  double* Hs = take(n_ctrl * 64); double* fs = take(n_ctrl * 8); double* Gxs = take(n_ctrl * 8 * 16);
  double* Bsave = take(n_ctrl * N * 5);
  double* plan = take(2 * 16); double* du_sys = take(16); double* misc = take(8);
  double* stage = take(256);
  double* ss = G.scen + size_t(scen) * kGenScenStride;

  for (int c = 0; c < n_ctrl; ++c) {
    const GenCtrl& K = P.c[c];
    double* gs = G.ctrl + (size_t(scen) * n_ctrl + c) * P.state_stride;
    double* dxg = gs + N;
    double* yold = gs + N + n_total;
    double* uold = yold + 4;
    // ---- ObserveAPosteriori with the C of the previous linearisation, x += dx, linearise -------
    if (t == 0) {
      double xh[N], dx[N + 4], ev[4], yv[4], uf[NIN];
      for (int i = 0; i < N; ++i) xh[i] = gs[i];
      for (int i = 0; i < N + 4; ++i) dx[i] = dxg[i];
      for (int r = 0; r < 4; ++r) {
        yv[r] = y[size_t(scen) * 4 + r];
        ev[r] = yv[r] - yold[r] - (plant_c_row_dot<PLANT>(xh, r, dx) + dx[N + r]);
      }
      for (int i = 0; i < N + 4; ++i) {
        double acc = dx[i];
        for (int r = 0; r < 4; ++r) acc = fma(K.M[i * 4 + r], ev[r], acc);
        dx[i] = acc;
        dxg[i] = acc;
      }
      for (int i = 0; i < N; ++i) {
        xh[i] += dx[i];
        gs[i] = xh[i];
      }
      for (int r = 0; r < 4; ++r) {
        yold[r] = yv[r];
        misc[r] = yv[r];
      }
      // u_full_old = GetPlantInput(u_old_, u_offset_)  (nerve_center.h:140)
      for (int i = 0; i < NIN; ++i) uf[i] = G.u_offset[size_t(scen) * NIN + i];
      uf[0] += ss[0]; uf[3] += ss[1]; uf[4] += ss[2]; uf[7] += ss[3];
      plant_linearize<PLANT>(xh, uf, A, B4, Cm, fc);
    }
    __syncthreads();
    // ---- DiscretizeRK4 ------------------------------------------------------------------
    for (int idx = t; idx < N * N; idx += kGenThreads) {
      const int i = idx / N, j = idx % N;
      double s = 0.0;
      for (int k = 0; k < N; ++k) s = fma(A[i * N + k], A[k * N + j], s);
      A2[idx] = s;
    }
    __syncthreads();
    for (int idx = t; idx < N * N; idx += kGenThreads) {
      const int i = idx / N, j = idx % N;
      double s = 0.0;
      for (int k = 0; k < N; ++k) s = fma(A2[i * N + k], A[k * N + j], s);
      A3[idx] = s;
    }
    __syncthreads();
    for (int idx = t; idx < N * N; idx += kGenThreads) {
      const int i = idx / N, j = idx % N;
      Acom[idx] = P.rk[0] * (i == j ? 1.0 : 0.0) + P.rk[1] * A[idx] + P.rk[2] * A2[idx] + P.rk[3] * A3[idx];
    }
    __syncthreads();
    for (int idx = t; idx < N * N + N * 4 + N; idx += kGenThreads) {
      if (idx < N * N) {
        const int i = idx / N, j = idx % N;
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], A[k * N + j], s);
        Ad[idx] = s;
      } else if (idx < N * N + N * 4) {
        const int e = idx - N * N, i = e / 4, j = e % 4;
        double s = 0.0;
        for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], B4[k * 4 + j], s);
        Bd4[e] = s;
      } else {
        const int i = idx - N * N - N * 4;
        double s = 0.0;
        for (int k = 0; k < N; ++k) s = fma(Acom[i * N + k], fc[k], s);
        fd[i] = s;
      }
    }
    __syncthreads();
    // ---- Update: B columns in this controller's input order (aug_lin_sys.cc:156-173); delay-line
    // contents relative to u_old (AdjustAllDelayedStates, aug_lin_sys.h:141-154) -------------------
    for (int idx = t; idx < N * 4; idx += kGenThreads) {
      const int i = idx / 4, j = idx % 4;
      const double v = Bd4[i * 4 + (K.reduced ? K.ctrl_idx[j] : j)];
      Bloc[idx] = v;
      Bsave[c * N * 5 + i * 5 + j] = v;
    }
    for (int i = t; i < N; i += kGenThreads) Bsave[c * N * 5 + i * 5 + 4] = fd[i];
    for (int idx = t; idx < 4 * dmax; idx += kGenThreads) {
      const int j = idx / dmax, tt = idx % dmax;
      double v = 0.0;
      if (tt < K.delay[j]) v = (tt == 0 ? dxg[K.head[j]] : dxg[K.chain[j] + tt - 1]) - uold[j];
      q[idx] = v;
    }
    __syncthreads();
    // ---- impulse responses of the four inputs (warps 0-3) and the free response (warp 4) -------
    if (warp < 5) {
      double* vb = vbuf + warp * 32;
      double adr[N];
      for (int k = 0; k < N; ++k) adr[k] = Ad[(lane < N ? lane : 0) * N + k];
      double v = 0.0;
      if (warp < 4 && lane < N) v = Bloc[lane * 4 + warp];
      int cur = 0;
      for (int k = 0; k < p; ++k) {
        if (warp == 4) {
          // x_{k+1} = Ad x_k + sum_j Bd_j q_j[k] + f_d; the delay chains only shift
          if (lane < N) vb[cur * 16 + lane] = v;
          __syncwarp();
          if (lane < N) {
            double s = fd[lane];
            for (int j = 0; j < 4; ++j)
              if (k < K.delay[j]) s = fma(Bloc[lane * 4 + j], q[j * dmax + k], s);
            for (int kk = 0; kk < N; ++kk) s = fma(adr[kk], vb[cur * 16 + kk], s);
            v = s;
            vb[(cur ^ 1) * 16 + lane] = v;
          }
          __syncwarp();
          if (lane < K.ny) {
            const int oy = K.out_idx[lane];
            double s = dxg[N + oy];   // the disturbance estimate of this output (C = [C_d | I])
            for (int kk = 0; kk < N; ++kk) s = fma(Cm[oy * N + kk], vb[(cur ^ 1) * 16 + kk], s);
            yfree[k * 4 + lane] = s;
          }
          __syncwarp();
        } else {
          if (lane < N) vb[cur * 16 + lane] = v;
          __syncwarp();
          if (lane < K.ny) {
            const int oy = K.out_idx[lane];
            double s = 0.0;
            for (int kk = 0; kk < N; ++kk) s = fma(Cm[oy * N + kk], vb[cur * 16 + kk], s);
            E[k * 16 + lane * 4 + warp] = s;
          }
          if (lane < N) {
            double s = 0.0;
            for (int kk = 0; kk < N; ++kk) s = fma(adr[kk], vb[cur * 16 + kk], s);
            v = s;
          }
          __syncwarp();
        }
        cur ^= 1;
      }
    }
    __syncthreads();
    // running sums of the impulse responses; w = free response - (y_ref - y)  (mpc_qp_solver.cc:31-37)
    if (t < 16) {
      double s = 0.0;
      for (int k = 0; k < p; ++k) {
        s += E[k * 16 + t];
        PE[k * 16 + t] = s;
      }
    }
    for (int idx = t; idx < p * K.ny; idx += kGenThreads) {
      const int i = idx / K.ny, yy = idx % K.ny;
      wv[i * 4 + yy] = yfree[i * 4 + yy] - (G.yref[K.yref_off + i * K.ny + yy] - misc[K.out_idx[yy]]);
    }
    __syncthreads();
    // ---- H = Su' Q Su + R, f = w' Q Su, Gx = (Q Su)' Su_other ---------------------------------
    {
      const int nv = K.nv, nu = K.nu, no = K.no, ny = K.ny, nvo = K.nvo;
      auto su = [&](int i, int yy, int mv, int ql) -> double {   // ql: local input
        const int k = i - (mv < m - 1 ? mv : m - 1) - K.delay[ql];
        if (k < 0) return 0.0;
        return (mv < m - 1 ? E : PE)[k * 16 + yy * 4 + ql];
      };
      const int n_entries = nv * nv + nv * nvo + nv;
      for (int e = t; e < n_entries; e += kGenThreads) {
        int kind, a = 0, b;
        if (e < nv * nv) { kind = 0; a = e / nv; b = e % nv; }
        else if (e < nv * nv + nv * nvo) { kind = 1; b = (e - nv * nv) / nvo; a = (e - nv * nv) % nvo; }
        else { kind = 2; b = e - nv * nv - nv * nvo; }
        const int mvb = b / nu, qb = b % nu;
        int mva = 0, qa = 0;
        if (kind == 0) { mva = a / nu; qa = a % nu; }
        if (kind == 1) { mva = a / no; qa = nu + a % no; }
        double acc = 0.0;
        for (int i = 0; i < p; ++i) {
          double sb[4];
          for (int y2 = 0; y2 < ny; ++y2) sb[y2] = su(i, y2, mvb, qb);
          for (int yy = 0; yy < ny; ++yy) {
            double qs = 0.0;
            for (int y2 = 0; y2 < ny; ++y2) qs = fma(K.Q[yy * ny + y2], sb[y2], qs);
            const double left = (kind == 2) ? wv[i * 4 + yy] : su(i, yy, mva, qa);
            acc = fma(left, qs, acc);
          }
        }
        if (kind == 0) {
          if (mva == mvb) acc += K.R[qa * nu + qb];   // u_weight_ = I_m (x) uwt  (mpc_qp_solver.h:62-80)
          Hs[c * 64 + a * nv + b] = acc;
          G.qpH[(size_t(scen) * n_ctrl + c) * 64 + a * nv + b] = acc;
        } else if (kind == 1) {
          Gxs[c * 128 + b * 16 + a] = acc;
        } else {
          fs[c * 8 + b] = acc;
          G.qpf[(size_t(scen) * n_ctrl + c) * 8 + b] = acc;
        }
      }
    }
    __syncthreads();
  }

  // ---- Jacobi sweeps: warp c solves sub-controller c (one thread; the QPs are tiny) -------------
  for (int i = t; i < P.n_pred; i += kGenThreads) plan[i] = ss[4 + i];   // du_prev = du_old_
  __syncthreads();
  {
    const int c = warp;
    const bool solver = lane == 0 && c < n_ctrl;
    QpData<8> qd;
    double Hp[64], fi[8], z[8];
    unsigned wset = kQpNoGuess, act = 0;
    double obj = 0.0;
    int status = 3;
    bool pd = false;
    if (solver) {
      const GenCtrl& K = P.c[c];
      const double* uold = G.ctrl + (size_t(scen) * n_ctrl + c) * P.state_stride + P.n + n_total + 4;
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) Hp[i * 8 + j] = (i < K.nv && j < K.nv) ? Hs[c * 64 + i * K.nv + j] : (i == j ? 1.0 : 0.0);
      for (int k = 0; k < 8; ++k) {
        const int i = k % K.nu;
        const bool on = k < K.nv;
        qd.lb[k] = on ? K.lower[i] - uold[i] : -1e30;
        qd.ub[k] = on ? K.upper[i] - uold[i] : 1e30;
        qd.lbA[k] = on ? K.rate_lower[i] : -1e30;
        qd.ubA[k] = on ? K.rate_upper[i] : 1e30;
      }
      pd = qp_invert_spd<8>(Hp, qd.J);
      wset = G.guess[size_t(scen) * n_ctrl + c];
    }
    int cur = 0;
    for (int it = 0; it < P.n_iter; ++it) {
      if (solver) {
        const GenCtrl& K = P.c[c];
        const double* dp = plan + cur * 16;
        for (int b = 0; b < 8; ++b) fi[b] = b < K.nv ? fs[c * 8 + b] : 0.0;
        if (K.reduced) {
          // du_other: the plans of the other controllers in controller order (nerve_center.h:281-286),
          // paired with the columns of Su_other by position (distributed_solver.h:109-115)
          int j = 0;
          for (int i = 0; i < P.n_pred; ++i) {
            if (i >= K.pred_off && i < K.pred_off + K.nv) continue;
            const double d = dp[i];
            for (int b = 0; b < K.nv; ++b) fi[b] = fma(Gxs[c * 128 + b * 16 + j], d, fi[b]);
            ++j;
          }
        }
        status = 3;
        if (pd) {
          switch (K.nu) {
            case 1: status = gen_qp<1>(qd, Hp, fi, &wset, z, &act, &obj); break;
            case 2: status = gen_qp<2>(qd, Hp, fi, &wset, z, &act, &obj); break;
            case 3: status = gen_qp<3>(qd, Hp, fi, &wset, z, &act, &obj); break;
            default: status = gen_qp<4>(qd, Hp, fi, &wset, z, &act, &obj); break;
          }
        }
        for (int k = 0; k < K.nv; ++k) plan[(cur ^ 1) * 16 + K.pred_off + k] = status == 0 ? z[k] : 0.0;   // zeros on failure
      }
      __syncthreads();
      cur ^= 1;
    }
    if (solver) {
      const GenCtrl& K = P.c[c];
      const size_t rec = size_t(scen) * n_ctrl + c;
      if (status == 0) G.guess[rec] = wset;
      // the solver numbers the one-sided constraints with a stride of 8 variables; reported with the
      // controller's own stride (cmpc.h: [0,nv) lower, [nv,2nv) upper, [2nv,3nv) rate >=, [3nv,4nv) rate <=)
      unsigned rep = 0;
      for (int kind = 0; kind < 4; ++kind)
        for (int i = 0; i < K.nv; ++i)
          if ((act >> (kind * 8 + i)) & 1u) rep |= 1u << (kind * K.nv + i);
      G.status[rec] = status;
      G.active[rec] = status == 0 ? rep : 0u;
      G.objective[rec] = status == 0 ? obj : 0.0;
    }
    // ---- du_old_, u_old_ (nerve_center.h:160-167,313-319) ----------------------------------------
    if (t == 0) {
      const double* dp = plan + cur * 16;
      for (int i = 0; i < P.n_pred; ++i) ss[4 + i] = dp[i];
      double un[4];
      for (int i = 0; i < 4; ++i) un[i] = ss[i];
      for (int cc = 0; cc < n_ctrl; ++cc)
        for (int i = 0; i < P.c[cc].nu; ++i) un[P.c[cc].in_off + i] += dp[P.c[cc].pred_off + i];
      for (int i = 0; i < 4; ++i) {
        du_sys[i] = -ss[i] + un[i];
        ss[i] = un[i];
        u[size_t(scen) * 4 + i] = un[i];
      }
    }
    __syncthreads();
  }
  // ---- SendU -> UpdateU -> ObserveAPriori for every sub-controller --------------------------------
  for (int c = 0; c < n_ctrl; ++c) {
    const GenCtrl& K = P.c[c];
    double* gs = G.ctrl + (size_t(scen) * n_ctrl + c) * P.state_stride;
    double* dxg = gs + N;
    double* uold = gs + N + n_total + 4;
    const double* Bl = Bsave + c * N * 5;
    double dul[4];
    for (int i = 0; i < 4; ++i) dul[i] = i < K.nu ? du_sys[K.ctrl_idx[i]] : 0.0;   // nerve_center.h:322-328
    for (int idx = t; idx < n_total; idx += kGenThreads) {
      double v = dxg[idx];   // disturbance states stay
      if (idx < N) {
        v = Bl[idx * 5 + 4];
        for (int i = 0; i < 4; ++i) {
          if (K.delay[i] == 0) v = fma(Bl[idx * 5 + i], dul[i], v);
          else v = fma(Bl[idx * 5 + i], dxg[K.head[i]] - uold[i], v);
        }
      }
      for (int i = 0; i < 4; ++i) {
        if (K.delay[i] == 0) continue;
        const int d = K.delay[i];
        if (idx == K.head[i]) v = dxg[K.chain[i]];
        else if (idx >= K.chain[i] && idx < K.chain[i] + d - 1) {
          const int tt = idx - K.chain[i];
          v = (tt == d - 2) ? uold[i] + dul[i] : dxg[idx + 1];
        }
      }
      stage[idx] = v;
    }
    __syncthreads();
    for (int idx = t; idx < n_total; idx += kGenThreads) dxg[idx] = stage[idx];
    if (t < 4) uold[t] += dul[t];
    __syncthreads();
  }
}

// NerveCenter::Initialize + DistributedController::Initialize (nerve_center.h:98-104,186-203).
__global__ void gen_init_kernel(const GenParams* __restrict__ Pp, GenState G, int NIN, const double* __restrict__ x_init,
                                const double* __restrict__ u_init, const double* __restrict__ u_init_full,
                                const double* __restrict__ y_init) {
  const GenParams& P = *Pp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  for (int c = 0; c < P.n_ctrl; ++c) {
    double* gs = G.ctrl + (size_t(b) * P.n_ctrl + c) * P.state_stride;
    for (int i = 0; i < P.state_stride; ++i) gs[i] = 0.0;
    for (int i = 0; i < P.n; ++i) gs[i] = x_init[size_t(b) * P.n + i];
    for (int i = 0; i < 4; ++i) {
      gs[P.n + P.n_total + i] = y_init[size_t(b) * 4 + i];
      gs[P.n + P.n_total + 4 + i] = u_init[size_t(b) * 4 + P.c[c].ctrl_idx[i]];
    }
    const size_t rec = size_t(b) * P.n_ctrl + c;
    G.guess[rec] = kQpNoGuess;
    G.status[rec] = 0;
    G.active[rec] = 0;
    G.objective[rec] = 0.0;
  }
  for (int i = 0; i < kGenScenStride; ++i) G.scen[size_t(b) * kGenScenStride + i] = 0.0;
  for (int i = 0; i < NIN; ++i) G.u_offset[size_t(b) * NIN + i] = u_init_full[size_t(b) * NIN + i];
}

// Start of a closed-loop run: x = x0, y = GetOutput(x0), empty actuator delay lines.
template <int PLANT>
__global__ void gen_start_kernel(const GenParams* __restrict__ Pp, const double* __restrict__ x0, ClosedLoopArrays A,
                                 double* u_init, double* u_init_full) {
  const GenParams& P = *Pp;
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  double x[N], yv[4];
  for (int i = 0; i < N; ++i) {
    x[i] = x0[size_t(b) * N + i];
    A.x[size_t(b) * N + i] = x[i];
  }
  plant_output<PLANT>(x, yv);
  for (int i = 0; i < 4; ++i) {
    A.y[size_t(b) * 4 + i] = yv[i];
    u_init[size_t(b) * 4 + i] = 0.0;
  }
  for (int i = 0; i < P.ring_total; ++i) A.ring[size_t(b) * P.ring_total + i] = 0.0;
  constexpr double udef_par[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  constexpr double udef_ser[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  for (int i = 0; i < NIN; ++i) u_init_full[size_t(b) * NIN + i] = PLANT == 0 ? udef_par[i] : udef_ser[i];
}

// Plant side of closed-loop record k (SimulationSystem::{SetOffset,SetInput,Integrate}
// simulation_system.h:66-116; TimeDelay::GetDelayedInput time_delay.h:41-58 with a ring of d
// entries per delayed input), one thread per scenario.
template <int PLANT>
__global__ void __launch_bounds__(64)
gen_advance_kernel(const GenParams* __restrict__ Pp, GenState G, int k, double t_k, ClosedLoopArrays A) {
  const GenParams& P = *Pp;
  constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN, REC = 1 + N + 8;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  double x[N], uc[4], yv[4], up[NIN];
  for (int i = 0; i < N; ++i) x[i] = A.x[size_t(b) * N + i];
  for (int i = 0; i < 4; ++i) uc[i] = A.u[size_t(b) * 4 + i];
  if (A.traj) {
    double* r = A.traj + (size_t(b) * A.n_steps + (k - A.rec_base)) * REC;
    r[0] = t_k;
    for (int i = 0; i < N; ++i) r[1 + i] = x[i];
    for (int i = 0; i < 4; ++i) r[1 + N + i] = uc[i];
    for (int i = 0; i < 4; ++i) r[1 + N + 4 + i] = A.y[size_t(b) * 4 + i];
  }
  for (int c = 0; c < P.n_ctrl; ++c) {
    const size_t oo = (size_t(b) * A.n_steps + (k - A.rec_base)) * P.n_ctrl + c;
    if (A.qp_active) A.qp_active[oo] = G.active[size_t(b) * P.n_ctrl + c];
    if (A.qp_objective) A.qp_objective[oo] = G.objective[size_t(b) * P.n_ctrl + c];
    if (A.qp_status) A.qp_status[oo] = G.status[size_t(b) * P.n_ctrl + c];
  }
  int blk = 0;
  while (blk + 1 < A.n_blocks && k >= A.block_end[b * A.n_blocks + blk]) ++blk;
  constexpr double udef_par[9] = {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
  constexpr double udef_ser[8] = {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  const double* off = A.block_off + (size_t(b) * A.n_blocks + blk) * NIN;
  for (int i = 0; i < NIN; ++i) up[i] = (PLANT == 0 ? udef_par[i] : udef_ser[i]) + off[i];
  constexpr int kInputIndex[4] = {0, 3, 4, 7};   // InputIndices of both plants
  for (int i = 0; i < 4; ++i) {
    double ud = uc[i];
    const int d = P.delays_sys[i];
    if (d > 0) {
      double* ring = A.ring + size_t(b) * P.ring_total + P.ring_off[i];
      ud = ring[k % d];
      ring[k % d] = uc[i];
    }
    up[kInputIndex[i]] += ud;
  }
  integrate_interval<PLANT>(up, x, P.Ts);
  plant_output<PLANT>(x, yv);
  for (int i = 0; i < N; ++i) A.x[size_t(b) * N + i] = x[i];
  for (int i = 0; i < 4; ++i) A.y[size_t(b) * 4 + i] = yv[i];
}

}  // namespace cmpc
