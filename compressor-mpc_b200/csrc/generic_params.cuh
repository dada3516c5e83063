// Parameter block and HBM arrays of the general configuration path (generic_kernels.cuh).
#pragma once

namespace cmpc {

constexpr int kGenMaxCtrl = 4, kGenMaxNv = 8, kGenMaxPred = 16, kGenThreads = 192;
constexpr int kGenScenStride = 24;   // u_old (4, system order) | du_old (<= 16)
constexpr int kGenMaxP = 256, kGenMaxDelay = 128;

struct GenCtrl {
  int nu, ny, nv, no, nvo, reduced;
  int out_idx[4], ctrl_idx[4];
  int delay[4];    // by LOCAL input position: the Delays argument of this controller's AugmentedLinearizedSystem
  int head[4];     // index in dx_aug of the delayed-input state of local input i (-1: not delayed)
  int chain[4];    // index in dx_aug of the first of its delay - 1 chain states
  int pred_off;    // offset of the own plan in the plan vector du (nerve_center.h:275-296)
  int in_off;      // offset of the own inputs in the system input vector (nerve_center.h:313-319)
  int yref_off;    // offset of the own p x ny reference in GenState::yref
  double Q[16], R[16], lower[4], upper[4], rate_lower[4], rate_upper[4], M[15 * 4];
};

struct GenParams {
  int plant, p, m, n_iter, batch, n_ctrl, n_pred;
  int n, n_obs, n_total, dmax;
  int state_stride;   // doubles per (scenario, controller): x_hat[n] | dx_aug[n_total] | y_old[4] | u_old[4]
  int delays_sys[4];  // per system control input: the plant side's TimeDelay
  int ring_off[4], ring_total;
  double Ts, rk[4];
  GenCtrl c[kGenMaxCtrl];
};

struct GenState {
  double* ctrl;       // [B][n_ctrl][state_stride]
  unsigned* guess;    // [B][n_ctrl] warm-start working set
  double* scen;       // [B][kGenScenStride]
  double* u_offset;   // [B][NIN]
  double* qpH;        // [B][n_ctrl][64]  (nv x nv row-major at the front)
  double* qpf;        // [B][n_ctrl][8]
  int* status;        // [B][n_ctrl]
  unsigned* active;
  double* objective;
  const double* yref; // per controller p x ny, at GenCtrl::yref_off
};

// Shared-memory footprint of gen_step_kernel in doubles (host and device agree through this).
__host__ __device__ inline int gen_smem_doubles(int n, int p, int dmax, int n_ctrl) {
  const int nn = n * n;
  return 7 * nn            // A, A2, A3, Acom, Ad + 2 spare
         + 4 * (n * 4)     // B4, Bd4, Bloc, Cm
         + 4 * n           // fc, fd, 2 spare
         + 4 * dmax        // delay-line contents
         + 5 * 2 * 16      // recurrence vectors
         + 2 * p * 16      // E, PE
         + 2 * p * 4       // free response, w
         + n_ctrl * (64 + 8 + 8 * 16 + n * 5)   // H, f, Gx, [Bloc | fd] kept for the a-priori update
         + 2 * 16 + 16 + 8   // plans, du_sys etc.
         + 256;            // a-priori staging (n_total <= 11 + 4 + 4 * 127 is capped by the host at 256)
}

}  // namespace cmpc
