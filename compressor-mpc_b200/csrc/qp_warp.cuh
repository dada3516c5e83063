// Warp-cooperative solve of the two sub-controllers' 4-variable QPs during the Jacobi sweeps
// (include/nerve_center.h:146-158,275-296; libs/mpc_qp_solver.cc:45-75).
//
// One warp, 16 lanes per sub-controller.  Lane l of a half warp is element (i, j) =
// (l >> 2, l & 3) of every 4 x 4 matrix; the same 16 lanes are also the 16 one-sided
// constraints of that QP (qp_dev.cuh numbering).  Everything lives in registers and moves by
// shuffles: H^-1 by Gauss-Jordan, the reduced system of the warm-start working set W
//     x = c - P f,  lambda = lam0 + Lam f      (see qp_dev.cuh, QpFastLayout)
// once per control step, and per sweep two 4 x 4 mat-vecs, the KKT check of all 16 constraints
// and the exchange of the plans between the two halves (Jacobi: both read the old plans).
// When W is no longer optimal (rare), lane 0 of that half runs the general dual active-set
// solver of qp_dev.cuh and the reduced system is rebuilt for the new W.
#pragma once
#include <cuda_runtime.h>

#include "qp_dev.cuh"

namespace cmpc {

constexpr unsigned kFullMask = 0xffffffffu;

struct QpLane {
  int half, i, j, base;  // base = first lane of this half
};

// shuffle within the half warp: value held by element (r, c)
__device__ __forceinline__ double qw_get(double v, const QpLane& L, int r, int c) {
  return __shfl_sync(kFullMask, v, L.base + 4 * r + c);
}
// sum over j (the 4 lanes of one row); result on all lanes of the row
__device__ __forceinline__ double qw_row_sum(double v) {
  v += __shfl_xor_sync(kFullMask, v, 1);
  v += __shfl_xor_sync(kFullMask, v, 2);
  return v;
}
// C = A * B
__device__ __forceinline__ double qw_mm(double a, double b, const QpLane& L) {
  double s = 0.0;
#pragma unroll
  for (int l = 0; l < 4; ++l) s = fma(qw_get(a, L, L.i, l), qw_get(b, L, l, L.j), s);
  return s;
}
// C = A * B'
__device__ __forceinline__ double qw_mm_nt(double a, double b, const QpLane& L) {
  double s = 0.0;
#pragma unroll
  for (int l = 0; l < 4; ++l) s = fma(qw_get(a, L, L.i, l), qw_get(b, L, L.j, l), s);
  return s;
}
// C = A' * B
__device__ __forceinline__ double qw_mm_tn(double a, double b, const QpLane& L) {
  double s = 0.0;
#pragma unroll
  for (int l = 0; l < 4; ++l) s = fma(qw_get(a, L, l, L.i), qw_get(b, L, l, L.j), s);
  return s;
}
// In-place Gauss-Jordan inverse of an SPD 4 x 4 matrix (no pivoting); ok = all pivots > 0.
__device__ __forceinline__ double qw_inverse(double a, const QpLane& L, bool* ok) {
  double b = (L.i == L.j) ? 1.0 : 0.0;
  bool good = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double piv = qw_get(a, L, k, k);
    if (!(piv > 0.0)) good = false;
    const double inv = 1.0 / piv;
    const double akj = qw_get(a, L, k, L.j) * inv;
    const double bkj = qw_get(b, L, k, L.j) * inv;
    const double aik = qw_get(a, L, L.i, k);
    if (L.i == k) {
      a = akj;
      b = bkj;
    } else {
      a = fma(-aik, akj, a);
      b = fma(-aik, bkj, b);
    }
  }
  *ok = good;
  return b;
}

// Per-half state of the reduced system (one matrix element or vector entry per lane).
struct QpReduced {
  double P, Lam;    // element (i, j)
  double c_row;     // c[i]      (same on the 4 lanes of row i)
  double lam0_row;  // lam0[i]
  int q;            // |W|
};

// Build the reduced system for working set `wset` (bitmask over the 16 constraints, at most 4
// members, ordered by index).  J = H^-1 element (i, j); bound[l] = right-hand side b of
// constraint l (this lane's own constraint, a'z >= b form).  Returns false if S is singular.
__device__ __forceinline__ bool qw_prepare(double J, double bnd, unsigned wset, const QpLane& L, QpReduced* out) {
  // row i of N = normal of the i-th member of W (zero row if i >= q)
  const int q = __popc(wset & 0xffffu);
  int idx = -1;
  {
    unsigned m = wset & 0xffffu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int b = __ffs(m) - 1;
      if (w == L.i && m) idx = b;
      m &= m - 1;
    }
  }
  const bool row_on = idx >= 0;
  double n = 0.0, b_row = 0.0;
  {
    const int kind = idx >> 2, ii = idx & 3;
    const double sgn_ = (kind & 1) ? -1.0 : 1.0;
    if (row_on) {
      if (L.j == ii) n = sgn_;
      if (kind >= 2 && ii >= 2 && L.j == ii - 2) n = -sgn_;
    }
    const double bsrc = __shfl_sync(kFullMask, bnd, L.base + (row_on ? idx : 0));
    b_row = row_on ? bsrc : 0.0;
  }
  const double NJ = qw_mm(n, J, L);            // (N J)[w][k], rows w >= q are zero
  double S = qw_mm_nt(NJ, n, L);               // N J N'
  if (L.i == L.j && L.i >= q) S = 1.0;         // identity padding keeps S invertible
  bool ok;
  const double Sinv = qw_inverse(S, L, &ok);
  const double Lam = qw_mm(Sinv, NJ, L);       // rows w >= q are zero
  // lam0[w] = sum_e Sinv[w][e] b[e]: lane (w, e) needs b[e] = b_row of row e
  const double b_col = qw_get(b_row, L, L.j, 0);
  const double lam0 = qw_row_sum(Sinv * b_col);
  const double P = J - qw_mm_tn(NJ, Lam, L);   // J - J N' S^-1 N J
  // c[k] = sum_w NJ[w][k] lam0[w]: lane (k, w) takes NJ[w][k] and lam0 of row w
  const double c = qw_row_sum(qw_get(NJ, L, L.j, L.i) * qw_get(lam0, L, L.j, 0));
  out->P = P;
  out->Lam = Lam;
  out->c_row = c;
  out->lam0_row = (L.i < q) ? lam0 : 0.0;
  out->q = q;
  return ok;
}

// One sweep on the reduced system: x = c - P f, lambda = lam0 + Lam f, then the KKT check.
// f_row: f[i] on the lanes of row i.  Returns (per half) whether W is still optimal.
__device__ __forceinline__ bool qw_eval(const QpReduced& r, double f_row, double bnd, unsigned wset,
                                        const QpLane& L, double* x_row, double* lam_row) {
  const double f_col = qw_get(f_row, L, L.j, 0);
  const double x = r.c_row - qw_row_sum(r.P * f_col);
  const double lam = r.lam0_row + qw_row_sum(r.Lam * f_col);
  *x_row = x;
  *lam_row = lam;
  const int l = L.i * 4 + L.j, kind = l >> 2, ii = l & 3;
  double ax = qw_get(x, L, ii, 0);
  const double xm = qw_get(x, L, (ii + 2) & 3, 0);
  if (kind >= 2 && ii >= 2) ax -= xm;
  const double slack = ((kind & 1) ? -ax : ax) - bnd;
  const bool in_w = (wset >> l) & 1u;
  const bool ok = (in_w || slack >= -kQpPrimalTol) && (L.i >= r.q || lam >= 0.0);
  const unsigned bal = __ballot_sync(kFullMask, ok);
  return ((bal >> L.base) & 0xffffu) == 0xffffu;
}

}  // namespace cmpc
