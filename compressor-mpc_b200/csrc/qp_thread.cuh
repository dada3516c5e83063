// One thread per 4-variable QP, everything in registers (fully unrolled), for the Jacobi sweeps of
// the distributed controllers (include/nerve_center.h:146-158,275-296; libs/mpc_qp_solver.cc:45-75).
// The two sub-controllers of a scenario sit in neighbouring lanes and swap plans with one shuffle
// per sweep, so a warp serves 16 scenarios.
//
// H is fixed during the sweeps: with J = H^-1 and the warm-start working set W (at most 4 of the 16
// one-sided constraints, numbering of qp_dev.cuh) the equality-constrained minimiser and its
// multipliers are affine in f,
//     x = c - P f,   lambda = lam0 + Lam f,
//     P = J - J N' S^-1 N J,  c = J N' S^-1 b_W,  Lam = S^-1 N J,  lam0 = S^-1 b_W,  S = N J N'
// (built once per step); a sweep is two 4x4 mat-vecs plus the KKT check of all 16 constraints.
// When W is no longer optimal the general dual active-set solver of qp_dev.cuh runs (out of line)
// and the reduced system is rebuilt.
#pragma once
#include <cuda_runtime.h>

#include "qp_dev.cuh"

namespace cmpc {

// In-place Gauss-Jordan inverse of an SPD 4x4 matrix; false if a pivot is not positive.
__device__ __forceinline__ bool qt_inverse(double (&a)[4][4]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double piv = a[k][k];
    if (!(piv > 0.0)) ok = false;
    const double inv = 1.0 / piv;
    double rowk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rowk[j] = (j == k) ? inv : a[k][j] * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i][j] = (j == k) ? -f * inv : fma(-f, rowk[j], a[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[k][j] = rowk[j];
  }
  return ok;
}

struct QtReduced {
  double P[4][4], Lam[4][4], c[4], lam0[4];
  int q;
};

// bnd[j]: right-hand side b of constraint j (a'z >= b form), j = 4*kind + variable.
__device__ __forceinline__ bool qt_prepare(const double (&J)[4][4], const double (&bnd)[16], unsigned wset,
                                           QtReduced& r) {
  double Nm[4][4], NJ[4][4], S[4][4], bw[4];
  unsigned m = wset & 0xffffu;
  const int q = __popc(m);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int idx = __ffs(m) - 1;          // -1 when the set is exhausted
    m &= m - 1;
    const bool on = idx >= 0;
    const int kind = idx >> 2, ii = idx & 3;
    const double sg = (kind & 1) ? -1.0 : 1.0;
    double b = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) b = (on && j == idx) ? bnd[j] : b;
    bw[w] = b;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      double n = 0.0;
      if (on && l == ii) n = sg;
      if (on && kind >= 2 && ii >= 2 && l == ii - 2) n = -sg;
      Nm[w][l] = n;
    }
  }
#pragma unroll
  for (int w = 0; w < 4; ++w)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 4; ++l) s = fma(Nm[w][l], J[l][k], s);
      NJ[w][k] = s;
    }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fma(NJ[a][k], Nm[c][k], s);
      S[a][c] = (a == c && a >= q) ? 1.0 : s;   // identity padding keeps S invertible
    }
  const bool ok = qt_inverse(S);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    double l0 = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) l0 = fma(S[w][e], bw[e], l0);
    r.lam0[w] = (w < q) ? l0 : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double s = 0.0;
#pragma unroll
      for (int e = 0; e < 4; ++e) s = fma(S[w][e], NJ[e][k], s);
      r.Lam[w][k] = s;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double ck = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) ck = fma(NJ[w][k], r.lam0[w], ck);
    r.c[k] = ck;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      double s = J[k][l];
#pragma unroll
      for (int w = 0; w < 4; ++w) s = fma(-NJ[w][k], r.Lam[w][l], s);
      r.P[k][l] = s;
    }
  }
  r.q = q;
  return ok;
}

// x = c - P f, lambda = lam0 + Lam f, KKT check of the full QP.
__device__ __forceinline__ bool qt_eval(const QtReduced& r, const double (&f)[4], const double (&bnd)[16],
                                        unsigned wset, double (&x)[4], double (&lam)[4]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double s = r.c[k], l = r.lam0[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s = fma(-r.P[k][j], f[j], s);
      l = fma(r.Lam[k][j], f[j], l);
    }
    x[k] = s;
    lam[k] = l;
    if (k < r.q && !(l >= 0.0)) ok = false;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int kind = j >> 2, i = j & 3;
    double ax = x[i];
    if (kind >= 2 && i >= 2) ax -= x[i - 2];
    const double slack = ((kind & 1) ? -ax : ax) - bnd[j];
    if (!((wset >> j) & 1u) && slack < -kQpPrimalTol) ok = false;
  }
  return ok;
}

// One repair move on the working set after a failed KKT check: drop the member with the most
// negative multiplier, otherwise add the most violated constraint.  Returns the new set (the
// same set when nothing can be done, e.g. four members and still infeasible).
__device__ __forceinline__ unsigned qt_repair(const QtReduced& r, const double (&x)[4], const double (&lam)[4],
                                              const double (&bnd)[16], unsigned wset) {
  double lmin = 0.0;
  int drop = -1;
  unsigned m = wset & 0xffffu;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int idx = __ffs(m) - 1;
    m &= m - 1;
    if (idx >= 0 && lam[w] < lmin) {
      lmin = lam[w];
      drop = idx;
    }
  }
  if (drop >= 0) return wset & ~(1u << drop);
  double smin = -kQpPrimalTol;
  int add = -1;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int kind = j >> 2, i = j & 3;
    double ax = x[i];
    if (kind >= 2 && i >= 2) ax -= x[i - 2];
    const double slack = ((kind & 1) ? -ax : ax) - bnd[j];
    if (!((wset >> j) & 1u) && slack < smin) {
      smin = slack;
      add = j;
    }
  }
  if (add >= 0 && r.q < 4) return wset | (1u << add);
  return wset;
}

}  // namespace cmpc
