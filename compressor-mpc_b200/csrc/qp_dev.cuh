// Device-side exact solve of the MPC QP that the reference hands to qpOASES
// (libs/mpc_qp_solver.cc:45-75, constraint rows include/mpc_qp_solver.h:108-123):
//     min 1/2 z'Hz + f'z   s.t.  lb <= z <= ub,   lbA <= Ain z <= ubA.
// One GPU thread per QP; nv = 4 (sub-controller) or 8 (centralised).  H is strictly
// convex, so the minimiser is unique: a warm start from the previous optimal working
// set is verified through the KKT conditions (the common case: one small solve), and
// only when that fails a cold Goldfarb-Idnani dual active-set pass runs.
// Constraint numbering (4*NV one-sided constraints), used for active-set reporting:
//   [0,NV) z_j >= lb_j   [NV,2NV) z_j <= ub_j   [2NV,3NV) (Ain z)_j >= lbA_j   [3NV,4NV) (Ain z)_j <= ubA_j
#pragma once
#include <cuda_runtime.h>

namespace cmpc {

constexpr int kQpIterationCap = 200;
constexpr double kQpPrimalTol = 1e-11;
constexpr unsigned kQpNoGuess = 0xFFFFFFFFu;

template <int NV>
struct QpData {
  double J[NV][NV];  // H^-1
  double lb[NV], ub[NV], lbA[NV], ubA[NV];
};

// a_j' x - b_j for constraint j
template <int NV, int NU>
__device__ __forceinline__ double qp_slack(const QpData<NV>& P, int j, const double* x) {
  const int kind = j / NV, i = j % NV;
  double ax = x[i];
  if (kind >= 2 && i >= NU) ax -= x[i - NU];
  switch (kind) {
    case 0: return ax - P.lb[i];
    case 1: return P.ub[i] - ax;
    case 2: return ax - P.lbA[i];
    default: return P.ubA[i] - ax;
  }
}
template <int NV, int NU>
__device__ __forceinline__ void qp_normal(const QpData<NV>& P, int j, double* a, double* b) {
  const int kind = j / NV, i = j % NV;
#pragma unroll
  for (int k = 0; k < NV; ++k) a[k] = 0.0;
  const double s = (kind & 1) ? -1.0 : 1.0;
  a[i] = s;
  if (kind >= 2 && i >= NU) a[i - NU] = -s;
  *b = (kind == 0) ? P.lb[i] : (kind == 1) ? -P.ub[i] : (kind == 2) ? P.lbA[i] : -P.ubA[i];
}

// J = H^-1 through Cholesky; false when H is not positive definite.
template <int NV>
__device__ bool qp_invert_spd(const double* H /*NV*NV row-major*/, double J[NV][NV]) {
  double L[NV][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) L[i][j] = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = H[i * NV + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double Li[NV][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) Li[i][j] = 0.0;
#pragma unroll
  for (int c = 0; c < NV; ++c)
#pragma unroll
    for (int i = c; i < NV; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = c; k < i; ++k) s -= L[i][k] * Li[k][c];
      Li[i][c] = s / L[i][i];
    }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (k >= i && k >= j) s += Li[k][i] * Li[k][j];
      J[i][j] = s;
    }
  return true;
}

template <int NV>
struct QpWorkingSet {
  int q;
  int idx[NV];
  double N[NV][NV], b[NV], JN[NV][NV], S[NV][NV];
};

template <int NV, int NU>
__device__ void qp_build_ws(const QpData<NV>& P, QpWorkingSet<NV>& W) {
  for (int w = 0; w < W.q; ++w) {
    qp_normal<NV, NU>(P, W.idx[w], W.N[w], &W.b[w]);
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int l = 0; l < NV; ++l) s += P.J[k][l] * W.N[w][l];
      W.JN[w][k] = s;
    }
  }
  for (int a = 0; a < W.q; ++a)
    for (int c = 0; c < W.q; ++c) {
      double s = 0.0;
      for (int k = 0; k < NV; ++k) s += W.N[a][k] * W.JN[c][k];
      W.S[a][c] = s;
    }
}

// Solve S r = rhs (S SPD, q x q) by Cholesky.
template <int NV>
__device__ bool qp_solve_spd(int q, const double S[NV][NV], const double* rhs, double* r) {
  double L[NV][NV];
  for (int i = 0; i < q; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = S[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double y[NV];
  for (int i = 0; i < q; ++i) {
    double s = rhs[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = q - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < q; ++k) s -= L[k][i] * r[k];
    r[i] = s / L[i][i];
  }
  return true;
}

template <int NV>
__device__ __forceinline__ void qp_drop(QpWorkingSet<NV>& W, double* u, int l) {
  for (int w = l; w + 1 < W.q; ++w) {
    W.idx[w] = W.idx[w + 1];
    u[w] = u[w + 1];
  }
  W.q--;
}

// Returns 0 ok, 1 iteration cap, 2 infeasible, 3 factorisation failure.  On failure z = 0
// (the reference returns zeros whenever qpOASES does not report success).
// guess: previous working set (kQpNoGuess = cold); updated to the final working set.
template <int NV, int NU>
__device__ __noinline__ int qp_solve(const QpData<NV>& P, const double* H, const double* f, unsigned* guess,
                                     double* z, unsigned* active, double* objective) {
  constexpr int NC = 4 * NV;
  double x0[NV], x[NV], u[NV + 1];
  double fmax = 1.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) s -= P.J[i][k] * f[k];
    x0[i] = s;
    fmax = fmax > fabs(f[i]) ? fmax : fabs(f[i]);
  }
  QpWorkingSet<NV> W;
  bool solved = false;
  int status = 0;

  if (*guess != kQpNoGuess) {
    W.q = 0;
    for (int j = 0; j < NC && W.q < NV; ++j)
      if ((*guess >> j) & 1u) W.idx[W.q++] = j;
    qp_build_ws<NV, NU>(P, W);
    double rhs[NV];
    for (int w = 0; w < W.q; ++w) {
      double s = W.b[w];
      for (int k = 0; k < NV; ++k) s -= W.N[w][k] * x0[k];
      rhs[w] = s;
    }
    bool ok = (W.q == 0) || qp_solve_spd<NV>(W.q, W.S, rhs, u);
    if (ok) {
      for (int k = 0; k < NV; ++k) {
        double s = x0[k];
        for (int w = 0; w < W.q; ++w) s += W.JN[w][k] * u[w];
        x[k] = s;
      }
      unsigned inW = 0;
      for (int w = 0; w < W.q; ++w) {
        if (!(u[w] >= 0.0)) ok = false;
        inW |= 1u << W.idx[w];
      }
      for (int j = 0; j < NC; ++j) {
        if ((inW >> j) & 1u) continue;
        if (qp_slack<NV, NU>(P, j, x) < -kQpPrimalTol) ok = false;
      }
      solved = ok;
    }
  }

  if (!solved) {
    W.q = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = x0[k];
    int iter = 0;
    for (;;) {
      unsigned inW = 0;
      for (int w = 0; w < W.q; ++w) inW |= 1u << W.idx[w];
      int p = -1;
      double sp = -kQpPrimalTol;
      for (int j = 0; j < NC; ++j) {
        if ((inW >> j) & 1u) continue;
        const double s = qp_slack<NV, NU>(P, j, x);
        if (s < sp) {
          sp = s;
          p = j;
        }
      }
      if (p < 0) break;
      double ap[NV], bp;
      qp_normal<NV, NU>(P, p, ap, &bp);
      double up = 0.0;
      for (;;) {
        if (++iter > kQpIterationCap) { status = 1; break; }
        qp_build_ws<NV, NU>(P, W);
        double d[NV], zd[NV], r[NV];
        for (int k = 0; k < NV; ++k) {
          double s = 0.0;
          for (int l = 0; l < NV; ++l) s += P.J[k][l] * ap[l];
          d[k] = s;
          zd[k] = s;
          r[k] = 0.0;
        }
        bool dependent = (W.q >= NV);
        if (W.q > 0) {
          double rhs[NV];
          for (int w = 0; w < W.q; ++w) {
            double s = 0.0;
            for (int k = 0; k < NV; ++k) s += W.N[w][k] * d[k];
            rhs[w] = s;
          }
          if (!qp_solve_spd<NV>(W.q, W.S, rhs, r)) { status = 3; break; }
          for (int k = 0; k < NV; ++k) {
            double s = d[k];
            for (int w = 0; w < W.q; ++w) s -= W.JN[w][k] * r[w];
            zd[k] = s;
          }
        }
        double zn = 0.0, dn = 0.0;
        for (int k = 0; k < NV; ++k) {
          zn += zd[k] * ap[k];
          dn += d[k] * ap[k];
        }
        if (zn <= 1e-13 * dn) dependent = true;
        double t1 = CUDART_INF;
        int l = -1;
        for (int w = 0; w < W.q; ++w)
          if (r[w] > 0.0 && u[w] / r[w] < t1) {
            t1 = u[w] / r[w];
            l = w;
          }
        const double t2 = dependent ? CUDART_INF : -sp / zn;
        const double t = t1 < t2 ? t1 : t2;
        if (!(t < CUDART_INF)) { status = 2; break; }
        for (int w = 0; w < W.q; ++w) u[w] -= t * r[w];
        up += t;
        if (dependent) {
          qp_drop<NV>(W, u, l);
          continue;
        }
        for (int k = 0; k < NV; ++k) x[k] += t * zd[k];
        if (t2 <= t1) {
          W.idx[W.q] = p;
          u[W.q] = up;
          W.q++;
          break;
        }
        qp_drop<NV>(W, u, l);
        sp = qp_slack<NV, NU>(P, p, x);
      }
      if (status != 0) break;
    }
  }

  if (status != 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) z[i] = 0.0;
    *active = 0;
    *objective = 0.0;
    return status;
  }
  unsigned wset = 0, act = 0;
  for (int w = 0; w < W.q; ++w) {
    wset |= 1u << W.idx[w];
    if (u[w] > 1e-9 * fmax) act |= 1u << W.idx[w];
  }
  *guess = wset;
  double obj = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) s += H[i * NV + k] * x[k];
    obj += x[i] * (0.5 * s + f[i]);
    z[i] = x[i];
  }
  *active = act;
  *objective = obj;
  return 0;
}

}  // namespace cmpc
