// ORACLE — TEST INFRASTRUCTURE ONLY (see plant.hpp).
//
// Exact solve of the reference's MPC QP (libs/mpc_qp_solver.cc:45-75):
//     min 1/2 z'Hz + f'z   s.t.  lb <= z <= ub,   lbA <= Ain z <= ubA
// with Ain from include/mpc_qp_solver.h:108-123 (row i: +1 at i, -1 at i-n_u for
// the second and later moves).  The reference hands this to qpOASES 3.2.0
// (SQProblem::hotstart, nWSR <= 10; un-vendored dependency, CMakeLists.txt:21).
// qpOASES is an exact active-set method and H is strictly convex, so its primal
// solution is THE unique minimiser; this file computes the same minimiser with
// the Goldfarb-Idnani dual active-set method (Math. Prog. 27, 1983), warm
// started from the previous optimal working set like qpOASES's hotstart.
// "parity unpinned" at QP level (the reference ships no QP known-answer test);
// pinned end-to-end by the golden closed-loop trajectories.
//
// One-sided constraint numbering used for active-set reporting (4*nv bits):
//   [0,nv)     z_j >= lb_j          [nv,2nv)   z_j <= ub_j
//   [2nv,3nv)  (Ain z)_j >= lbA_j   [3nv,4nv)  (Ain z)_j <= ubA_j
#pragma once
#include <cmath>
#include <limits>

namespace oracle {

constexpr int kQpMaxVars = 8;
constexpr int kQpIterationCap = 200;
constexpr double kQpPrimalTol = 1e-11;

struct QpWorkspace {
  bool has_guess = false;
  unsigned guess = 0;  // working set of the previous successful solve
  void Reset() {
    has_guess = false;
    guess = 0;
  }
};

namespace qpdetail {

struct Problem {
  int nv, nu;
  const double *H, *f, *lb, *ub, *lbA, *ubA;
  double J[kQpMaxVars][kQpMaxVars];  // H^-1
};

// normal a_j (as dense row) and right-hand side b_j of  a_j'z >= b_j
inline void Constraint(const Problem& P, int j, double a[kQpMaxVars], double* b) {
  const int nv = P.nv, kind = j / nv, i = j % nv;
  for (int k = 0; k < nv; ++k) a[k] = 0;
  const double sgn = (kind & 1) ? -1.0 : 1.0;
  a[i] = sgn;
  if (kind >= 2 && i >= P.nu) a[i - P.nu] = -sgn;
  switch (kind) {
    case 0: *b = P.lb[i]; break;
    case 1: *b = -P.ub[i]; break;
    case 2: *b = P.lbA[i]; break;
    default: *b = -P.ubA[i]; break;
  }
}

inline bool InvertSpd(int n, const double* H, double J[kQpMaxVars][kQpMaxVars]) {
  double L[kQpMaxVars][kQpMaxVars] = {};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = H[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i][i] = std::sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double Li[kQpMaxVars][kQpMaxVars] = {};  // L^-1 (lower)
  for (int c = 0; c < n; ++c)
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i][k] * Li[k][c];
      Li[i][c] = s / L[i][i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = (i > j ? i : j); k < n; ++k) s += Li[k][i] * Li[k][j];
      J[i][j] = s;
    }
  return true;
}

// Solve S r = rhs for SPD S (q<=8) by Cholesky; false when not positive definite.
inline bool SolveSpd(int q, double S[kQpMaxVars][kQpMaxVars], const double* rhs, double* r) {
  double L[kQpMaxVars][kQpMaxVars] = {};
  for (int i = 0; i < q; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = S[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i][i] = std::sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double y[kQpMaxVars];
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

struct WorkingSet {
  int q = 0;
  int idx[kQpMaxVars];
  double N[kQpMaxVars][kQpMaxVars];   // normals (rows)
  double b[kQpMaxVars];
  double JN[kQpMaxVars][kQpMaxVars];  // J * N^T  (column w = J a_w), stored [w][k]
  double S[kQpMaxVars][kQpMaxVars];   // N J N^T
};

inline void BuildWorkingSet(const Problem& P, WorkingSet* W) {
  const int nv = P.nv;
  for (int w = 0; w < W->q; ++w) {
    Constraint(P, W->idx[w], W->N[w], &W->b[w]);
    for (int k = 0; k < nv; ++k) {
      double s = 0;
      for (int l = 0; l < nv; ++l) s += P.J[k][l] * W->N[w][l];
      W->JN[w][k] = s;
    }
  }
  for (int a = 0; a < W->q; ++a)
    for (int c = 0; c < W->q; ++c) {
      double s = 0;
      for (int k = 0; k < nv; ++k) s += W->N[a][k] * W->JN[c][k];
      W->S[a][c] = s;
    }
}

inline void DropFromWorkingSet(WorkingSet* W, double* u, int l) {
  for (int w = l; w + 1 < W->q; ++w) {
    W->idx[w] = W->idx[w + 1];
    u[w] = u[w + 1];
  }
  W->q--;
}

}  // namespace qpdetail

// Returns 0 on success, 1 iteration cap, 2 infeasible, 3 H not positive definite.
// active: bitmask of constraints with a strictly positive multiplier.
inline int SolveMpcQp(int nv, int nu, const double* H, const double* f, const double* lb,
                      const double* ub, const double* lbA, const double* ubA, QpWorkspace* ws,
                      double* z_out, unsigned* active, double* objective, int* iterations) {
  using namespace qpdetail;
  Problem P{nv, nu, H, f, lb, ub, lbA, ubA, {}};
  *active = 0;
  *objective = 0;
  *iterations = 0;
  if (!InvertSpd(nv, H, P.J)) return 3;
  const int nc = 4 * nv;
  double x0[kQpMaxVars];
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int k = 0; k < nv; ++k) s -= P.J[i][k] * f[k];
    x0[i] = s;
  }
  double fmax = 1.0;
  for (int i = 0; i < nv; ++i) fmax = std::fmax(fmax, std::fabs(f[i]));

  double x[kQpMaxVars], u[kQpMaxVars + 1];
  WorkingSet W;
  bool solved = false;

  // ---- warm start: is the previous working set still optimal? ----
  if (ws && ws->has_guess) {
    W.q = 0;
    for (int j = 0; j < nc && W.q < nv; ++j)
      if (ws->guess >> j & 1u) W.idx[W.q++] = j;
    BuildWorkingSet(P, &W);
    double rhs[kQpMaxVars];
    for (int w = 0; w < W.q; ++w) {
      double s = W.b[w];
      for (int k = 0; k < nv; ++k) s -= W.N[w][k] * x0[k];
      rhs[w] = s;
    }
    bool ok = (W.q == 0) || SolveSpd(W.q, W.S, rhs, u);
    if (ok) {
      for (int k = 0; k < nv; ++k) {
        double s = x0[k];
        for (int w = 0; w < W.q; ++w) s += W.JN[w][k] * u[w];
        x[k] = s;
      }
      for (int w = 0; w < W.q; ++w)
        if (!(u[w] >= 0)) ok = false;
      unsigned inW = 0;
      for (int w = 0; w < W.q; ++w) inW |= 1u << W.idx[w];
      for (int j = 0; j < nc && ok; ++j) {
        if (inW >> j & 1u) continue;
        double a[kQpMaxVars], b;
        Constraint(P, j, a, &b);
        double s = -b;
        for (int k = 0; k < nv; ++k) s += a[k] * x[k];
        if (s < -kQpPrimalTol) ok = false;
      }
      solved = ok;
    }
  }

  // ---- cold Goldfarb-Idnani ----
  if (!solved) {
    W.q = 0;
    for (int k = 0; k < nv; ++k) x[k] = x0[k];
    int iter = 0;
    for (;;) {
      unsigned inW = 0;
      for (int w = 0; w < W.q; ++w) inW |= 1u << W.idx[w];
      int p = -1;
      double sp = -kQpPrimalTol;
      double ap[kQpMaxVars], bp = 0;
      for (int j = 0; j < nc; ++j) {
        if (inW >> j & 1u) continue;
        double a[kQpMaxVars], b;
        Constraint(P, j, a, &b);
        double s = -b;
        for (int k = 0; k < nv; ++k) s += a[k] * x[k];
        if (s < sp) {
          sp = s;
          p = j;
        }
      }
      if (p < 0) break;  // primal feasible: optimal
      Constraint(P, p, ap, &bp);
      double up = 0;
      for (;;) {
        if (++iter > kQpIterationCap) return 1;
        BuildWorkingSet(P, &W);
        double d[kQpMaxVars], zdir[kQpMaxVars], r[kQpMaxVars] = {};
        for (int k = 0; k < nv; ++k) {
          double s = 0;
          for (int l = 0; l < nv; ++l) s += P.J[k][l] * ap[l];
          d[k] = s;
          zdir[k] = s;
        }
        bool dependent = (W.q >= nv);
        if (W.q > 0) {
          double rhs[kQpMaxVars];
          for (int w = 0; w < W.q; ++w) {
            double s = 0;
            for (int k = 0; k < nv; ++k) s += W.N[w][k] * d[k];
            rhs[w] = s;
          }
          if (!SolveSpd(W.q, W.S, rhs, r)) return 3;
          for (int k = 0; k < nv; ++k) {
            double s = d[k];
            for (int w = 0; w < W.q; ++w) s -= W.JN[w][k] * r[w];
            zdir[k] = s;
          }
        }
        double zn = 0, dn = 0;
        for (int k = 0; k < nv; ++k) {
          zn += zdir[k] * ap[k];
          dn += d[k] * ap[k];
        }
        if (zn <= 1e-13 * dn) dependent = true;
        double t1 = std::numeric_limits<double>::infinity();
        int l = -1;
        for (int w = 0; w < W.q; ++w)
          if (r[w] > 0 && u[w] / r[w] < t1) {
            t1 = u[w] / r[w];
            l = w;
          }
        const double t2 = dependent ? std::numeric_limits<double>::infinity() : -sp / zn;
        const double t = t1 < t2 ? t1 : t2;
        if (!(t < std::numeric_limits<double>::infinity())) return 2;
        for (int w = 0; w < W.q; ++w) u[w] -= t * r[w];
        up += t;
        if (dependent) {
          DropFromWorkingSet(&W, u, l);
          continue;
        }
        for (int k = 0; k < nv; ++k) x[k] += t * zdir[k];
        if (t2 <= t1) {
          W.idx[W.q] = p;
          u[W.q] = up;
          W.q++;
          break;
        }
        DropFromWorkingSet(&W, u, l);
        sp = -bp;
        for (int k = 0; k < nv; ++k) sp += ap[k] * x[k];
      }
    }
    *iterations = iter;
  }

  unsigned wset = 0, act = 0;
  for (int w = 0; w < W.q; ++w) {
    wset |= 1u << W.idx[w];
    if (u[w] > 1e-9 * fmax) act |= 1u << W.idx[w];
  }
  if (ws) {
    ws->has_guess = true;
    ws->guess = wset;
  }
  double obj = 0;
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int k = 0; k < nv; ++k) s += H[i * nv + k] * x[k];
    obj += x[i] * (0.5 * s + f[i]);
    z_out[i] = x[i];
  }
  *active = act;
  *objective = obj;
  return 0;
}

}  // namespace oracle
