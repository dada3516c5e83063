// ORACLE — TEST INFRASTRUCTURE ONLY (see plant.hpp).  Literal CPU restatement
// of the reference's per-sample-time control step.  Matrices are materialised
// exactly as the reference does (Su, Sx, Sf, Su_other, YSu) so that each
// function can be compared one-to-one; no algebraic shortcuts are taken here.
//
// Restates, without Eigen/qpOASES:
//   libs/aug_lin_sys.cc:145-177,232-255,260-334   Update / DiscretizeRK4 / GeneratePrediction
//   libs/aug_lin_sys.cc:27-57,62-86,91-113,118-140,182-227   AComposite / BComposite
//   include/aug_lin_sys.h:129-163,235-252         Adjust* helpers, AComposite::operator*
//   include/mpc_qp_solver.h:62-80,108-123          SetWeights / GetConstraintMatrix
//   libs/mpc_qp_solver.cc:19-40,45-75              GenerateQP / SolveQP
//   include/distributed_solver.h:83-121            GenerateDistributedQP / ApplyOtherInput
//   libs/observer.cc:6-19,24-40                    ObserveAPriori / ObserveAPosteriori
//   libs/distributed_controller.cc:27-67,72-108    Initialize / GenerateInitialQP
//   include/distributed_controller.h:146-152,206-226   UpdateU / GetInput
//   include/nerve_center.h:98-122,134-182,186-328  NerveCenter
//
// The QP solve itself lives in qpOASES 3.2.0 (CMakeLists.txt:21), which is not
// vendored in the reference.  H is strictly convex (R > 0), so the minimiser
// is unique; qp.hpp restates it as an exact dual active-set (Goldfarb-Idnani)
// solve.  Parity across that boundary is pinned end-to-end by the reference's
// six golden closed-loop trajectories (tests/golden/).
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <vector>

#include "plant.hpp"
#include "qp.hpp"

namespace oracle {

// Column-major dense matrix, like Eigen::MatrixXd (prediction.h:11-17).
struct MatX {
  int rows = 0, cols = 0;
  std::vector<double> a;
  void Resize(int r, int c) {
    rows = r;
    cols = c;
    a.assign(static_cast<size_t>(r) * c, 0.0);
  }
  double& operator()(int i, int j) { return a[static_cast<size_t>(j) * rows + i]; }
  double operator()(int i, int j) const { return a[static_cast<size_t>(j) * rows + i]; }
};

struct Prediction {
  MatX Su, Sx, Sf, Su_other;
};

// Compile-time configuration of one sub-controller in the reference
// (parallel_compressors_constants.h:70-93, serial_compressors_constants.h:84-109),
// turned into a runtime struct.
struct ControllerConfig {
  int n_sub_control_inputs = 2;      // own inputs (4 => centralised, not "reduced")
  int control_input_indices[4] = {0, 1, 2, 3};  // permutation of system control inputs
  int n_controlled_outputs = 3;
  int controlled_output_indices[4] = {0, 1, 3, 0};
  // The Delays template argument of this controller's AugmentedLinearizedSystem, by LOCAL input
  // position (aug_lin_sys.h:34-40, aug_lin_sys.cc:156-173).  The reference's two plants hand the
  // same ConstexprArray<0,40,0,40> to every sub-controller because their permutations map delayed
  // inputs onto delayed inputs; a configuration with unequal delays instantiates each
  // AugmentedLinearizedSystem with the delays of its own input order.  delays[0] < 0: use
  // SystemConfig::delays as they are (the reference's own instantiations).
  int delays[4] = {-1, -1, -1, -1};
};

constexpr int kMaxControllers = 4;   // NerveCenter takes a parameter pack of any size (nerve_center.h:19-38)

struct SystemConfig {
  PlantKind plant = kParallel;
  int delays[4] = {0, 40, 0, 40};
  int n_disturbance_states = 4;
  int p = 100, m = 2;
  double Ts = 0.05;
  int n_controllers = 2;
  int n_solver_iterations = 9;
  ControllerConfig ctrl[kMaxControllers];
};

// ---------------------------------------------------------------------------
// AugmentedLinearizedSystem  (include/aug_lin_sys.h:34-223)
// ---------------------------------------------------------------------------
class AugLinSys {
 public:
  int n_states, n_control_inputs = 4, n_outputs = 4;
  int n_delay_states, n_delayed_inputs, n_disturbance_states;
  int n_sub_control_inputs, n_other_control_inputs;
  int n_aug_states, n_obs_states, n_total_states;
  bool is_reduced;
  int n_delay_[4];
  int ctrl_idx_[4];

  // AComposite / BComposite (aug_lin_sys.h:166-204)
  std::vector<double> Aorig;   // n×n row-major
  std::vector<double> Adelay;  // n×n_delayed_inputs row-major
  std::vector<int> Aaug;       // n_aug
  std::vector<double> Borig;   // n×(n_ci - n_delayed) row-major
  int Baug[4];
  std::vector<double> C;       // n_outputs × n_obs row-major
  std::vector<double> f;       // n
  Plant sys_;
  double sampling_time_;

  AugLinSys(const SystemConfig& sc, const ControllerConfig& cc)
      : sys_(sc.plant), sampling_time_(sc.Ts) {
    n_states = sys_.n_states;
    n_disturbance_states = sc.n_disturbance_states;
    n_delay_states = 0;
    n_delayed_inputs = 0;
    for (int i = 0; i < 4; ++i) {
      n_delay_[i] = cc.delays[0] < 0 ? sc.delays[i] : cc.delays[i];
      ctrl_idx_[i] = cc.control_input_indices[i];
      n_delay_states += n_delay_[i];
      if (n_delay_[i] != 0) n_delayed_inputs++;
    }
    n_sub_control_inputs = cc.n_sub_control_inputs;
    is_reduced = n_sub_control_inputs != n_control_inputs;
    n_other_control_inputs = n_control_inputs - n_sub_control_inputs;
    n_aug_states = n_disturbance_states + n_delay_states;
    n_obs_states = n_states + n_disturbance_states;
    n_total_states = n_aug_states + n_states;
    const int n = n_states;
    Aorig.assign(n * n, 0.0);
    Adelay.assign(n * n_delayed_inputs, 0.0);
    Borig.assign(n * (n_control_inputs - n_delayed_inputs), 0.0);
    f.assign(n, 0.0);
    // ctor: C = [0 | I]  (aug_lin_sys.cc:11-22)
    C.assign(n_outputs * n_obs_states, 0.0);
    for (int i = 0; i < n_outputs && i < n_disturbance_states; ++i)
      C[i * n_obs_states + n_states + i] = 1.0;
    // AComposite ctor (aug_lin_sys.cc:27-57)
    Aaug.assign(n_aug_states, 0);
    for (int i = 0; i < n_disturbance_states; ++i) Aaug[i] = i;
    for (int i = 0; i < n_delayed_inputs; ++i) Aaug[n_disturbance_states + i] = -1;
    int index_delay_states = n_disturbance_states + n_delayed_inputs;
    int index_delayed_inputs = n_disturbance_states;
    for (int i = 0; i < n_control_inputs; ++i) {
      if (n_delay_[i] != 0) {
        const int size_block = n_delay_[i] - 1;
        Aaug[index_delay_states] = index_delayed_inputs;
        for (int j = 1; j < size_block; ++j) Aaug[index_delay_states + j] = index_delay_states + j - 1;
        index_delay_states += n_delay_[i] - 1;
        index_delayed_inputs++;
      }
    }
    // BComposite ctor (aug_lin_sys.cc:182-199)
    int idx = n_delayed_inputs;
    for (int i = 0; i < n_control_inputs; ++i) {
      if (n_delay_[i] != 0) {
        idx += n_delay_[i] - 1;
        Baug[i] = n_states + n_disturbance_states + idx - 1;
      } else {
        Baug[i] = -1;
      }
    }
  }

  // aug_lin_sys.cc:232-255
  static void DiscretizeRK4(const PlantLin& c, double Ts, PlantLin* d) {
    const int n = c.n;
    d->Resize(n);
    std::vector<double> A2(n * n, 0.0), A3(n * n, 0.0), Ac(n * n, 0.0);
    auto mm = [n](const std::vector<double>& X, const std::vector<double>& Y, int yc,
                  std::vector<double>* Z) {
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < yc; ++j) {
          double s = 0;
          for (int k = 0; k < n; ++k) s += X[i * n + k] * Y[k * yc + j];
          (*Z)[i * yc + j] = s;
        }
    };
    mm(c.A, c.A, n, &A2);
    mm(A2, c.A, n, &A3);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        Ac[i * n + j] = Ts * (i == j ? 1.0 : 0.0) + Ts * Ts / 2.0 * c.A[i * n + j] +
                        Ts * Ts * Ts / 6.0 * A2[i * n + j] +
                        Ts * Ts * Ts * Ts / 24.0 * A3[i * n + j];
    mm(Ac, c.A, n, &d->A);
    for (int i = 0; i < n; ++i) d->A[i * n + i] += 1.0;
    mm(Ac, c.B, 4, &d->B);
    d->C = c.C;
    mm(Ac, c.f, 1, &d->f);
  }

  // aug_lin_sys.cc:145-177
  void Update(const double* x, const double* u) {
    PlantLin cont, disc;
    sys_.GetLinearizedSystem(x, u, &cont);
    DiscretizeRK4(cont, sampling_time_, &disc);
    const int n = n_states;
    Aorig = disc.A;
    int index_delayed_inputs = 0, index_inputs = 0;
    const int n_nd = n_control_inputs - n_delayed_inputs;
    for (int i = 0; i < n_control_inputs; ++i) {
      const int index = is_reduced ? ctrl_idx_[i] : i;
      if (n_delay_[i] == 0) {
        for (int r = 0; r < n; ++r) Borig[r * n_nd + index_inputs] = disc.B[r * 4 + index];
        index_inputs++;
      } else {
        for (int r = 0; r < n; ++r)
          Adelay[r * n_delayed_inputs + index_delayed_inputs] = disc.B[r * 4 + index];
        index_delayed_inputs++;
      }
    }
    for (int r = 0; r < n_outputs; ++r)
      for (int q = 0; q < n; ++q) C[r * n_obs_states + q] = disc.C[r * n + q];
    f = disc.f;
  }

  // AComposite::MultiplyC, C *= A  (aug_lin_sys.cc:62-86).  c is n_y × n_total row-major.
  void A_MultiplyC(int n_y, std::vector<double>* c) const {
    const int n = n_states, nt = n_total_states;
    std::vector<double> temp(n_y * n), temp2(n_y * n_aug_states);
    for (int r = 0; r < n_y; ++r) {
      for (int q = 0; q < n; ++q) temp[r * n + q] = (*c)[r * nt + q];
      for (int q = 0; q < n_aug_states; ++q) temp2[r * n_aug_states + q] = (*c)[r * nt + n + q];
    }
    for (int r = 0; r < n_y; ++r)
      for (int q = 0; q < n; ++q) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += temp[r * n + k] * Aorig[k * n + q];
        (*c)[r * nt + q] = s;
      }
    for (int i = 0; i < n_aug_states; ++i)
      for (int r = 0; r < n_y; ++r)
        (*c)[r * nt + n + i] = (Aaug[i] >= 0) ? temp2[r * n_aug_states + Aaug[i]] : 0.0;
    for (int r = 0; r < n_y; ++r)
      for (int j = 0; j < n_delayed_inputs; ++j) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += temp[r * n + k] * Adelay[k * n_delayed_inputs + j];
        (*c)[r * nt + n_obs_states + j] += s;
      }
  }

  // BComposite::MultiplyC, out = C·B  (aug_lin_sys.cc:91-113).  out n_y×4 row-major.
  void B_MultiplyC(int n_y, const std::vector<double>& c, std::vector<double>* out) const {
    const int n = n_states, nt = n_total_states;
    const int n_nd = n_control_inputs - n_delayed_inputs;
    int index_inputs = 0;
    for (int i = 0; i < n_control_inputs; ++i) {
      if (n_delay_[i] == 0) {
        for (int r = 0; r < n_y; ++r) {
          double s = 0;
          for (int k = 0; k < n; ++k) s += c[r * nt + k] * Borig[k * n_nd + index_inputs];
          (*out)[r * 4 + i] = s;
        }
        index_inputs++;
      } else {
        for (int r = 0; r < n_y; ++r) (*out)[r * 4 + i] = c[r * nt + Baug[i]];
      }
    }
  }

  // AComposite::TimesAugmentedOnly (aug_lin_sys.cc:118-140); x has n_aug entries.
  void A_TimesAugmentedOnly(const double* x, double* x_out) const {
    for (int i = 0; i < n_total_states; ++i) x_out[i] = 0;
    for (int i = 0; i < n_aug_states; ++i)
      if (Aaug[i] >= 0) x_out[n_states + Aaug[i]] = x[i];
    for (int r = 0; r < n_states; ++r) {
      double s = 0;
      for (int j = 0; j < n_delayed_inputs; ++j)
        s += Adelay[r * n_delayed_inputs + j] * x[n_disturbance_states + j];
      x_out[r] += s;
    }
  }

  // AComposite::operator* (aug_lin_sys.h:235-252)
  void A_Times(const double* x, double* x_out) const {
    A_TimesAugmentedOnly(x + n_states, x_out);
    for (int r = 0; r < n_states; ++r) {
      double s = 0;
      for (int k = 0; k < n_states; ++k) s += Aorig[r * n_states + k] * x[k];
      x_out[r] += s;
    }
  }

  // BComposite::operator* (aug_lin_sys.cc:204-227)
  void B_Times(const double* u, double* x_out) const {
    for (int i = 0; i < n_total_states; ++i) x_out[i] = 0;
    const int n_nd = n_control_inputs - n_delayed_inputs;
    int index_inputs = 0;
    for (int i = 0; i < n_control_inputs; ++i) {
      if (n_delay_[i] == 0) {
        for (int r = 0; r < n_states; ++r) x_out[r] += Borig[r * n_nd + index_inputs] * u[i];
        index_inputs++;
      } else {
        x_out[Baug[i]] = u[i];
      }
    }
  }

  // aug_lin_sys.h:129-138
  void AdjustFirstDelayedStates(double* x, const double* u) const {
    int index_delayed_inputs = n_obs_states;
    for (int i = 0; i < n_control_inputs; ++i)
      if (n_delay_[i] != 0) {
        x[index_delayed_inputs] -= u[i];
        index_delayed_inputs++;
      }
  }
  // aug_lin_sys.h:141-154
  void AdjustAllDelayedStates(double* x, const double* u) const {
    int index_delay_states = n_obs_states + n_delayed_inputs;
    int index_delayed_inputs = n_obs_states;
    for (int i = 0; i < n_control_inputs; ++i)
      if (n_delay_[i] != 0) {
        x[index_delayed_inputs] -= u[i];
        for (int j = 1; j < n_delay_[i]; ++j) x[index_delay_states + j - 1] -= u[i];
        index_delay_states += n_delay_[i] - 1;
        index_delayed_inputs++;
      }
  }
  // aug_lin_sys.h:157-163
  void AdjustAppliedInput(double* du, const double* u) const {
    for (int i = 0; i < n_control_inputs; ++i)
      if (n_delay_[i] != 0) du[i] += u[i];
  }

  // aug_lin_sys.cc:260-334
  void GeneratePrediction(const int* controlled, int n_y, MatX* Su, MatX* Sx, MatX* Sf,
                          MatX* Su_other, int p, int m) const {
    const int n = n_states, nt = n_total_states;
    std::vector<double> c(n_y * nt, 0.0);
    for (int i = 0; i < n_y; ++i)
      for (int q = 0; q < n_obs_states; ++q) c[i * nt + q] = C[controlled[i] * n_obs_states + q];
    Su->Resize(p * n_y, m * n_sub_control_inputs);
    Sx->Resize(p * n_y, n_aug_states);
    Sf->Resize(p * n_y, n);
    if (is_reduced) Su_other->Resize(p * n_y, m * n_other_control_inputs);
    std::vector<double> to_add(n_y * 4);
    for (int r = 0; r < n_y; ++r)
      for (int q = 0; q < n; ++q) (*Sf)(r, q) = c[r * nt + q];
    for (int i = 0; i < p; ++i) {
      if (i > 0)
        for (int r = 0; r < n_y; ++r)
          for (int q = 0; q < n; ++q)
            (*Sf)(i * n_y + r, q) = (*Sf)((i - 1) * n_y + r, q) + c[r * nt + q];
      B_MultiplyC(n_y, c, &to_add);
      for (int j = 0; j < p - i; ++j) {
        const int ind_row = i + j;
        const int ind_col = (j < m) ? j : m - 1;
        for (int r = 0; r < n_y; ++r) {
          for (int q = 0; q < n_sub_control_inputs; ++q)
            (*Su)(ind_row * n_y + r, ind_col * n_sub_control_inputs + q) += to_add[r * 4 + q];
          if (is_reduced)
            for (int q = 0; q < n_other_control_inputs; ++q)
              (*Su_other)(ind_row * n_y + r, ind_col * n_other_control_inputs + q) +=
                  to_add[r * 4 + n_sub_control_inputs + q];
        }
      }
      A_MultiplyC(n_y, &c);
      for (int r = 0; r < n_y; ++r)
        for (int q = 0; q < n_aug_states; ++q) (*Sx)(i * n_y + r, q) = c[r * nt + n + q];
    }
  }
};

// ---------------------------------------------------------------------------
// Observer  (include/observer.h, libs/observer.cc)
// ---------------------------------------------------------------------------
class Observer {
 public:
  std::vector<double> M_;      // n_obs × n_outputs row-major
  std::vector<double> y_old_;  // n_outputs
  std::vector<double> dx_aug_; // n_total
  const AugLinSys* p_ = nullptr;

  // observer.cc:6-19
  void ObserveAPriori(const double* du_in, const double* u_old) {
    const AugLinSys& s = *p_;
    double du[4];
    for (int i = 0; i < 4; ++i) du[i] = du_in[i];
    std::vector<double> dx = dx_aug_;
    for (int i = 0; i < s.n_states; ++i) dx[i] = 0;
    s.AdjustFirstDelayedStates(dx.data(), u_old);
    s.AdjustAppliedInput(du, u_old);
    std::vector<double> bu(s.n_total_states), ax(s.n_total_states);
    s.B_Times(du, bu.data());
    s.A_Times(dx.data(), ax.data());
    for (int i = 0; i < s.n_total_states; ++i) dx_aug_[i] = bu[i] + ax[i];
    for (int i = 0; i < s.n_states; ++i) dx_aug_[i] += s.f[i];
  }

  // observer.cc:24-40; returns dx_aug_.head(n_states)
  void ObserveAPosteriori(const double* y_in, double* dx_head) {
    const AugLinSys& s = *p_;
    const int no = s.n_obs_states;
    double e[4];
    for (int r = 0; r < s.n_outputs; ++r) {
      double cy = 0;
      for (int q = 0; q < no; ++q) cy += s.C[r * no + q] * dx_aug_[q];
      e[r] = y_in[r] - y_old_[r] - cy;
    }
    for (int i = 0; i < no; ++i) {
      double acc = 0;
      for (int r = 0; r < s.n_outputs; ++r) acc += M_[i * s.n_outputs + r] * e[r];
      dx_aug_[i] = dx_aug_[i] + acc;
    }
    for (int r = 0; r < s.n_outputs; ++r) y_old_[r] = y_in[r];
    for (int i = 0; i < s.n_states; ++i) dx_head[i] = dx_aug_[i];
  }
};

// ---------------------------------------------------------------------------
// MpcQpSolver + DistributedSolver  (include/mpc_qp_solver.h, distributed_solver.h)
// ---------------------------------------------------------------------------
struct InputConstraints {
  double lower_bound[4], upper_bound[4], lower_rate_bound[4], upper_rate_bound[4];
};

struct QP {
  int nv = 0;
  std::vector<double> H;  // nv×nv row-major (mpc_qp_solver.h:45-50)
  std::vector<double> f;  // nv
};

class DistributedSolver {
 public:
  int n_outputs, n_control_inputs, p, m, nv;
  std::vector<double> y_ref_;     // p*n_outputs
  std::vector<double> u_weight_;  // nv×nv
  std::vector<double> ywt_;       // n_outputs×n_outputs; y_weight_ = I_p ⊗ ywt (block diag)
  InputConstraints u_constraints_;
  MatX y_pred_weight_;            // Q·Su
  QpWorkspace ws_;                // warm-start state of the exact active-set solver
  int last_status_ = 0;
  unsigned last_active_ = 0;
  double last_objective_ = 0;
  int last_iterations_ = 0;

  DistributedSolver(int n_out, int n_ci, int p_in, int m_in)
      : n_outputs(n_out), n_control_inputs(n_ci), p(p_in), m(m_in), nv(m_in * n_ci) {
    y_ref_.assign(p * n_outputs, 0.0);
    u_weight_.assign(nv * nv, 0.0);
    ywt_.assign(n_outputs * n_outputs, 0.0);
    for (int i = 0; i < n_outputs; ++i) ywt_[i * n_outputs + i] = 1.0;
    for (int i = 0; i < nv; ++i) u_weight_[i * nv + i] = 1.0;
  }

  // mpc_qp_solver.h:62-80.  uwt n_ci×n_ci, ywt n_out×n_out (row-major here).
  void SetWeights(const double* uwt, const double* ywt) {
    for (int i = 0; i < n_outputs * n_outputs; ++i) ywt_[i] = ywt[i];
    std::fill(u_weight_.begin(), u_weight_.end(), 0.0);
    for (int b = 0; b < m; ++b)
      for (int i = 0; i < n_control_inputs; ++i)
        for (int j = 0; j < n_control_inputs; ++j)
          u_weight_[(b * n_control_inputs + i) * nv + b * n_control_inputs + j] =
              uwt[i * n_control_inputs + j];
  }

  // y_weight_ * X for the block-diagonal y_weight_
  void ApplyYWeight(const MatX& X, MatX* out) const {
    out->Resize(X.rows, X.cols);
    for (int c = 0; c < X.cols; ++c)
      for (int i = 0; i < p; ++i)
        for (int r = 0; r < n_outputs; ++r) {
          double s = 0;
          for (int q = 0; q < n_outputs; ++q)
            s += ywt_[r * n_outputs + q] * X(i * n_outputs + q, c);
          (*out)(i * n_outputs + r, c) = s;
        }
  }

  // distributed_solver.h:83-94 -> mpc_qp_solver.cc:19-40
  void GenerateDistributedQP(QP* qp, const MatX& Su, const MatX& Sx, const MatX& Sf,
                             const double* delta_x0, int n_total_states, int n_aug_states,
                             const double* y_prev) {
    ApplyYWeight(Su, &y_pred_weight_);
    const MatX& W = y_pred_weight_;
    const int rows = p * n_outputs;
    qp->nv = nv;
    qp->H.assign(nv * nv, 0.0);
    qp->f.assign(nv, 0.0);
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) {
        double s = 0;
        for (int r = 0; r < rows; ++r) s += Su(r, i) * W(r, j);
        qp->H[i * nv + j] = s + u_weight_[i * nv + j];
      }
    const int n_states = n_total_states - n_aug_states;
    std::vector<double> t_f(rows), t_x(rows), dy_ref(rows);
    for (int r = 0; r < rows; ++r) {
      double s = 0;
      for (int q = 0; q < n_states; ++q) s += delta_x0[q] * Sf(r, q);
      t_f[r] = s;
      double s2 = 0;
      for (int q = 0; q < n_aug_states; ++q) s2 += delta_x0[n_states + q] * Sx(r, q);
      t_x[r] = s2;
      dy_ref[r] = y_ref_[r] - y_prev[r % n_outputs];
    }
    for (int j = 0; j < nv; ++j) {
      double a = 0, b = 0, c = 0;
      for (int r = 0; r < rows; ++r) {
        a += t_f[r] * W(r, j);
        b += dy_ref[r] * W(r, j);
        c += t_x[r] * W(r, j);
      }
      qp->f[j] = a - b + c;
    }
  }

  // distributed_solver.h:109-115:  f += (Su_other du_other)^T (Q Su)
  void ApplyOtherInput(QP* qp, const double* du_other, const MatX& Su_other) const {
    const int rows = Su_other.rows;
    std::vector<double> t(rows);
    for (int r = 0; r < rows; ++r) {
      double s = 0;
      for (int c = 0; c < Su_other.cols; ++c) s += du_other[c] * Su_other(r, c);
      t[r] = s;
    }
    for (int j = 0; j < nv; ++j) {
      double s = 0;
      for (int r = 0; r < rows; ++r) s += t[r] * y_pred_weight_(r, j);
      qp->f[j] += s;
    }
  }

  // mpc_qp_solver.cc:45-75 (constraint rows: mpc_qp_solver.h:108-123).
  // Returns zeros when the solver does not succeed, like the reference.
  void SolveQP(const QP& qp, const double* u_old, double* z_out) {
    double lb[8], ub[8], lbA[8], ubA[8];
    for (int b = 0; b < m; ++b)
      for (int i = 0; i < n_control_inputs; ++i) {
        const int k = b * n_control_inputs + i;
        lb[k] = u_constraints_.lower_bound[i] - u_old[i];
        ub[k] = u_constraints_.upper_bound[i] - u_old[i];
        lbA[k] = u_constraints_.lower_rate_bound[i];
        ubA[k] = u_constraints_.upper_rate_bound[i];
      }
    last_status_ = SolveMpcQp(nv, n_control_inputs, qp.H.data(), qp.f.data(), lb, ub, lbA, ubA,
                              &ws_, z_out, &last_active_, &last_objective_, &last_iterations_);
    if (last_status_ != 0) {
      for (int i = 0; i < nv; ++i) z_out[i] = 0.0;
    }
  }
};

// ---------------------------------------------------------------------------
// DistributedController  (include/distributed_controller.h, libs/distributed_controller.cc)
// ---------------------------------------------------------------------------
class DistributedController {
 public:
  ControllerConfig cfg_;
  AugLinSys auglinsys_;
  Observer observer_;
  DistributedSolver qp_solver_;
  std::vector<double> x_;  // n_states
  double u_old_[4];        // FullControlInput in this controller's own ordering
  MatX su_other_;
  QP qp_;
  Prediction pred;
  int p, m;
  bool is_reduced;
  int n_control_inputs;  // own

  DistributedController(const SystemConfig& sc, int index)
      : cfg_(sc.ctrl[index]),
        auglinsys_(sc, sc.ctrl[index]),
        qp_solver_(sc.ctrl[index].n_controlled_outputs, sc.ctrl[index].n_sub_control_inputs, sc.p,
                   sc.m),
        p(sc.p),
        m(sc.m) {
    is_reduced = auglinsys_.is_reduced;
    n_control_inputs = auglinsys_.n_sub_control_inputs;
    x_.assign(auglinsys_.n_states, 0.0);
    for (double& v : u_old_) v = 0;
    observer_.M_.assign(auglinsys_.n_obs_states * 4, 0.0);
    observer_.y_old_.assign(4, 0.0);
    observer_.dx_aug_.assign(auglinsys_.n_total_states, 0.0);
  }
  DistributedController(const DistributedController&) = delete;

  void BuildDeltaX0(std::vector<double>* delta_x0) const {
    const AugLinSys& s = auglinsys_;
    delta_x0->assign(s.n_total_states, 0.0);
    for (int i = 0; i < s.n_states; ++i) (*delta_x0)[i] = s.f[i];
    for (int i = 0; i < s.n_aug_states; ++i)
      (*delta_x0)[s.n_states + i] = observer_.dx_aug_[s.n_states + i];
    s.AdjustAllDelayedStates(delta_x0->data(), u_old_);
  }

  // distributed_controller.cc:27-67 (the qpOASES cold start it ends with only
  // seeds that library's homotopy and is not needed by an exact solver)
  void Initialize(const double* x_init, const double* u_init, const double* full_u_old,
                  const double* y_init, const double* dx_init) {
    auglinsys_.Update(x_init, full_u_old);
    for (int i = 0; i < 4; ++i) u_old_[i] = u_init[i];
    for (int i = 0; i < auglinsys_.n_states; ++i) x_[i] = x_init[i];
    observer_.p_ = &auglinsys_;
    for (int i = 0; i < 4; ++i) observer_.y_old_[i] = y_init[i];
    for (int i = 0; i < auglinsys_.n_total_states; ++i)
      observer_.dx_aug_[i] = dx_init ? dx_init[i] : 0.0;
    qp_solver_.ws_.Reset();
  }

  // distributed_controller.cc:72-108
  void GenerateInitialQP(const double* y, const double* full_u_old) {
    std::vector<double> dx_head(auglinsys_.n_states);
    observer_.ObserveAPosteriori(y, dx_head.data());
    for (int i = 0; i < auglinsys_.n_states; ++i) x_[i] += dx_head[i];
    auglinsys_.Update(x_.data(), full_u_old);
    std::vector<double> delta_x0;
    BuildDeltaX0(&delta_x0);
    auglinsys_.GeneratePrediction(cfg_.controlled_output_indices, cfg_.n_controlled_outputs,
                                  &pred.Su, &pred.Sx, &pred.Sf, &su_other_, p, m);
    double y_controlled[4];
    for (int i = 0; i < cfg_.n_controlled_outputs; ++i)
      y_controlled[i] = y[cfg_.controlled_output_indices[i]];
    qp_solver_.GenerateDistributedQP(&qp_, pred.Su, pred.Sx, pred.Sf, delta_x0.data(),
                                     auglinsys_.n_total_states, auglinsys_.n_aug_states,
                                     y_controlled);
  }

  // distributed_controller.h:206-226
  void GetInput(double* u_solution, const double* du_last) {
    if (is_reduced) {
      QP qp_new = qp_;
      qp_solver_.ApplyOtherInput(&qp_new, du_last, su_other_);
      qp_solver_.SolveQP(qp_new, u_old_, u_solution);
    } else {
      qp_solver_.SolveQP(qp_, u_old_, u_solution);
    }
  }

  // distributed_controller.h:146-152
  void UpdateU(const double* du) {
    observer_.ObserveAPriori(du, u_old_);
    for (int i = 0; i < 4; ++i) u_old_[i] += du[i];
  }
};

// ---------------------------------------------------------------------------
// NerveCenter  (include/nerve_center.h)
// ---------------------------------------------------------------------------
class NerveCenter {
 public:
  SystemConfig sc_;
  Plant plant_;
  std::vector<DistributedController*> sub_;
  double u_old_[4];
  std::vector<double> du_old_;   // n_prediction_control_inputs
  std::vector<double> u_offset_;
  int n_pred_ = 0;

  explicit NerveCenter(const SystemConfig& sc) : sc_(sc), plant_(sc.plant) {
    for (int c = 0; c < sc.n_controllers; ++c) {
      sub_.push_back(new DistributedController(sc, c));
      n_pred_ += sc.m * sc.ctrl[c].n_sub_control_inputs;
    }
    for (double& v : u_old_) v = 0;
    du_old_.assign(n_pred_, 0.0);
    u_offset_.assign(plant_.n_inputs, 0.0);
  }
  ~NerveCenter() {
    for (auto* c : sub_) delete c;
  }
  NerveCenter(const NerveCenter&) = delete;

  void SetObserverGain(int c, const double* M) {
    auto& o = sub_[c]->observer_;
    for (size_t i = 0; i < o.M_.size(); ++i) o.M_[i] = M[i];
  }
  void SetConstraints(int c, const InputConstraints& ic) { sub_[c]->qp_solver_.u_constraints_ = ic; }

  // nerve_center.h:113-116,225-234: per-controller ywt, uwt sub-matrix of the full 4×4
  void SetWeights(const double* uwt_full, const double* const* ywts) {
    for (size_t c = 0; c < sub_.size(); ++c) {
      const int nu = sub_[c]->n_control_inputs;
      double uwt_sub[16];
      for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nu; ++j)
          uwt_sub[i * nu + j] = uwt_full[sub_[c]->cfg_.control_input_indices[i] * 4 +
                                         sub_[c]->cfg_.control_input_indices[j]];
      sub_[c]->qp_solver_.SetWeights(uwt_sub, ywts[c]);
    }
  }

  // nerve_center.h:119-122,237-249.  y_ref is p × n_outputs(4), row per prediction step.
  void SetOutputReference(const double* y_ref) {
    for (auto* c : sub_) {
      const int ny = c->cfg_.n_controlled_outputs;
      for (int i = 0; i < sc_.p; ++i)
        for (int r = 0; r < ny; ++r)
          c->qp_solver_.y_ref_[i * ny + r] = y_ref[i * 4 + c->cfg_.controlled_output_indices[r]];
    }
  }

  // nerve_center.h:98-104,186-203
  void Initialize(const double* x_init, const double* u_init, const double* u_init_full,
                  const double* y_init) {
    for (auto* c : sub_) {
      double u_init_sub[4];
      for (int i = 0; i < 4; ++i) u_init_sub[i] = u_init[c->cfg_.control_input_indices[i]];
      c->Initialize(x_init, u_init_sub, u_init_full, y_init, nullptr);
    }
    for (int i = 0; i < plant_.n_inputs; ++i) u_offset_[i] = u_init_full[i];
    for (int i = 0; i < 4; ++i) u_old_[i] = 0;  // ctor value (nerve_center.h:93)
    std::fill(du_old_.begin(), du_old_.end(), 0.0);
  }

  // nerve_center.h:134-182 (timing code omitted).  y has 4 entries; returns u_old_ in u.
  void GetNextInput(const double* y, double* u) {
    std::vector<double> u_full_old(plant_.n_inputs);
    plant_.GetPlantInput(u_old_, u_offset_.data(), u_full_old.data());
    for (auto* c : sub_) c->GenerateInitialQP(y, u_full_old.data());

    std::vector<double> du_prev = du_old_, du_new(n_pred_, 0.0), du_other(n_pred_);
    for (int it = 0; it < sc_.n_solver_iterations; ++it) {
      int prediction_index = 0;
      for (auto* c : sub_) {
        // nerve_center.h:275-296
        const int own = sc_.m * c->n_control_inputs;
        int k = 0;
        for (int i = 0; i < prediction_index; ++i) du_other[k++] = du_prev[i];
        for (int i = prediction_index + own; i < n_pred_; ++i) du_other[k++] = du_prev[i];
        c->GetInput(&du_new[prediction_index], du_other.data());
        prediction_index += own;
      }
      du_prev = du_new;
    }
    du_old_ = du_prev;
    double du[4];
    for (int i = 0; i < 4; ++i) du[i] = -u_old_[i];
    int prediction_index = 0, input_index = 0;
    for (auto* c : sub_) {  // nerve_center.h:313-319
      for (int i = 0; i < c->n_control_inputs; ++i)
        u_old_[input_index + i] += du_old_[prediction_index + i];
      prediction_index += sc_.m * c->n_control_inputs;
      input_index += c->n_control_inputs;
    }
    for (int i = 0; i < 4; ++i) du[i] += u_old_[i];
    for (auto* c : sub_) {  // nerve_center.h:322-328
      double du_reordered[4] = {0, 0, 0, 0};
      for (int i = 0; i < c->n_control_inputs; ++i)
        du_reordered[i] = du[c->cfg_.control_input_indices[i]];
      c->UpdateU(du_reordered);
    }
    for (int i = 0; i < 4; ++i) u[i] = u_old_[i];
  }
};

}  // namespace oracle
