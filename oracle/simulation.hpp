// ORACLE — TEST INFRASTRUCTURE ONLY (see plant.hpp).
//
// Closed-loop harness around the control step:
//   include/simulation_system.h:69-116   SetInput / operator() / Integrate
//       (boost::odeint runge_kutta_dopri5 inside controlled_runge_kutta,
//        rel/abs 1e-6; the inf-norm is overridden by a 2-norm, :119-133)
//   include/time_delay.h:41-58           TimeDelay::GetDelayedInput
// and the reconstructed driver loop of the missing tests/common-simulation.inc
// (SURVEY.md §3.1): per sample  y = GetOutput(x); u = GetNextInput(y); record;
// SetInput(u); integrate [t, t+Ts]; t += Ts.
//
// boost::odeint is an un-vendored dependency; its controlled Dormand-Prince
// stepper is restated from the published algorithm (default_error_checker /
// default_step_adjuster of odeint v2): err_i = |xerr_i| / (eps_abs + eps_rel *
// (|x_i| + dt |dxdt_i|)); reject and shrink by max(0.9 err^(-1/3), 0.2) when
// err > 1; after an accepted step grow by 0.9 err^(-1/5) (err floored at 5^-5)
// when err < 0.5.  Every 50 ms interval starts a fresh stepper with dt = Ts
// (simulation_system.h:112-115 builds the stepper inside Integrate).
#pragma once
#include <cmath>
#include <vector>

#include "controller.hpp"

namespace oracle {

// include/time_delay.h:13-76
class TimeDelay {
 public:
  int n_delay_[4];
  std::vector<double> u_delay_;
  int current_input_[4];
  explicit TimeDelay(const int delays[4]) {
    int sum = 0;
    for (int i = 0; i < 4; ++i) {
      n_delay_[i] = delays[i];
      current_input_[i] = sum;
      sum += delays[i];
    }
    u_delay_.assign(sum, 0.0);
  }
  void GetDelayedInput(const double u_next[4], double u_out[4]) {
    int index_delay_states = 0;
    for (int i = 0; i < 4; ++i) {
      if (n_delay_[i] == 0) {
        u_out[i] = u_next[i];
      } else {
        index_delay_states += n_delay_[i];
        u_out[i] = u_delay_[current_input_[i]];
        u_delay_[current_input_[i]] = u_next[i];
        current_input_[i]++;
        if (current_input_[i] == index_delay_states) current_input_[i] -= n_delay_[i];
      }
    }
  }
};

// One try_step of controlled_runge_kutta<runge_kutta_dopri5> (FSAL).
// Returns true on success (x, dxdt, t advanced; dt = suggestion for the next step),
// false on rejection (dt reduced, nothing else changed).
inline bool Dopri5TryStep(const Plant& plant, const double* u, int n, double* x, double* dxdt,
                          double* t, double* dt, double eps_abs, double eps_rel,
                          int* n_rhs_evals) {
  constexpr double a2 = 1.0 / 5, a3 = 3.0 / 10, a4 = 4.0 / 5, a5 = 8.0 / 9;
  (void)a2; (void)a3; (void)a4; (void)a5;  // autonomous system: stage times unused
  constexpr double b21 = 1.0 / 5;
  constexpr double b31 = 3.0 / 40, b32 = 9.0 / 40;
  constexpr double b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9;
  constexpr double b51 = 19372.0 / 6561, b52 = -25360.0 / 2187, b53 = 64448.0 / 6561,
                   b54 = -212.0 / 729;
  constexpr double b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247,
                   b64 = 49.0 / 176, b65 = -5103.0 / 18656;
  constexpr double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784,
                   c6 = 11.0 / 84;
  constexpr double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                   dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  const double h = *dt;
  double k2[16], k3[16], k4[16], k5[16], k6[16], k7[16], xt[16], xn[16];
  const double* k1 = dxdt;
  for (int i = 0; i < n; ++i) xt[i] = x[i] + h * b21 * k1[i];
  plant.GetDerivative(xt, u, k2);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + h * (b31 * k1[i] + b32 * k2[i]);
  plant.GetDerivative(xt, u, k3);
  for (int i = 0; i < n; ++i) xt[i] = x[i] + h * (b41 * k1[i] + b42 * k2[i] + b43 * k3[i]);
  plant.GetDerivative(xt, u, k4);
  for (int i = 0; i < n; ++i)
    xt[i] = x[i] + h * (b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i]);
  plant.GetDerivative(xt, u, k5);
  for (int i = 0; i < n; ++i)
    xt[i] = x[i] + h * (b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i]);
  plant.GetDerivative(xt, u, k6);
  for (int i = 0; i < n; ++i)
    xn[i] = x[i] + h * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
  plant.GetDerivative(xn, u, k7);
  *n_rhs_evals += 6;
  // error estimate and its weighted 2-norm (simulation_system.h:119-133)
  double sumsq = 0;
  for (int i = 0; i < n; ++i) {
    const double xerr =
        h * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
    const double e =
        std::fabs(xerr) / (eps_abs + eps_rel * (std::fabs(x[i]) + std::fabs(h) * std::fabs(k1[i])));
    sumsq += e * e;
  }
  double err = std::sqrt(sumsq);
  if (err > 1.0) {
    *dt = h * std::fmax(0.9 * std::pow(err, -1.0 / 3.0), 0.2);
    return false;
  }
  *t += h;
  for (int i = 0; i < n; ++i) {
    x[i] = xn[i];
    dxdt[i] = k7[i];
  }
  if (err < 0.5) {
    err = std::fmax(std::pow(5.0, -5.0), err);
    *dt = h * 9.0 / 10.0 * std::pow(err, -1.0 / 5.0);
  }
  return true;
}

// odeint's integrate_adaptive puts no bound on the number of accepted steps: a plant state that runs
// away (unphysical inputs) makes it take ever smaller steps for ever.  The GPU path must not do
// that, so both sides stop an interval after this many accepted steps (the nominal count is 1-4)
// and leave the state where it is; the records of such a scenario are no longer meaningful.
constexpr int kMaxStepsPerInterval = 4000;

// integrate_adaptive over one sampling interval [t0, t0+Ts] starting with dt = Ts.
inline int IntegrateInterval(const Plant& plant, const double* u, double* x, double t0, double Ts,
                             int* n_rhs_evals) {
  const int n = plant.n_states;
  double dxdt[16];
  plant.GetDerivative(x, u, dxdt);
  *n_rhs_evals += 1;
  double t = t0, dt = Ts;
  const double t_end = t0 + Ts;
  int steps = 0, fails = 0;
  const double eps = std::numeric_limits<double>::epsilon();
  // less_with_sign(t, t_end, dt): (t_end - t) > eps
  while (t_end - t > eps && steps < kMaxStepsPerInterval) {
    if ((t + dt) - t_end > eps) dt = t_end - t;
    while (!Dopri5TryStep(plant, u, n, x, dxdt, &t, &dt, 1e-6, 1e-6, n_rhs_evals)) {
      if (++fails > 500) return -1;
    }
    fails = 0;
    ++steps;
  }
  return steps;   // kMaxStepsPerInterval: the interval was cut short (runaway plant state)
}

// SimulationSystem (simulation_system.h:17-117) reduced to what the driver uses.
class SimulationSystem {
 public:
  const Plant* p_sys_;
  std::vector<double> x_, u_offset_, u_;
  TimeDelay delayed_inputs_;
  int rhs_evals_ = 0;
  SimulationSystem(const Plant* sys, const std::vector<double>& u_offset,
                   const std::vector<double>& x_in, const int delays[4])
      : p_sys_(sys), x_(x_in), u_offset_(u_offset), u_(u_offset), delayed_inputs_(delays) {}
  void SetOffset(const double* u_in) {
    for (int i = 0; i < p_sys_->n_inputs; ++i) u_offset_[i] = u_in[i];
  }
  // simulation_system.h:69-71,80-86
  void SetInput(const double u[4]) {
    double ud[4];
    delayed_inputs_.GetDelayedInput(u, ud);
    for (int i = 0; i < p_sys_->n_inputs; ++i) u_[i] = u_offset_[i];
    for (int i = 0; i < 4; ++i) u_[Plant::ControlInputIndex(i)] += ud[i];
  }
  int IntegrateOneSample(double t0, double Ts) {
    return IntegrateInterval(*p_sys_, u_.data(), x_.data(), t0, Ts, &rhs_evals_);
  }
};

}  // namespace oracle
