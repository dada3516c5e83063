// ORACLE — TEST INFRASTRUCTURE ONLY (see plant.hpp).  extern "C" surface so
// tests/ and bench.py's cpu_baseline leg can drive the CPU restatement through
// ctypes.  Built by oracle/Makefile into oracle/libcmpc_oracle.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "controller.hpp"
#include "simulation.hpp"

using namespace oracle;

namespace {

struct Setup {
  SystemConfig sc;
  double uwt[16];
  double ywt[kMaxControllers][16];
  InputConstraints ic[kMaxControllers];
  std::vector<double> M[kMaxControllers];
  std::vector<double> yref;  // p × 4
};

struct Handle {
  Setup s;
  NerveCenter* nc = nullptr;
  ~Handle() { delete nc; }
};

void DefaultConfig(int plant, int mode, SystemConfig* sc) {
  sc->plant = plant == 0 ? kParallel : kSerial;
  const bool par = plant == 0;
  if (mode == 0) {  // centralised
    sc->n_controllers = 1;
    sc->n_solver_iterations = 1;
    sc->ctrl[0].n_sub_control_inputs = 4;
    const int idx[4] = {0, 1, 2, 3};
    std::memcpy(sc->ctrl[0].control_input_indices, idx, sizeof idx);
    if (par) {
      sc->ctrl[0].n_controlled_outputs = 3;
      const int o[4] = {0, 1, 3, 0};
      std::memcpy(sc->ctrl[0].controlled_output_indices, o, sizeof o);
    } else {
      sc->ctrl[0].n_controlled_outputs = 4;
      const int o[4] = {0, 1, 2, 3};
      std::memcpy(sc->ctrl[0].controlled_output_indices, o, sizeof o);
    }
    return;
  }
  sc->n_controllers = 2;
  sc->n_solver_iterations = 9;
  const int idx1[4] = {0, 1, 2, 3}, idx2[4] = {2, 3, 0, 1};
  std::memcpy(sc->ctrl[0].control_input_indices, idx1, sizeof idx1);
  std::memcpy(sc->ctrl[1].control_input_indices, idx2, sizeof idx2);
  for (int c = 0; c < 2; ++c) sc->ctrl[c].n_sub_control_inputs = 2;
  int o[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int ny = 0;
  if (par && mode == 1) { ny = 3; int t[4] = {0, 1, 3, 0}; std::memcpy(o[0], t, sizeof t); std::memcpy(o[1], t, sizeof t); }
  if (par && mode == 2) { ny = 2; int t0[4] = {0, 3, 0, 0}, t1[4] = {1, 3, 0, 0}; std::memcpy(o[0], t0, sizeof t0); std::memcpy(o[1], t1, sizeof t1); }
  if (!par && mode == 1) { ny = 4; int t[4] = {0, 1, 2, 3}; std::memcpy(o[0], t, sizeof t); std::memcpy(o[1], t, sizeof t); }
  if (!par && mode == 2) { ny = 2; int t0[4] = {0, 1, 0, 0}, t1[4] = {2, 3, 0, 0}; std::memcpy(o[0], t0, sizeof t0); std::memcpy(o[1], t1, sizeof t1); }
  // SERIAL_CTRL_NONCOOP_OLD{1,2} (serial_compressors_constants.h:47-59,103-104)
  if (!par && mode == 3) { ny = 3; int t0[4] = {0, 1, 2, 0}, t1[4] = {2, 3, 1, 0}; std::memcpy(o[0], t0, sizeof t0); std::memcpy(o[1], t1, sizeof t1); }
  for (int c = 0; c < 2; ++c) {
    sc->ctrl[c].n_controlled_outputs = ny;
    std::memcpy(sc->ctrl[c].controlled_output_indices, o[c], sizeof o[c]);
  }
}

NerveCenter* MakeNerveCenter(const Setup& s) {
  NerveCenter* nc = new NerveCenter(s.sc);
  const double* ywts[kMaxControllers] = {s.ywt[0], s.ywt[1], s.ywt[2], s.ywt[3]};
  nc->SetWeights(s.uwt, ywts);
  nc->SetOutputReference(s.yref.data());
  for (int c = 0; c < s.sc.n_controllers; ++c) {
    nc->SetConstraints(c, s.ic[c]);
    nc->SetObserverGain(c, s.M[c].data());
  }
  return nc;
}

}  // namespace

extern "C" {

static void* FinishCreate(Handle* h);

void* orc_create(int plant, int mode, int p, int n_iter) {
  Handle* h = new Handle;
  DefaultConfig(plant, mode, &h->s.sc);
  if (p > 0) h->s.sc.p = p;
  if (n_iter > 0) h->s.sc.n_solver_iterations = n_iter;
  return FinishCreate(h);
}

// Runtime form of the reference's template configuration (constexpr_array.h,
// {parallel,serial}_compressors_constants.h, nerve_center.h:19-38): any prediction / move
// horizon, per-input delays (system order; each sub-controller's AugmentedLinearizedSystem gets
// them in its own input order), up to four sub-controllers with their own input counts, output
// partitions and input permutations.
void* orc_create_config(int plant, int p, int m, int n_iter, const int* delays, int n_controllers,
                        const int* n_sub_control_inputs, const int* n_controlled_outputs,
                        const int* controlled_output_indices /* n_ctrl x 4 */,
                        const int* control_input_indices /* n_ctrl x 4 */) {
  Handle* h = new Handle;
  SystemConfig& sc = h->s.sc;
  sc.plant = plant == 0 ? kParallel : kSerial;
  sc.p = p;
  sc.m = m;
  sc.n_solver_iterations = n_iter;
  sc.n_controllers = n_controllers;
  for (int i = 0; i < 4; ++i) sc.delays[i] = delays[i];
  for (int c = 0; c < n_controllers; ++c) {
    ControllerConfig& cc = sc.ctrl[c];
    cc.n_sub_control_inputs = n_sub_control_inputs[c];
    cc.n_controlled_outputs = n_controlled_outputs[c];
    for (int i = 0; i < 4; ++i) {
      cc.controlled_output_indices[i] = controlled_output_indices[c * 4 + i];
      cc.control_input_indices[i] = control_input_indices[c * 4 + i];
      // a controller that is not "reduced" keeps the system order (aug_lin_sys.cc:158)
      cc.delays[i] = cc.n_sub_control_inputs == 4 ? delays[i] : delays[cc.control_input_indices[i]];
    }
  }
  return FinishCreate(h);
}

static void* FinishCreate(Handle* h) {
  Plant pl(h->s.sc.plant);
  const int n_obs = pl.n_states + h->s.sc.n_disturbance_states;
  std::memset(h->s.uwt, 0, sizeof h->s.uwt);
  std::memset(h->s.ywt, 0, sizeof h->s.ywt);
  for (int i = 0; i < 4; ++i) h->s.uwt[i * 4 + i] = 1;
  for (int c = 0; c < kMaxControllers; ++c) {
    const int ny = h->s.sc.ctrl[c].n_controlled_outputs;
    for (int i = 0; i < ny; ++i) h->s.ywt[c][i * ny + i] = 1;
    for (int i = 0; i < 4; ++i) {
      h->s.ic[c].lower_bound[i] = -1e30;
      h->s.ic[c].upper_bound[i] = 1e30;
      h->s.ic[c].lower_rate_bound[i] = -1e30;
      h->s.ic[c].upper_rate_bound[i] = 1e30;
    }
    // default observer gain [0; I]
    h->s.M[c].assign(n_obs * 4, 0.0);
    for (int i = 0; i < 4; ++i) h->s.M[c][(pl.n_states + i) * 4 + i] = 1;
  }
  h->s.yref.assign(h->s.sc.p * 4, 0.0);
  return h;
}
void orc_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int orc_n_states(void* hv) { return Plant(static_cast<Handle*>(hv)->s.sc.plant).n_states; }
int orc_n_inputs(void* hv) { return Plant(static_cast<Handle*>(hv)->s.sc.plant).n_inputs; }
int orc_n_controllers(void* hv) { return static_cast<Handle*>(hv)->s.sc.n_controllers; }

void orc_set_weights(void* hv, int ctrl, const double* uwt_full, const double* ywt) {
  Handle* h = static_cast<Handle*>(hv);
  if (uwt_full) std::memcpy(h->s.uwt, uwt_full, sizeof h->s.uwt);
  const int ny = h->s.sc.ctrl[ctrl].n_controlled_outputs;
  if (ywt) std::memcpy(h->s.ywt[ctrl], ywt, sizeof(double) * ny * ny);
}
void orc_set_constraints(void* hv, int ctrl, const double* lo, const double* up, const double* rlo,
                         const double* rup) {
  Handle* h = static_cast<Handle*>(hv);
  const int nu = h->s.sc.ctrl[ctrl].n_sub_control_inputs;
  for (int i = 0; i < nu; ++i) {
    h->s.ic[ctrl].lower_bound[i] = lo[i];
    h->s.ic[ctrl].upper_bound[i] = up[i];
    h->s.ic[ctrl].lower_rate_bound[i] = rlo[i];
    h->s.ic[ctrl].upper_rate_bound[i] = rup[i];
  }
}
void orc_set_observer_gain(void* hv, int ctrl, const double* M) {
  Handle* h = static_cast<Handle*>(hv);
  std::memcpy(h->s.M[ctrl].data(), M, sizeof(double) * h->s.M[ctrl].size());
}
void orc_set_output_reference(void* hv, const double* yref) {
  Handle* h = static_cast<Handle*>(hv);
  std::memcpy(h->s.yref.data(), yref, sizeof(double) * h->s.yref.size());
}

void orc_initialize(void* hv, const double* x_init, const double* u_init, const double* u_init_full,
                    const double* y_init) {
  Handle* h = static_cast<Handle*>(hv);
  delete h->nc;
  h->nc = MakeNerveCenter(h->s);
  h->nc->Initialize(x_init, u_init, u_init_full, y_init);
}
void orc_get_next_input(void* hv, const double* y, double* u) {
  static_cast<Handle*>(hv)->nc->GetNextInput(y, u);
}

// ---- fine-grained parity hooks ------------------------------------------------
void orc_plant_defaults(int plant, double* x0, double* u0) {
  Plant pl(plant == 0 ? kParallel : kSerial);
  auto x = pl.GetDefaultState();
  auto u = pl.GetDefaultInput();
  std::memcpy(x0, x.data(), sizeof(double) * x.size());
  std::memcpy(u0, u.data(), sizeof(double) * u.size());
}
void orc_plant_derivative(int plant, const double* x, const double* u, double* dxdt) {
  Plant(plant == 0 ? kParallel : kSerial).GetDerivative(x, u, dxdt);
}
void orc_plant_output(int plant, const double* x, double* y) {
  Plant(plant == 0 ? kParallel : kSerial).GetOutput(x, y);
}
void orc_plant_linearize(int plant, const double* x, const double* u, double* A, double* B,
                         double* C, double* f) {
  Plant pl(plant == 0 ? kParallel : kSerial);
  PlantLin lin;
  pl.GetLinearizedSystem(x, u, &lin);
  std::memcpy(A, lin.A.data(), sizeof(double) * lin.A.size());
  std::memcpy(B, lin.B.data(), sizeof(double) * lin.B.size());
  std::memcpy(C, lin.C.data(), sizeof(double) * lin.C.size());
  std::memcpy(f, lin.f.data(), sizeof(double) * lin.f.size());
}
void orc_plant_discretize(int plant, const double* x, const double* u, double Ts, double* Ad,
                          double* Bd, double* Cd, double* fd) {
  Plant pl(plant == 0 ? kParallel : kSerial);
  PlantLin lin, d;
  pl.GetLinearizedSystem(x, u, &lin);
  AugLinSys::DiscretizeRK4(lin, Ts, &d);
  std::memcpy(Ad, d.A.data(), sizeof(double) * d.A.size());
  std::memcpy(Bd, d.B.data(), sizeof(double) * d.B.size());
  std::memcpy(Cd, d.C.data(), sizeof(double) * d.C.size());
  std::memcpy(fd, d.f.data(), sizeof(double) * d.f.size());
}
// integrate one sampling interval with plant input u (already offset+delayed); returns #accepted steps
int orc_plant_integrate(int plant, double* x, const double* u, double t0, double Ts) {
  Plant pl(plant == 0 ? kParallel : kSerial);
  int evals = 0;
  return IntegrateInterval(pl, u, x, t0, Ts, &evals);
}

// current linearisation of controller `ctrl` (row-major)
void orc_get_linearization(void* hv, int ctrl, double* Aorig, double* Borig, double* Adelay,
                           double* C, double* f) {
  const AugLinSys& s = static_cast<Handle*>(hv)->nc->sub_[ctrl]->auglinsys_;
  if (Aorig) std::memcpy(Aorig, s.Aorig.data(), sizeof(double) * s.Aorig.size());
  if (Borig) std::memcpy(Borig, s.Borig.data(), sizeof(double) * s.Borig.size());
  if (Adelay) std::memcpy(Adelay, s.Adelay.data(), sizeof(double) * s.Adelay.size());
  if (C) std::memcpy(C, s.C.data(), sizeof(double) * s.C.size());
  if (f) std::memcpy(f, s.f.data(), sizeof(double) * s.f.size());
}
// prediction matrices of the last GenerateInitialQP, column-major like Eigen::MatrixXd
void orc_get_prediction(void* hv, int ctrl, double* Su, double* Sx, double* Sf, double* Su_other) {
  DistributedController* c = static_cast<Handle*>(hv)->nc->sub_[ctrl];
  if (Su) std::memcpy(Su, c->pred.Su.a.data(), sizeof(double) * c->pred.Su.a.size());
  if (Sx) std::memcpy(Sx, c->pred.Sx.a.data(), sizeof(double) * c->pred.Sx.a.size());
  if (Sf) std::memcpy(Sf, c->pred.Sf.a.data(), sizeof(double) * c->pred.Sf.a.size());
  if (Su_other && c->is_reduced)
    std::memcpy(Su_other, c->su_other_.a.data(), sizeof(double) * c->su_other_.a.size());
}
void orc_get_qp(void* hv, int ctrl, double* H, double* f) {
  DistributedController* c = static_cast<Handle*>(hv)->nc->sub_[ctrl];
  std::memcpy(H, c->qp_.H.data(), sizeof(double) * c->qp_.H.size());
  std::memcpy(f, c->qp_.f.data(), sizeof(double) * c->qp_.f.size());
}
void orc_get_ctrl_state(void* hv, int ctrl, double* x_hat, double* dx_aug, double* y_old,
                        double* u_old) {
  DistributedController* c = static_cast<Handle*>(hv)->nc->sub_[ctrl];
  if (x_hat) std::memcpy(x_hat, c->x_.data(), sizeof(double) * c->x_.size());
  if (dx_aug)
    std::memcpy(dx_aug, c->observer_.dx_aug_.data(), sizeof(double) * c->observer_.dx_aug_.size());
  if (y_old) std::memcpy(y_old, c->observer_.y_old_.data(), sizeof(double) * 4);
  if (u_old) std::memcpy(u_old, c->u_old_, sizeof(double) * 4);
}
void orc_get_plan(void* hv, double* du_old) {
  NerveCenter* nc = static_cast<Handle*>(hv)->nc;
  std::memcpy(du_old, nc->du_old_.data(), sizeof(double) * nc->du_old_.size());
}
void orc_last_qp_info(void* hv, int ctrl, int* status, unsigned* active, double* objective) {
  const DistributedSolver& q = static_cast<Handle*>(hv)->nc->sub_[ctrl]->qp_solver_;
  *status = q.last_status_;
  *active = q.last_active_;
  *objective = q.last_objective_;
}

// standalone QP solve; guess_io: in = working-set guess (or 0xFFFFFFFF for none), out = final set
int orc_solve_qp(int nv, int nu, const double* H, const double* f, const double* lb,
                 const double* ub, const double* lbA, const double* ubA, unsigned* guess_io,
                 double* z, unsigned* active, double* objective, int* iterations) {
  QpWorkspace ws;
  if (guess_io && *guess_io != 0xFFFFFFFFu) {
    ws.has_guess = true;
    ws.guess = *guess_io;
  }
  const int st = SolveMpcQp(nv, nu, H, f, lb, ub, lbA, ubA, &ws, z, active, objective, iterations);
  if (guess_io) *guess_io = ws.guess;
  if (st != 0)
    for (int i = 0; i < nv; ++i) z[i] = 0;
  return st;
}

// ---- closed loop ----------------------------------------------------------------
// One scenario: blocks of (offset vector added to u_def, last record index exclusive).
// traj: n_steps × (1 + n + 4 + 4) = [t, x, u, y]; qp_active/qp_obj/qp_status: n_steps × n_ctrl
// (last solver iteration of each step); step_ns: per-step wall time of GetNextInput.
static void RunScenario(const Setup& s, const double* x0, int n_blocks, const int* block_end,
                        const double* block_off, int n_steps, double* traj, uint32_t* qp_active,
                        double* qp_obj, int32_t* qp_status, double* step_ns) {
  Plant plant(s.sc.plant);
  const int n = plant.n_states, ni = plant.n_inputs, rec = 1 + n + 8;
  std::vector<double> x(x0, x0 + n), u_def = plant.GetDefaultInput();
  SimulationSystem sim(&plant, u_def, x, s.sc.delays);
  NerveCenter* nc = MakeNerveCenter(s);
  double y0[4], u_init[4] = {0, 0, 0, 0};
  plant.GetOutput(x.data(), y0);
  nc->Initialize(x.data(), u_init, u_def.data(), y0);
  double t = 0;
  int blk = 0;
  std::vector<double> off(ni);
  for (int k = 0; k < n_steps; ++k) {
    while (blk + 1 < n_blocks && k >= block_end[blk]) ++blk;
    for (int i = 0; i < ni; ++i) off[i] = u_def[i] + block_off[blk * ni + i];
    sim.SetOffset(off.data());
    double y[4], u[4];
    plant.GetOutput(sim.x_.data(), y);
    auto t0 = std::chrono::steady_clock::now();
    nc->GetNextInput(y, u);
    auto t1 = std::chrono::steady_clock::now();
    if (step_ns) step_ns[k] = std::chrono::duration<double, std::nano>(t1 - t0).count();
    if (traj) {
      double* r = traj + static_cast<size_t>(k) * rec;
      r[0] = t;
      for (int i = 0; i < n; ++i) r[1 + i] = sim.x_[i];
      for (int i = 0; i < 4; ++i) r[1 + n + i] = u[i];
      for (int i = 0; i < 4; ++i) r[1 + n + 4 + i] = y[i];
    }
    for (int c = 0; c < s.sc.n_controllers; ++c) {
      const DistributedSolver& q = nc->sub_[c]->qp_solver_;
      if (qp_active) qp_active[k * s.sc.n_controllers + c] = q.last_active_;
      if (qp_obj) qp_obj[k * s.sc.n_controllers + c] = q.last_objective_;
      if (qp_status) qp_status[k * s.sc.n_controllers + c] = q.last_status_;
    }
    sim.SetInput(u);
    sim.IntegrateOneSample(t, s.sc.Ts);
    t += s.sc.Ts;
  }
  delete nc;
}

// Batch of B independent scenarios over n_threads host threads (scenario-major arrays).
void orc_run_closed_loop(void* hv, int B, int n_steps, const double* x0, int n_blocks,
                         const int* block_end, const double* block_off, double* traj,
                         uint32_t* qp_active, double* qp_obj, int32_t* qp_status, double* step_ns,
                         int n_threads) {
  Handle* h = static_cast<Handle*>(hv);
  Plant plant(h->s.sc.plant);
  const int n = plant.n_states, ni = plant.n_inputs, rec = 1 + n + 8, ncz = h->s.sc.n_controllers;
  auto work = [&](int tid) {
    for (int b = tid; b < B; b += n_threads) {
      RunScenario(h->s, x0 + static_cast<size_t>(b) * n, n_blocks, block_end + b * n_blocks,
                  block_off + static_cast<size_t>(b) * n_blocks * ni, n_steps,
                  traj ? traj + static_cast<size_t>(b) * n_steps * rec : nullptr,
                  qp_active ? qp_active + static_cast<size_t>(b) * n_steps * ncz : nullptr,
                  qp_obj ? qp_obj + static_cast<size_t>(b) * n_steps * ncz : nullptr,
                  qp_status ? qp_status + static_cast<size_t>(b) * n_steps * ncz : nullptr,
                  step_ns ? step_ns + static_cast<size_t>(b) * n_steps : nullptr);
    }
  };
  if (n_threads <= 1) {
    n_threads = 1;
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
  for (auto& t : th) t.join();
}

}  // extern "C"
