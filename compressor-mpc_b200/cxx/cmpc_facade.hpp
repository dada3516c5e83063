// C++ host-side mirror of the reference's controller interface on top of the C ABI (cmpc.h).
// Same class and method names, same call order (ctor -> SetWeights -> SetOutputReference ->
// Initialize -> GetNextInput*) and the same error behaviour as the reference:
//   ControllerInterface<System>::GetNextInput        include/controller_interface.h:46
//   NerveCenter::{SetWeights,SetOutputReference,Initialize,GetNextInput}   include/nerve_center.h:89-182
//   InputConstraints                                  include/input_constraints.h:11-27
//   ReadString / ReadNumbers                          include/read_files.h:13-81
// Eigen is not required; when <Eigen/Eigen> is available the fixed-size overloads below accept the
// reference's own vector types unchanged.  Plain arrays are row-major doubles.
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cmpc.h"

#if defined(__has_include)
#if __has_include(<Eigen/Eigen>)
#include <Eigen/Eigen>
#define CMPC_HAVE_EIGEN 1
#endif
#endif

namespace cmpc_host {

inline void Check(int rc) {
  if (rc != CMPC_OK) throw std::runtime_error(std::string("cmpc: ") + cmpc_last_error());
}

template <int n_control_inputs>
struct InputConstraints {
  std::array<double, n_control_inputs> lower_bound, upper_bound, lower_rate_bound, upper_rate_bound;
  bool use_rate_constraints = false;  // never read by the reference either (mpc_qp_solver.cc:57-64)
};

/// Pure-virtual controller interface, batched: y holds batch x 4 outputs, u batch x 4 inputs.
class ControllerInterface {
 public:
  virtual ~ControllerInterface() {}
  virtual void GetNextInput(const double* y, double* u) = 0;
};

/// NerveCenter for `batch` independent plant scenarios.
class NerveCenter : public ControllerInterface {
 public:
  NerveCenter(int plant, int mode, int n_solver_iterations, int batch = 1, int p = 100, int device = 0) {
    Check(cmpc_default_config(plant, mode, batch, &cfg_));
    cfg_.n_iterations = n_solver_iterations;
    cfg_.p = p;
    Check(cmpc_plant_dims(plant, &n_states_, &n_inputs_));
    Check(cmpc_create(&cfg_, device, &h_));
  }
  /// A configuration outside the reference's own instantiations (cmpc_config filled by hand: other
  /// delays, move horizon, output partitions, up to four sub-controllers; see include/cmpc.h).
  explicit NerveCenter(const cmpc_config& cfg, int device = 0) : cfg_(cfg) {
    Check(cmpc_plant_dims(cfg.plant, &n_states_, &n_inputs_));
    Check(cmpc_create(&cfg_, device, &h_));
  }
  ~NerveCenter() override { cmpc_destroy(h_); }
  NerveCenter(const NerveCenter&) = delete;
  NerveCenter& operator=(const NerveCenter&) = delete;

  int n_states() const { return n_states_; }
  int n_inputs() const { return n_inputs_; }
  int n_controllers() const { return cfg_.n_controllers; }
  int n_sub_control_inputs() const { return cfg_.n_sub_control_inputs; }
  int n_controlled_outputs(int c) const { return cfg_.n_controlled_outputs[c]; }
  int batch() const { return cfg_.batch; }
  cmpc_handle* handle() { return h_; }

  /// uwt: full 4 x 4 input weight; ywts[c]: n_y x n_y output weight of sub-controller c
  /// (tuple overload, nerve_center.h:113-116; the sub-matrix of uwt is taken as in :225-234).
  void SetWeights(const double* uwt, const std::vector<std::vector<double>>& ywts) {
    for (int c = 0; c < cfg_.n_controllers; ++c) {
      const int nu = cfg_.n_sub_control_inputs_per[c] ? cfg_.n_sub_control_inputs_per[c] : cfg_.n_sub_control_inputs;
      std::vector<double> sub(nu * nu);
      for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nu; ++j)
          sub[i * nu + j] = uwt[cfg_.control_input_indices[c][i] * 4 + cfg_.control_input_indices[c][j]];
      Check(cmpc_set_weights(h_, c, sub.data(), ywts[c].data()));
    }
  }
  /// y_ref: p x 4 (nerve_center.h:119-122)
  void SetOutputReference(const double* y_ref) { Check(cmpc_set_output_reference(h_, y_ref)); }
  /// one reference value per output, replicated over the horizon like the reference driver does
  void SetOutputReferenceConstant(const double yref4[4]) {
    std::vector<double> r(static_cast<size_t>(cfg_.p) * 4);
    for (int i = 0; i < cfg_.p; ++i) std::copy(yref4, yref4 + 4, r.begin() + 4 * i);
    SetOutputReference(r.data());
  }
  template <int NU>
  void SetConstraints(int c, const InputConstraints<NU>& k) {
    Check(cmpc_set_constraints(h_, c, k.lower_bound.data(), k.upper_bound.data(), k.lower_rate_bound.data(),
                               k.upper_rate_bound.data()));
  }
  void SetConstraints(int c, const double* lo, const double* up, const double* rlo, const double* rup) {
    Check(cmpc_set_constraints(h_, c, lo, up, rlo, rup));
  }
  void SetObserverGain(int c, const double* M) { Check(cmpc_set_observer_gain(h_, c, M)); }
  /// nerve_center.h:98-104; arrays are batch-major
  void Initialize(const double* x_init, const double* u_init, const double* u_init_full, const double* y_init) {
    Check(cmpc_initialize(h_, x_init, u_init, u_init_full, y_init));
  }
  void GetNextInput(const double* y, double* u) override { Check(cmpc_get_next_input(h_, y, u)); }
  /// nerve_center.h:134-182: the time covers QP generation, the first n_timing_iterations sweeps
  /// and what follows the sweeps (boost::timer::nanosecond_type is a 64-bit integer)
  void GetNextInputWithTiming(const double* y, double* u, int n_timing_iterations = -1, int64_t* time_out = nullptr) {
    int64_t ns = 0;
    Check(cmpc_get_next_input_timed(h_, y, u, n_timing_iterations, &ns));
    if (time_out) *time_out = ns;
  }

#ifdef CMPC_HAVE_EIGEN
  template <int N>
  Eigen::Matrix<double, 4, 1> GetNextInput(const Eigen::Matrix<double, N, 1>& y) {
    static_assert(N == 4, "the compressor plants have four outputs");
    Eigen::Matrix<double, 4, 1> u;
    GetNextInput(y.data(), u.data());
    return u;
  }
#endif

 private:
  cmpc_config cfg_;
  cmpc_handle* h_ = nullptr;
  int n_states_ = 0, n_inputs_ = 0;
};

// ---- setup files (include/read_files.h:13-81 semantics) -----------------------------------
class SetupReader {
 public:
  explicit SetupReader(const std::string& path) : in_(path) {
    if (!in_) throw std::runtime_error("cannot open setup file " + path);
  }
  std::string ReadString() {
    std::string line;
    while (NextLine(&line)) {
      std::istringstream ls(line);
      std::string out, rest;
      if (!(ls >> out)) throw std::runtime_error("Error reading setup file at line number" + std::to_string(line_));
      if (ls >> rest) std::cerr << "Extra text \"" << rest << "\" in line " << line_ << " being ignored." << std::endl;
      return out;
    }
    throw std::runtime_error("Error reading setup file at line number" + std::to_string(line_));
  }
  int ReadNumbers(double* out, int n, bool throw_error = true) {
    int i = 0;
    std::string line, rest;
    std::istringstream ls;
    while (i < n) {
      if (!NextLine(&line)) {
        if (throw_error)
          throw std::runtime_error("Error reading setup file at line number" + std::to_string(line_) +
                                   " (probably not enough entries given)");
        break;
      }
      ls = std::istringstream(line);
      while (i < n && ls >> out[i]) ++i;
    }
    if (ls >> rest) std::cerr << "Extra text \"" << rest << "\" in line " << line_ << " being ignored." << std::endl;
    return i;
  }
  void ExpectKey(const std::string& key) {
    const std::string k = ReadString();
    if (k != key) throw std::runtime_error("setup file: expected key '" + key + "', found '" + k + "'");
  }

 private:
  bool NextLine(std::string* line) {
    while (std::getline(in_, *line)) {
      ++line_;
      if (line->empty() || (*line)[0] == '#' ||
          std::all_of(line->begin(), line->end(), [](unsigned char ch) { return std::isspace(ch); }))
        continue;
      return true;
    }
    return false;
  }
  std::ifstream in_;
  int line_ = 0;
};

}  // namespace cmpc_host
