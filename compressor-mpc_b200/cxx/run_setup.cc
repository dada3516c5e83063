// Driver in the shape of the reference's tests/*-with-timing.cc executables: reads a setup file,
// runs the closed loop (controller + plant + actuator delay) for every simulation block on the
// GPU and writes the reference's .dat records.
//   cmpc_run_setup <setup-file> <parallel|serial> <centralized|cooperative|noncoop|noncoop-old> [batch]
// Record layout and key order: SURVEY.md 3.1 (reconstructed tests/common-simulation.inc).
#include <cstdio>
#include <cstring>

#include "cmpc_facade.hpp"

using namespace cmpc_host;

static std::string FormatRow(const double* v, int n) {
  std::vector<std::string> t(n);
  size_t w = 0;
  for (int i = 0; i < n; ++i) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%g", v[i]);
    t[i] = buf;
    w = std::max(w, t[i].size());
  }
  std::string s;
  for (int i = 0; i < n; ++i) s += std::string(w - t[i].size(), ' ') + t[i] + (i + 1 < n ? " " : "");
  return s;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <setup-file> <parallel|serial> <centralized|cooperative|noncoop|noncoop-old> [batch]\n", argv[0]);
    return 2;
  }
  try {
    const int plant = std::strcmp(argv[2], "serial") == 0 ? CMPC_PLANT_SERIAL : CMPC_PLANT_PARALLEL;
    const int mode = std::strcmp(argv[3], "centralized") == 0 ? CMPC_MODE_CENTRALIZED
                     : std::strcmp(argv[3], "cooperative") == 0 ? CMPC_MODE_COOPERATIVE
                     : std::strcmp(argv[3], "noncoop-old") == 0 ? CMPC_MODE_NONCOOPERATIVE_OLD : CMPC_MODE_NONCOOPERATIVE;
    const int batch = argc > 4 ? std::atoi(argv[4]) : 1;
    SetupReader rd(argv[1]);
    double v[64];
    rd.ExpectKey("n-iterations"); rd.ReadNumbers(v, 1); const int n_iter = int(v[0]);
    rd.ExpectKey("n-timing-iterations"); rd.ReadNumbers(v, 1); const int n_timing = int(v[0]);
    rd.ExpectKey("folder-name"); const std::string folder = rd.ReadString();
    rd.ExpectKey("output-filename"); const std::string fname = rd.ReadString();
    NerveCenter nc(plant, mode, n_iter, batch);
    const int n = nc.n_states(), ni = nc.n_inputs(), nu = nc.n_sub_control_inputs();
    double yref[4], uwt[16];
    rd.ExpectKey("yref"); rd.ReadNumbers(yref, 4);
    rd.ExpectKey("uwt"); rd.ReadNumbers(uwt, 16);
    rd.ExpectKey("ywt");
    std::vector<std::vector<double>> ywts(nc.n_controllers());
    for (int c = 0; c < nc.n_controllers(); ++c) {
      const int ny = nc.n_controlled_outputs(c);
      ywts[c].resize(ny * ny);
      rd.ReadNumbers(ywts[c].data(), ny * ny);
    }
    double lo[4], up[4], rlo[4], rup[4];
    rd.ExpectKey("constraints-lower"); rd.ReadNumbers(lo, nu);
    rd.ExpectKey("constraints-upper"); rd.ReadNumbers(up, nu);
    rd.ExpectKey("constraints-rate-lower"); rd.ReadNumbers(rlo, nu);
    rd.ExpectKey("constraints-rate-upper"); rd.ReadNumbers(rup, nu);
    rd.ExpectKey("simulation");
    std::vector<double> offs;
    std::vector<double> tends;
    for (;;) {
      const int got = rd.ReadNumbers(v, ni + 1, false);
      if (got < ni + 1) break;
      offs.insert(offs.end(), v, v + ni);
      tends.push_back(v[ni]);
    }
    if (tends.empty()) throw std::runtime_error("setup file: no simulation block");
    nc.SetWeights(uwt, ywts);
    nc.SetOutputReferenceConstant(yref);
    for (int c = 0; c < nc.n_controllers(); ++c) nc.SetConstraints(c, lo, up, rlo, rup);
    // block ends in records: the reference accumulates t += Ts and runs a block while t < t_end
    const int nb = int(tends.size());
    std::vector<int32_t> block_end(nb);
    double t = 0;
    int k = 0;
    for (int b = 0; b < nb; ++b) {
      while (t < tends[b]) { t += 0.05; ++k; }
      block_end[b] = k;
    }
    const int T = block_end[nb - 1], rec = 1 + n + 8;
    std::vector<double> x_def(n), u_def(ni);
    Check(cmpc_plant_defaults(plant, x_def.data(), u_def.data()));
    std::vector<double> x0(size_t(batch) * n), bo(size_t(batch) * nb * ni);
    std::vector<int32_t> be(size_t(batch) * nb);
    for (int b = 0; b < batch; ++b) {
      std::copy(x_def.begin(), x_def.end(), x0.begin() + size_t(b) * n);
      std::copy(offs.begin(), offs.end(), bo.begin() + size_t(b) * nb * ni);
      std::copy(block_end.begin(), block_end.end(), be.begin() + size_t(b) * nb);
    }
    std::vector<double> traj(size_t(batch) * T * rec);
    // every record carries the time of its own control step inside the reference's timing window
    // (GetNextInputWithTiming(y, n_timing_iterations, &time), nerve_center.h:134-182)
    std::vector<int64_t> step_ns(T);
    Check(cmpc_run_closed_loop_timed(nc.handle(), T, x0.data(), nb, be.data(), bo.data(), traj.data(), nullptr, nullptr,
                                     nullptr, n_timing, step_ns.data()));
    double mean_ns = 0;
    for (int r = 0; r < T; ++r) mean_ns += double(step_ns[r]) / T;
    std::ofstream out(folder + "/" + fname);
    if (!out) throw std::runtime_error("cannot write " + folder + "/" + fname + " (does the folder exist?)");
    for (int r = 0; r < T; ++r) {
      const double* q = traj.data() + size_t(r) * rec;   // scenario 0
      char tb[32];
      std::snprintf(tb, sizeof tb, "%g", q[0]);
      out << tb << "\n" << FormatRow(q + 1, n) << "\n" << FormatRow(q + 1 + n, 4) << "\n" << FormatRow(q + 5 + n, 4)
          << "\n" << (long long)step_ns[r] << "\n\n";
    }
    std::printf("%d records x %d scenario(s) -> %s/%s, %.1f us per batched control step\n", T, batch, folder.c_str(),
                fname.c_str(), mean_ns / 1e3);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
