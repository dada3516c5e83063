// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's plant
// models (katie-jones/compressor-mpc).  Nothing under oracle/ is linked into or
// called by the product path (compressor-mpc_b200/); only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
//
// Restates, without Eigen/Boost:
//   include/valve_eqs.h:25-59            valve mass flow and its pressure derivative
//   systems/compressor.cc:14-66          Compressor<b>::GetDerivative
//   systems/compressor.cc:68-78          CompressorBase::GetOutput
//   systems/compressor.cc:80-175         Compressor<b>::GetLinearizedSystem
//   systems/compressor.cc:177-221        CompressorBase::Parameters
//   systems/tank.cc:10-49                Tank
//   systems/parallel_compressors.cc:9-127, include/parallel_compressors.h:74-99
//   systems/serial_compressors.cc:8-117,  include/serial_compressors.h:85-127
//   include/dynamic_system.h:54-59       GetPlantInput (offset + scatter-add)
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace oracle {

constexpr double kPi = 3.14159265358979323846;
constexpr double kSpeedSound = 340.0;

inline double sign_of(double v) { return (v > 0) - (v < 0); }

// include/valve_eqs.h:16-29
inline double ValveDerivative(double p_in, double p_out, double u_valve,
                              const double C[8], double volume) {
  const double M[4] = {u_valve * u_valve * u_valve, u_valve * u_valve, u_valve, 1.0};
  double dot = 0;
  for (int i = 0; i < 4; ++i) dot += M[i] * C[i];
  return kSpeedSound * kSpeedSound / volume * 1e-5 *
         (sign_of(p_in - p_out) / 2. * 100 / std::sqrt(std::fabs(p_in * 100 - p_out * 100))) *
         dot;
}

// include/valve_eqs.h:39-52
inline double ValveMassFlow(double p_in, double p_out, double u_valve,
                            const double C[8], double m_offset) {
  const double dp_sqrt = 10 * std::sqrt(std::fabs(p_in - p_out)) * sign_of(p_in - p_out);
  const double M3[8] = {dp_sqrt * u_valve * u_valve * u_valve,
                        dp_sqrt * u_valve * u_valve,
                        dp_sqrt * u_valve,
                        dp_sqrt,
                        u_valve * u_valve * u_valve,
                        u_valve * u_valve,
                        u_valve,
                        1.0};
  double dot = 0;
  for (int i = 0; i < 8; ++i) dot += C[i] * M3[i];
  return dot + m_offset;
}

// systems/compressor.cc:177-221
struct CompressorParams {
  double J, tau_r, m_in_c, m_out_c, torque_drive_c, delta_bar, n_bar;
  double V1, V2, AdivL, SD_multiplier;
  double C[8], D[8], A[12], m_rec_ss_c[2], SD_c[2], T_ss_c[3];
  CompressorParams() {
    J = (0.4 + 0.2070) * 0.4;
    tau_r = 1 / 0.5;
    const double a[12] = {0.000299749505193654, -0.000171254191089237, 3.57321648097597e-05,
                          -9.1783572200945e-07, -0.252701086129365,    0.136885752773673,
                          -0.02642368327081,    0.00161012740365743,   54.8046725371143,
                          -29.9550791497765,    5.27827499839098,      0.693826282579158};
    const double c[8] = {-0.423884232813775, 0.626400271518973, -0.0995040168384753,
                         0.0201535563630318, -0.490814924104294, 0.843580880467905,
                         -0.423103455111209, 0.0386841406482887};
    const double d[8] = {-0.0083454, -0.0094965, 0.16826, -0.032215,
                         -0.61199,   0.94175,    -0.48522, 0.10369};
    std::memcpy(A, a, sizeof a);
    std::memcpy(C, c, sizeof c);
    std::memcpy(D, d, sizeof d);
    m_in_c = 0.0051;
    m_rec_ss_c[0] = 0.0047;
    m_rec_ss_c[1] = 0.0263;
    m_out_c = 0.017;
    T_ss_c[0] = 2.5543945754982;
    T_ss_c[1] = 47.4222669576423;
    T_ss_c[2] = 0.6218;
    SD_c[0] = 5.55;
    SD_c[1] = 0.66;
    SD_multiplier = 100;
    torque_drive_c = 15000;
    delta_bar = 0.1;
    n_bar = 1e2;
    V1 = 2 * kPi * (0.60 / 2.0) * (0.60 / 2.0) * 2.0 + kPi * (0.08 / 2.0) * (0.08 / 2.0) * 8.191;
    V2 = kPi * (0.60 / 2.0) * (0.60 / 2.0) * 2.0 + kPi * (0.08 / 2.0) * (0.08 / 2.0) * 5.940;
    AdivL = kPi * (0.08 / 2) * (0.08 / 2) / 3 * 0.1;
  }
};

// Linearised 5-state compressor (row-major like dynamic_system.h:33-38).
struct CompLin {
  double A[5][5], B[5][2], C[2][5], f[5];
};

struct Compressor {
  CompressorParams params_;
  bool has_input_tank;
  explicit Compressor(bool with_tank = true) : has_input_tank(with_tank) {}

  // systems/compressor.cc:14-66.  x = [p1,p2,mc,wc,mr]; u = [td,u_in,u_out,u_rec,p_in|m_in,p_out]
  void GetDerivative(double* m_out, const double x[5], const double u[6], double dxdt[5]) const {
    const CompressorParams& P = params_;
    const double p1 = x[0], p2 = x[1], mc = x[2], wc = x[3], mr = x[4];
    const double td = u[0] * P.torque_drive_c / wc;
    const double u_input = u[1], u_out = u[2], u_rec = u[3], p_out = u[5];
    double m_in;
    if (has_input_tank) {
      m_in = ValveMassFlow(u[4], p1, u_input, P.C, P.m_in_c);
    } else {
      m_in = u[4];
    }
    *m_out = ValveMassFlow(p2, p_out, u_out, P.D, P.m_out_c);
    const double m_rec_ss =
        (P.m_rec_ss_c[0] * (std::sqrt(p2 * 1e5 - p1 * 1e5) * u_rec) + P.m_rec_ss_c[1] * 1.0) *
        (u_rec > 1e-2 ? 1.0 : 0.0);
    const double mc2 = mc * mc, mc3 = mc * mc2, wc2 = wc * wc;
    const double M[12] = {wc2 * mc3, wc2 * mc2, wc2 * mc, wc2, wc * mc3, wc * mc2,
                          wc * mc,   wc,        mc3,      mc2, mc,       1.0};
    double p_ratio = 0;
    for (int i = 0; i < 12; ++i) p_ratio += P.A[i] * M[i];
    const double T_ss_model = P.T_ss_c[0] + P.T_ss_c[1] * mc + P.T_ss_c[2];
    dxdt[0] = kSpeedSound * kSpeedSound / P.V1 * (m_in + mr - mc) * 1e-5;
    dxdt[1] = kSpeedSound * kSpeedSound / P.V2 * (mc - mr - *m_out) * 1e-5;
    dxdt[2] = P.AdivL * (p_ratio * p1 - p2) * 1e5;
    dxdt[3] = (td - T_ss_model) / P.J;
    dxdt[4] = P.tau_r * (m_rec_ss - mr);
  }

  // systems/compressor.cc:68-78
  void GetOutput(const double x[5], double y[2]) const {
    const CompressorParams& P = params_;
    const double p1 = x[0], p2 = x[1], mass_flow = x[2];
    y[0] = p2;
    y[1] = P.SD_multiplier * (-(p2 / p1) / P.SD_c[0] + P.SD_c[1] / P.SD_c[0] + mass_flow);
  }

  // systems/compressor.cc:80-175
  void GetLinearizedSystem(double* m_out, const double x[5], const double u[6],
                           CompLin* lin) const {
    const CompressorParams& P = params_;
    std::memset(lin, 0, sizeof *lin);
    const double p1 = x[0], p2 = x[1], mc = x[2], wc = x[3];
    const double td_in = u[0], u_input = u[1], u_out = u[2], u_rec = u[3], p_out = u[5];
    const double k1 = kSpeedSound * kSpeedSound / P.V1 * 1e-5;
    const double k2 = kSpeedSound * kSpeedSound / P.V2 * 1e-5;
    lin->A[0][0] = -1; lin->A[0][2] = -k1; lin->A[0][4] = k1;
    if (has_input_tank) lin->A[0][0] = -ValveDerivative(u[4], p1, u_input, P.C, P.V1);
    lin->A[1][1] = -ValveDerivative(p2, p_out, u_out, P.D, P.V2);
    lin->A[1][2] = k2; lin->A[1][4] = -k2;
    const double wc2 = wc * wc, mc2 = mc * mc, mc3 = mc * mc2;
    const double dM_dm[12] = {3 * wc2 * mc2, 2 * wc2 * mc, wc2, 0, 3 * wc * mc2, 2 * wc * mc,
                              wc,            0,            3 * mc2, 2 * mc, 1,   0};
    const double dM_dw[12] = {2 * wc * mc3, 2 * wc * mc2, 2 * wc * mc, 2 * wc, mc3, mc2,
                              mc,           1,            0,           0,      0,   0};
    const double M[12] = {wc2 * mc3, wc2 * mc2, wc2 * mc, wc2, wc * mc3, wc * mc2,
                          wc * mc,   wc,        mc3,      mc2, mc,       1.0};
    double p_ratio = 0, a_dm = 0, a_dw = 0;
    for (int i = 0; i < 12; ++i) {
      p_ratio += P.A[i] * M[i];
      a_dm += P.A[i] * dM_dm[i];
      a_dw += P.A[i] * dM_dw[i];
    }
    lin->A[2][0] = P.AdivL * (p_ratio * 1e5);
    lin->A[2][1] = -P.AdivL * 1e5;
    lin->A[2][2] = P.AdivL * (p1 * 1e5) * a_dm;
    lin->A[2][3] = P.AdivL * (p1 * 1e5) * a_dw;
    lin->A[3][2] = -1.0 / P.J * P.T_ss_c[1];
    lin->A[3][3] = -1.0 / P.J * td_in * P.torque_drive_c / wc2;
    const double dsq = std::sqrt(p2 * 1e5 - p1 * 1e5);
    const double a40 = P.tau_r * (P.m_rec_ss_c[0] * 1 / 2 * u_rec / dsq * 1e5);
    lin->A[4][0] = -a40; lin->A[4][1] = a40; lin->A[4][4] = -P.tau_r;
    // B: recycle dead zone smoothed with exponentials (compressor.cc:150-167)
    double dmr_ur = P.tau_r * P.m_rec_ss_c[0] * dsq;
    constexpr double x0 = 1e-2;
    if (u_rec < 2 * x0) {
      double a;
      if (u_rec >= x0) {
        a = P.delta_bar + (1 - P.delta_bar) * std::exp(P.n_bar * (u_rec - x0));
      } else {
        a = 2 - (1 - P.delta_bar) * std::exp(-P.n_bar * u_rec);
      }
      dmr_ur = a * dmr_ur;
    }
    lin->B[3][0] = 1.0 / P.J * P.torque_drive_c / wc;
    lin->B[4][1] = dmr_ur;
    lin->C[0][1] = 1;
    lin->C[1][0] = 100 * p2 / (P.SD_c[0] * p1 * p1);
    lin->C[1][1] = -100. / (P.SD_c[0] * p1);
    lin->C[1][2] = 100;
    GetDerivative(m_out, x, u, lin->f);
  }
};

// systems/tank.cc:10-49
struct Tank {
  double volume, D[8], m_out_c;
  Tank() {
    volume = 20 * kPi * (0.60 / 2) * (0.60 / 2) * 2 + kPi * (0.08 / 2) * (0.08 / 2) * 5.940;
    const double d[8] = {-0.0083454, -0.0094965, 0.16826, -0.032215,
                         -0.61199,   0.94175,    -0.48522, 0.10369};
    std::memcpy(D, d, sizeof d);
    m_out_c = 0.017;
  }
  double GetDerivative(double p_d, const double u[3]) const {
    const double m_out = ValveMassFlow(p_d, u[1], u[0], D, m_out_c);
    return kSpeedSound * kSpeedSound / volume * (u[2] - m_out) * 1e-5;
  }
  double LinearizedA(double p_d, const double u[3]) const {
    return -ValveDerivative(p_d, u[1], u[0], D, volume);
  }
};

enum PlantKind { kParallel = 0, kSerial = 1 };

// Row-major linearisation of the full plant (dynamic_system.h:33-38).
struct PlantLin {
  int n = 0;
  std::vector<double> A, B, C, f;  // A n×n, B n×4, C 4×n, f n
  void Resize(int n_states) {
    n = n_states;
    A.assign(n * n, 0.0);
    B.assign(n * 4, 0.0);
    C.assign(4 * n, 0.0);
    f.assign(n, 0.0);
  }
};

// Two-compressor plant, parallel (+tank) or serial.  4 control inputs = plant
// inputs {0,3,4,7} in both (parallel_compressors.h:28, serial_compressors.h:30).
struct Plant {
  PlantKind kind;
  int n_states, n_inputs;
  static constexpr int n_outputs = 4;
  static constexpr int n_control_inputs = 4;
  Compressor comp0{true}, comp1;
  Tank tank;
  double p_in_ = 1.0, p_out_ = 1.0;

  explicit Plant(PlantKind k) : kind(k), comp1(k == kParallel) {
    n_states = (k == kParallel) ? 11 : 10;
    n_inputs = (k == kParallel) ? 9 : 8;
  }
  static int ControlInputIndex(int i) {
    static const int idx[4] = {0, 3, 4, 7};
    return idx[i];
  }

  // parallel_compressors.h:74-86 / serial_compressors.h:85-95
  std::vector<double> GetDefaultState() const {
    if (kind == kParallel) return {0.916, 1.145, 0.152, 440, 0, 0.916, 1.145, 0.152, 440, 0, 1.12};
    return {0.867, 1.03, 0.176, 395, 0, 0.999, 1.19, 0.176, 395, 0};
  }
  std::vector<double> GetDefaultInput() const {
    if (kind == kParallel) return {0.304, 0.43, 1.0, 0, 0.304, 0.43, 1.0, 0, 0.7};
    return {0.304, 0.405, 1, 0, 0.304, -1, 0.393, 0};
  }

  // dynamic_system.h:54-59 (ExpandArray adds, constexpr_array.h:92-96)
  void GetPlantInput(const double u_control[4], const double* u_offset, double* u) const {
    for (int i = 0; i < n_inputs; ++i) u[i] = u_offset[i];
    for (int i = 0; i < 4; ++i) u[ControlInputIndex(i)] += u_control[i];
  }

  // parallel_compressors.h:94-98 / serial_compressors.h:104-126
  void CompressorInput(const double* u_in, int i, const double* x, double u[6]) const {
    for (int j = 0; j < 4; ++j) u[j] = u_in[i * 4 + j];
    if (kind == kParallel) {
      u[4] = p_in_;
      u[5] = x[n_states - 1];
    } else {
      u[4] = (i == 0) ? p_in_ : -1;
      u[5] = (i == 1) ? p_out_ : x[(i + 1) * 5];
    }
  }

  // parallel_compressors.cc:9-26 / serial_compressors.cc:8-26
  void GetDerivative(const double* x, const double* u, double* dxdt) const {
    double uc[6];
    if (kind == kParallel) {
      double mass_flow, total = 0;
      const Compressor* c[2] = {&comp0, &comp1};
      for (int i = 0; i < 2; ++i) {
        CompressorInput(u, i, x, uc);
        c[i]->GetDerivative(&mass_flow, x + 5 * i, uc, dxdt + 5 * i);
        total += mass_flow;
      }
      const double ut[3] = {u[n_inputs - 1], p_out_, total};
      dxdt[10] = tank.GetDerivative(x[10], ut);
    } else {
      double m_out = -1;
      const Compressor* c[2] = {&comp0, &comp1};
      for (int i = 0; i < 2; ++i) {
        CompressorInput(u, i, x, uc);
        if (i > 0) uc[4] = m_out;
        c[i]->GetDerivative(&m_out, x + 5 * i, uc, dxdt + 5 * i);
      }
    }
  }

  // parallel_compressors.cc:112-127 / serial_compressors.cc:108-117
  void GetOutput(const double* x, double y[4]) const {
    double y0[2], y1[2];
    comp0.GetOutput(x, y0);
    comp1.GetOutput(x + 5, y1);
    if (kind == kParallel) {
      y[0] = y0[1];
      y[1] = y1[1];
      y[2] = y0[0] - y1[0];
      y[3] = x[10];
    } else {
      y[0] = y0[0]; y[1] = y0[1]; y[2] = y1[0]; y[3] = y1[1];
    }
  }

  // parallel_compressors.cc:28-110 / serial_compressors.cc:28-106
  void GetLinearizedSystem(const double* x, const double* u, PlantLin* lin) const {
    const int n = n_states;
    lin->Resize(n);
    auto A = [&](int i, int j) -> double& { return lin->A[i * n + j]; };
    auto B = [&](int i, int j) -> double& { return lin->B[i * 4 + j]; };
    auto C = [&](int i, int j) -> double& { return lin->C[i * n + j]; };
    double uc[6];
    CompLin cl[2];
    if (kind == kParallel) {
      const Compressor* c[2] = {&comp0, &comp1};
      double mass_flow_total = 0, m_out;
      for (int i = 0; i < 2; ++i) {
        CompressorInput(u, i, x, uc);
        c[i]->GetLinearizedSystem(&m_out, x + 5 * i, uc, &cl[i]);
        for (int r = 0; r < 5; ++r) {
          for (int q = 0; q < 5; ++q) A(5 * i + r, 5 * i + q) = cl[i].A[r][q];
          for (int q = 0; q < 2; ++q) B(5 * i + r, 2 * i + q) = cl[i].B[r][q];
          lin->f[5 * i + r] = cl[i].f[r];
        }
        for (int q = 0; q < 5; ++q) C(i, 5 * i + q) = cl[i].C[1][q];  // surge distances
        mass_flow_total += ValveMassFlow(x[5 * i + 1], x[n - 1], u[4 * i + 2], c[i]->params_.D,
                                         c[i]->params_.m_out_c);
        // effect of compressor i on tank / of tank on compressor i / tank self term
        A(10, 5 * i + 1) =
            ValveDerivative(x[5 * i + 1], x[n - 1], u[4 * i + 2], c[i]->params_.D, tank.volume);
        A(5 * i + 1, 10) = ValveDerivative(x[5 * i + 1], x[n - 1], u[4 * i + 2], c[i]->params_.D,
                                           c[i]->params_.V2);
        A(10, 10) += -ValveDerivative(x[5 * i + 1], x[n - 1], u[4 * i + 2], c[i]->params_.D,
                                      tank.volume);
      }
      const double ut[3] = {u[n_inputs - 1], p_out_, mass_flow_total};
      A(10, 10) += tank.LinearizedA(x[10], ut);
      for (int q = 0; q < 5; ++q) {
        C(2, q) = cl[0].C[0][q];
        C(2, 5 + q) = -cl[1].C[0][q];
      }
      C(3, 10) = 1;
      lin->f[10] = tank.GetDerivative(x[10], ut);
    } else {
      double m_out = 0;
      CompressorInput(u, 0, x, uc);
      comp0.GetLinearizedSystem(&m_out, x, uc, &cl[0]);
      for (int r = 0; r < 5; ++r) {
        for (int q = 0; q < 5; ++q) A(r, q) = cl[0].A[r][q];
        for (int q = 0; q < 2; ++q) B(r, q) = cl[0].B[r][q];
      }
      for (int r = 0; r < 2; ++r)
        for (int q = 0; q < 5; ++q) C(r, q) = cl[0].C[r][q];
      // effect of first compressor's p2 on second's p1 (serial_compressors.cc:51-52)
      A(5, 1) = ValveDerivative(x[1], x[5], u[2], comp0.params_.D, comp1.params_.V1);
      // follower: linearised with the m_in = -1 placeholder (serial_compressors.cc:60-62)
      CompressorInput(u, 1, x, uc);
      comp1.GetLinearizedSystem(&m_out, x + 5, uc, &cl[1]);
      for (int r = 0; r < 5; ++r) {
        for (int q = 0; q < 5; ++q) A(5 + r, 5 + q) = cl[1].A[r][q];
        for (int q = 0; q < 2; ++q) B(5 + r, 2 + q) = cl[1].B[r][q];
      }
      A(5, 5) = -ValveDerivative(x[1], x[5], u[2], comp0.params_.D, comp1.params_.V1);
      A(1, 5) = ValveDerivative(x[1], x[5], u[2], comp0.params_.D, comp0.params_.V2);
      for (int r = 0; r < 2; ++r)
        for (int q = 0; q < 5; ++q) C(2 + r, 5 + q) = cl[1].C[r][q];
      GetDerivative(x, u, lin->f.data());  // serial_compressors.cc:102
    }
  }
};

}  // namespace oracle
