// Device-side plant models: two centrifugal compressors in parallel (+ discharge tank)
// or in series.  Same physics and fitted maps as the reference plant classes:
//   systems/compressor.cc:14-66 (derivative), :68-78 (output), :80-175 (Jacobians),
//   :177-221 (parameters); systems/tank.cc:10-49; include/valve_eqs.h:16-52;
//   systems/parallel_compressors.cc:9-127; systems/serial_compressors.cc:8-117.
// Written for one GPU thread per compressor: every routine is branch-light scalar FP64.
#pragma once
#include <cuda_runtime.h>

namespace cmpc {

constexpr double kPi = 3.14159265358979323846;
constexpr double kSS2 = 340.0 * 340.0 * 1e-5;  // speed_sound^2 * 1e-5

// --- compressor parameters (compressor.cc:177-221) ---
constexpr double kJ = (0.4 + 0.2070) * 0.4;
constexpr double kTauR = 1 / 0.5;
constexpr double kMinC = 0.0051;
constexpr double kMoutC = 0.017;
constexpr double kTorqueDriveC = 15000;
constexpr double kDeltaBar = 0.1;
constexpr double kNBar = 1e2;
constexpr double kV1 = 2 * kPi * (0.60 / 2.0) * (0.60 / 2.0) * 2.0 + kPi * (0.08 / 2.0) * (0.08 / 2.0) * 8.191;
constexpr double kV2 = kPi * (0.60 / 2.0) * (0.60 / 2.0) * 2.0 + kPi * (0.08 / 2.0) * (0.08 / 2.0) * 5.940;
constexpr double kAdivL = kPi * (0.08 / 2) * (0.08 / 2) / 3 * 0.1;
constexpr double kSDc0 = 5.55, kSDc1 = 0.66, kSDmult = 100;
constexpr double kMrec0 = 0.0047, kMrec1 = 0.0263;
constexpr double kTss0 = 2.5543945754982, kTss1 = 47.4222669576423, kTss2 = 0.6218;
constexpr double kTankVolume = 20 * kPi * (0.60 / 2) * (0.60 / 2) * 2 + kPi * (0.08 / 2) * (0.08 / 2) * 5.940;

// pressure-ratio map, 12 coefficients over [wc^2, wc, 1] x [mc^3, mc^2, mc, 1]
__device__ __forceinline__ double map_a(int i) {
  constexpr double a[12] = {0.000299749505193654, -0.000171254191089237, 3.57321648097597e-05,
                            -9.1783572200945e-07, -0.252701086129365,    0.136885752773673,
                            -0.02642368327081,    0.00161012740365743,   54.8046725371143,
                            -29.9550791497765,    5.27827499839098,      0.693826282579158};
  return a[i];
}
// inlet valve (C) and outlet valve (D) maps, 8 coefficients each
struct ValveC {
  static __device__ __forceinline__ double c(int i) {
    constexpr double v[8] = {-0.423884232813775, 0.626400271518973, -0.0995040168384753,
                             0.0201535563630318, -0.490814924104294, 0.843580880467905,
                             -0.423103455111209, 0.0386841406482887};
    return v[i];
  }
};
struct ValveD {
  static __device__ __forceinline__ double c(int i) {
    constexpr double v[8] = {-0.0083454, -0.0094965, 0.16826, -0.032215,
                             -0.61199,   0.94175,    -0.48522, 0.10369};
    return v[i];
  }
};

__device__ __forceinline__ double sgn(double v) { return (v > 0.0) - (v < 0.0); }

// ---- straight-line square root and division for the plant integrator ---------------------------
// sqrt() and / compile to a fast path plus a branch to a special-case routine each; the branches cut
// the derivative into basic blocks, so its five independent sqrt / div chains run one after the
// other.  The integrator is one warp per block in a one-wave kernel, i.e. bound by exactly that
// dependent latency.  These forms are the fast path alone (hardware seed, two coupled Newton steps,
// one residual correction with fused multiply-adds, which rounds like the IEEE operation for operands
// in the normal range) and only RECORD in `bad` that an operand was outside [1e-290, 1e290]; the
// caller then redoes the whole derivative with the standard operations (never in a sane plant state).
__device__ __forceinline__ double sqrt_inrange(double x, bool& bad) {
  bad |= !(x >= 1e-290 && x <= 1e290);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
}
__device__ __forceinline__ double div_inrange(double a, double b, bool& bad) {
  bad |= !(fabs(b) >= 1e-290 && fabs(b) <= 1e290 && fabs(a) <= 1e290 && (fabs(a) >= 1e-290 || a == 0.0));
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
  double e = fma(-b, y, 1.0);
  y = fma(y, e, y);
  e = fma(-b, y, 1.0);
  y = fma(y, e, y);
  const double q = a * y;
  return fma(fma(-b, q, a), y, q);
}

// valve_eqs.h:39-52
template <class V>
__device__ __forceinline__ double valve_mass_flow(double p_in, double p_out, double u, double m_offset) {
  const double dp = 10.0 * sqrt(fabs(p_in - p_out)) * sgn(p_in - p_out);
  const double u2 = u * u, u3 = u2 * u;
  const double poly_dp = V::c(0) * u3 + V::c(1) * u2 + V::c(2) * u + V::c(3);
  const double poly_0 = V::c(4) * u3 + V::c(5) * u2 + V::c(6) * u + V::c(7);
  return dp * poly_dp + poly_0 + m_offset;
}
// the same with the straight-line square root (lane-pair integrator)
template <class V>
__device__ __forceinline__ double valve_mass_flow_fast(double p_in, double p_out, double u, double m_offset, bool& bad) {
  const double dp = 10.0 * sqrt_inrange(fabs(p_in - p_out), bad) * sgn(p_in - p_out);
  const double u2 = u * u, u3 = u2 * u;
  const double poly_dp = V::c(0) * u3 + V::c(1) * u2 + V::c(2) * u + V::c(3);
  const double poly_0 = V::c(4) * u3 + V::c(5) * u2 + V::c(6) * u + V::c(7);
  return dp * poly_dp + poly_0 + m_offset;
}
// valve_eqs.h:16-29 (only the first four map coefficients enter)
template <class V>
__device__ __forceinline__ double valve_derivative(double p_in, double p_out, double u, double volume) {
  const double u2 = u * u, u3 = u2 * u;
  const double poly = V::c(0) * u3 + V::c(1) * u2 + V::c(2) * u + V::c(3);
  return (340.0 * 340.0) / volume * 1e-5 *
         (sgn(p_in - p_out) / 2. * 100 / sqrt(fabs(p_in * 100 - p_out * 100))) * poly;
}

// compressor.cc:14-66.  x = [p1,p2,mc,wc,mr]; u = [td,u_in,u_out,u_rec]; p_or_m_in = inlet
// pressure (HAS_TANK) or inlet mass flow; p_out = downstream pressure.
template <bool HAS_TANK>
__device__ __forceinline__ void compressor_derivative(const double x[5], const double u[4],
                                                      double p_or_m_in, double p_out,
                                                      double dxdt[5], double* m_out) {
  const double p1 = x[0], p2 = x[1], mc = x[2], wc = x[3], mr = x[4];
  const double td = u[0] * kTorqueDriveC / wc;
  const double m_in = HAS_TANK ? valve_mass_flow<ValveC>(p_or_m_in, p1, u[1], kMinC) : p_or_m_in;
  *m_out = valve_mass_flow<ValveD>(p2, p_out, u[2], kMoutC);
  const double m_rec_ss =
      (u[3] > 1e-2) ? (kMrec0 * (sqrt(p2 * 1e5 - p1 * 1e5) * u[3]) + kMrec1) : 0.0;
  const double mc2 = mc * mc, mc3 = mc * mc2, wc2 = wc * wc;
  const double q2 = map_a(0) * mc3 + map_a(1) * mc2 + map_a(2) * mc + map_a(3);
  const double q1 = map_a(4) * mc3 + map_a(5) * mc2 + map_a(6) * mc + map_a(7);
  const double q0 = map_a(8) * mc3 + map_a(9) * mc2 + map_a(10) * mc + map_a(11);
  const double p_ratio = wc2 * q2 + wc * q1 + q0;
  const double T_ss = kTss0 + kTss1 * mc + kTss2;
  dxdt[0] = (340.0 * 340.0) / kV1 * (m_in + mr - mc) * 1e-5;
  dxdt[1] = (340.0 * 340.0) / kV2 * (mc - mr - *m_out) * 1e-5;
  dxdt[2] = kAdivL * (p_ratio * p1 - p2) * 1e5;
  dxdt[3] = (td - T_ss) / kJ;
  dxdt[4] = kTauR * (m_rec_ss - mr);
}

// Same with the inlet mass flow given (used by the lane-pair integrator).  FAST: straight-line square
// roots and division that flag out-of-range operands in `bad` instead of handling them.
template <bool FAST>
__device__ __forceinline__ void compressor_derivative_core(const double x[5], const double u[4], double m_in,
                                                           double p_out, double dxdt[5], double* m_out, bool& bad) {
  const double p1 = x[0], p2 = x[1], mc = x[2], wc = x[3], mr = x[4];
  const double td = FAST ? div_inrange(u[0] * kTorqueDriveC, wc, bad) : u[0] * kTorqueDriveC / wc;
  *m_out = FAST ? valve_mass_flow_fast<ValveD>(p2, p_out, u[2], kMoutC, bad) : valve_mass_flow<ValveD>(p2, p_out, u[2], kMoutC);
  double m_rec_ss = 0.0;
  if (FAST) {
    // the root is only used with the recycle valve open: an out-of-range operand behind a closed valve is no fault
    bool bad_r = false;
    const double rt = sqrt_inrange(p2 * 1e5 - p1 * 1e5, bad_r);
    const bool open = u[3] > 1e-2;
    bad |= open && bad_r;
    m_rec_ss = open ? (kMrec0 * (rt * u[3]) + kMrec1) : 0.0;
  } else {
    m_rec_ss = (u[3] > 1e-2) ? (kMrec0 * (sqrt(p2 * 1e5 - p1 * 1e5) * u[3]) + kMrec1) : 0.0;
  }
  const double mc2 = mc * mc, mc3 = mc * mc2, wc2 = wc * wc;
  const double q2 = map_a(0) * mc3 + map_a(1) * mc2 + map_a(2) * mc + map_a(3);
  const double q1 = map_a(4) * mc3 + map_a(5) * mc2 + map_a(6) * mc + map_a(7);
  const double q0 = map_a(8) * mc3 + map_a(9) * mc2 + map_a(10) * mc + map_a(11);
  const double p_ratio = wc2 * q2 + wc * q1 + q0;
  const double T_ss = kTss0 + kTss1 * mc + kTss2;
  dxdt[0] = (340.0 * 340.0) / kV1 * (m_in + mr - mc) * 1e-5;
  dxdt[1] = (340.0 * 340.0) / kV2 * (mc - mr - *m_out) * 1e-5;
  dxdt[2] = kAdivL * (p_ratio * p1 - p2) * 1e5;
  dxdt[3] = (td - T_ss) / kJ;
  dxdt[4] = kTauR * (m_rec_ss - mr);
}

// compressor.cc:68-78
__device__ __forceinline__ void compressor_output(const double x[5], double* p2_out, double* sd_out) {
  *p2_out = x[1];
  *sd_out = kSDmult * (-(x[1] / x[0]) / kSDc0 + kSDc1 / kSDc0 + x[2]);
}

// Non-trivial entries of one compressor's Jacobians (compressor.cc:80-175).
struct CompJac {
  double a00, a11;            // valve derivatives (a00 = -1 placeholder when !HAS_TANK)
  double a20, a22, a23;       // pressure-ratio map row (a21 = -AdivL*1e5 constant)
  double a33;                 // torque term (a32 constant)
  double a40;                 // recycle row: A[4][0] = -a40, A[4][1] = +a40
  double b30, b41;            // B[3][0] (torque), B[4][1] (recycle valve)
  double c10, c11;            // surge-distance row of C (c12 = 100)
};
constexpr double kA02 = -(340.0 * 340.0) / kV1 * 1e-5;  // A[0][2]; A[0][4] = -kA02
constexpr double kA12 = (340.0 * 340.0) / kV2 * 1e-5;   // A[1][2]; A[1][4] = -kA12
constexpr double kA21 = -kAdivL * 1e5;
constexpr double kA32 = -1.0 / kJ * kTss1;
constexpr double kA44 = -kTauR;

template <bool HAS_TANK>
__device__ __forceinline__ void compressor_jacobian(const double x[5], const double u[4],
                                                    double p_in, double p_out, CompJac* j) {
  const double p1 = x[0], p2 = x[1], mc = x[2], wc = x[3];
  j->a00 = HAS_TANK ? -valve_derivative<ValveC>(p_in, p1, u[1], kV1) : -1.0;
  j->a11 = -valve_derivative<ValveD>(p2, p_out, u[2], kV2);
  const double wc2 = wc * wc, mc2 = mc * mc, mc3 = mc * mc2;
  const double q2 = map_a(0) * mc3 + map_a(1) * mc2 + map_a(2) * mc + map_a(3);
  const double q1 = map_a(4) * mc3 + map_a(5) * mc2 + map_a(6) * mc + map_a(7);
  const double q0 = map_a(8) * mc3 + map_a(9) * mc2 + map_a(10) * mc + map_a(11);
  const double dq2 = 3 * map_a(0) * mc2 + 2 * map_a(1) * mc + map_a(2);
  const double dq1 = 3 * map_a(4) * mc2 + 2 * map_a(5) * mc + map_a(6);
  const double dq0 = 3 * map_a(8) * mc2 + 2 * map_a(9) * mc + map_a(10);
  const double p_ratio = wc2 * q2 + wc * q1 + q0;
  j->a20 = kAdivL * (p_ratio * 1e5);
  j->a22 = kAdivL * (p1 * 1e5) * (wc2 * dq2 + wc * dq1 + dq0);
  j->a23 = kAdivL * (p1 * 1e5) * (2 * wc * q2 + q1);
  j->a33 = -1.0 / kJ * u[0] * kTorqueDriveC / wc2;
  const double dsq = sqrt(p2 * 1e5 - p1 * 1e5);
  j->a40 = kTauR * (kMrec0 * 0.5 * u[3] / dsq * 1e5);
  double dmr_ur = kTauR * kMrec0 * dsq;
  constexpr double x0 = 1e-2;
  if (u[3] < 2 * x0) {  // dead zone smoothed with exponentials (compressor.cc:150-167)
    const double a = (u[3] >= x0) ? kDeltaBar + (1 - kDeltaBar) * exp(kNBar * (u[3] - x0))
                                  : 2 - (1 - kDeltaBar) * exp(-kNBar * u[3]);
    dmr_ur = a * dmr_ur;
  }
  j->b30 = 1.0 / kJ * kTorqueDriveC / wc;
  j->b41 = dmr_ur;
  j->c10 = 100 * p2 / (kSDc0 * p1 * p1);
  j->c11 = -100. / (kSDc0 * p1);
}

template <int PLANT>
struct PlantDims {
  static constexpr int N = PLANT == 0 ? 11 : 10;
  static constexpr int NIN = PLANT == 0 ? 9 : 8;
};

// Full-plant derivative (parallel_compressors.cc:9-26 / serial_compressors.cc:8-26).
template <int PLANT>
__device__ __forceinline__ void plant_derivative(const double* x, const double* u, double* dxdt) {
  if (PLANT == 0) {
    double m0, m1;
    compressor_derivative<true>(x, u, 1.0, x[10], dxdt, &m0);
    compressor_derivative<true>(x + 5, u + 4, 1.0, x[10], dxdt + 5, &m1);
    const double m_out_tank = valve_mass_flow<ValveD>(x[10], 1.0, u[8], kMoutC);
    dxdt[10] = (340.0 * 340.0) / kTankVolume * ((m0 + m1) - m_out_tank) * 1e-5;
  } else {
    double m0, m1;
    compressor_derivative<true>(x, u, 1.0, x[5], dxdt, &m0);
    compressor_derivative<false>(x + 5, u + 4, m0, 1.0, dxdt + 5, &m1);
  }
}

// parallel_compressors.cc:112-127 / serial_compressors.cc:108-117
template <int PLANT>
__device__ __forceinline__ void plant_output(const double* x, double y[4]) {
  double p20, sd0, p21, sd1;
  compressor_output(x, &p20, &sd0);
  compressor_output(x + 5, &p21, &sd1);
  if (PLANT == 0) {
    y[0] = sd0; y[1] = sd1; y[2] = p20 - p21; y[3] = x[10];
  } else {
    y[0] = p20; y[1] = sd0; y[2] = p21; y[3] = sd1;
  }
}

// Row `r` of the plant's C matrix at state x (only p1, p2 of each compressor enter).
// Returns the three possibly non-zero entries and their first column.
template <int PLANT>
__device__ __forceinline__ void plant_c_entry(const double* x, double* Cx /*4 x N row-major*/) {
  constexpr int N = PlantDims<PLANT>::N;
  for (int i = 0; i < 4 * N; ++i) Cx[i] = 0.0;
  const double c10a = 100 * x[1] / (kSDc0 * x[0] * x[0]), c11a = -100. / (kSDc0 * x[0]);
  const double c10b = 100 * x[6] / (kSDc0 * x[5] * x[5]), c11b = -100. / (kSDc0 * x[5]);
  if (PLANT == 0) {
    Cx[0 * N + 0] = c10a; Cx[0 * N + 1] = c11a; Cx[0 * N + 2] = 100;
    Cx[1 * N + 5] = c10b; Cx[1 * N + 6] = c11b; Cx[1 * N + 7] = 100;
    Cx[2 * N + 1] = 1; Cx[2 * N + 6] = -1;
    Cx[3 * N + 10] = 1;
  } else {
    Cx[0 * N + 1] = 1;
    Cx[1 * N + 0] = c10a; Cx[1 * N + 1] = c11a; Cx[1 * N + 2] = 100;
    Cx[2 * N + 6] = 1;
    Cx[3 * N + 5] = c10b; Cx[3 * N + 6] = c11b; Cx[3 * N + 7] = 100;
  }
}

// Continuous-time linearisation of the full plant written densely (row-major):
// A N x N, Bc N x 4 (system control-input order), C 4 x N, f N.
// parallel_compressors.cc:28-110 / serial_compressors.cc:28-106.
template <int PLANT>
__device__ void plant_linearize(const double* x, const double* u, double* A, double* Bc, double* C,
                                double* f) {
  constexpr int N = PlantDims<PLANT>::N;
  for (int i = 0; i < N * N; ++i) A[i] = 0.0;
  for (int i = 0; i < N * 4; ++i) Bc[i] = 0.0;
  plant_c_entry<PLANT>(x, C);
  plant_derivative<PLANT>(x, u, f);
  CompJac j0, j1;
  if (PLANT == 0) {
    compressor_jacobian<true>(x, u, 1.0, x[10], &j0);
    compressor_jacobian<true>(x + 5, u + 4, 1.0, x[10], &j1);
  } else {
    compressor_jacobian<true>(x, u, 1.0, x[5], &j0);
    compressor_jacobian<false>(x + 5, u + 4, -1.0, 1.0, &j1);
  }
  const CompJac* js[2] = {&j0, &j1};
  for (int i = 0; i < 2; ++i) {
    const CompJac& j = *js[i];
    const int o = 5 * i;
    A[(o + 0) * N + o + 0] = j.a00; A[(o + 0) * N + o + 2] = kA02; A[(o + 0) * N + o + 4] = -kA02;
    A[(o + 1) * N + o + 1] = j.a11; A[(o + 1) * N + o + 2] = kA12; A[(o + 1) * N + o + 4] = -kA12;
    A[(o + 2) * N + o + 0] = j.a20; A[(o + 2) * N + o + 1] = kA21;
    A[(o + 2) * N + o + 2] = j.a22; A[(o + 2) * N + o + 3] = j.a23;
    A[(o + 3) * N + o + 2] = kA32;  A[(o + 3) * N + o + 3] = j.a33;
    A[(o + 4) * N + o + 0] = -j.a40; A[(o + 4) * N + o + 1] = j.a40; A[(o + 4) * N + o + 4] = kA44;
    Bc[(o + 3) * 4 + 2 * i + 0] = j.b30;
    Bc[(o + 4) * 4 + 2 * i + 1] = j.b41;
  }
  if (PLANT == 0) {
    double a1010 = 0;
    for (int i = 0; i < 2; ++i) {
      const double p2 = x[5 * i + 1], uo = u[4 * i + 2];
      const double dv_tank = valve_derivative<ValveD>(p2, x[10], uo, kTankVolume);
      A[10 * N + 5 * i + 1] = dv_tank;
      A[(5 * i + 1) * N + 10] = valve_derivative<ValveD>(p2, x[10], uo, kV2);
      a1010 -= dv_tank;
    }
    a1010 += -valve_derivative<ValveD>(x[10], 1.0, u[8], kTankVolume);
    A[10 * N + 10] = a1010;
  } else {
    const double dv1 = valve_derivative<ValveD>(x[1], x[5], u[2], kV1);
    A[5 * N + 1] = dv1;
    A[5 * N + 5] = -dv1;
    A[1 * N + 5] = valve_derivative<ValveD>(x[1], x[5], u[2], kV2);
  }
}

// The same linearisation split over three threads (part 0/1: one compressor each, part 2: the
// coupling terms and the tank).  Writes only the entries that can be non-zero: A (stride lda),
// X = [B (columns permuted by inv: system input -> column) | f] (stride ldx), C (4 x N dense).
// The destination's other entries must already be zero (they never change).
template <int PLANT>
__device__ void plant_linearize_part_x(int part, const double* x, const double* u, double* A, int lda,
                                       double* X, int ldx, const int* inv, double* C) {
  constexpr int N = PlantDims<PLANT>::N;
  if (part < 2) {
    const int i = part, o = 5 * part;
    CompJac j;
    double m_out, fl[5];
    if (PLANT == 0) {
      compressor_jacobian<true>(x + o, u + 4 * i, 1.0, x[10], &j);
      compressor_derivative<true>(x + o, u + 4 * i, 1.0, x[10], fl, &m_out);
    } else if (i == 0) {
      compressor_jacobian<true>(x, u, 1.0, x[5], &j);
      compressor_derivative<true>(x, u, 1.0, x[5], fl, &m_out);
    } else {
      compressor_jacobian<false>(x + 5, u + 4, -1.0, 1.0, &j);
      const double m0 = valve_mass_flow<ValveD>(x[1], x[5], u[2], kMoutC);  // first compressor's outflow
      compressor_derivative<false>(x + 5, u + 4, m0, 1.0, fl, &m_out);
    }
    A[(o + 0) * lda + o + 0] = j.a00; A[(o + 0) * lda + o + 2] = kA02; A[(o + 0) * lda + o + 4] = -kA02;
    A[(o + 1) * lda + o + 1] = j.a11; A[(o + 1) * lda + o + 2] = kA12; A[(o + 1) * lda + o + 4] = -kA12;
    A[(o + 2) * lda + o + 0] = j.a20; A[(o + 2) * lda + o + 1] = kA21;
    A[(o + 2) * lda + o + 2] = j.a22; A[(o + 2) * lda + o + 3] = j.a23;
    A[(o + 3) * lda + o + 2] = kA32;  A[(o + 3) * lda + o + 3] = j.a33;
    A[(o + 4) * lda + o + 0] = -j.a40; A[(o + 4) * lda + o + 1] = j.a40; A[(o + 4) * lda + o + 4] = kA44;
    X[(o + 3) * ldx + inv[2 * i + 0]] = j.b30;
    X[(o + 4) * ldx + inv[2 * i + 1]] = j.b41;
#pragma unroll
    for (int r = 0; r < 5; ++r) X[(o + r) * ldx + 4] = fl[r];
    // this compressor's rows of C
    if (PLANT == 0) {
      C[i * N + o + 0] = j.c10; C[i * N + o + 1] = j.c11; C[i * N + o + 2] = 100;
    } else {
      C[(2 * i) * N + o + 1] = 1;
      C[(2 * i + 1) * N + o + 0] = j.c10; C[(2 * i + 1) * N + o + 1] = j.c11; C[(2 * i + 1) * N + o + 2] = 100;
    }
    if (PLANT == 1 && i == 1) {
      // serial coupling: the valve between the compressors (serial_compressors.cc:51-52,70-85);
      // written by this thread after its own block so that A(5,5) is not overwritten
      const double dv1 = valve_derivative<ValveD>(x[1], x[5], u[2], kV1);
      A[5 * lda + 1] = dv1;
      A[5 * lda + 5] = -dv1;
      A[1 * lda + 5] = valve_derivative<ValveD>(x[1], x[5], u[2], kV2);
    }
  } else if (PLANT == 0) {
    double a1010 = 0, m_total = 0;
    for (int i = 0; i < 2; ++i) {
      const double p2 = x[5 * i + 1], uo = u[4 * i + 2];
      const double dv_tank = valve_derivative<ValveD>(p2, x[10], uo, kTankVolume);
      A[10 * lda + 5 * i + 1] = dv_tank;
      A[(5 * i + 1) * lda + 10] = valve_derivative<ValveD>(p2, x[10], uo, kV2);
      a1010 -= dv_tank;
      m_total += valve_mass_flow<ValveD>(p2, x[10], uo, kMoutC);
    }
    a1010 += -valve_derivative<ValveD>(x[10], 1.0, u[8], kTankVolume);
    A[10 * lda + 10] = a1010;
    C[2 * N + 1] = 1; C[2 * N + 6] = -1;
    C[3 * N + 10] = 1;
    const double m_out_tank = valve_mass_flow<ValveD>(x[10], 1.0, u[8], kMoutC);
    X[10 * ldx + 4] = (340.0 * 340.0) / kTankVolume * (m_total - m_out_tank) * 1e-5;
  }
}

// ---- lane-pair form of the plant ODE --------------------------------------------------------
// Two neighbouring lanes (parity c = lane & 1) integrate one scenario: lane parity c holds
// compressor c's five states in xs[0..4]; for the parallel plant both also hold the tank pressure
// in xs[5] and compute its (identical) derivative.  uc = that compressor's four inputs, u_tank =
// tank valve (parallel plant).  All 32 lanes of the warp must call it together (the exchange inside
// the pair is a whole-warp shuffle; see dopri5_try_step_pair for why).
template <int PLANT, bool FAST>
__device__ __forceinline__ void pair_derivative_impl(int c, const double xs[6], const double uc[4],
                                                     double u_tank, double d[6], bool& bad) {
  constexpr unsigned full = 0xffffffffu;
  if (PLANT == 0) {
    const double m_in = FAST ? valve_mass_flow_fast<ValveC>(1.0, xs[0], uc[1], kMinC, bad)
                             : valve_mass_flow<ValveC>(1.0, xs[0], uc[1], kMinC);
    double m_out;
    compressor_derivative_core<FAST>(xs, uc, m_in, xs[5], d, &m_out, bad);
    const double m_other = __shfl_xor_sync(full, m_out, 1);
    const double m0 = c == 0 ? m_out : m_other, m1 = c == 0 ? m_other : m_out;
    const double m_out_tank = FAST ? valve_mass_flow_fast<ValveD>(xs[5], 1.0, u_tank, kMoutC, bad)
                                   : valve_mass_flow<ValveD>(xs[5], 1.0, u_tank, kMoutC);
    d[5] = (340.0 * 340.0) / kTankVolume * ((m0 + m1) - m_out_tank) * 1e-5;
  } else {
    // serial: compressor 0 discharges into compressor 1's inlet volume
    const double o_p1 = __shfl_xor_sync(full, xs[0], 1);   // the other compressor's p1
    const double o_p2 = __shfl_xor_sync(full, xs[1], 1);   // the other compressor's p2
    const double o_uout = __shfl_xor_sync(full, uc[2], 1); // the other compressor's outlet valve
    double m_in, p_out;
    if (c == 0) {
      m_in = FAST ? valve_mass_flow_fast<ValveC>(1.0, xs[0], uc[1], kMinC, bad) : valve_mass_flow<ValveC>(1.0, xs[0], uc[1], kMinC);
      p_out = o_p1;
    } else {
      m_in = FAST ? valve_mass_flow_fast<ValveD>(o_p2, xs[0], o_uout, kMoutC, bad)
                  : valve_mass_flow<ValveD>(o_p2, xs[0], o_uout, kMoutC);   // first compressor's outflow
      p_out = 1.0;
    }
    double m_out;
    compressor_derivative_core<FAST>(xs, uc, m_in, p_out, d, &m_out, bad);
    d[5] = 0.0;
  }
}
// The derivative as the integrator uses it: straight-line arithmetic first; a lane pair with an
// operand outside the range takes the result of the standard operations instead (computed by the
// whole warp, used by that pair only, so a scenario's numbers never depend on its neighbours).
template <int PLANT>
__device__ __forceinline__ void pair_derivative(int c, const double xs[6], const double uc[4], double u_tank, double d[6]) {
  bool bad = false;
#ifdef CMPC_NO_FAST_DERIV
  pair_derivative_impl<PLANT, false>(c, xs, uc, u_tank, d, bad);
#else
  pair_derivative_impl<PLANT, true>(c, xs, uc, u_tank, d, bad);
  const int bad_other = __shfl_xor_sync(0xffffffffu, int(bad), 1);   // (every lane takes part: no short circuit)
  const bool pair_bad = bad || bad_other != 0;
  if (__any_sync(0xffffffffu, pair_bad)) {
    double ds[6];
    bool unused = false;
    pair_derivative_impl<PLANT, false>(c, xs, uc, u_tank, ds, unused);
    if (pair_bad) {
#pragma unroll
      for (int i = 0; i < 6; ++i) d[i] = ds[i];
    }
  }
#endif
}

}  // namespace cmpc
