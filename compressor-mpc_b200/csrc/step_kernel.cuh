// The batched control step as three back-to-back launches on one stream (no host round trip),
// each shaped after the parallelism its part of the step offers:
//   lin_kernel       observer a-posteriori + plant linearisation: scalar code, 4 threads per
//                    (scenario, sub-controller), tens of thousands of threads in flight
//   assemble_kernel  discretisation, prediction and QP assembly: small dense matrix algebra on the
//                    FP64 tensor cores, one 64-thread group per (scenario, sub-controller) -- a CTA of
//                    its own at p = 100, two groups per CTA otherwise
//   solve_kernel     Jacobi sweeps of the QP solves, first move, a-priori observer update: one
//                    lane pair per scenario (lane = sub-controller), the QP in registers
// In the closed loop (plant_kernels.cuh) the plant kernel that follows also does lin_kernel's work
// for the next record, and all kernels are chained by programmatic dependent launch.
//
// Reference path replaced: NerveCenter::GetNextInputWithTiming
// (include/nerve_center.h:134-182) -> DistributedController::GenerateInitialQP
// (libs/distributed_controller.cc:72-108) -> Observer::ObserveAPosteriori
// (libs/observer.cc:24-40), AugmentedLinearizedSystem::Update / DiscretizeRK4 /
// GeneratePrediction (libs/aug_lin_sys.cc:145-177,232-255,260-334),
// DistributedSolver::GenerateDistributedQP (include/distributed_solver.h:83-94 ->
// libs/mpc_qp_solver.cc:19-40), then n_iterations Jacobi sweeps of
// DistributedController::GetInput (include/distributed_controller.h:206-226,
// distributed_solver.h:98-121, mpc_qp_solver.cc:45-75) and UpdateU / ObserveAPriori
// (distributed_controller.h:146-152, observer.cc:6-19).
//
// The prediction matrices Su/Sx/Sf/Su_other are never materialised.  Everything
// the QP needs follows from the impulse-response table
//     E[k][y][c] = C~ Ad^k [Bd | fd],   k = 0..p-1
// (C~ = controlled rows of C, Bd columns in this controller's input order):
//     G_k = E_k for undelayed inputs, E_{k-40} for delayed ones  (C~ A_aug^k B_aug)
//     Su[r] = [G_r | sum_{k<r} G_k],  Sf[r] fd = sum_{k<=r} E_k[fd],
//     Sx[r] x_aug = d + sum_t E_{r-t}[delayed] q[t]   (q = delay-line contents)
// and E itself is built as a product L R of baby steps L_a = C~ Ad^a (a < 8) and giant
// steps R_b = Ad^(8b) [Bd | fd], both obtained by repeated squaring/doubling, so the
// sequential depth is ~log2(p) small matrix products instead of p.  The table is stored as its
// running sum inside each block of 8 rows (left factor: the running sum of the baby steps); the
// prefix sums over the whole horizon are that plus a per-block prefix, E_k a difference of neighbours.
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "plant_dev.cuh"
#include "qp_dev.cuh"
#include "qp_thread.cuh"

namespace cmpc {

constexpr int kDelay = 40;      // Delays = {0,40,0,40} (parallel/serial_compressors_constants.h)
constexpr int kNDist = 4;       // n_disturbance_states
constexpr int kNAug = kNDist + 2 * kDelay;  // 84
constexpr int kBaby = 8;        // baby steps a = 0..7
constexpr int kNC = 6;          // columns of the giant-step blocks: [Bd (4) | fd | X40]
constexpr int kNS = 5;          // of which need prefix sums over the horizon: [Bd | fd]
constexpr int kLD = 12;         // row stride of the N x N matrices (N <= 12), conflict-free for DMMA loads
constexpr int kNNP = kLD * kLD; // storage of one N x N matrix
constexpr int kLDV = 20;        // row stride of V (N x 16)
constexpr int kCtrlStateStride = 128;  // doubles per (scenario, controller) in global memory
constexpr int kScenStateStride = 16;   // doubles per scenario
// hand-over record of one (scenario, controller): continuous A (12 x 12, zero padded),
// [Bc | fc] in the controller's input order (12 x 12, zero padded), C (4 x N), [Bd | fd] (N x 6)
constexpr int kWorkStride = 512, kWAc = 0, kWXc = 144, kWCc = 288, kWBF = 336;
// kWBF holds what the a-priori observer update needs: base[N] (free response of the state part),
// then the Bd columns of the undelayed inputs 0 and 2 (N each)
// kWPlan: hand-over between the launches of a solve that is split for the timing window
// (n-timing-iterations): the plan (up to 8), the working set, the status of the last sweep
constexpr int kWPlan = 384, kWPlanSet = 392;
constexpr int kRing = kDelay - 1;   // slots of one delay ring (the head is kept separately)
constexpr int kMaxStageTilesLadder = 18;  // squaring (2) + up to 16 blocks x 6 columns / 8 per giant-step stage
constexpr int kMaxPow = 8;           // stages of the power ladder: horizons up to 8 * 2^5 = 256
constexpr int kMaxStageTiles = 20;  // E tiles one warp may keep in registers (aliased E)
// The 2x horizon of the sweep (p = 200) runs with 3 CTAs per SM instead of 4: that leaves 168 registers
// per thread, enough to keep all 27 (36 for four outputs) table tiles of a warp in registers, so the
// table can overwrite its own inputs there as well (65 KB of shared memory per CTA instead of 100 KB,
// which allowed only 2 CTAs per SM and spilled at 128 registers).
#ifdef CMPC_P200_LEGACY   // the round-1 layout, kept for A/B measurements
__host__ __device__ constexpr int stage_tiles_for(int) { return kMaxStageTiles; }
__host__ __device__ constexpr int assemble_min_blocks(int) { return 4; }
#else
__host__ __device__ constexpr int stage_tiles_for(int p) { return p == 200 ? 36 : kMaxStageTiles; }
#ifndef CMPC_ASM_MINB100
#define CMPC_ASM_MINB100 4
#endif
__host__ __device__ constexpr int assemble_min_blocks(int pct) { return pct == 200 ? 3 : pct == 100 ? CMPC_ASM_MINB100 : 4; }
#endif

// assemble_kernel runs one CTA per scenario with a 64-thread group per sub-controller, or one 64-thread CTA
// per (scenario, sub-controller): the groups never meet inside the kernel.  A 64-thread CTA hands back its
// registers and shared memory as soon as ITS sub-controller is done, which is worth 2 % at p = 100
// (103.2 -> 101.2 us per launch of 4096 scenarios); at p = 200 the 128-thread form is 1.3 % ahead
// (2450 vs 2484 us), so the split is tied to the horizon.  -DCMPC_NO_SPLIT_CTA builds the 128-thread form
// everywhere (A/B builds).
template <class S>
__host__ __device__ constexpr int assemble_ctas_per_scenario(int pct) {
#ifdef CMPC_NO_SPLIT_CTA
  (void)pct;
  return 1;
#else
  return pct == 100 ? S::NCTRL : 1;
#endif
}
template <class S>
__host__ __device__ constexpr int assemble_block_threads(int pct) { return S::NCTRL * S::TPC / assemble_ctas_per_scenario<S>(pct); }

// offsets inside one controller's global state record
constexpr int kOffXhat = 0, kOffDx = 16, kOffYold = 112, kOffUold = 116;

template <int PLANT_, int NY_, int NU_, int NCTRL_>
struct Shape {
  static constexpr int PLANT = PLANT_, NY = NY_, NU = NU_, NCTRL = NCTRL_;
  static constexpr int N = PlantDims<PLANT>::N, NIN = PlantDims<PLANT>::NIN;
  static constexpr int NO = 4 - NU, NV = 2 * NU, NVO = 2 * NO;
  static constexpr int NOBS = N + kNDist, NTOT = N + kNAug;
  static constexpr int NH = NV * (NV + 1) / 2;     // upper triangle of H
  static constexpr int NACC = NH + NV * NVO + NV;  // H | Gx | f
  static constexpr int NCH = NY * kNC;             // doubles per row of the impulse-response table
  static constexpr int NSC = NY * kNS;             // scan channels
  static constexpr int TPC = 64;                   // threads per controller group
  static constexpr int WPC = TPC / 32;             // warps per controller group
};

struct CtrlParams {
  int out_idx[4];     // ControlledOutputIndices
  int ctrl_idx[4];    // ControlInputIndices (local -> system control input)
  double Q[16];       // ywt, NY x NY row-major (symmetric)
  double R[16];       // uwt sub-matrix, NU x NU row-major
  double lower[4], upper[4], rate_lower[4], rate_upper[4];
  double M[15 * 4];   // observer gain, NOBS x 4 row-major
};

struct StepParams {
  int p, b_max, b_full, n_pow, n_iter, batch, ldr;   // b_full: full_blocks(p, b_max)
  int ring_pos;   // position of the oldest entry in the 39-slot delay rings (same for all scenarios)
  int obs_states_free;   // 1: no observer gain row of a plant state is non-zero (the reference's M = [0; I]),
                         //    so the a-posteriori state estimate, and with it the linearisation point, does
                         //    not depend on the new measurement
  double Ts;
  double rk[4];   // Ts, Ts^2/2, Ts^3/6, Ts^4/24: the RK4 polynomial of DiscretizeRK4, evaluated once on the host
  const double* yref;   // [NCTRL][NY][p]: the lanes of a warp own consecutive rows, so this is unit stride for them
  CtrlParams c[2];
};

// Global (HBM) arrays of one handle.
struct DeviceState {
  double* ctrl;        // [B][NCTRL][kCtrlStateStride]
  unsigned* guess;     // [B][NCTRL]
  double* scen;        // [B][kScenStateStride]: u_old (4, system order), du_old (8)
  double* u_offset;    // [B][NIN]
  double* work;        // [B][NCTRL][kWorkStride] hand-over between the three kernels
  // results / parity hooks of the last step
  double* qpH;         // [B][NCTRL][NV*NV]
  double* qpf;         // [B][NCTRL][NV]
  double* qpG;         // [B][NCTRL][NV*NVO]
  double* lin;         // [B][NCTRL][N*N + N*5]  (Ad | [Bd fd])   (capture only)
  double* etab;        // [B][NCTRL][p*NY*5]                       (capture only)
  int* status;         // [B][NCTRL]
  unsigned* active;    // [B][NCTRL]
  double* objective;   // [B][NCTRL]
  long long* ticks;    // [B][32] per-phase clock64() stamps (CMPC_PHASE_TIMING builds only)
};

// Programmatic dependent launch: the kernels of a step are launched with programmatic stream
// serialisation, so the next kernel's CTAs may become resident while this grid drains.  pdl_wait()
// (first statement of every kernel) blocks until the preceding grid has completed and its writes are
// visible; pdl_trigger() tells the scheduler that this CTA no longer minds company.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Barrier of one controller group.  The id sits in a register, so ptxas reserves all 16 named barriers for
// the CTA ("used 16 barriers"), and an SM has 64: at most FOUR such CTAs per SM, whatever their size.  The
// 128-thread form of assemble_kernel is at 3 or 4 CTAs per SM anyway; the 64-thread form (eight per SM) must
// use a barrier with an immediate id (cta_sync).
__device__ __forceinline__ void group_sync(int g, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

// Columns of the giant-step matrix R.  Table rows k >= p - 40 are never read through a delayed
// input column (those read 40 rows back) and rows k >= p - 39 never through the X40 column, so the
// giant-step blocks b >= b_full keep only the three columns that are: cc = 0, 2 (undelayed inputs)
// and 4 (f_d).  Block b < b_full: columns 6 b + cc; block b >= b_full: 6 b_full + 3 (b - b_full) + cc/2.
__host__ __device__ constexpr int full_blocks(int p, int b_max) {
  int f = (p - (kDelay - 1) + kBaby - 1) / kBaby;
  if (f < 1) f = 1;   // block 0 also feeds the short convolution of the first 39 rows
  return f > b_max ? b_max : f;
}
__host__ __device__ constexpr int giant_cols(int b_max, int b_full) { return kNC * b_full + 3 * (b_max - b_full); }
__host__ __device__ constexpr int giant_col(int b, int cc, int b_full) {
  return b < b_full ? kNC * b + cc : kNC * b_full + 3 * (b - b_full) + (cc >> 1);
}
// Row stride of R: >= the column count and = 4 or 12 mod 16 so that the DMMA B-fragment loads
// (4 rows x 4 columns per half warp) hit 16 banks.
__host__ __device__ constexpr int giant_stride(int cols) {
  int ld = cols;
  while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
  return ld;
}

// Stages of the power ladder for b_max giant steps: 3 baby doublings + ceil(log2(b_max)).
__host__ __device__ constexpr int ladder_stages(int b_max) {
  int lg = 0;
  while ((1 << lg) < b_max) ++lg;
  return 3 + lg;
}

// Shared-memory footprint of one controller group, in doubles.  The big region is used three
// times: RK4 scratch -> powers Ad^(2^j) + L + R + V -> impulse-response table E -> reduction buffer.
template <class S>
struct SmemLayout {
  int yv = 0, dxd = 0, q = 0, Cc = 0, BF = 0, carry = 0, U = 0, cz = 0, mbar = 0, bp = 0, ldBP = 0, lt = 0, region = 0,
      L = 0, R = 0, V = 0, lr_end = 0, E = 0, ldE = 0, total = 0;
  bool e_alias = false;
  __host__ __device__ static constexpr int take(int& o, int n) {
    const int r = o;
    o += (n + 1) & ~1;
    return r;
  }
  // constexpr: with a compile-time horizon the kernel gets every offset as a constant (evaluated at run time
  // by every thread, the strides' search loops were 3 % of the kernel's samples)
  __host__ __device__ constexpr SmemLayout(int p, int b_max, int n_pow, int max_stage_tiles) {
    int o = 0;
    yv = take(o, 4);
    dxd = take(o, 4);
    q = take(o, 2 * kDelay);
    Cc = take(o, 4 * S::N);
    BF = take(o, S::N * kNC);
    carry = take(o, S::WPC * S::NSC);
    U = take(o, 6 * kLD);
    cz = take(o, kDelay * S::NY);
    mbar = take(o, 2);   // transaction barrier of the bulk-copy staging
    ldBP = b_max;
    bp = take(o, S::NSC * ldBP);          // block prefixes of the scan channels (prefix-table form of phase 6)
    region = o;
    (void)n_pow;
    // RK4: Ac, A2, A3, Acom, Xc; then the powers Ad^(2^j) alternate between the A2 and A3 slots and the running
    // sums of the baby steps (8 NY rows) take the Acom and Xc slots, plus a sixth one for four outputs
    const int n_scr = (kBaby * S::NY * kLD > 2 * kNNP ? 6 : 5) * kNNP;
    // E is channel-major: E[(y kNC + c) ldE + r].  ldE = 2 mod 4 keeps the 16-byte row-pair loads of
    // phase 6 aligned and spreads the four column pairs of a DMMA output tile over all banks.
    ldE = kBaby * b_max + 2;
    const int e_size = (S::NCH * ldE + 1) & ~1;
    const int red_size = S::WPC * 48;   // exchange buffer of the final reduction
    // E always starts where the (by then dead) powers are.  It can also overwrite its own inputs
    // L, R, V when every warp can hold its output tiles in registers; for longer horizons L, R, V
    // are placed behind it.
    const int n_nt = (giant_cols(b_max, full_blocks(p, b_max)) + 7) / 8;
    e_alias = (n_nt + S::WPC - 1) / S::WPC <= max_stage_tiles / S::NY;
    E = region;
    L = region + ((e_alias || n_scr > e_size) ? n_scr : e_size);
    R = L + kBaby * S::NY * kLD;
    const int r_cols = giant_cols(b_max, full_blocks(p, b_max));
    V = R + kLD * giant_stride(r_cols);     // 12 rows: rows >= N stay zero (K padding)
    lr_end = V + kLD * kLDV + 8;           // + slack for fragment reads past the last row
    // cumulative baby steps (prefix-table form): in dead RK4 scratch when E is written after a barrier
    // (e_alias: the table only overwrites the scratch once every warp holds its fragments), else behind V
    lt = region + 3 * kNNP;   // from the Acom slot on: dead once the ladder runs
    if (!e_alias) {
      lt = lr_end;
      lr_end += kBaby * S::NY * kLD;
    }
    int end = lr_end > region + e_size ? lr_end : region + e_size;
    if (end < E + red_size) end = E + red_size;
    total = (end + 1) & ~1;
  }
};

// ---- FP64 tensor-core tiles -----------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8), one warp.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)+{0,1}].
__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Operand fragments of one 8-row (A) or 8-column (B) block for K = 12 (three k-tiles).  The
// operands live zero-padded to K = 12 in shared memory (the pads are exact zeros: planted with the
// seeds and reproduced by every product, see the load phase of assemble_kernel), so the loads need
// no predicates; rows or columns beyond the 12 only feed output rows/columns that are never stored.
// Which row of its 8-row block a lane holds (A fragment and result alike) is ours to choose, as long as
// the A loads and the use of the result agree: lanes 4i..4i+3 take row frag_row = i with its two low
// bits swapped (0 2 1 3 4 6 5 7).  A quarter warp -- the unit in which 16-byte accesses are served --
// then stores rows r and r + 2 of a result tile instead of r and r + 1, which with the row strides used
// here (12, 20, 68, ... doubles: 6, 10, 34 sixteen-byte slots, i.e. 6 or 2 modulo the 8 slots of a
// wavefront) touch 8 different slots: the tile stores lose their two-way bank conflicts (5 % of the kernel's
// shared-memory wavefronts).  The 8-byte fragment loads see the same four rows per half warp as before.
// Measured: assemble_kernel -0.4 % at p = 100, -1.4 % at p = 200 (tools/ab.sh, -DCMPC_FRAG_ROW_IDENTITY).
__device__ __forceinline__ int frag_row(int lane) {
  const int r = lane >> 2;
#ifdef CMPC_FRAG_ROW_IDENTITY   // (A/B builds)
  return r;
#else
  return (r & 4) | ((r & 1) << 1) | ((r >> 1) & 1);
#endif
}
__device__ __forceinline__ void frag_a(const double* A, int lda, int mt, int lane, double (&a)[3]) {
  const double* p = A + (8 * mt + frag_row(lane)) * lda + (lane & 3);
  a[0] = p[0];
  a[1] = p[4];
  a[2] = p[8];
}
__device__ __forceinline__ void frag_b(const double* B, int ldb, int nt, int lane, double (&b)[3]) {
  const double* p = B + (lane & 3) * ldb + 8 * nt + (lane >> 2);
  b[0] = p[0];
  b[1] = p[4 * ldb];
  b[2] = p[8 * ldb];
}
// The same for the N x N matrices (12 x 12 with their padding): the second row block (A) or column block (B) has
// four rows / columns, held by lanes 0..15 (frag_row keeps bit 2 of the row).  The other half of the warp stays
// out of the load -- one shared-memory wavefront per instruction instead of two -- and feeds zeros, which only
// reach output rows / columns that are never stored.
__device__ __forceinline__ void frag_a12(const double* A, int mt, int lane, double (&a)[3]) {
  a[0] = a[1] = a[2] = 0.0;
  if (mt == 0 || lane < 16) frag_a(A, kLD, mt, lane, a);
}
__device__ __forceinline__ void frag_b12(const double* B, int nt, int lane, double (&b)[3]) {
  b[0] = b[1] = b[2] = 0.0;
  if (nt == 0 || lane < 16) frag_b(B, kLD, nt, lane, b);
}
__device__ __forceinline__ void mma3(double (&c)[2], const double (&a)[3], const double (&b)[3]) {
  c[0] = 0.0;
  c[1] = 0.0;
  dmma_884(c, a[0], b[0]);
  dmma_884(c, a[1], b[1]);
  dmma_884(c, a[2], b[2]);
}
// NT independent tiles sharing the A fragments (one row block times NT column blocks): the DMMAs
// of different tiles are interleaved so that no instruction waits for the previous one.
// Tiles [LO, HI) are computed, as straight-line code: a run-time tile count must be turned into
// one of these by a switch (mma3_shared_a / _b below).  A predicate per DMMA would let the compiler
// speculate the masked tiles, and every DMMA occupies the FP64 pipe of its SM sub-partition for
// ~16 cycles whether its result is used or not.
template <int NT, int LO, int HI>
__device__ __forceinline__ void mma3_shared_a_range(double (&c)[NT][2], const double (&a)[3], const double (&b)[NT][3]) {
#pragma unroll
  for (int i = LO; i < HI; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int kt = 0; kt < 3; ++kt)
#pragma unroll
    for (int i = LO; i < HI; ++i) dmma_884(c[i], a[kt], b[i][kt]);
}
template <int NT>
__device__ __forceinline__ void mma3_shared_a(double (&c)[NT][2], const double (&a)[3], const double (&b)[NT][3], int n_on = NT) {
  static_assert(NT <= 6, "tile count");
  switch (n_on) {
    case 1: mma3_shared_a_range<NT, 0, 1>(c, a, b); break;
    case 2: if constexpr (NT >= 2) mma3_shared_a_range<NT, 0, 2>(c, a, b); break;
    case 3: if constexpr (NT >= 3) mma3_shared_a_range<NT, 0, 3>(c, a, b); break;
    case 4: if constexpr (NT >= 4) mma3_shared_a_range<NT, 0, 4>(c, a, b); break;
    case 5: if constexpr (NT >= 5) mma3_shared_a_range<NT, 0, 5>(c, a, b); break;
    case 6: if constexpr (NT >= 6) mma3_shared_a_range<NT, 0, 6>(c, a, b); break;
    default: break;
  }
}
// NT independent tiles sharing the B fragments (NT row blocks times one column block)
template <int NT, int HI>
__device__ __forceinline__ void mma3_shared_b_range(double (&c)[NT][2], const double (&a)[NT][3], const double (&b)[3]) {
#pragma unroll
  for (int i = 0; i < HI; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int kt = 0; kt < 3; ++kt)
#pragma unroll
    for (int i = 0; i < HI; ++i) dmma_884(c[i], a[i][kt], b[kt]);
}
template <int NT>
__device__ __forceinline__ void mma3_shared_b(double (&c)[NT][2], const double (&a)[NT][3], const double (&b)[3], int n_on = NT) {
  static_assert(NT <= 4, "tile count");
  switch (n_on) {
    case 1: mma3_shared_b_range<NT, 1>(c, a, b); break;
    case 2: if constexpr (NT >= 2) mma3_shared_b_range<NT, 2>(c, a, b); break;
    case 3: if constexpr (NT >= 3) mma3_shared_b_range<NT, 3>(c, a, b); break;
    case 4: if constexpr (NT >= 4) mma3_shared_b_range<NT, 4>(c, a, b); break;
    default: break;
  }
}

// v[i] of every lane summed over the warp, for W values at once: each of the log2(W) rounds
// sends half of the remaining values to the partner lane and keeps the other half, so lane l ends
// up with the warp total of value l mod W (W - 1 shuffles instead of 5 W).
template <int W>
__device__ __forceinline__ double warp_transpose_reduce(double (&v)[W], int lane) {
  static_assert(W == 32 || W == 16, "value count");
#pragma unroll
  for (int h = W / 2; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double keep = up ? v[i + h] : v[i];
      const double send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  double r = v[0];
  if (W == 16) r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

// Store a tile into a row-major matrix C (even stride, even column offset) as one 16-byte store
// per lane.  Rows >= mc and column pairs starting at >= nc_pad are dropped (nc_pad even; a pad
// column inside the pair receives an exact zero because the B operand's pad column is zero).
template <bool ADD_EYE = false>
__device__ __forceinline__ void tile_store(double* C, int ldc, int row_off, int col_off, int mc, int nc_pad,
                                           int mt, int nt, int lane, const double (&c)[2], int n_eye = 0) {
  const int r = 8 * mt + frag_row(lane), cc = 8 * nt + 2 * (lane & 3);
  if (r < mc && cc < nc_pad) {
    double2 v;
    v.x = c[0];
    v.y = c[1];
    if (ADD_EYE) {   // + I_(n_eye)
      if (r == cc && r < n_eye) v.x += 1.0;
      if (r == cc + 1 && r < n_eye) v.y += 1.0;
    }
    *reinterpret_cast<double2*>(C + (row_off + r) * ldc + col_off + cc) = v;
  }
}

// C[r] . dx[0..N) for plant output row r at state x (the non-zeros of C, compressor.cc:169-170).
template <int PLANT>
__device__ __forceinline__ double plant_c_row_dot(const double* x, int r, const double* v) {
  if (PLANT == 0) {
    if (r < 2) {
      const double* xc = x + 5 * r;
      const double* vc = v + 5 * r;
      return 100 * xc[1] / (kSDc0 * xc[0] * xc[0]) * vc[0] - 100. / (kSDc0 * xc[0]) * vc[1] + 100 * vc[2];
    }
    return r == 2 ? v[1] - v[6] : v[10];
  } else {
    const int c = r >> 1;
    const double* xc = x + 5 * c;
    const double* vc = v + 5 * c;
    if ((r & 1) == 0) return vc[1];
    return 100 * xc[1] / (kSDc0 * xc[0] * xc[0]) * vc[0] - 100. / (kSDc0 * xc[0]) * vc[1] + 100 * vc[2];
  }
}

// One control step for the scenario owned by this CTA.  y4: the new measurement (4 doubles).
// u_out: 4 doubles.  All threads of the CTA must call it.
#ifndef CMPC_MIN_BLOCKS
#define CMPC_MIN_BLOCKS 4
#endif

// ---- K0: Observer::ObserveAPosteriori (observer.cc:24-40) with the C of the previous
// linearisation (same x_hat), x_ += dx (distributed_controller.cc:80), then the plant is
// linearised at (x_hat, u_full_old) (aug_lin_sys.cc:147).  Four threads per (scenario,
// controller): each repeats the tiny observer update in registers, three of them fill one part
// of the continuous-time matrices in the hand-over record (whose zero pattern never changes).
template <class S>
__device__ __forceinline__ void lin_part(const StepParams& P, const DeviceState& G, int scen, int g, int part,
                                         const double (&yv)[4], unsigned sync_mask) {
  constexpr int N = S::N, NOBS = S::NOBS, NIN = S::NIN;
  double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  const double* ss = G.scen + size_t(scen) * kScenStateStride;
  double* wk = G.work + (size_t(scen) * S::NCTRL + g) * kWorkStride;
  double xh[N], dx[NOBS], ev[4];
#pragma unroll
  for (int i = 0; i < N; ++i) xh[i] = gs[kOffXhat + i];
#pragma unroll
  for (int i = 0; i < NOBS; ++i) dx[i] = gs[kOffDx + i];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    ev[r] = yv[r] - gs[kOffYold + r] - (plant_c_row_dot<S::PLANT>(xh, r, dx) + dx[N + r]);
#pragma unroll
  for (int i = 0; i < NOBS; ++i) {
    double acc = dx[i];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc = fma(P.c[g].M[i * 4 + r], ev[r], acc);
    dx[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) xh[i] += dx[i];
  if (sync_mask) __syncwarp(sync_mask);   // every part has read the old observer state before part 3 replaces it
                                          // (0: the caller orders them with a block barrier)
  if (part == 3) {
#pragma unroll
    for (int i = 0; i < N; ++i) gs[kOffXhat + i] = xh[i];
#pragma unroll
    for (int i = 0; i < NOBS; ++i) gs[kOffDx + i] = dx[i];
#pragma unroll
    for (int r = 0; r < 4; ++r) gs[kOffYold + r] = yv[r];
    return;
  }
  // u_full_old = GetPlantInput(u_old_, u_offset_)  (nerve_center.h:140)
  double uf[NIN];
#pragma unroll
  for (int i = 0; i < NIN; ++i) uf[i] = G.u_offset[size_t(scen) * NIN + i];
  uf[0] += ss[0]; uf[3] += ss[1]; uf[4] += ss[2]; uf[7] += ss[3];
  int inv[4];   // system control input -> this controller's local input (aug_lin_sys.cc:156-173)
#pragma unroll
  for (int c = 0; c < 4; ++c) inv[P.c[g].ctrl_idx[c]] = c;
  plant_linearize_part_x<S::PLANT>(part, xh, uf, wk + kWAc, kLD, wk + kWXc, kLD, inv, wk + kWCc);
}

// The same for the closed loop when StepParams::obs_states_free holds: the gain rows of the plant
// states are zero, so x_hat+ = x_hat + dx[0..N) whatever the new measurement is, and the
// linearisation can run while the plant is still being integrated.  lin_part_early fills part
// `part` (0..2) of the hand-over record; the part-0 lane also returns what the observer update
// still owes (ObsPending), and lin_finish completes and stores it once the measurement exists.
// Every value is computed by the same expressions, in the same order, as in lin_part.
template <class S>
struct ObsPending {
  double xh[S::N], dxs[S::N], dxn[4], yold[4], s[4];
};

template <class S>
__device__ __forceinline__ void lin_part_early(const StepParams& P, const DeviceState& G, int scen, int g, int part,
                                               ObsPending<S>& pend, bool apriori) {
  constexpr int N = S::N, NOBS = S::NOBS, NIN = S::NIN, NU = S::NU;
  double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  const double* ss = G.scen + size_t(scen) * kScenStateStride;
  double* wk = G.work + (size_t(scen) * S::NCTRL + g) * kWorkStride;
  double xh[N], dx[NOBS];
#pragma unroll
  for (int i = 0; i < N; ++i) xh[i] = gs[kOffXhat + i];
#pragma unroll
  for (int i = N; i < NOBS; ++i) dx[i] = gs[kOffDx + i];
  if (apriori) {
    // UpdateU / ObserveAPriori of the record just solved (distributed_controller.h:146-152,
    // observer.cc:6-19), left to this kernel by solve_kernel<S, false>: same expressions as
    // apriori_update.  Every part forms the predicted state part it linearises around; part 0 also
    // moves the delay lines (ring_pos has already been advanced by the host for the next record).
    const double* wb = wk + kWBF;
    double du[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) du[k] = (k < NU) ? ss[4 + g * 4 + k] : 0.0;   // first move of the plan (du_old_)
#pragma unroll
    for (int i = 0; i < N; ++i) dx[i] = fma(wb[N + i], du[0], fma(wb[2 * N + i], du[2], wb[i]));
    if (part == 0) {
      double uold[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) uold[i] = gs[kOffUold + i];
      const int ring0 = kOffDx + NOBS + 2 + (P.ring_pos + kRing - 1) % kRing, ring1 = ring0 + kRing;
      const double old0 = gs[ring0], old1 = gs[ring1];
      gs[kOffDx + NOBS + 0] = old0;
      gs[kOffDx + NOBS + 1] = old1;
      gs[ring0] = uold[1] + du[1];
      gs[ring1] = uold[3] + du[3];
#pragma unroll
      for (int i = 0; i < 4; ++i) gs[kOffUold + i] = uold[i] + du[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dx[i] = gs[kOffDx + i];
  }
  if (part == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      pend.s[r] = plant_c_row_dot<S::PLANT>(xh, r, dx) + dx[N + r];
      pend.yold[r] = gs[kOffYold + r];
      pend.dxn[r] = dx[N + r];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    xh[i] += dx[i];
    if (part == 0) {
      pend.xh[i] = xh[i];
      pend.dxs[i] = dx[i];
    }
  }
  double uf[NIN];
#pragma unroll
  for (int i = 0; i < NIN; ++i) uf[i] = G.u_offset[size_t(scen) * NIN + i];
  uf[0] += ss[0]; uf[3] += ss[1]; uf[4] += ss[2]; uf[7] += ss[3];
  int inv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) inv[P.c[g].ctrl_idx[c]] = c;
  plant_linearize_part_x<S::PLANT>(part, xh, uf, wk + kWAc, kLD, wk + kWXc, kLD, inv, wk + kWCc);
}

template <class S>
__device__ __forceinline__ void lin_finish(const StepParams& P, const DeviceState& G, int scen, int g,
                                           const ObsPending<S>& pend, const double (&yv)[4]) {
  constexpr int N = S::N;
  double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  double ev[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) ev[r] = yv[r] - pend.yold[r] - pend.s[r];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    gs[kOffXhat + i] = pend.xh[i];
    gs[kOffDx + i] = pend.dxs[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = pend.dxn[i];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc = fma(P.c[g].M[(N + i) * 4 + r], ev[r], acc);
    gs[kOffDx + N + i] = acc;
    gs[kOffYold + i] = yv[i];
  }
}

template <class S>
__global__ void __launch_bounds__(128)
lin_kernel(StepParams P, DeviceState G, double* __restrict__ y, const double* __restrict__ y_from) {
  pdl_wait();
  pdl_trigger();   // single wave: the next grid may queue up behind it at once
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int part = tid & 3, quad = tid >> 2;
  const int scen = quad / S::NCTRL, g = quad % S::NCTRL;
  const bool on = scen < P.batch;
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (y_from) {
    // cmpc_get_next_input with a page-locked measurement buffer: the block brings its scenarios' rows in
    // itself, one contiguous chunk of mapped host memory read once, and leaves them where the assemble
    // kernel expects them (instead of a copy on the stream in front of the step)
    constexpr int kScenPerBlock = 32 / S::NCTRL;
    const int s0 = blockIdx.x * kScenPerBlock;
    const int n_here = P.batch - s0 < kScenPerBlock ? P.batch - s0 : kScenPerBlock;
    if (int(threadIdx.x) < n_here * 4) y[size_t(s0) * 4 + threadIdx.x] = y_from[size_t(s0) * 4 + threadIdx.x];
    __syncthreads();
  }
  if (!on) return;
  double yv[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) yv[r] = y[size_t(scen) * 4 + r];
  lin_part<S>(P, G, scen, g, part, yv, m);
}

#ifdef CMPC_PHASE_TIMING
#define CMPC_TICK(i) do { if (threadIdx.x == 0 && g == 0) { G.ticks[size_t(scen) * 32 + (i)] = clock64() - tick_t0_; } } while (0)
#else
#define CMPC_TICK(i) do { } while (0)
#endif
// the same for the one-wave kernels: slot i of scenario `sc`, relative to t0 (stamped by one lane per scenario)
#ifdef CMPC_PHASE_TIMING
#define CMPC_TICK_AT(sc, i, t0) do { G.ticks[size_t(sc) * 32 + (i)] = clock64() - (t0); } while (0)
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define CMPC_GTIME_AT(sc, i) do { G.ticks[size_t(sc) * 32 + (i)] = gtime_ns(); } while (0)
#else
#define CMPC_TICK_AT(sc, i, t0) do { } while (0)
#define CMPC_GTIME_AT(sc, i) do { } while (0)
#endif

// ---- K1: discretisation, prediction and QP assembly of one scenario per CTA --------------------
// PCT > 0 fixes the prediction horizon at compile time (the reference's p = 100 and the 2x sweep):
// every tile count, stride and loop bound below then folds to a constant.  PCT = 0 reads them
// from the parameters.
template <class S, int RPT, int PCT>
__global__ void __launch_bounds__(assemble_block_threads<S>(PCT), assemble_min_blocks(PCT) * assemble_ctas_per_scenario<S>(PCT))
assemble_kernel(StepParams P, DeviceState G, const double* __restrict__ y) {
  extern __shared__ __align__(16) double smem[];
  constexpr int kCtasPerScen = assemble_ctas_per_scenario<S>(PCT);
#ifdef CMPC_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x % kCtasPerScen == 0 && blockIdx.x / kCtasPerScen < P.batch) CMPC_GTIME_AT(blockIdx.x / kCtasPerScen, 16);
#endif
  pdl_wait();
#ifdef CMPC_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x % kCtasPerScen == 0 && blockIdx.x / kCtasPerScen < P.batch) CMPC_GTIME_AT(blockIdx.x / kCtasPerScen, 17);
#endif
  constexpr int N = S::N, NY = S::NY, NU = S::NU, NV = S::NV, NVO = S::NVO, NO = S::NO;
  constexpr int TPC = S::TPC, WPC = S::WPC, NOBS = S::NOBS, NSC = S::NSC, NH = S::NH;
#ifdef CMPC_PHASE_TIMING
  const long long tick_t0_ = clock64();
#endif
  const int scen = blockIdx.x / kCtasPerScen;
  if (scen >= P.batch) return;
  // one CTA per scenario with a thread group per sub-controller, or (kCtasPerScen = NCTRL) one CTA per
  // (scenario, sub-controller): the groups never meet, so either works
  const int g = kCtasPerScen > 1 ? int(blockIdx.x % kCtasPerScen) : int(threadIdx.x / TPC), t = threadIdx.x % TPC;
  const int lane = t & 31, warp = t >> 5;
#ifdef CMPC_POISON_SMEM
  // test builds: start from NaN-filled shared memory, so that any read of a word this launch did
  // not write shows up in the parity tests
  {
    unsigned dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    for (unsigned i = threadIdx.x; i < dyn_bytes / 8; i += blockDim.x) smem[i] = __longlong_as_double(0x7ff8dead0000beefLL);
    __syncthreads();
  }
#endif
  constexpr bool CT = PCT > 0;
  constexpr int kBmaxCt = (PCT + kBaby - 1) / kBaby;
  const int p = CT ? PCT : P.p, b_max = CT ? kBmaxCt : P.b_max;
  constexpr int kBfullCt = full_blocks(PCT > 0 ? PCT : 1, kBmaxCt > 0 ? kBmaxCt : 1);
  const int b_full = CT ? kBfullCt : P.b_full;
  const int r_cols_all = giant_cols(b_max, b_full);   // columns of R
  const int ldr = CT ? giant_stride(giant_cols(kBmaxCt, kBfullCt)) : P.ldr, n_pow = CT ? ladder_stages(kBmaxCt) : P.n_pow;
  constexpr int kStageTiles = stage_tiles_for(PCT);
  constexpr SmemLayout<S> lay_ct(CT ? PCT : 1, CT ? kBmaxCt : 1, CT ? ladder_stages(kBmaxCt > 0 ? kBmaxCt : 1) : 3, kStageTiles);
  const SmemLayout<S> lay = CT ? lay_ct : SmemLayout<S>(p, b_max, n_pow, kStageTiles);
  const int ldE = lay.ldE;
  double* sm = smem + (kCtasPerScen > 1 ? 0 : g * lay.total);
  auto gsync = [&]() {
    if constexpr (kCtasPerScen > 1) cta_sync();
    else group_sync(g, TPC);
  };
  const double* gs = G.ctrl + (size_t(scen) * S::NCTRL + g) * kCtrlStateStride;
  double* wk = G.work + (size_t(scen) * S::NCTRL + g) * kWorkStride;

  double* yv = sm + lay.yv; double* dxd = sm + lay.dxd; double* q = sm + lay.q; double* Cc = sm + lay.Cc;
  double* BF = sm + lay.BF; double* scr = sm + lay.region; double* U = sm + lay.U;
  double* L = sm + lay.L; double* R = sm + lay.R; double* V = sm + lay.V; double* E = sm + lay.E;
  double* carry = sm + lay.carry; double* CZ = sm + lay.cz;
  double* BP = sm + lay.bp; double* LT = sm + lay.lt;
  const int ldBP = lay.ldBP;
#ifdef CMPC_SCAN_BY_SHUFFLE   // (A/B builds) prefix sums over the horizon by a register / shuffle scan
  constexpr bool kPT = false;
#else
  constexpr bool kPT = true;
#endif
  double* Ac = scr;                 // continuous A (stride kLD)
  double* A2 = scr + kNNP;
  double* A3 = scr + 2 * kNNP;
  double* Acom = scr + 3 * kNNP;
  double* Xc = scr + 4 * kNNP;      // [Bc (local input order) | fc], N x 5, stride kLD

  // ---- load: linearisation from K0, delay-line contents, measurement -----------------------
  // all global loads are issued first so that they are in flight together
  constexpr int NLD = (kNNP + TPC - 1) / TPC;
  double a_v[NLD], x_v[NLD], q_v[2], cc_v = 0.0, yd_v = 0.0;
#ifndef CMPC_NO_TMA_STAGING
  // The three dense pieces of the hand-over record (A, [B|f], C: 2.6 KB) come in as bulk asynchronous
  // copies (TMA, cp.async.bulk = UBLKCP in SASS) that complete on a transaction barrier, instead of
  // ld.global -> registers -> st.shared.  Measured neutral at p = 100 (103.0 vs 103.1 us per launch)
  // and 1 % faster at p = 200 (2484 vs 2508 us); CMPC_NO_TMA_STAGING builds the register path.
  const unsigned mbar_s = unsigned(__cvta_generic_to_shared(sm + lay.mbar));
  if (t == 0) {
    // barrier init, expected bytes and the copies by the same thread, in program order: the other threads meet
    // the initialised barrier behind the group barrier further down, when their own loads are on their way
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr unsigned kBytesA = kNNP * 8, kBytesC = 4 * N * 8;
    static_assert(kBytesA % 16 == 0 && kBytesC % 16 == 0, "bulk copies move multiples of 16 bytes");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(2 * kBytesA + kBytesC) : "memory");
    auto bulk = [&](double* dst, const double* src, unsigned bytes) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src), "r"(bytes), "r"(mbar_s) : "memory");
    };
    bulk(Ac, wk + kWAc, kBytesA);
    bulk(Xc, wk + kWXc, kBytesA);
    bulk(Cc, wk + kWCc, kBytesC);
  }
#else
#pragma unroll
  for (int k = 0; k < NLD; ++k) {
    const int i = t + k * TPC;
    a_v[k] = (i < kNNP) ? wk[kWAc + i] : 0.0;
    x_v[k] = (i < kNNP) ? wk[kWXc + i] : 0.0;
  }
#endif
  // delay-line contents relative to u_old (AdjustAllDelayedStates, aug_lin_sys.h:141-154)
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = t + k * TPC;
    q_v[k] = 0.0;
    if (i < 2 * kDelay) {
      const int d = i / kDelay, tt = i % kDelay;
      // head, then the ring in logical order (oldest first)
      const int slot = (tt == 0) ? (NOBS + d) : (NOBS + 2 + d * kRing + (P.ring_pos + tt - 1) % kRing);
      q_v[k] = gs[kOffDx + slot] - gs[kOffUold + 1 + 2 * d];
    }
  }
#ifdef CMPC_NO_TMA_STAGING
  if (t < 4 * N) cc_v = wk[kWCc + t];
#endif
  if (t < 4) yd_v = y[size_t(scen) * 4 + t];
  else if (t < 8) yd_v = gs[kOffDx + N + t - 4];
  // Zero padding without clearing the region: the hand-over record arrives zero-padded, every
  // product below stores all kLD rows/columns of its result (a pad row of A or pad column of B gives
  // an exact zero), and the pads of the three seeds (L, R, V) are cleared where they are planted.
#ifdef CMPC_NO_TMA_STAGING
#pragma unroll
  for (int k = 0; k < NLD; ++k) {
    const int i = t + k * TPC;
    if (i < kNNP) {
      Ac[i] = a_v[k];
      Xc[i] = x_v[k];
    }
  }
  if (t < 4 * N) Cc[t] = cc_v;
#else
  (void)a_v; (void)x_v; (void)cc_v;
  gsync();
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CMPC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra CMPC_DONE;\n"
      "bra CMPC_WAIT;\n"
      "CMPC_DONE:\n"
      "}\n" ::"r"(mbar_s) : "memory");
#endif
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (t + k * TPC < 2 * kDelay) q[t + k * TPC] = q_v[k];
  if (t < 4) yv[t] = yd_v;
  else if (t < 8) dxd[t - 4] = yd_v;
  gsync();

  CMPC_TICK(0);
  // ---- phase 3: DiscretizeRK4 (aug_lin_sys.cc:232-255) on the FP64 tensor cores -------------
  // every N x N product is 2 x 2 output tiles of 8 x 8 with K = 12 (3 DMMA per tile); warp w of
  // the group owns the row block mt = w, so its A fragments are loaded once per product.
  const int mt_w = warp & 1;
#ifdef CMPC_RK4_THREE_LEVELS   // (A/B builds) the products as the reference orders them: A^2, A^3, Acom, Acom [A | X]
  {
    double a[3], bb[3][3], cc[3][2];
    frag_a12(Ac, mt_w, lane, a);
    frag_b(Ac, kLD, 0, lane, bb[0]);
    frag_b12(Ac, 1, lane, bb[1]);
    frag_b(Xc, kLD, 0, lane, bb[2]);
    mma3_shared_a<3>(cc, a, bb, 2);
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cc[0]);
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cc[1]);
    gsync();
    frag_a12(A2, mt_w, lane, a);
    mma3_shared_a<3>(cc, a, bb, 2);
    tile_store(A3, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cc[0]);
    tile_store(A3, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cc[1]);
    gsync();
    const double k1 = P.rk[0], k2 = P.rk[1], k3 = P.rk[2], k4 = P.rk[3];
    for (int idx = t; idx < kNNP; idx += TPC) {   // pads included: they come out as exact zeros
      const int i = idx / kLD, j = idx % kLD;
      Acom[idx] = k1 * ((i == j && i < N) ? 1.0 : 0.0) + k2 * Ac[idx] + k3 * A2[idx] + k4 * A3[idx];
    }
    gsync();
    // Ad = I + Acom Ac (into the A2 slot), [Bd | fd] = Acom Xc
    frag_a12(Acom, mt_w, lane, a);
    mma3_shared_a<3>(cc, a, bb, 3);
    tile_store<true>(A2, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cc[0], N);
    tile_store<true>(A2, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cc[1], N);
    tile_store(BF, kNC, 0, 0, N, kNC, mt_w, 0, lane, cc[2]);
    gsync();
  }
#else
  // The same polynomial with two dependent product levels instead of three (Paterson-Stockmeyer):
  //   Ad = I + k1 A + A^2 (k2 I + k3 A + k4 A^2),   [Bd | fd] = k1 X + k2 A X + A^2 (k3 X + k4 A X)
  // level 1 forms A^2 and A X, level 2 multiplies A^2 by M = k2 I + k3 A + k4 A^2 and Y = k3 X + k4 A X; the
  // lower-order terms ride in registers from one epilogue to the next (same lane, same tile position).
  {
    double a[3], bb[3][3], cc[3][2];
    const double k1 = P.rk[0], k2 = P.rk[1], k3 = P.rk[2], k4 = P.rk[3];
    double* Mb = A3;     // M (stride kLD)
    double* Yb = Acom;   // Y (stride kLD)
    frag_a12(Ac, mt_w, lane, a);
    frag_b(Ac, kLD, 0, lane, bb[0]);
    frag_b12(Ac, 1, lane, bb[1]);
    frag_b(Xc, kLD, 0, lane, bb[2]);
    mma3_shared_a<3>(cc, a, bb, 3);
    const int r = 8 * mt_w + frag_row(lane), c0 = 2 * (lane & 3);
    double eA[2][2], tB[2];   // I + k1 A at this lane's positions of the two column tiles, k1 X + k2 A X
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cc[0]);
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cc[1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = 8 * nt + c0;
      double2 av = make_double2(0.0, 0.0);
      if (r < kLD && col < kLD) av = *reinterpret_cast<const double2*>(Ac + r * kLD + col);
      const double e0 = (r == col && r < N) ? 1.0 : 0.0, e1 = (r == col + 1 && r < N) ? 1.0 : 0.0;
      eA[nt][0] = fma(k1, av.x, e0);
      eA[nt][1] = fma(k1, av.y, e1);
      double m[2];
      m[0] = fma(k4, cc[nt][0], fma(k3, av.x, k2 * e0));
      m[1] = fma(k4, cc[nt][1], fma(k3, av.y, k2 * e1));
      tile_store(Mb, kLD, 0, 0, kLD, kLD, mt_w, nt, lane, m);
    }
    {
      double2 xv = make_double2(0.0, 0.0);
      if (r < kLD) xv = *reinterpret_cast<const double2*>(Xc + r * kLD + c0);
      tB[0] = fma(k2, cc[2][0], k1 * xv.x);
      tB[1] = fma(k2, cc[2][1], k1 * xv.y);
      double yv2[2];
      yv2[0] = fma(k4, cc[2][0], k3 * xv.x);
      yv2[1] = fma(k4, cc[2][1], k3 * xv.y);
      tile_store(Yb, kLD, 0, 0, kLD, 8, mt_w, 0, lane, yv2);
    }
    gsync();
    frag_a12(A2, mt_w, lane, a);
    frag_b(Mb, kLD, 0, lane, bb[0]);
    frag_b12(Mb, 1, lane, bb[1]);
    frag_b(Yb, kLD, 0, lane, bb[2]);
    mma3_shared_a<3>(cc, a, bb, 3);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      cc[nt][0] += eA[nt][0];
      cc[nt][1] += eA[nt][1];
    }
    cc[2][0] += tB[0];
    cc[2][1] += tB[1];
    // Ad goes to the A2 slot: a warp has read its own row block of A^2 before it writes it, and the other
    // warp's row block is not touched
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cc[0]);
    tile_store(A2, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cc[1]);
    tile_store(BF, kNC, 0, 0, N, kNC, mt_w, 0, lane, cc[2]);
    gsync();
  }
#endif
  double* Pw = scr + kNNP;  // Ad lives in the A2 slot; the powers Ad^(2^j) alternate between this slot and the next
  if (G.lin) {
    double* gl = G.lin + (size_t(scen) * S::NCTRL + g) * (N * N + N * 5);
    for (int idx = t; idx < N * N + N * 5; idx += TPC)
      gl[idx] = (idx < N * N) ? Pw[(idx / N) * kLD + idx % N] : BF[((idx - N * N) / 5) * kNC + (idx - N * N) % 5];
  }

  CMPC_TICK(1);
  // ---- phase 4: powers Ad^(2^j) with baby (L), giant (R) and delay (V) steps by doubling ----
  // L_a = C~ Ad^a (a < 8), stored output-major (row 8 y + a): rows a in [2^j, 2^(j+1)) =
  //       rows a in [0, 2^j) * Ad^(2^j), j = 0..2
  // V_a = Ad^a Bd[:, delayed] (a < 8), same doubling on the left
  // R_b = Ad^(8b) [Bd fd X40]: blocks [2^j, 2^(j+1)) = Ad^(8*2^j) * blocks [0, 2^j), j = 0..
  // X40 = state reached after the 40 queued delayed inputs have been applied (free response
  // of the delay line): conv[r] = C~ Ad^(r-39) X40 for r >= 39 comes out of the same table.
  for (int idx = t; idx < NY * kLD; idx += TPC) {
    const int yy = idx / kLD, j = idx % kLD;
    L[yy * kBaby * kLD + j] = (j < N) ? Cc[P.c[g].out_idx[yy] * N + j] : 0.0;
  }
  for (int idx = t; idx < kLD * 8; idx += TPC) {   // the whole first column tile of R (column 5 = X40 comes later)
    const int i = idx >> 3, j = idx & 7;
    R[i * ldr + j] = (i < N && j < 5) ? BF[i * kNC + j] : 0.0;
  }
  for (int idx = t; idx < kLD * 2; idx += TPC)
    V[(idx >> 1) * kLDV + (idx & 1)] = ((idx >> 1) < N) ? BF[(idx >> 1) * kNC + 1 + 2 * (idx & 1)] : 0.0;
  gsync();
  // The stages are unrolled with a compile-time index: every stage but the last one of the run has
  // compile-time tile counts, so its DMMAs are straight-line code with nothing to predicate.
  auto stage = [&](auto Jc) {
    constexpr int j = decltype(Jc)::value, s = j + 1;
    if (s > n_pow + (j == 3 ? 1 : 0)) return;   // (the delay-line block of stage 3 below runs even if stage 3 does not)
    const double* Pm = Pw + (j & 1) * kNNP;   // a stage only reads Ad^(2^j) and writes Ad^(2^(j+1)): two slots take turns
    double* Pn = Pw + (s & 1) * kNNP;
    if constexpr (j == 3) {
      CMPC_TICK(9);
      if constexpr (kPT) {
        // the baby steps are complete: their running sums LT_a = sum_{a' <= a} L_a' (left factor of the table
        // product, phase 5) go to two RK4 slots nobody uses any more; the barriers of the ladder stages that
        // follow order them in front of phase 5
        for (int idx = t; idx < NY * kLD; idx += TPC) {
          const int yy = idx / kLD, jj = idx % kLD;
          double lv[kBaby];
#pragma unroll
          for (int a = 0; a < kBaby; ++a) lv[a] = L[(yy * kBaby + a) * kLD + jj];
#pragma unroll
          for (int a = 1; a < kBaby; ++a) lv[a] += lv[a - 1];
#pragma unroll
          for (int a = 0; a < kBaby; ++a) LT[(yy * kBaby + a) * kLD + jj] = lv[a];
        }
      }
      // X40 = P(P(P(P U_0 + U_1) + U_2) + U_3) + U_4 with U_b = sum_{a,d} V_(7-a)[:, d] q_d[8b + a],
      // P = Ad^8 = Pm.  Warp 0 of the group, lanes 0..N-1 hold one state each.
      // U (N x 5) = V (N x 16) Qm (16 x 5) on the tensor cores, Qm[2 va + d][b] = q_d[8 b + 7 - va]:
      // warp w owns row block w, K = 16 is four k-tiles
      {
        double cu[2] = {0.0, 0.0};
        const int n = lane >> 2;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = 4 * kk + (lane & 3);
          const int vr = 8 * mt_w + (lane >> 2);   // rows past the matrix only feed rows that are dropped
          const double av = (vr < kLD) ? V[vr * kLDV + k] : 0.0;
          const double bv = (n < 5) ? q[(k & 1) * kDelay + 8 * n + 7 - (k >> 1)] : 0.0;
          dmma_884(cu, av, bv);
        }
        const int i = 8 * mt_w + (lane >> 2), b = 2 * (lane & 3);
        if (i < kLD) {
          if (b < 5) U[b * kLD + i] = cu[0];
          if (b + 1 < 5) U[(b + 1) * kLD + i] = cu[1];
        }
      }
      gsync();
      if (warp == 0) {
        // Z_b = state at the start of delay block b (Z_0 = 0) goes to V[:, b]: phase 5 turns it
        // into C~ Ad^a Z_b for the rows that still see a partially drained delay line
        double z = (lane < N) ? U[lane] : 0.0;
        double pr[N];   // this lane's row of P, reused by every step
#pragma unroll
        for (int k = 0; k < N; ++k) pr[k] = 0.0;
        if (lane < N) {
#pragma unroll
          for (int k = 0; k < N; ++k) pr[k] = Pm[lane * kLD + k];
        }
        if (lane < kLD) V[lane * kLDV] = 0.0;
        for (int b = 1; b < 5; ++b) {
          if (lane < kLD) {
            U[5 * kLD + lane] = (lane < N) ? z : 0.0;
            V[lane * kLDV + b] = (lane < N) ? z : 0.0;
          }
          __syncwarp();
          if (lane < N) {
            double a0 = U[b * kLD + lane], a1 = 0.0;
#pragma unroll
            for (int k = 0; k < N; k += 2) {
              a0 = fma(pr[k], U[5 * kLD + k], a0);
              if (k + 1 < N) a1 = fma(pr[k + 1], U[5 * kLD + k + 1], a1);
            }
            z = a0 + a1;
          }
          __syncwarp();
        }
        if (lane < N) R[lane * ldr + 5] = z;
      }
      gsync();
      CMPC_TICK(8);
    }
    if (s > n_pow) return;   // horizons of at most 8 rows have no giant step: the ladder ends after the baby stages
    // this warp's row block of Pm multiplies: Pm (squaring), V (j < 3), R (j >= 3);
    // its column block of Pm is multiplied by the rows of L (j < 3).  All products of a stage
    // are independent: their DMMAs are issued interleaved.
    double a[3];
    frag_a12(Pm, mt_w, lane, a);
    if constexpr (j < 3) {
      constexpr int vc = 2 << j, l_cnt = 1 << j, n_lm = (l_cnt * NY + 7) >> 3;
      // tiles 0,1: squaring; 2: V.  While V has at most 4 columns (j < 2) they ride in the unused half
      // of the squaring's second column tile (Pm has columns 8..11 to offer there).
      constexpr bool v_merged = j < 2;
      constexpr int n_sq_tiles = v_merged ? 2 : 3;
      double bq[3][3], cq[3][2];
      frag_b(Pm, kLD, 0, lane, bq[0]);
      if constexpr (v_merged) {
        const int nl = lane >> 2;
        const double* pb = (nl < 4) ? Pm + (lane & 3) * kLD + 8 + nl : V + (lane & 3) * kLDV + (nl - 4 < vc ? nl - 4 : 0);
        const int ldb = (nl < 4) ? kLD : kLDV;
        bq[1][0] = pb[0];
        bq[1][1] = pb[4 * ldb];
        bq[1][2] = pb[8 * ldb];
      } else {
        frag_b12(Pm, 1, lane, bq[1]);
        frag_b(V, kLDV, 0, lane, bq[2]);
      }
      double al[2][3], bl[3], cl[2][2];
      frag_b12(Pm, mt_w, lane, bl);      // column block nt = w of Pm
      // fragment row i of the doubling is (y, a) = (i >> j, i mod 2^j), i.e. row 8 y + a of L
      int l_row[2];
#pragma unroll
      for (int mt = 0; mt < n_lm; ++mt) {
        const int i = 8 * mt + (lane >> 2), yy = i >> j;
        l_row[mt] = (yy < NY) ? kBaby * yy + (i & (l_cnt - 1)) : -1;
        al[mt][0] = al[mt][1] = al[mt][2] = 0.0;
        if (l_row[mt] >= 0) {   // lanes past the rows in use stay out of the load (fewer wavefronts)
          const double* pl = L + l_row[mt] * kLD + (lane & 3);
          al[mt][0] = pl[0];
          al[mt][1] = pl[4];
          al[mt][2] = pl[8];
        }
      }
      mma3_shared_a_range<3, 0, n_sq_tiles>(cq, a, bq);
      mma3_shared_b_range<2, n_lm>(cl, al, bl);
      if constexpr (!v_merged) tile_store(V, kLDV, 0, vc, kLD, vc, mt_w, 0, lane, cq[2]);
#pragma unroll
      for (int mt = 0; mt < n_lm; ++mt) {
        const int cc = 8 * mt_w + 2 * (lane & 3);
        if (l_row[mt] >= 0 && cc < kLD) {
          double2 v;
          v.x = cl[mt][0];
          v.y = cl[mt][1];
          *reinterpret_cast<double2*>(L + (l_row[mt] + l_cnt) * kLD + cc) = v;
        }
      }
      tile_store(Pn, kLD, 0, 0, kLD, kLD, mt_w, 0, lane, cq[0]);
      if constexpr (v_merged) {
        const int r = 8 * mt_w + frag_row(lane), q = lane & 3;
        if (r < kLD) {
          double2 v;
          v.x = cq[1][0];
          v.y = cq[1][1];
          if (q < 2) *reinterpret_cast<double2*>(Pn + r * kLD + 8 + 2 * q) = v;                      // Pm^2 columns 8..11
          else if (2 * (q - 2) < vc) *reinterpret_cast<double2*>(V + r * kLDV + vc + 2 * (q - 2)) = v;   // the new V columns
        }
      } else {
        tile_store(Pn, kLD, 0, 0, kLD, kLD, mt_w, 1, lane, cq[1]);
      }
    } else {
      // giant steps: blocks [r_base, r_base + cnt) = Pm * blocks [0, cnt).  The first n_f targets are
      // full blocks (their sources are the contiguous columns [0, 6 n_f)); the remaining n_p keep
      // three columns each, gathered from columns cc = 0, 2, 4 of their source blocks.  All tiles of
      // the stage (squaring, full, pruned) go through the tensor cores six at a time; with a
      // compile-time horizon every count below is a constant and the code is straight-line.
      constexpr int r_base = 1 << (j - 3);
      const bool do_sq = s < n_pow;   // the last stage needs no further power
      const int cnt = do_sq ? r_base : b_max - r_base;
      int n_f = b_full - r_base;
      n_f = n_f < 0 ? 0 : (n_f > cnt ? cnt : n_f);
      const int n_p = cnt - n_f;
      const int f_cols = n_f * kNC, f_tiles = (f_cols + 7) >> 3;
      const int p_cols = 3 * n_p, p_tiles = (p_cols + 7) >> 3;
      const int p_col0 = giant_col(r_base + n_f, 0, b_full);   // first target column of the pruned part
      const int sq_tiles = do_sq ? 2 : 0, n_tiles = sq_tiles + f_tiles + p_tiles;
      auto load_tile = [&](int k, double (&bt)[3]) {
        if (k < sq_tiles) {
          frag_b12(Pm, k, lane, bt);
        } else if (k < sq_tiles + f_tiles) {
          frag_b(R, ldr, k - sq_tiles, lane, bt);
        } else {
          const int m = 8 * (k - sq_tiles - f_tiles) + (lane >> 2);   // pruned-part column of this lane
          const int col = m < p_cols ? giant_col(n_f + m / 3, 2 * (m % 3), b_full) : 0;
          const double* pb = R + (lane & 3) * ldr + col;
          bt[0] = pb[0];
          bt[1] = pb[4 * ldr];
          bt[2] = pb[8 * ldr];
        }
      };
      auto store_tile = [&](int k, const double (&ct)[2]) {
        if (k < sq_tiles) {
          tile_store(Pn, kLD, 0, 0, kLD, kLD, mt_w, k, lane, ct);
        } else if (k < sq_tiles + f_tiles) {
          tile_store(R, ldr, 0, r_base * kNC, kLD, f_cols, mt_w, k - sq_tiles, lane, ct);
        } else {
          // the pruned part may start at an odd column: 8-byte stores
          const int r = 8 * mt_w + frag_row(lane), m = 8 * (k - sq_tiles - f_tiles) + 2 * (lane & 3);
          if (r < kLD) {
            if (m < p_cols) R[r * ldr + p_col0 + m] = ct[0];
            if (m + 1 < p_cols) R[r * ldr + p_col0 + m + 1] = ct[1];
          }
        }
      };
#pragma unroll
      for (int k0 = 0; k0 < kMaxStageTilesLadder; k0 += 6) {
        if (k0 < n_tiles) {
          double bw[6][3], cw[6][2];
#pragma unroll
          for (int i = 0; i < 6; ++i)
            if (k0 + i < n_tiles) load_tile(k0 + i, bw[i]);
          mma3_shared_a<6>(cw, a, bw, n_tiles - k0 < 6 ? n_tiles - k0 : 6);
#pragma unroll
          for (int i = 0; i < 6; ++i)
            if (k0 + i < n_tiles) store_tile(k0 + i, cw[i]);
        }
      }
    }
    gsync();
  };
  stage(std::integral_constant<int, 0>{});
  stage(std::integral_constant<int, 1>{});
  stage(std::integral_constant<int, 2>{});
  stage(std::integral_constant<int, 3>{});
  stage(std::integral_constant<int, 4>{});
  stage(std::integral_constant<int, 5>{});
  stage(std::integral_constant<int, 6>{});
  stage(std::integral_constant<int, 7>{});
  static_assert(kMaxPow == 8, "one stage call per power");

  CMPC_TICK(2);
  // ---- phase 5: E[a + 8b][y][c] = (L_a R_b)[y][c]: (8 NY x N) (N x b_max kNC) on the tensor cores ----
  // warp w takes the column blocks nt = w, w + WPC, ...; the NY row blocks of L stay in registers.
  // Prefix-table form (kPT): the left factor is the running sum of the baby steps, LT_a = sum_{a' <= a} L_a', so
  // the table holds T[8 b + a] = sum_{k = 8 b}^{8 b + a} E_k, the prefix sum of the impulse response INSIDE its block
  // of 8 rows.  E_r is then a difference of two neighbouring entries, and the prefix sums over the whole horizon
  // that Su and Sf need are T plus the block prefix BP[b] = sum_{b' < b} T[8 b' + 7] (13 serial additions per
  // channel), instead of a 15-channel shuffle scan over the 64 threads of the group (150 shuffles per warp).
  // (LT is formed inside the ladder, as soon as the baby steps are complete.)
  {
    const int n_nt = (r_cols_all + 7) >> 3;
    double al[NY][3];
#pragma unroll
    for (int mt = 0; mt < NY; ++mt) frag_a(kPT ? LT : L, kLD, mt, lane, al[mt]);
    // row block mt of L is output mt, its rows are the baby steps a: a lane's two results are
    // channels (cc, cc + 1) of row a + 8 b (cc is even, so both stay inside block b)
    auto store_e = [&](int mt, int nt, const double (&c)[2]) {
      const int a = frag_row(lane), n = 8 * nt + 2 * (lane & 3);
      if (n < kNC * b_full) {
        const int b = n / kNC, cc = n % kNC;
        double* e = E + (mt * kNC + cc) * ldE + kBaby * b + a;
        e[0] = c[0];
        e[ldE] = c[1];
      } else {
        // blocks that keep cc = 0, 2, 4 only (6 b_full is even: a pair never straddles the two parts)
#pragma unroll
        for (int el = 0; el < 2; ++el) {
          const int m = n + el - kNC * b_full;
          if (n + el < r_cols_all) E[(mt * kNC + 2 * (m % 3)) * ldE + kBaby * (b_full + m / 3) + a] = c[el];
        }
      }
    };
    if (warp == WPC - 1) {
      // CZ[r][y] = (L_a Z_b)[y] with r + 1 = 8 b + a: the block-state part of Sx x_aug for r < 39
      double bz[3], czz[NY][2];
      frag_b(V, kLDV, 0, lane, bz);
      if constexpr (kPT) {
        double alz[NY][3];   // the baby steps themselves
#pragma unroll
        for (int mt = 0; mt < NY; ++mt) frag_a(L, kLD, mt, lane, alz[mt]);
        mma3_shared_b<NY>(czz, alz, bz);
      } else {
        mma3_shared_b<NY>(czz, al, bz);
      }
#pragma unroll
      for (int mt = 0; mt < NY; ++mt) {
        const double (&cz)[2] = czz[mt];
        const int a = frag_row(lane), y = mt, n = 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = 8 * (n + e) + a - 1;
          if (n + e < 5 && r >= 0 && r < kDelay - 1) CZ[r * NY + y] = cz[e];
        }
      }
    }
    if (lay.e_alias) {
      constexpr int MAXNT = kStageTiles / NY;   // column blocks per warp held in registers
      double acc[MAXNT][NY][2];
#pragma unroll
      for (int i = 0; i < MAXNT; ++i) {
        const int nt = warp + i * WPC;
        if (nt < n_nt) {
          double b[3];
          frag_b(R, ldr, nt, lane, b);
          mma3_shared_b<NY>(acc[i], al, b);
        }
      }
      gsync();   // L, R and the powers are dead: E overwrites them
#pragma unroll
      for (int i = 0; i < MAXNT; ++i) {
        const int nt = warp + i * WPC;
        if (nt < n_nt) {
#pragma unroll
          for (int mt = 0; mt < NY; ++mt) store_e(mt, nt, acc[i][mt]);
        }
      }
    } else {
      for (int nt = warp; nt < n_nt; nt += WPC) {
        double b[3], c[NY][2];
        frag_b(R, ldr, nt, lane, b);
        mma3_shared_b<NY>(c, al, b);
#pragma unroll
        for (int mt = 0; mt < NY; ++mt) store_e(mt, nt, c[mt]);
      }
    }
  }
  gsync();
  if constexpr (kPT) {
    if (warp == 0) {
      // Rows r < 39 see a partially drained delay line: with r + 1 = 8 bb + aa the a-priori free response is
      // C~ Ad^aa Z_bb (CZ, phase 5) plus the aa inputs of the current block, a short convolution with the first rows
      // of the delayed-input columns.  One lane per (block, output) forms the seven sums of its block from
      // 16-byte loads -- the taps E_k = T[k] - T[k-1] of both delayed columns, the eight queue entries of the
      // block -- and adds them to CZ (a loop over the taps inside the Gram phase costs 360 wavefronts per
      // scenario, this 110).
      for (int it = lane; it < 5 * NY; it += 32) {
        const int bb = it / NY, yy = it % NY;
        double tap[2][kBaby], qd[2][kBaby];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const double2* tp = reinterpret_cast<const double2*>(E + (yy * kNC + 1 + 2 * d) * ldE);
          const double2* qp = reinterpret_cast<const double2*>(q + d * kDelay + kBaby * bb);
#pragma unroll
          for (int k = 0; k < kBaby / 2; ++k) {
            const double2 tv = tp[k], qv = qp[k];
            tap[d][2 * k] = tv.x;
            tap[d][2 * k + 1] = tv.y;
            qd[d][2 * k] = qv.x;
            qd[d][2 * k + 1] = qv.y;
          }
#pragma unroll
          for (int k = kBaby - 1; k > 0; --k) tap[d][k] -= tap[d][k - 1];
        }
#pragma unroll
        for (int aa = 1; aa < kBaby; ++aa) {
          double v0 = 0.0, v1 = 0.0;
#pragma unroll
          for (int i = 0; i < aa; ++i) {
            v0 = fma(tap[0][aa - 1 - i], qd[0][i], v0);
            v1 = fma(tap[1][aa - 1 - i], qd[1][i], v1);
          }
          const int r = kBaby * bb + aa - 1;
          if (r < kDelay - 1) CZ[r * NY + yy] += v0 + v1;
        }
      }
    } else {
      for (int idx = t - 32; idx < NSC; idx += TPC - 32) {
        const int yy = idx / kNS, c = idx % kNS;
        const double* e = E + (yy * kNC + c) * ldE + (kBaby - 1);
        const int nb = (c == 1 || c == 3) ? b_full : b_max;   // delayed-input columns stop at b_full
        // eight blocks at a time: the loads first (a store to BP might alias them for all the compiler knows)
        double sacc = 0.0;
        for (int b0 = 0; b0 < nb; b0 += 8) {
          double ev[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ev[i] = (b0 + i < nb) ? e[kBaby * (b0 + i)] : 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (b0 + i < nb) BP[idx * ldBP + b0 + i] = sacc;
            sacc += ev[i];
          }
        }
      }
    }
    gsync();
  }
  if (G.etab) {
    double* ge = G.etab + (size_t(scen) * S::NCTRL + g) * (size_t(p) * NY * 5);
    for (int idx = t; idx < p * NY * 5; idx += TPC) {
      const int r = idx / (5 * NY), cc = idx % 5;
      // delayed-input columns do not exist in the blocks b >= b_full (and nothing reads them there)
      const double* e = E + (((idx / 5) % NY) * kNC + cc) * ldE + r;
      ge[idx] = ((cc & 1) && r >= kBaby * b_full) ? 0.0 : (kPT && (r & 7)) ? e[0] - e[-1] : e[0];
    }
  }

  CMPC_TICK(3);
  // ---- phase 6: QP assembly.  Thread t owns prediction rows [t*rpt, (t+1)*rpt). ---------------
  //   G_r[y][c]: c < 4 input columns (delayed ones read 40 rows back), c = 4 the fd column
  //   prefix sums over r by a register scan: local totals -> warp scan -> carry across warps
  //   w_r = Sf fd + Sx x_aug - (y_ref - y)   (mpc_qp_solver.cc:31-37)
  //   H = Su' Q Su + R, Gx = Su' Q Su_other, f = Su' Q w accumulated per thread, then reduced.
  // RPT = rows per thread (compile time): 2 covers p <= 128, 4 covers p <= 256
  double acc[S::NACC];
#pragma unroll
  for (int i = 0; i < S::NACC; ++i) acc[i] = 0.0;
  // one prediction row's share of H = Su' Q Su, Gx = Su' Q Su_other and f = Su' Q w
  auto gram_row = [&](const double (&su)[NY][NV], const double (&so)[NY][NVO > 0 ? NVO : 1], const double (&wv)[NY]) {
    double qs[NY][NV];
#pragma unroll
    for (int y = 0; y < NY; ++y)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        double s = 0.0;
#pragma unroll
        for (int y2 = 0; y2 < NY; ++y2) s = fma(P.c[g].Q[y * NY + y2], su[y2][v], s);
        qs[y][v] = s;
      }
    int hi = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int v2 = v; v2 < NV; ++v2, ++hi)
#pragma unroll
        for (int y = 0; y < NY; ++y) acc[hi] = fma(su[y][v], qs[y][v2], acc[hi]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int vo = 0; vo < NVO; ++vo)
#pragma unroll
        for (int y = 0; y < NY; ++y)
          acc[NH + v * NVO + vo] = fma(so[y][vo], qs[y][v], acc[NH + v * NVO + vo]);
#pragma unroll
      for (int y = 0; y < NY; ++y) acc[NH + NV * NVO + v] = fma(wv[y], qs[y][v], acc[NH + NV * NVO + v]);
    }
  };
  if constexpr (!kPT) {
    const int r0 = t * RPT;
    // the thread's RPT rows of every channel come in as 16-byte row pairs (r0 and kDelay are
    // multiples of RPT, so the pairs of the delayed columns are aligned too)
    auto load_g = [&](double (&gv)[RPT][NSC]) {
      const bool on = r0 < p, del = r0 >= kDelay;
#pragma unroll
      for (int y = 0; y < NY; ++y)
#pragma unroll
        for (int c = 0; c < kNS; ++c) {
          const bool delayed = (c == 1 || c == 3);
          const bool ld = on && (!delayed || del);
          const double* src = E + (y * kNC + c) * ldE + r0 - ((delayed && del) ? kDelay : 0);
#pragma unroll
          for (int j = 0; j < RPT; j += 2) {
            double2 v = make_double2(0.0, 0.0);
            if (ld) v = *reinterpret_cast<const double2*>(src + j);
            gv[j][y * kNS + c] = v.x;
            gv[j + 1][y * kNS + c] = v.y;
          }
        }
    };
    double off[NSC];
    {
      double tot[NSC];
#pragma unroll
      for (int c = 0; c < NSC; ++c) tot[c] = 0.0;
      {
        double gv[RPT][NSC];
        load_g(gv);
#pragma unroll
        for (int j = 0; j < RPT; ++j)
          if (r0 + j < p) {
#pragma unroll
            for (int c = 0; c < NSC; ++c) tot[c] += gv[j][c];
          }
      }
#pragma unroll
      for (int c = 0; c < NSC; ++c) {
        double v = tot[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += up;
        }
        off[c] = v - tot[c];                       // exclusive prefix inside the warp
        if (lane == 31) carry[warp * NSC + c] = v;  // warp total
      }
      gsync();
#pragma unroll
      for (int c = 0; c < NSC; ++c)
        for (int w = 0; w < warp; ++w) off[c] += carry[w * NSC + c];
    }
    CMPC_TICK(4);
    // Sx x_aug, delay-line part.  Rows r >= 39 read C~ Ad^(r-39) X40 from the table.  Rows r < 39
    // see a partially drained delay line: with r + 1 = 8 b + a the state is Ad^a Z_b plus the a
    // inputs of the current block, i.e. CZ[r] plus a short convolution (fewer than 8 taps).
    double conv[RPT][NY];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int r = r0 + j;
#pragma unroll
      for (int y = 0; y < NY; ++y) conv[j][y] = 0.0;
      if (r < p) {
        if (r >= kDelay - 1) {
#pragma unroll
          for (int y = 0; y < NY; ++y) conv[j][y] = E[(y * kNC + 5) * ldE + r - (kDelay - 1)];
        } else {
          const int bb = (r + 1) >> 3, aa = (r + 1) & 7;
#pragma unroll
          for (int y = 0; y < NY; ++y) conv[j][y] = CZ[r * NY + y];
          for (int i = 0; i < aa; ++i) {
            const double* Ek = E + (aa - 1 - i);
            const double q0 = q[8 * bb + i], q1 = q[kDelay + 8 * bb + i];
#pragma unroll
            for (int y = 0; y < NY; ++y)
              conv[j][y] = fma(Ek[(y * kNC + 1) * ldE], q0, fma(Ek[(y * kNC + 3) * ldE], q1, conv[j][y]));
          }
        }
      }
    }
    CMPC_TICK(5);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int r = r0 + j;
      if (r < p) {
        // one row at a time (8-byte loads): both rows of a pair would not fit the register budget
        double gv[NSC], su[NY][NV], so[NY][NVO > 0 ? NVO : 1], wv[NY];
        {
          const bool del = r >= kDelay;
#pragma unroll
          for (int y = 0; y < NY; ++y)
#pragma unroll
            for (int c = 0; c < kNS; ++c) {
              const bool delayed = (c == 1 || c == 3);
              gv[y * kNS + c] = (!delayed || del) ? E[(y * kNC + c) * ldE + r - (delayed ? kDelay : 0)] : 0.0;
            }
        }
#pragma unroll
        for (int y = 0; y < NY; ++y) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < NU) {
              su[y][i] = gv[y * kNS + i];
              su[y][NU + i] = off[y * kNS + i];
            } else if (NVO > 0) {
              so[y][i - NU] = gv[y * kNS + i];
              so[y][NO + i - NU] = off[y * kNS + i];
            }
          }
          const int oy = P.c[g].out_idx[y];
          const double yref = P.yref[(size_t(g) * NY + y) * p + r];
          wv[y] = (off[y * kNS + 4] + gv[y * kNS + 4]) + dxd[oy] + conv[j][y] - (yref - yv[oy]);
        }
#pragma unroll
        for (int c = 0; c < NSC; ++c) off[c] += gv[c];
        gram_row(su, so, wv);
      }
    }
  } else {
    // Prefix-table form: thread t owns rows t, t + 64, ... (unit-stride, conflict-free 8-byte loads).  Per row and
    // channel: T[r], T[r - 1] (inside the block) and the block prefix give E_r = T[r] - T[r - 1], the exclusive
    // prefix BP + T[r - 1] and, for the f_d column, the inclusive one BP + T[r].
    CMPC_TICK(4);
    CMPC_TICK(5);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int r = t + j * TPC;
      if (r < p) {
        // Sx x_aug, delay-line part (see the scan form above): rows r >= 39 read C~ Ad^(r-39) X40 from the table,
        // rows r < 39 take CZ[r], which holds block state and short convolution by now
        double cv[NY];
        if (r >= kDelay - 1) {
          const int k = r - (kDelay - 1);
#pragma unroll
          for (int y = 0; y < NY; ++y) {
            const double* e5 = E + (y * kNC + 5) * ldE + k;
            cv[y] = e5[0] - ((k & 7) ? e5[-1] : 0.0);
          }
        } else {
#pragma unroll
          for (int y = 0; y < NY; ++y) cv[y] = CZ[r * NY + y];
        }
        double su[NY][NV], so[NY][NVO > 0 ? NVO : 1], wv[NY];
        const bool del = r >= kDelay;
#pragma unroll
        for (int y = 0; y < NY; ++y) {
#pragma unroll
          for (int c = 0; c < kNS; ++c) {
            const bool delayed = (c == 1 || c == 3);
            double gvv = 0.0, ex = 0.0, inc = 0.0;
            if (!delayed || del) {
              const int rr = r - (delayed ? kDelay : 0);
              const double* e = E + (y * kNC + c) * ldE + rr;
              const double cur = e[0], prev = (rr & 7) ? e[-1] : 0.0, bpv = BP[(y * kNS + c) * ldBP + (rr >> 3)];
              gvv = cur - prev;
              ex = bpv + prev;
              inc = bpv + cur;
            }
            if (c < 4) {
              if (c < NU) {
                su[y][c] = gvv;
                su[y][NU + c] = ex;
              } else if (NVO > 0) {
                so[y][c - NU] = gvv;
                so[y][NO + c - NU] = ex;
              }
            } else {
              const int oy = P.c[g].out_idx[y];
              const double yref = P.yref[(size_t(g) * NY + y) * p + r];
              wv[y] = inc + dxd[oy] + cv[y] - (yref - yv[oy]);
            }
          }
        }
        gram_row(su, so, wv);
      }
    }
  }
  CMPC_TICK(6);
  // only the reduction is left: once the last wave is here, the solve grid may move in.  (Measured:
  // triggering at the top of the kernel instead costs 12 us per step.)
  pdl_trigger();
  // Sum the accumulators over the group: a butterfly transpose-reduce in registers leaves the warp
  // total of accumulator l in lane l (accumulators 32.. in lanes 0..15 of a second pass), then the
  // warps of the group meet through a few words of shared memory.
  {
    static_assert(S::NACC <= 48, "accumulator count");
    constexpr int RS = 48;
    double lo[32], tot_hi = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) lo[i] = (i < S::NACC) ? acc[i] : 0.0;
    const double tot_lo = warp_transpose_reduce<32>(lo, lane);
    if constexpr (S::NACC > 32) {
      double hi[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) hi[i] = (32 + i < S::NACC) ? acc[32 + i] : 0.0;
      tot_hi = warp_transpose_reduce<16>(hi, lane);
    }
    gsync();  // E is dead: reuse it as the exchange buffer
    double* red = E;
    red[warp * RS + lane] = tot_lo;
    if (S::NACC > 32 && lane < 16) red[warp * RS + 32 + lane] = tot_hi;
    gsync();
    if (t < S::NACC) {
      double v = red[t];
#pragma unroll
      for (int w = 1; w < WPC; ++w) v += red[w * RS + t];
      if (t < NH) {
        int a = 0, rem = t;
        while (rem >= NV - a) { rem -= NV - a; ++a; }
        const int b = a + rem;
        double hv = v;
        if (a / NU == b / NU) hv += P.c[g].R[(a % NU) * NU + (b % NU)];  // u_weight_ = I_m (x) uwt
        double* gH = G.qpH + (size_t(scen) * S::NCTRL + g) * NV * NV;
        gH[a * NV + b] = hv;
        gH[b * NV + a] = hv;
      } else if (t < NH + NV * NVO) {
        G.qpG[(size_t(scen) * S::NCTRL + g) * NV * (NVO > 0 ? NVO : 1) + (t - NH)] = v;
      } else {
        const int v_i = t - NH - NV * NVO;
        G.qpf[(size_t(scen) * S::NCTRL + g) * NV + v_i] = v;
      }
    }
  }
  CMPC_TICK(7);
  // what the a-priori observer update (observer.cc:6-19) needs in K2: the part of the new
  // state estimate that does not depend on the move, B (head - u_old) + f_d, and the Bd columns
  // of the undelayed inputs
  if (t < N) {
    wk[kWBF + t] = fma(BF[t * kNC + 1], q[0], fma(BF[t * kNC + 3], q[kDelay], BF[t * kNC + 4]));
    wk[kWBF + N + t] = BF[t * kNC + 0];
    wk[kWBF + 2 * N + t] = BF[t * kNC + 2];
  }
#ifdef CMPC_PHASE_TIMING
  if (threadIdx.x == 0 && g == 0) CMPC_GTIME_AT(scen, 18);
#endif
}

// General solve for one 4-variable QP when the warm-start working set is no longer optimal.
// Kept out of line: it is rare, and its arrays must not cost the sweep loop any registers.
// Everything is passed by value (no pointer into the kernel parameters may escape, or the
// compiler would spill the whole parameter block to local memory in every thread).
struct QpBounds4 {
  double lo0, lo1, up0, up1, rlo0, rlo1, rup0, rup1;
};
static __device__ __noinline__ int qp_fallback4(const double* H, double f0, double f1, double f2, double f3, QpBounds4 b,
                                         unsigned* wset_out) {
  QpData<4> qd;
  qd.lb[0] = qd.lb[2] = b.lo0; qd.lb[1] = qd.lb[3] = b.lo1;
  qd.ub[0] = qd.ub[2] = b.up0; qd.ub[1] = qd.ub[3] = b.up1;
  qd.lbA[0] = qd.lbA[2] = b.rlo0; qd.lbA[1] = qd.lbA[3] = b.rlo1;
  qd.ubA[0] = qd.ubA[2] = b.rup0; qd.ubA[1] = qd.ubA[3] = b.rup1;
  double fi[4] = {f0, f1, f2, f3}, zs[4], obj_;
  unsigned act_, gset = kQpNoGuess;
  const int st = qp_invert_spd<4>(H, qd.J) ? qp_solve<4, 2>(qd, H, fi, &gset, zs, &act_, &obj_) : 3;
  if (st == 0) *wset_out = gset;
  return st;
}

// UpdateU / ObserveAPriori of one controller (distributed_controller.h:146-152, observer.cc:6-19),
// one thread: new state part = base + Bd[:,0] du0 (+ Bd[:,2] du2), the disturbance estimate stays,
// each delay line hands its oldest entry to the head and takes the newly commanded input
// (the chains are rings: position P.ring_pos is the oldest slot and becomes the newest).
template <class S>
__device__ __forceinline__ void apriori_update(const StepParams& P, const DeviceState& G, size_t ctrl_rec,
                                               const double (&du)[4]) {
  constexpr int N = S::N, NOBS = S::NOBS;
  double* gs = G.ctrl + ctrl_rec * kCtrlStateStride;
  const double* wb = G.work + ctrl_rec * kWorkStride + kWBF;
  double base[N], c0[N], c2[N], uold[4];
#pragma unroll
  for (int i = 0; i < N; ++i) { base[i] = wb[i]; c0[i] = wb[N + i]; c2[i] = wb[2 * N + i]; }
#pragma unroll
  for (int i = 0; i < 4; ++i) uold[i] = gs[kOffUold + i];
  const int ring0 = kOffDx + NOBS + 2 + P.ring_pos, ring1 = ring0 + kRing;
  const double old0 = gs[ring0], old1 = gs[ring1];
#pragma unroll
  for (int i = 0; i < N; ++i) gs[kOffDx + i] = fma(c0[i], du[0], fma(c2[i], du[2], base[i]));
  gs[kOffDx + NOBS + 0] = old0;
  gs[kOffDx + NOBS + 1] = old1;
  gs[ring0] = uold[1] + du[1];
  gs[ring1] = uold[3] + du[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) gs[kOffUold + i] = uold[i] + du[i];
}

// ---- K2: Jacobi sweeps (nerve_center.h:146-158,275-296), first move (nerve_center.h:162-167,
// 313-319), UpdateU / ObserveAPriori.  Distributed controllers: one thread per sub-controller,
// the pair of a scenario in neighbouring lanes (16 scenarios per warp).  Centralised: one thread
// per scenario with the general solver.
#ifndef CMPC_SOLVE_MIN_BLOCKS
#define CMPC_SOLVE_MIN_BLOCKS 1
#endif
// MODE 0: the whole of it (the production path).  The other two exist for the reference's timing
// window (GetNextInputWithTiming, nerve_center.h:134-179: sweeps i >= n_timing_iterations are left
// out of the measured time): MODE 1 runs sweeps [it_begin, it_end) and parks plan, working set and
// status in the hand-over record; MODE 2 is what follows the sweeps (first move, UpdateU /
// ObserveAPriori).  Sweeps 0..n-1 followed by n..n_iter-1 and MODE 2 leave exactly the state MODE 0 does.
// APRIORI = false leaves UpdateU / ObserveAPriori to the plant kernel that follows in the closed loop
// (lin_part_early), where the lanes that linearise the next record have time to spare.
template <class S, int MODE, bool APRIORI = true>
__device__ __forceinline__ void solve_body(const StepParams& P, const DeviceState& G, double* __restrict__ u,
                                           int it_begin, int it_end) {
  constexpr int NU = S::NU, NV = S::NV, NVO = S::NVO, NCTRL = S::NCTRL;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (NV == 4 && NCTRL == 2) {
    // Lanes past the end of the batch work along on its last scenario and store nothing: the plan
    // exchange below is a whole-warp shuffle (one mask per lane pair would send it through
    // WARPSYNC.COLLECTIVE, group by group, in every sweep), so the warp must stay whole.
    const bool valid = (tid >> 1) < P.batch;
    const int scen = valid ? (tid >> 1) : P.batch - 1, c = tid & 1;
    double* ss = G.scen + size_t(scen) * kScenStateStride;
    const size_t rec = size_t(scen) * 2 + c;
    double* plan = G.work + rec * kWorkStride + kWPlan;
    if constexpr (MODE == 2) {
      if (!valid) return;
      double zf[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) zf[k] = plan[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) ss[4 + c * 4 + k] = zf[k];
      const double un0 = ss[2 * c] + zf[0], un1 = ss[2 * c + 1] + zf[1];
      ss[2 * c] = un0;
      ss[2 * c + 1] = un1;
      u[size_t(scen) * 4 + 2 * c] = un0;
      u[size_t(scen) * 4 + 2 * c + 1] = un1;
      const double duf[4] = {zf[0], zf[1], 0.0, 0.0};
      apriori_update<S>(P, G, rec, duf);
      return;
    }
    const double* gH = G.qpH + rec * 16;
    const double* gf = G.qpf + rec * 4;
    const double* gG = G.qpG + rec * 16;
    const double* uo = G.ctrl + rec * kCtrlStateStride + kOffUold;
    double J[4][4], Gx[4][4], f0[4], z[4], bnd[16];
    // What the assemble kernel of this record does not write (own previous plan, applied inputs, warm
    // start: all left by the previous record's solve kernel) is fetched BEFORE the dependent-launch wait
    // of the production kernel, together with the first touch of the parameter block.
#pragma unroll
    for (int i = 0; i < 4; ++i)
      z[i] = (MODE == 1 && it_begin > 0) ? plan[i] : ss[4 + c * 4 + i];   // du_prev = du_old_ (own plan)
    const double uo0 = uo[0], uo1 = uo[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double uold = (i & 1) ? uo1 : uo0;
      bnd[i] = P.c[c].lower[i & 1] - uold;
      bnd[4 + i] = -(P.c[c].upper[i & 1] - uold);
      bnd[8 + i] = P.c[c].rate_lower[i & 1];
      bnd[12 + i] = -P.c[c].rate_upper[i & 1];
    }
    unsigned wset = G.guess[rec];
    if (MODE == 1 && it_begin > 0) wset = unsigned(__double_as_longlong(plan[kWPlanSet - kWPlan]));
    if (wset == kQpNoGuess) wset = 0;     // no warm start: begin from the unconstrained minimiser
    const double us0 = ss[2 * c], us1 = ss[2 * c + 1];   // NerveCenter's u_old_ of this controller's inputs
    if constexpr (MODE == 0) {
      pdl_wait();
      pdl_trigger();   // single wave
#ifdef CMPC_PHASE_TIMING
      if (c == 0) CMPC_GTIME_AT(scen, 20);
#endif
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f0[i] = gf[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        J[i][j] = gH[i * 4 + j];
        Gx[i][j] = gG[i * 4 + j];
      }
    }
#ifdef CMPC_PHASE_TIMING
    const long long tk0_ = clock64();
#endif
    if (c == 0 && MODE == 0) CMPC_TICK_AT(scen, 10, tk0_ + (J[0][0] != J[0][0] ? 1 : 0));   // loads have arrived
    const bool pd = qt_inverse(J);
    if (c == 0 && MODE == 0) CMPC_TICK_AT(scen, 30, tk0_ + (J[0][0] != J[0][0] ? 1 : 0));   // inverse done
    // H^-1 is only needed when a reduced system is built (once per step, and again after a change of the
    // working set): it waits in shared memory instead of holding 32 registers through the sweeps, where
    // the reduced system, the cross term and the bounds already fill the register file
    __shared__ double J_sh[16 * 64];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) J_sh[(i * 4 + j) * 64 + threadIdx.x] = J[i][j];
    QtReduced red;
    bool red_ok = false, need_prep = true;   // red belongs to wset once prepared
    double x[4] = {0.0, 0.0, 0.0, 0.0}, lam[4] = {0.0, 0.0, 0.0, 0.0}, fi[4] = {0.0, 0.0, 0.0, 0.0};
    int status = pd ? 0 : 3;
    const int it_lo = MODE == 1 ? it_begin : 0, it_hi = MODE == 1 ? it_end : P.n_iter;
    for (int it = it_lo; it < it_hi; ++it) {
      double zo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) zo[k] = __shfl_xor_sync(0xffffffffu, z[k], 1);   // the other controller's previous plan
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double s = f0[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) s = fma(Gx[i][k], zo[k], s);
        fi[i] = s;
      }
      // One copy of prepare / evaluate / repair (they are large, fully unrolled register code and
      // this kernel is bound by instruction fetch).  The warm-start working set is evaluated first;
      // when it is not optimal it is repaired one constraint at a time in registers (each accepted
      // result satisfies the KKT conditions of the full QP, i.e. is the unique minimiser), and
      // only if that does not settle, the general dual active-set solve supplies the set.
      if (pd) {
        bool fell_back = false;
        int n_rep = 0;
        for (;;) {
          if (need_prep) {
            double Jl[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) Jl[i][j] = J_sh[(i * 4 + j) * 64 + threadIdx.x];
            red_ok = qt_prepare(Jl, bnd, wset, red);
            need_prep = false;
          }
          const bool ok = qt_eval(red, fi, bnd, wset, x, lam) && red_ok;
          if (ok) {
            status = 0;
            break;
          }
          if (fell_back) {
            status = 1;   // should not happen: KKT of the set the general solver returned
            break;
          }
          if (red_ok && n_rep < 6) {
            const unsigned nw = qt_repair(red, x, lam, bnd, wset);
            ++n_rep;
            if (nw != wset) {
              wset = nw;
              need_prep = true;
              continue;
            }
          }
          QpBounds4 qb;
          qb.lo0 = bnd[0]; qb.lo1 = bnd[1]; qb.up0 = -bnd[4]; qb.up1 = -bnd[5];
          qb.rlo0 = bnd[8]; qb.rlo1 = bnd[9]; qb.rup0 = -bnd[12]; qb.rup1 = -bnd[13];
          unsigned new_w = wset;
          status = qp_fallback4(gH, fi[0], fi[1], fi[2], fi[3], qb, &new_w);
          fell_back = true;
          if (status != 0) break;
          wset = new_w;
          need_prep = true;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) z[k] = (status == 0) ? x[k] : 0.0;   // mpc_qp_solver.cc:66-69: zeros on failure
      if (c == 0 && MODE == 0 && it == 0) CMPC_TICK_AT(scen, 31, tk0_ + (z[0] != z[0] ? 1 : 0));   // first sweep done
    }
    if (!valid) return;   // (after the last shuffle)
    if (c == 0 && MODE == 0) CMPC_TICK_AT(scen, 11, tk0_ + (z[0] != z[0] ? 1 : 0));   // sweeps done
    // report of the last sweep: active constraints (strictly positive multiplier), objective
    double fmax = 1.0, obj = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      fmax = fmax > fabs(fi[i]) ? fmax : fabs(fi[i]);
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fma(gH[i * 4 + k], z[k], s);   // (re-read: H is not kept through the sweeps)
      obj += z[i] * (0.5 * s + fi[i]);
    }
    unsigned act = 0, m = wset & 0xffffu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int idx = __ffs(m) - 1;
      m &= m - 1;
      if (idx >= 0 && lam[w] > 1e-9 * fmax) act |= 1u << idx;
    }
    if (status == 0 && (MODE == 0 || it_end == P.n_iter)) G.guess[rec] = wset;
    G.status[rec] = status;
    G.active[rec] = status == 0 ? act : 0u;
    G.objective[rec] = status == 0 ? obj : 0.0;
    if constexpr (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) plan[k] = z[k];
      plan[kWPlanSet - kWPlan] = __longlong_as_double((long long)wset);
    } else {
      // du_old_ = du_prev; u_old_ += first move of each controller's plan (system order = ctrl 0, ctrl 1)
#pragma unroll
      for (int k = 0; k < 4; ++k) ss[4 + c * 4 + k] = z[k];
      const double un0 = us0 + z[0], un1 = us1 + z[1];
      ss[2 * c] = un0;
      ss[2 * c + 1] = un1;
      u[size_t(scen) * 4 + 2 * c] = un0;
      u[size_t(scen) * 4 + 2 * c + 1] = un1;
      // each controller sees only its own inputs move (nerve_center.h:322-328)
      const double du[4] = {z[0], z[1], 0.0, 0.0};
      if (APRIORI) apriori_update<S>(P, G, rec, du);
      if (c == 0 && MODE == 0) CMPC_TICK_AT(scen, 12, tk0_);
    }
  } else {
    // centralised controller: a single controller has no plan to exchange, every sweep solves the
    // same QP (distributed_controller.h:215-218), so one solve gives the result of all sweeps
    static_assert(NVO == 0 && NCTRL == 1, "generic path is the centralised controller");
    if constexpr (MODE == 0) {
      pdl_wait();
      pdl_trigger();   // single wave
    }
    const int scen = tid;
    if (scen >= P.batch) return;
    double* ss = G.scen + size_t(scen) * kScenStateStride;
    const size_t rec = size_t(scen);
    double* plan = G.work + rec * kWorkStride + kWPlan;
    if constexpr (MODE == 1) {
      if (it_begin > 0 || it_end <= it_begin) return;   // every sweep solves the same QP: the first one has the answer
    }
    if constexpr (MODE == 2) {
      double du[4];
#pragma unroll
      for (int i = 0; i < NV; ++i) ss[4 + i] = plan[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        du[k] = plan[k];
        const double un = ss[k] + du[k];
        ss[k] = un;
        u[size_t(scen) * 4 + k] = un;
      }
      apriori_update<S>(P, G, rec, du);
      return;
    }
    const double* gH = G.qpH + rec * NV * NV;
    const double* gf = G.qpf + rec * NV;
    const double* uo = G.ctrl + rec * kCtrlStateStride + kOffUold;
    QpData<NV> qd;
    double f0[NV], z[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      f0[i] = gf[i];
      z[i] = 0.0;
      qd.lb[i] = P.c[0].lower[i % NU] - uo[i % NU];
      qd.ub[i] = P.c[0].upper[i % NU] - uo[i % NU];
      qd.lbA[i] = P.c[0].rate_lower[i % NU];
      qd.ubA[i] = P.c[0].rate_upper[i % NU];
    }
    unsigned wset = G.guess[rec], act = 0;
    double obj = 0.0;
    int status = 3;
    if (qp_invert_spd<NV>(gH, qd.J)) status = qp_solve<NV, NU>(qd, gH, f0, &wset, z, &act, &obj);
    if (status != 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) z[i] = 0.0;
    } else {
      G.guess[rec] = wset;
    }
    G.status[rec] = status;
    G.active[rec] = status == 0 ? act : 0u;
    G.objective[rec] = status == 0 ? obj : 0.0;
    if constexpr (MODE == 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) plan[i] = z[i];
    } else {
      double du[4];
#pragma unroll
      for (int i = 0; i < NV; ++i) ss[4 + i] = z[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        du[k] = z[k];
        const double un = ss[k] + z[k];
        ss[k] = un;
        u[size_t(scen) * 4 + k] = un;
      }
      if (APRIORI) apriori_update<S>(P, G, rec, du);
    }
  }
}

template <class S, bool APRIORI>
__global__ void __launch_bounds__(64, CMPC_SOLVE_MIN_BLOCKS)
solve_kernel(StepParams P, DeviceState G, double* __restrict__ u) {
#ifdef CMPC_PHASE_TIMING
  const int sc_ = (blockIdx.x * blockDim.x + threadIdx.x) / S::NCTRL;
  const bool st_ = (threadIdx.x & 31) == 0 && sc_ < P.batch;
  if (st_) CMPC_GTIME_AT(sc_, 19);
#endif
  solve_body<S, 0, APRIORI>(P, G, u, 0, 0);   // (waits for the assemble kernel inside, after its own preparations)
#ifdef CMPC_PHASE_TIMING
  if (st_) CMPC_GTIME_AT(sc_, 21);
#endif
}

// The same in pieces, for the timing window only.
template <class S>
__global__ void __launch_bounds__(64, CMPC_SOLVE_MIN_BLOCKS)
solve_sweeps_kernel(StepParams P, DeviceState G, int it_begin, int it_end) {
  solve_body<S, 1>(P, G, nullptr, it_begin, it_end);
}
template <class S>
__global__ void __launch_bounds__(64, CMPC_SOLVE_MIN_BLOCKS)
solve_finish_kernel(StepParams P, DeviceState G, double* __restrict__ u) {
  solve_body<S, 2>(P, G, u, 0, 0);
}

}  // namespace cmpc
