/* cmpc.h — C ABI of the B200-native compressor-MPC control step.
 *
 * Drop-in boundary for the per-sample-time control hot path of
 * katie-jones/compressor-mpc, batched over B independent plant scenarios.
 * The reference has no FFI layer; its seam is the C++ interface
 *   ControllerInterface<System>::GetNextInput(const Output& y)
 *       (include/controller_interface.h:46), implemented by
 *   NerveCenter (include/nerve_center.h:89-182).
 * Every entry point below names the reference method it replaces.  All arrays
 * are plain row-major doubles/ints, scenario-major (leading dimension B); "host"
 * entry points take host pointers and do the host<->device copies themselves,
 * "_device" entry points take device pointers (same layout) and a CUDA stream.
 * The handle owns all device memory.  One handle per host thread / stream
 * (same threading contract as the reference: not thread-safe).
 * All functions return CMPC_OK (0) or a CMPC_ERR_* code; cmpc_last_error()
 * gives the text.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with CMPC_ERR_CUDA.
 */
#ifndef CMPC_H
#define CMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CMPC_OK = 0, CMPC_ERR_ARG = 1, CMPC_ERR_CUDA = 2, CMPC_ERR_UNSUPPORTED = 3, CMPC_ERR_STATE = 4 };
enum { CMPC_PLANT_PARALLEL = 0, CMPC_PLANT_SERIAL = 1 };                 /* systems/{parallel,serial}_compressors.cc */
enum { CMPC_MODE_CENTRALIZED = 0, CMPC_MODE_COOPERATIVE = 1, CMPC_MODE_NONCOOPERATIVE = 2,
       /* serial plant only: the earlier non-cooperative output partition {0,1,2} / {2,3,1}
        * (SERIAL_CTRL_NONCOOP_OLD{1,2}, serial_compressors_constants.h:47-59,103-104) */
       CMPC_MODE_NONCOOPERATIVE_OLD = 3 };

#define CMPC_N_CONTROL_INPUTS 4   /* plant inputs {0,3,4,7}: torque1, recycle1, torque2, recycle2 */
#define CMPC_N_OUTPUTS 4
#define CMPC_MAX_CONTROLLERS 4   /* NerveCenter takes a parameter pack of sub-controllers (nerve_center.h:19-38) */

typedef struct cmpc_handle cmpc_handle;

/* Runtime form of the reference's compile-time configuration
 * (include/constexpr_array.h:10-142, include/parallel_compressors_constants.h:70-93,
 *  include/serial_compressors_constants.h:84-109, include/common-variables.h:20-112,
 *  the sub-controller pack of include/nerve_center.h:19-38).
 * The reference's own instantiations (m = 2, Delays = {0,40,0,40}, its seven controller shapes) run
 * on kernels tuned for them; every other configuration runs on a general path with the same
 * results contract:
 *   delays     per control input {torque1, recycle1, torque2, recycle2}: 0 or 2..128 samples; every
 *              sub-controller's AugmentedLinearizedSystem sees them in its own input order
 *   m          1..4 with m * (own inputs) <= 8 per sub-controller;  p in [m, 256]
 *   partitions 1..4 sub-controllers, each with its own input count (n_sub_control_inputs_per, summing
 *              to 4; the own inputs of sub-controller c are the system inputs that follow those of
 *              the sub-controllers before it, as NerveCenter assumes, nerve_center.h:313-328) and its
 *              own controlled outputs (1..4 of the plant's 4)
 * With more than two sub-controllers the plans of the others reach a sub-controller exactly as in
 * the reference: concatenated in controller order and paired with the columns of Su_other by
 * position (nerve_center.h:281-286, distributed_solver.h:109-115). */
typedef struct cmpc_config {
  int32_t plant;                 /* CMPC_PLANT_* */
  int32_t mode;                  /* CMPC_MODE_* (only read by cmpc_default_config) */
  int32_t p;                     /* prediction horizon (reference: 100; sweep: 200) */
  int32_t m;                     /* move horizon (reference: 2) */
  double Ts;                     /* sampling time (0.05 s) */
  int32_t n_iterations;          /* solver sweeps per step: setup key n-iterations (1 cent, 9 distributed) */
  int32_t batch;                 /* B independent scenarios */
  int32_t delays[4];             /* Delays (reference: {0,40,0,40}), per control input in plant order */
  int32_t n_disturbance_states;  /* 4 */
  int32_t n_controllers;         /* 1..CMPC_MAX_CONTROLLERS */
  int32_t n_sub_control_inputs;  /* own inputs per sub-controller: 4 (cent) or 2 */
  int32_t n_controlled_outputs[CMPC_MAX_CONTROLLERS];
  int32_t controlled_output_indices[CMPC_MAX_CONTROLLERS][4]; /* ControlledOutputIndices */
  int32_t control_input_indices[CMPC_MAX_CONTROLLERS][4];     /* ControlInputIndices: own inputs first */
  int32_t n_sub_control_inputs_per[CMPC_MAX_CONTROLLERS];     /* own inputs of each sub-controller; 0 = n_sub_control_inputs */
} cmpc_config;

/* Fill cfg with the reference's constants for one of its six workflows. */
int cmpc_default_config(int plant, int mode, int batch, cmpc_config* cfg);

/* Plant dimensions for a plant kind: n_states (11|10), n_inputs (9|8). */
int cmpc_plant_dims(int plant, int* n_states, int* n_inputs);
/* {Parallel,Serial}Compressors::GetDefaultState / GetDefaultInput
 * (include/parallel_compressors.h:74-86, include/serial_compressors.h:85-95). */
int cmpc_plant_defaults(int plant, double* x_default, double* u_default);

/* NerveCenter ctor + DistributedController ctors (nerve_center.h:89-95,
 * distributed_controller.cc:6-22).  Allocates all per-scenario state on `device`. */
int cmpc_create(const cmpc_config* cfg, int device, cmpc_handle** out);
int cmpc_destroy(cmpc_handle* h);
const char* cmpc_last_error(void);

/* NerveCenter::SetWeights, tuple overload (nerve_center.h:113-116 -> mpc_qp_solver.h:62-80).
 * uwt: n_u x n_u (this controller's sub-matrix of the full input weight, n_u its own inputs),
 * ywt: n_controlled_outputs^2.  Both must be symmetric (CMPC_ERR_UNSUPPORTED otherwise: H is kept
 * as a symmetric matrix) and finite. */
int cmpc_set_weights(cmpc_handle* h, int ctrl, const double* uwt, const double* ywt);
/* NerveCenter::SetOutputReference (nerve_center.h:119-122): yref is p x 4 (all plant outputs). */
int cmpc_set_output_reference(cmpc_handle* h, const double* yref);
/* InputConstraints (include/input_constraints.h:11-27), one entry per own input each.  +-infinity
 * means unbounded; NaN is CMPC_ERR_ARG.  lower > upper is taken as it is, like the reference does:
 * every QP is infeasible and every step applies the zero move (mpc_qp_solver.cc:66-69). */
int cmpc_set_constraints(cmpc_handle* h, int ctrl, const double* lower, const double* upper,
                         const double* rate_lower, const double* rate_upper);
/* Observer ctor's ObserverMatrix M, (n_states+n_dist) x 4 row-major (observer.h:31-33).
 * Default [0; I]. */
int cmpc_set_observer_gain(cmpc_handle* h, int ctrl, const double* M);

/* NerveCenter::Initialize (nerve_center.h:98-104): x_init B x n_states, u_init B x 4,
 * u_init_full B x n_inputs (becomes u_offset_), y_init B x 4.  Host pointers. */
int cmpc_initialize(cmpc_handle* h, const double* x_init, const double* u_init,
                    const double* u_init_full, const double* y_init);

/* ControllerInterface::GetNextInput / NerveCenter::GetNextInputWithTiming
 * (controller_interface.h:46, nerve_center.h:125-182), batched: y B x 4 -> u B x 4.
 * Host buffers; H2D of y and D2H of u happen inside (page-locked buffers that are mapped into the
 * device -- u 16-byte aligned -- are read and written by the kernels themselves, with no copy on the stream).
 * CMPC_ERR_STATE after a closed-loop run on the same handle (cmpc_run_closed_loop*): that loop
 * has already consumed the next measurement (its plant kernel runs the observer update and the
 * linearisation of the following record), so the handle needs cmpc_initialize first. */
int cmpc_get_next_input(cmpc_handle* h, const double* y, double* u);
/* NerveCenter::GetNextInputWithTiming (nerve_center.h:134-182): the same step; *time_ns gets the
 * time of what the reference's timer covers, measured with CUDA events on the handle's stream: the
 * copy of y, QP generation, the first n_timing_iterations sweeps and, after the remaining sweeps
 * (not measured, nerve_center.h:151), everything that follows them (first move, a-priori observer
 * update, copy of u).  n_timing_iterations < 0 or >= n_iterations: the whole call. */
int cmpc_get_next_input_timed(cmpc_handle* h, const double* y, double* u, int n_timing_iterations,
                              int64_t* time_ns);
/* Same with device-resident y/u (B x 4 doubles each) on `stream` (a cudaStream_t), no sync. */
int cmpc_get_next_input_device(cmpc_handle* h, const double* y_dev, double* u_dev, void* stream);

/* Deviation from the reference at the QP solver, by design: the reference hands every QP to qpOASES
 * with a working-set-recalculation budget of 10 (n_wsr_max, mpc_qp_solver.h:24, mpc_qp_solver.cc:50-69)
 * and applies a zero move when qpOASES needs more.  The solver here has no such budget: it returns the
 * exact minimiser whenever the QP is feasible and H is positive definite, and the zero move only when
 * it is not (status != 0).  Steps on which the reference would have run out of its budget therefore
 * differ; none of the reference's six recorded runs contains one (they are reproduced over all
 * 10 000 records), and the budget depends on qpOASES's homotopy path, which is not restated.
 *
 * Per-scenario result of the last step's final solver sweep (mpc_qp_solver.cc:62-75):
 * status B x n_ctrl (0 ok; !=0 => that controller applied zeros), active B x n_ctrl
 * (bitmask over 4*nv one-sided constraints: [0,nv) z>=lb, [nv,2nv) z<=ub, [2nv,3nv) rate>=,
 * [3nv,4nv) rate<=), objective B x n_ctrl (1/2 z'Hz + f'z).  Any pointer may be NULL. */
int cmpc_get_step_info(cmpc_handle* h, int32_t* status, uint32_t* active, double* objective);

/* Closed loop on the device (reconstructed driver loop; plant side:
 * SimulationSystem::{SetOffset,SetInput,Integrate} simulation_system.h:66-116,
 * TimeDelay::GetDelayedInput time_delay.h:41-58).  Scenario b starts at x0[b] with the
 * controller initialised like the reference driver (u_init = 0, u_offset = default input,
 * y0 = GetOutput(x0)); records [block_end[b][i-1], block_end[b][i]) run with plant-input
 * offsets block_off[b][i][:] added to the default input.
 * traj: B x n_steps x (1+n_states+4+4) = [t, x, u, y] per record (may be NULL);
 * qp_active / qp_objective / qp_status: B x n_steps x n_ctrl (may be NULL).  Host pointers.
 * The plant integrator (Dormand-Prince with odeint's step control, simulation_system.h:66-133) ends
 * a sampling interval after 4000 accepted steps: a scenario whose plant state runs away under an
 * unphysical input then yields non-finite records and failed QPs instead of a call that never
 * returns (the reference's loop is unbounded there). */
int cmpc_run_closed_loop(cmpc_handle* h, int n_steps, const double* x0, int n_blocks,
                         const int32_t* block_end, const double* block_off, double* traj,
                         uint32_t* qp_active, double* qp_objective, int32_t* qp_status);
/* The same run with the reference's timing window around the control step of every record (the
 * setup key n-timing-iterations; what the *-with-timing programs write as the last value of a
 * record): step_ns n_steps values, one per record for the whole batch (CUDA events; the plant side
 * of the loop is not part of it, as in the reference).  step_ns NULL: cmpc_run_closed_loop. */
int cmpc_run_closed_loop_timed(cmpc_handle* h, int n_steps, const double* x0, int n_blocks,
                               const int32_t* block_end, const double* block_off, double* traj,
                               uint32_t* qp_active, double* qp_objective, int32_t* qp_status,
                               int n_timing_iterations, int64_t* step_ns);
/* Device-resident variant: same arrays in device memory, runs on `stream`, no sync.
 * Runs records [first_step, first_step + n_steps) of a run whose arrays hold total_steps
 * records per scenario; first_step == 0 (re)starts the scenarios from x0_dev, later calls
 * continue from the state the handle holds (CMPC_ERR_STATE if there is no closed loop to
 * continue: none was started, or cmpc_initialize has restarted the controller since). */
int cmpc_run_closed_loop_device(cmpc_handle* h, int first_step, int n_steps, int total_steps,
                                const double* x0_dev, int n_blocks, const int32_t* block_end_dev,
                                const double* block_off_dev, double* traj_dev,
                                uint32_t* qp_active_dev, double* qp_objective_dev,
                                int32_t* qp_status_dev, void* stream);
/* The same closed loop one record per call with HOST buffers: the body of the reference driver's
 * loop (SimulationSystem::SetOffset -> ControllerInterface::GetNextInput ->
 * SimulationSystem::SetInput -> Integrate, simulation_system.h:66-116).  cmpc_closed_loop_start
 * places scenario b at x0[b] (B x n_states) with the controller initialised like the reference
 * driver; every cmpc_closed_loop_step uploads this sample's plant-input offsets (B x n_inputs,
 * added to the default input), runs the control step and the plant advance on the device and
 * downloads the record [t, x, u, y] (B x (1+n_states+8)).  Blocking; runs on the handle's stream.
 * Buffers that are page-locked and mapped into the device (cudaHostAlloc / cudaHostRegister, pinned torch
 * tensors; the record 16-byte aligned) are read and written by the plant kernel itself, in contiguous
 * chunks, with no copy on the stream; any other host memory goes through two copies.
 * cmpc_closed_loop_pipeline(h, 1), called before the first step of a run, lets every step launch the control
 * step of the NEXT record behind its plant advance (a control step needs the measurement the plant advance
 * produces, not the next plant-input offsets), so that it runs while the caller consumes the record and
 * prepares the next call; records are identical either way.  While pipelined, the read-back hooks
 * (cmpc_get_step_info, cmpc_get_qp, ...) show the record that has been launched ahead. */
int cmpc_closed_loop_start(cmpc_handle* h, const double* x0);
int cmpc_closed_loop_pipeline(cmpc_handle* h, int on);
int cmpc_closed_loop_step(cmpc_handle* h, const double* plant_offset, double* record);
/* Per-kernel device timing: when on, every control step (linearise [host-facing step only],
 * assemble, solve) is bracketed by CUDA events on its stream (which also keeps its kernels from
 * overlapping by dependent launch: use it to explain a run, not to time one); cmpc_get_timing synchronises, returns the
 * number of timed control steps, the summed duration of whole control steps and of their
 * assemble kernels alone, and resets the counters. */
int cmpc_set_timing(cmpc_handle* h, int on);
int cmpc_get_timing(cmpc_handle* h, int64_t* n_steps, double* step_ms, double* assemble_ms);
/* Developer aid: per-scenario stamps of the last closed-loop record (B x 32 int64; only filled by builds
 * with -DCMPC_PHASE_TIMING, zeros otherwise): slots 0-15 clock64() phase stamps inside the kernels, slots
 * 16-31 %globaltimer (ns) stamps of the kernels' starts and ends (tools/ticks.py names them). */
int cmpc_debug_phase_ticks(cmpc_handle* h, long long* out);
/* Number of kernel launches issued by this handle so far (for bench accounting). */
int cmpc_launch_count(cmpc_handle* h, int64_t* n_launches);

/* ---- fine-grained parity hooks (same names as the reference methods) ---------------- */
/* Keep the last step's linearisation and impulse-response table in HBM so that
 * cmpc_get_linearization / cmpc_generate_prediction can read them (off by default: the
 * production step never writes them). */
int cmpc_set_capture(cmpc_handle* h, int on);
/* AugmentedLinearizedSystem::Update result of the last step (aug_lin_sys.cc:145-177):
 * Aorig B x n x n, Bd B x n x 4 (discretised B, columns in this controller's input order;
 * the reference splits them into Borig/Adelay), f B x n.  Host pointers, any may be NULL. */
int cmpc_get_linearization(cmpc_handle* h, int ctrl, double* Aorig, double* Bd, double* f);
/* MpcQpSolver::GenerateQP result of the last step (mpc_qp_solver.cc:19-40,
 * distributed_solver.h:83-94): H B x nv x nv, f B x nv (before ApplyOtherInput), and the
 * cross term Gx B x nv x nvo with f_it = f + Gx du_other (distributed_solver.h:109-115). */
int cmpc_get_qp(cmpc_handle* h, int ctrl, double* H, double* f, double* Gx);
/* AugmentedLinearizedSystem::GeneratePrediction (aug_lin_sys.cc:260-334), debug/parity only:
 * Su B x (p*ny) x nv and Su_other B x (p*ny) x nvo (row-major) rebuilt from the last step's
 * impulse-response table.  The production path never materialises Su/Sx/Sf. */
int cmpc_generate_prediction(cmpc_handle* h, int ctrl, double* Su, double* Su_other);
/* Controller state: x_hat B x n, dx_aug B x n_total, y_old B x 4, u_old B x 4 (local order). */
int cmpc_get_controller_state(cmpc_handle* h, int ctrl, double* x_hat, double* dx_aug,
                              double* y_old, double* u_old);
/* MpcQpSolver::SolveQP (mpc_qp_solver.cc:45-75) on caller-supplied QPs, one per batch entry:
 * H nq x nv x nv, f nq x nv, lb/ub/lbA/ubA nq x nv, guess_io nq (working-set bitmask in/out,
 * 0xFFFFFFFF = cold start).  nv = 4 or 8, nu = nv/2.  Host pointers. */
int cmpc_solve_qp(int device, int nq, int nv, const double* H, const double* f, const double* lb,
                  const double* ub, const double* lbA, const double* ubA, uint32_t* guess_io,
                  double* z, uint32_t* active, double* objective, int32_t* status);
/* DynamicSystem::{GetDerivative,GetOutput,GetLinearizedSystem} on the device
 * (systems/ sources): x nq x n, u nq x n_inputs -> dxdt nq x n, y nq x 4, A nq x n x n,
 * Bc nq x n x 4, C nq x 4 x n.  Host pointers, outputs may be NULL. */
int cmpc_plant_eval(int device, int plant, int nq, const double* x, const double* u, double* dxdt,
                    double* y, double* A, double* Bc, double* C);
/* SimulationSystem::Integrate over one sampling interval (simulation_system.h:108-116):
 * x nq x n in/out, u nq x n_inputs (already offset + delayed).  n_substeps nq (may be NULL). */
int cmpc_plant_integrate(int device, int plant, int nq, double* x, const double* u, double Ts,
                         int32_t* n_substeps);
/* The plant integrator's straight-line square root and division (sqrt_inrange / div_inrange in
 * csrc/plant_dev.cuh: the fast path of the IEEE routines without their branch to the special cases)
 * next to the standard operations, element by element: sqrt(a[i]) and a[i] / b[i] both ways, and
 * flagged[i] = 1 if the square root, 2 if the division (3: both) reported an operand outside its range --
 * the cases in which the integrator redoes its derivative with the standard operations. */
int cmpc_inrange_math(int device, int n, const double* a, const double* b, double* sqrt_fast, double* sqrt_std,
                      double* div_fast, double* div_std, int32_t* flagged);

/* ---- measurement ------------------------------------------------------------------- */
typedef struct cmpc_fp64_peak {
  double dfma_tflops;         /* vector FP64 fused multiply-add */
  double dmma_m8n8k4_tflops;  /* mma.sync m8n8k4 f64 */
  double dmma_m16n8k8_tflops; /* mma.sync m16n8k8 f64 */
  int32_t sm_count;
} cmpc_fp64_peak;
/* FP64 pipe peak measured on `device` (roofline denominator). */
int cmpc_measure_fp64_peak(int device, cmpc_fp64_peak* out);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_H */
