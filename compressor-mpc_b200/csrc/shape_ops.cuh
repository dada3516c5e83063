// Launch sequences of the control step for one controller shape; instantiated once per shape by
// the shape_*.cu translation units through CMPC_DEFINE_SHAPE_OPS.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "handle.cuh"

namespace cmpc {

// batches from this size on run the plant kernel's higher-occupancy instantiation (more than three
// waves of 16-scenario blocks at 2 blocks per SM)
constexpr int kAdvanceBigBatch = 16384;
#ifndef CMPC_ADV_BIG_MINB
#define CMPC_ADV_BIG_MINB 3
#endif

// Launch with programmatic stream serialisation (see pdl_wait / pdl_trigger in step_kernel.cuh).
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// The assemble_kernel instantiation for a prediction horizon: compile-time horizons for the
// reference's p = 100 and the 2x sweep, run-time horizon otherwise.
using AssembleFn = void (*)(StepParams, DeviceState, const double*);
template <class S>
AssembleFn assemble_variant(int p) {
  if (p == 100) return assemble_kernel<S, 2, 100>;
  if (p == 200) return assemble_kernel<S, 4, 200>;
  return p <= 2 * S::TPC ? assemble_kernel<S, 2, 0> : assemble_kernel<S, 4, 0>;
}

template <class S>
int shape_setup(cmpc_handle* h) {
  const SmemLayout<S> lay(h->P.p, h->P.b_max, h->P.n_pow, stage_tiles_for(h->P.p));
  h->smem_bytes = sizeof(double) * (size_t(S::NCTRL / assemble_ctas_per_scenario<S>(h->P.p)) * lay.total + 8);
#ifdef CMPC_PHASE_TIMING
  if (const char* e = getenv("CMPC_DEBUG_SMEM_MIN")) { size_t m = size_t(atol(e)); if (h->smem_bytes < m) h->smem_bytes = m; }  // occupancy experiments
#endif
  if (h->smem_bytes > 227 * 1024)
    return fail(CMPC_ERR_UNSUPPORTED, "prediction horizon too long for on-chip tables");
  const AssembleFn fn = assemble_variant<S>(h->P.p);
  // the limit belongs to the function, not to the handle: another handle with a longer horizon may
  // share this instantiation, so it is opened up to what the SM offers
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // ask for the largest shared-memory carveout: occupancy of this kernel is bounded by shared memory
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (getenv("CMPC_DEBUG")) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, assemble_block_threads<S>(h->P.p), h->smem_bytes);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, fn);
    fprintf(stderr, "[cmpc] assemble_kernel: smem %zu B/CTA, %d regs, %zu B local, occupancy %d CTAs/SM\n",
            h->smem_bytes, fa.numRegs, fa.localSizeBytes, nb);
  }
  return CMPC_OK;
}

template <class S>
int launch_init(cmpc_handle* h, const double* x, const double* u, const double* uf, const double* y,
                cudaStream_t st) {
  const int B = h->cfg.batch;
  h->P.ring_pos = 0;
  h->lin_ahead = false;
  init_kernel<S><<<(B + 127) / 128, 128, 0, st>>>(B, h->G, x, u, uf, y, h->P);
  h->launches++;
  CU(cudaGetLastError());
  return CMPC_OK;
}

template <class S>
int launch_step_impl(cmpc_handle* h, const double* y, double* u, cudaStream_t st, bool defer_apriori) {
  const int B = h->cfg.batch;
  cudaEvent_t* ev = nullptr;
  if (h->timing) {
    if (h->ev_used + 4 > h->ev.size()) {
      const size_t old = h->ev.size();
      h->ev.resize(old + 1024);
      for (size_t i = old; i < h->ev.size(); ++i) CU(cudaEventCreate(&h->ev[i]));
    }
    ev = &h->ev[h->ev_used];
    h->ev_used += 4;
    CU(cudaEventRecord(ev[0], st));
  }
  // K0: observer + linearisation, 4 threads per (scenario, controller).  Inside a closed-loop run
  // the previous record's plant kernel has already done this work (lin_ahead).
  if (!h->lin_ahead) {
    const int n_thr = B * S::NCTRL * 4;
    CU(launch_pdl(lin_kernel<S>, (n_thr + 127) / 128, 128, 0, st, h->P, h->G, const_cast<double*>(y), h->y_mapped_src));
    h->launches++;
  }
  if (ev) CU(cudaEventRecord(ev[1], st));
  // K1: discretisation, prediction, QP assembly; one CTA per scenario
  CU(launch_pdl(assemble_variant<S>(h->cfg.p), B * assemble_ctas_per_scenario<S>(h->P.p), assemble_block_threads<S>(h->P.p), h->smem_bytes, st, h->P, h->G, y));
  if (ev) CU(cudaEventRecord(ev[2], st));
  // K2: Jacobi sweeps + update; one lane pair per scenario
  const unsigned solve_grid = (B * S::NCTRL + 63) / 64;
  if (h->window_on && h->window_n >= 0 && h->window_n < h->P.n_iter) {
    // timing window (nerve_center.h:150-159): the sweeps from n_timing_iterations on are left out
    // of the measured time, what follows the sweeps is measured again
    cudaEvent_t* w = &h->win_ev[h->win_used];
    const int n_t = h->window_n;
    if (n_t > 0) {
      solve_sweeps_kernel<S><<<solve_grid, 64, 0, st>>>(h->P, h->G, 0, n_t);
      h->launches++;
    }
    CU(cudaEventRecord(w[1], st));
    solve_sweeps_kernel<S><<<solve_grid, 64, 0, st>>>(h->P, h->G, n_t, h->P.n_iter);
    CU(cudaEventRecord(w[2], st));
    solve_finish_kernel<S><<<solve_grid, 64, 0, st>>>(h->P, h->G, u);
    h->launches++;
  } else {
    // defer_apriori: the plant kernel of the closed loop does UpdateU / ObserveAPriori (lin_part_early)
    if (defer_apriori)
      CU(launch_pdl(solve_kernel<S, false>, solve_grid, 64, 0, st, h->P, h->G, u));
    else
      CU(launch_pdl(solve_kernel<S, true>, solve_grid, 64, 0, st, h->P, h->G, u));
    if (h->window_on) {   // the window is the whole step: the two inner events coincide
      CU(cudaEventRecord(h->win_ev[h->win_used + 1], st));
      CU(cudaEventRecord(h->win_ev[h->win_used + 2], st));
    }
  }
  h->P.ring_pos = (h->P.ring_pos + 1) % kRing;   // the oldest ring slot was consumed and refilled
  if (ev) CU(cudaEventRecord(ev[3], st));
  h->launches += 2;
  CU(cudaGetLastError());
  return CMPC_OK;
}

template <class S>
int launch_step(cmpc_handle* h, const double* y, double* u, cudaStream_t st) {
  return launch_step_impl<S>(h, y, u, st, false);
}

template <class S>
int launch_closed_loop(cmpc_handle* h, int first_step, int n_steps, const double* x0,
                       ClosedLoopArrays A, bool reinit, cudaStream_t st) {
  const int B = h->cfg.batch;
  if (reinit) {
    cl_start_kernel<S::PLANT><<<(B + 63) / 64, 64, 0, st>>>(B, x0, A, h->d_uinit, h->d_uinitfull);
    h->launches++;
    CU(cudaGetLastError());
    int rc = launch_init<S>(h, A.x, h->d_uinit, h->d_uinitfull, A.y, st);
    if (rc) return rc;
    h->initialized = true;
    h->loop_started = true;
  }
  double t = 0.0;
  for (int k = 0; k < first_step; ++k) t += h->cfg.Ts;  // the reference driver accumulates t += Ts
  for (int k = first_step; k < first_step + n_steps; ++k) {
    // with the reference's observer gain the plant kernel linearises early and takes the a-priori update
    // along (the timing window keeps the whole controller step inside its own launches)
    const bool k3_apriori = h->P.obs_states_free && !(h->window_on && h->window_n >= 0 && h->window_n < h->P.n_iter);
    if (A.phases & 1) {
      if (h->window_on) CU(cudaEventRecord(h->win_ev[h->win_used], st));
      int rc = launch_step_impl<S>(h, A.y, A.u, st, k3_apriori);
      if (rc) return rc;
      if (h->window_on) {
        CU(cudaEventRecord(h->win_ev[h->win_used + 3], st));
        h->win_used += 4;
      }
    }
    // plant side of record k, and the observer update + linearisation of record k + 1
    if (A.phases & 2) {
      if (B >= kAdvanceBigBatch)
        CU(launch_pdl(cl_advance_kernel<S, CMPC_ADV_BIG_MINB>, (B + 15) / 16, 128, 0, st, h->P, h->G, k, t, h->cfg.Ts, A, true, k3_apriori));
      else
        CU(launch_pdl(cl_advance_kernel<S, 1>, (B + 15) / 16, 128, 0, st, h->P, h->G, k, t, h->cfg.Ts, A, true, k3_apriori));
      h->lin_ahead = true;
      h->launches++;
    }
    t += h->cfg.Ts;
  }
  CU(cudaGetLastError());
  return CMPC_OK;
}

template <class S>
constexpr ShapeOps make_shape_ops() {
  return ShapeOps{S::PLANT, S::NY, S::NU, S::NCTRL, &shape_setup<S>, &launch_init<S>, &launch_step<S>,
                  &launch_closed_loop<S>};
}

}  // namespace cmpc

#define CMPC_DEFINE_SHAPE_OPS(name, ...) \
  namespace cmpc { extern const ShapeOps name = make_shape_ops<Shape<__VA_ARGS__>>(); }
