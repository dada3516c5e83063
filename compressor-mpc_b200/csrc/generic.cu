// Launch sequences of the general configuration path (generic_kernels.cuh), behind the same
// ShapeOps table as the tuned shapes.
#include "generic_kernels.cuh"
#include "handle.cuh"

namespace cmpc {

namespace {

template <int PLANT>
int gen_setup(cmpc_handle* h) {
  const GenParams& P = h->GP;
  h->smem_bytes = sizeof(double) * size_t(gen_smem_doubles(P.n, P.p, P.dmax, P.n_ctrl));
  if (h->smem_bytes > 227 * 1024) return fail(CMPC_ERR_UNSUPPORTED, "configuration too large for on-chip tables");
  CU(cudaFuncSetAttribute(gen_step_kernel<PLANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return CMPC_OK;
}

template <int PLANT>
int gen_init(cmpc_handle* h, const double* x, const double* u, const double* uf, const double* y, cudaStream_t st) {
  const int B = h->cfg.batch;
  h->lin_ahead = false;
  gen_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(h->d_genp, h->GS, h->NIN, x, u, uf, y);
  h->launches++;
  CU(cudaGetLastError());
  return CMPC_OK;
}

template <int PLANT>
int gen_step(cmpc_handle* h, const double* y, double* u, cudaStream_t st) {
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
    CU(cudaEventRecord(ev[1], st));
  }
  gen_step_kernel<PLANT><<<B, kGenThreads, h->smem_bytes, st>>>(h->d_genp, h->GS, y, u);
  h->launches++;
  if (ev) {
    CU(cudaEventRecord(ev[2], st));
    CU(cudaEventRecord(ev[3], st));
  }
  if (h->window_on) {   // one kernel does the whole step: the timing window is the whole step
    CU(cudaEventRecord(h->win_ev[h->win_used + 1], st));
    CU(cudaEventRecord(h->win_ev[h->win_used + 2], st));
  }
  CU(cudaGetLastError());
  return CMPC_OK;
}

template <int PLANT>
int gen_closed_loop(cmpc_handle* h, int first_step, int n_steps, const double* x0, ClosedLoopArrays A, bool reinit,
                    cudaStream_t st) {
  const int B = h->cfg.batch;
  if (reinit) {
    gen_start_kernel<PLANT><<<(B + 63) / 64, 64, 0, st>>>(h->d_genp, x0, A, h->d_uinit, h->d_uinitfull);
    h->launches++;
    CU(cudaGetLastError());
    if (int rc = gen_init<PLANT>(h, A.x, h->d_uinit, h->d_uinitfull, A.y, st)) return rc;
    h->initialized = true;
    h->loop_started = true;
  }
  double t = 0.0;
  for (int k = 0; k < first_step; ++k) t += h->cfg.Ts;
  for (int k = first_step; k < first_step + n_steps; ++k) {
    if (h->window_on) CU(cudaEventRecord(h->win_ev[h->win_used], st));
    if (int rc = gen_step<PLANT>(h, A.y, A.u, st)) return rc;
    if (h->window_on) {
      CU(cudaEventRecord(h->win_ev[h->win_used + 3], st));
      h->win_used += 4;
    }
    gen_advance_kernel<PLANT><<<(B + 63) / 64, 64, 0, st>>>(h->d_genp, h->GS, k, t, A);
    h->launches++;
    t += h->cfg.Ts;
  }
  CU(cudaGetLastError());
  return CMPC_OK;
}

}  // namespace

extern const ShapeOps kOps_generic_par = {0, 0, 0, 0, &gen_setup<0>, &gen_init<0>, &gen_step<0>, &gen_closed_loop<0>};
extern const ShapeOps kOps_generic_ser = {1, 0, 0, 0, &gen_setup<1>, &gen_init<1>, &gen_step<1>, &gen_closed_loop<1>};

}  // namespace cmpc
