// The handle behind include/cmpc.h and the helpers shared by the translation units of the library
// (one per controller shape, so that they compile in parallel, plus the C ABI in cmpc.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "cmpc.h"
#include "generic_params.cuh"
#include "plant_kernels.cuh"
#include "step_kernel.cuh"

namespace cmpc {
struct ShapeOps;
}

namespace cmpc {

// sets the text behind cmpc_last_error() (thread-local, defined in cmpc.cu) and returns `code`
int fail(int code, const std::string& msg);

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(CMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
  return e;
}

// The calling thread's current device is put back when an entry point returns: a host that
// drives several GPUs from one thread (torch, the tests) must not find it changed.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = (prev == dev) ? cudaSuccess : cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Temporary device array of an entry point: freed on every return path.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  DevBuf() = default;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t alloc(size_t n) { return dalloc(&p, n); }
  cudaError_t upload(const T* host, size_t n) {
    cudaError_t e = alloc(n);
    if (e == cudaSuccess) e = cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    return e;
  }
  cudaError_t download(T* host, size_t n) const {
    return cudaMemcpy(host, p, n * sizeof(T), cudaMemcpyDeviceToHost);
  }
};

}  // namespace cmpc


struct cmpc_handle {
  cmpc_config cfg;
  int device = 0;
  int shape = -1;  // index into the instantiated shapes
  int N = 0, NIN = 0, NV = 0, NVO = 0, NCTRL = 0;
  cmpc::StepParams P;
  cmpc::DeviceState G;
  // general configuration path (generic_kernels.cuh): its parameter block lives in device memory
  bool generic = false;
  cmpc::GenParams GP;
  cmpc::GenState GS;
  cmpc::GenParams* d_genp = nullptr;
  const cmpc::ShapeOps* ops = nullptr;   // launch functions of this handle's path
  double* d_yref = nullptr;
  double *d_y = nullptr, *d_u = nullptr;  // [B][4]
  double *d_xinit = nullptr, *d_uinit = nullptr, *d_uinitfull = nullptr, *d_yinit = nullptr;
  // closed loop
  double *d_x = nullptr, *d_ring = nullptr;
  int* d_block_end = nullptr;
  double* d_block_off = nullptr;
  size_t block_cap = 0;
  // closed loop one record per call with host I/O (cmpc_closed_loop_start / _step)
  int* d_step_end = nullptr;
  double *d_step_off = nullptr, *d_rec = nullptr;
  const double* y_mapped_src = nullptr;   // cmpc_get_next_input: device address of a mapped host measurement buffer
                                           // that lin_kernel reads itself (null: y is already on the device)
  bool stream_pipeline = false;   // cmpc_closed_loop_pipeline: the control step of the next record is launched ahead
  bool ctrl_ahead = false;        // ... and has been for record stream_next
  cudaEvent_t ev_plant = nullptr; // end of the plant advance of the record being returned
  int stream_next = -1;
  size_t smem_bytes = 0;
  int64_t launches = 0;
  bool initialized = false;
  bool capture = false;
  bool timing = false;
  bool lin_ahead = false;   // closed loop: the next record's observer update + linearisation are already done
  bool loop_started = false;   // the on-device plants (state, delay rings, measurement) belong to a running closed loop
  // where the running closed loop stands: a continuation must pick up exactly here
  int loop_next = 0, loop_total = 0, loop_blocks = 0;
  const void *loop_block_end = nullptr, *loop_block_off = nullptr, *loop_traj = nullptr;
  cudaStream_t stream = nullptr;   // the handle's own (non-blocking) stream: host-pointer entry points run on it
  std::vector<cudaEvent_t> ev;   // pairs (start, stop) around control-step launches
  size_t ev_used = 0;
  // timing window of the reference's GetNextInputWithTiming (n-timing-iterations): four events per
  // control step [start, end of the timed sweeps, start of what follows the sweeps, end]; the
  // caller of launch_step records the outer two
  bool window_on = false;
  int window_n = -1;
  std::vector<cudaEvent_t> win_ev;
  size_t win_used = 0;
};

namespace cmpc {

// Launch functions of one controller shape (a Shape<> instantiation), filled in by its own
// translation unit (shape_*.cu).
struct ShapeOps {
  int plant, ny, nu, nctrl;
  int (*setup)(cmpc_handle*);
  int (*init)(cmpc_handle*, const double* x, const double* u, const double* uf, const double* y, cudaStream_t);
  int (*step)(cmpc_handle*, const double* y, double* u, cudaStream_t);
  int (*closed_loop)(cmpc_handle*, int first_step, int n_steps, const double* x0, ClosedLoopArrays A, bool reinit,
                     cudaStream_t);
};
constexpr int kNumShapes = 7;
extern const ShapeOps* const kShapeOps[kNumShapes];
extern const ShapeOps kOps_generic_par, kOps_generic_ser;

}  // namespace cmpc
