"""ctypes binding of the C ABI in include/cmpc.h (libcmpc_b200.so).

This is the only place the package touches native code.  There is no CPU
implementation behind it: if the CUDA library is missing or no GPU is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("CMPC_B200_LIB", _PKG / "libcmpc_b200.so"))

CMPC_OK = 0
ERR_NAMES = {1: "CMPC_ERR_ARG", 2: "CMPC_ERR_CUDA", 3: "CMPC_ERR_UNSUPPORTED", 4: "CMPC_ERR_STATE"}

EXPORTED_SYMBOLS = [
    "cmpc_default_config", "cmpc_plant_dims", "cmpc_plant_defaults", "cmpc_create", "cmpc_destroy",
    "cmpc_last_error", "cmpc_set_weights", "cmpc_set_output_reference", "cmpc_set_constraints",
    "cmpc_set_observer_gain", "cmpc_initialize", "cmpc_get_next_input",
    "cmpc_get_next_input_device", "cmpc_get_next_input_timed", "cmpc_run_closed_loop_timed", "cmpc_get_step_info", "cmpc_run_closed_loop",
    "cmpc_run_closed_loop_device", "cmpc_closed_loop_start", "cmpc_closed_loop_step", "cmpc_closed_loop_pipeline", "cmpc_inrange_math", "cmpc_launch_count", "cmpc_set_capture", "cmpc_set_timing",
    "cmpc_get_timing", "cmpc_debug_phase_ticks",
    "cmpc_get_linearization", "cmpc_get_qp", "cmpc_generate_prediction",
    "cmpc_get_controller_state", "cmpc_solve_qp", "cmpc_plant_eval", "cmpc_plant_integrate",
    "cmpc_measure_fp64_peak",
]


class CmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


MAX_CONTROLLERS = 4


class Config(C.Structure):
    _fields_ = [
        ("plant", C.c_int32), ("mode", C.c_int32), ("p", C.c_int32), ("m", C.c_int32),
        ("Ts", C.c_double), ("n_iterations", C.c_int32), ("batch", C.c_int32),
        ("delays", C.c_int32 * 4), ("n_disturbance_states", C.c_int32),
        ("n_controllers", C.c_int32), ("n_sub_control_inputs", C.c_int32),
        ("n_controlled_outputs", C.c_int32 * MAX_CONTROLLERS),
        ("controlled_output_indices", (C.c_int32 * 4) * MAX_CONTROLLERS),
        ("control_input_indices", (C.c_int32 * 4) * MAX_CONTROLLERS),
        ("n_sub_control_inputs_per", C.c_int32 * MAX_CONTROLLERS),
    ]


class Fp64Peak(C.Structure):
    _fields_ = [("dfma_tflops", C.c_double), ("dmma_m8n8k4_tflops", C.c_double),
                ("dmma_m16n8k8_tflops", C.c_double), ("sm_count", C.c_int32)]


_lib = None


def lib():
    """Load libcmpc_b200.so (built by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  The control step has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH))
        L.cmpc_last_error.restype = C.c_char_p
        for name in EXPORTED_SYMBOLS:
            if name != "cmpc_last_error":
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != CMPC_OK:
        raise CmpcError(rc, lib().cmpc_last_error().decode())


def ptr(a):
    """Data pointer of a C-contiguous numpy array (None -> NULL), or a raw int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)  # keeps `a` alive for the duration of the call


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def default_config(plant: int, mode: int, batch: int) -> Config:
    cfg = Config()
    check(lib().cmpc_default_config(plant, mode, batch, C.byref(cfg)))
    return cfg


def plant_defaults(plant: int):
    n, nin = C.c_int(), C.c_int()
    check(lib().cmpc_plant_dims(plant, C.byref(n), C.byref(nin)))
    x = np.zeros(n.value)
    u = np.zeros(nin.value)
    check(lib().cmpc_plant_defaults(plant, ptr(x), ptr(u)))
    return x, u


def measure_fp64_peak(device: int = 0) -> dict:
    pk = Fp64Peak()
    check(lib().cmpc_measure_fp64_peak(device, C.byref(pk)))
    return dict(dfma_tflops=pk.dfma_tflops, dmma_m8n8k4_tflops=pk.dmma_m8n8k4_tflops,
                dmma_m16n8k8_tflops=pk.dmma_m16n8k8_tflops, sm_count=pk.sm_count)


def solve_qp(H, f, lb, ub, lbA, ubA, guess=None, device=0):
    """Batched MpcQpSolver::SolveQP on the GPU.  H (nq,nv,nv), others (nq,nv)."""
    H = f64(H); f = f64(f)
    nq, nv = f.shape
    g = np.full(nq, 0xFFFFFFFF, dtype=np.uint32) if guess is None else np.ascontiguousarray(guess, dtype=np.uint32).copy()
    z = np.zeros((nq, nv)); act = np.zeros(nq, dtype=np.uint32); obj = np.zeros(nq)
    st = np.zeros(nq, dtype=np.int32)
    check(lib().cmpc_solve_qp(device, nq, nv, ptr(H), ptr(f), ptr(f64(lb)), ptr(f64(ub)), ptr(f64(lbA)),
                              ptr(f64(ubA)), ptr(g), ptr(z), ptr(act), ptr(obj), ptr(st)))
    return dict(z=z, active=act, objective=obj, status=st, working_set=g)


def plant_eval(plant, x, u, device=0):
    x = f64(np.atleast_2d(x)); u = f64(np.atleast_2d(u))
    nq, n = x.shape
    out = dict(dxdt=np.zeros((nq, n)), y=np.zeros((nq, 4)), A=np.zeros((nq, n, n)),
               B=np.zeros((nq, n, 4)), C=np.zeros((nq, 4, n)))
    check(lib().cmpc_plant_eval(device, plant, nq, ptr(x), ptr(u), ptr(out["dxdt"]), ptr(out["y"]),
                                ptr(out["A"]), ptr(out["B"]), ptr(out["C"])))
    return out


def plant_integrate(plant, x, u, Ts=0.05, device=0):
    x = f64(np.atleast_2d(x)).copy(); u = f64(np.atleast_2d(u))
    ns = np.zeros(x.shape[0], dtype=np.int32)
    check(lib().cmpc_plant_integrate(device, plant, x.shape[0], ptr(x), ptr(u), C.c_double(Ts), ptr(ns)))
    return x, ns


def inrange_math(a, b, device=0):
    """sqrt(a) and a / b by the plant integrator's straight-line forms and by the standard operations."""
    a = f64(np.ravel(a)); b = f64(np.ravel(b))
    n = a.shape[0]
    out = [np.empty(n) for _ in range(4)]
    flagged = np.zeros(n, dtype=np.int32)
    check(lib().cmpc_inrange_math(device, n, ptr(a), ptr(b), *(ptr(o) for o in out), ptr(flagged)))
    return {"sqrt_fast": out[0], "sqrt_std": out[1], "div_fast": out[2], "div_std": out[3], "flagged": flagged}
