"""ctypes binding of the CPU oracle (oracle/libcmpc_oracle.so) — test infrastructure only."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
_LIB = None

dp = C.POINTER(C.c_double)


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _LIB
    if _LIB is None:
        so = ROOT / "oracle" / "libcmpc_oracle.so"
        if not so.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle")])
        _LIB = C.CDLL(str(so))
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_create.argtypes = [C.c_int] * 4
        _LIB.orc_create_config.restype = C.c_void_p
        for name in ("orc_destroy", "orc_set_weights", "orc_set_constraints", "orc_set_observer_gain",
                     "orc_set_output_reference", "orc_initialize", "orc_get_next_input",
                     "orc_get_linearization", "orc_get_prediction", "orc_get_qp",
                     "orc_get_ctrl_state", "orc_get_plan", "orc_last_qp_info",
                     "orc_run_closed_loop", "orc_plant_defaults", "orc_plant_derivative",
                     "orc_plant_output", "orc_plant_linearize", "orc_plant_discretize"):
            getattr(_LIB, name).restype = None
    return _LIB


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Oracle:
    """One reference-equivalent controller stack (NerveCenter + sub-controllers) on the CPU."""

    def __init__(self, setup, p: int = 100, n_iter: int | None = None):
        L = lib()
        self.setup = setup
        self.p = p
        self.h = C.c_void_p(L.orc_create(setup.plant, setup.mode, p,
                                         n_iter if n_iter is not None else setup.n_iterations))
        self.n = L.orc_n_states(self.h)
        self.n_in = L.orc_n_inputs(self.h)
        self.n_ctrl = L.orc_n_controllers(self.h)
        self.nu = setup.n_sub_control_inputs
        self.ny = [len(o) for o in setup.controlled_outputs]
        uwt = f64(setup.uwt)
        for c in range(self.n_ctrl):
            L.orc_set_weights(self.h, c, _p(uwt), _p(f64(setup.ywt[c])))
            L.orc_set_constraints(self.h, c, _p(f64(setup.lower)), _p(f64(setup.upper)),
                                  _p(f64(setup.rate_lower)), _p(f64(setup.rate_upper)))
        yref = f64(np.tile(np.asarray(setup.yref, dtype=np.float64), (p, 1)))
        L.orc_set_output_reference(self.h, _p(yref))

    @classmethod
    def from_configuration(cls, conf, uwt, ywts, constraints, yref):
        """A general configuration (compressor_mpc_b200.Configuration): uwt full 4 x 4, ywts one matrix
        per sub-controller, constraints one (lower, upper, rate_lower, rate_upper) per sub-controller,
        yref (4,) or (p, 4)."""
        L = lib()
        self = cls.__new__(cls)
        self.setup, self.p = None, conf.p
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        nc = len(conf.controllers)
        outs = np.zeros((nc, 4), dtype=np.int32)
        for c, sc in enumerate(conf.controllers):
            outs[c, :len(sc.controlled_outputs)] = sc.controlled_outputs
        self.h = C.c_void_p(L.orc_create_config(
            C.c_int(conf.plant), C.c_int(conf.p), C.c_int(conf.m), C.c_int(conf.n_iterations), _p(i32(conf.delays)),
            C.c_int(nc), _p(i32([sc.n_inputs for sc in conf.controllers])),
            _p(i32([len(sc.controlled_outputs) for sc in conf.controllers])), _p(outs),
            _p(i32(conf.input_permutations()))))
        self.n = L.orc_n_states(self.h)
        self.n_in = L.orc_n_inputs(self.h)
        self.n_ctrl = nc
        self.nu = conf.controllers[0].n_inputs
        self.ny = [len(sc.controlled_outputs) for sc in conf.controllers]
        uwt = f64(uwt)
        for c in range(nc):
            L.orc_set_weights(self.h, c, _p(uwt), _p(f64(ywts[c])))
            lo, up, rlo, rup = constraints[c]
            L.orc_set_constraints(self.h, c, _p(f64(lo)), _p(f64(up)), _p(f64(rlo)), _p(f64(rup)))
        yref = f64(yref)
        if yref.ndim == 1:
            yref = f64(np.tile(yref, (conf.p, 1)))
        L.orc_set_output_reference(self.h, _p(yref))
        return self

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def set_observer_gain(self, c, M):
        lib().orc_set_observer_gain(self.h, c, _p(f64(M)))

    def set_output_reference(self, yref):
        lib().orc_set_output_reference(self.h, _p(f64(yref)))

    def initialize(self, x0, u_init, u_full, y0):
        lib().orc_initialize(self.h, _p(f64(x0)), _p(f64(u_init)), _p(f64(u_full)), _p(f64(y0)))

    def get_next_input(self, y):
        u = np.zeros(4)
        lib().orc_get_next_input(self.h, _p(f64(y)), _p(u))
        return u

    def linearization(self, c):
        n, nd = self.n, 2
        A = np.zeros((n, n)); Bo = np.zeros((n, 4 - nd)); Ad = np.zeros((n, nd))
        Cm = np.zeros((4, n + 4)); f = np.zeros(n)
        lib().orc_get_linearization(self.h, c, _p(A), _p(Bo), _p(Ad), _p(Cm), _p(f))
        return A, Bo, Ad, Cm, f

    def prediction(self, c):
        rows = self.p * self.ny[c]
        nu, no = self.nu, 4 - self.nu
        Su = np.zeros((2 * nu, rows)); Sx = np.zeros((84, rows)); Sf = np.zeros((self.n, rows))
        Suo = np.zeros((2 * no, rows)) if no else None
        lib().orc_get_prediction(self.h, c, _p(Su), _p(Sx), _p(Sf), _p(Suo))
        return Su.T, Sx.T, Sf.T, (Suo.T if no else None)   # column-major -> (rows, cols)

    def qp(self, c):
        nv = 2 * self.nu
        H = np.zeros((nv, nv)); f = np.zeros(nv)
        lib().orc_get_qp(self.h, c, _p(H), _p(f))
        return H, f

    def ctrl_state(self, c):
        x = np.zeros(self.n); dx = np.zeros(self.n + 84); yo = np.zeros(4); uo = np.zeros(4)
        lib().orc_get_ctrl_state(self.h, c, _p(x), _p(dx), _p(yo), _p(uo))
        return x, dx, yo, uo

    def plan(self):
        du = np.zeros(2 * 4)
        lib().orc_get_plan(self.h, _p(du))
        return du

    def last_qp_info(self, c):
        st = C.c_int(); act = C.c_uint(); obj = C.c_double()
        lib().orc_last_qp_info(self.h, c, C.byref(st), C.byref(act), C.byref(obj))
        return st.value, act.value, obj.value

    def run_closed_loop(self, x0, block_end, block_off, n_steps, n_threads=1, want_ns=False):
        x0 = f64(np.atleast_2d(x0)); B = x0.shape[0]
        block_end = np.ascontiguousarray(np.atleast_2d(block_end), dtype=np.int32)
        block_off = f64(block_off).reshape(B, block_end.shape[1], self.n_in)
        rec = 1 + self.n + 8
        traj = np.zeros((B, n_steps, rec))
        act = np.zeros((B, n_steps, self.n_ctrl), dtype=np.uint32)
        obj = np.zeros((B, n_steps, self.n_ctrl))
        st = np.zeros((B, n_steps, self.n_ctrl), dtype=np.int32)
        ns = np.zeros((B, n_steps)) if want_ns else None
        lib().orc_run_closed_loop(self.h, B, n_steps, _p(x0), block_end.shape[1], _p(block_end),
                                  _p(block_off), _p(traj), _p(act), _p(obj), _p(st), _p(ns),
                                  n_threads)
        out = dict(traj=traj, active=act, objective=obj, status=st)
        if want_ns:
            out["step_ns"] = ns
        return out


def plant_defaults(plant):
    n = 11 if plant == 0 else 10
    x = np.zeros(n); u = np.zeros(n - 2)
    lib().orc_plant_defaults(plant, _p(x), _p(u))
    return x, u


def plant_derivative(plant, x, u):
    d = np.zeros(len(x))
    lib().orc_plant_derivative(plant, _p(f64(x)), _p(f64(u)), _p(d))
    return d


def plant_output(plant, x):
    y = np.zeros(4)
    lib().orc_plant_output(plant, _p(f64(x)), _p(y))
    return y


def plant_linearize(plant, x, u):
    n = len(x)
    A = np.zeros((n, n)); B = np.zeros((n, 4)); Cm = np.zeros((4, n)); f = np.zeros(n)
    lib().orc_plant_linearize(plant, _p(f64(x)), _p(f64(u)), _p(A), _p(B), _p(Cm), _p(f))
    return A, B, Cm, f


def plant_discretize(plant, x, u, Ts=0.05):
    n = len(x)
    A = np.zeros((n, n)); B = np.zeros((n, 4)); Cm = np.zeros((4, n)); f = np.zeros(n)
    lib().orc_plant_discretize(plant, _p(f64(x)), _p(f64(u)), C.c_double(Ts), _p(A), _p(B), _p(Cm), _p(f))
    return A, B, Cm, f


def plant_integrate(plant, x, u, t0=0.0, Ts=0.05):
    x = f64(x).copy()
    L = lib()
    L.orc_plant_integrate.restype = C.c_int
    steps = L.orc_plant_integrate(plant, _p(x), _p(f64(u)), C.c_double(t0), C.c_double(Ts))
    return x, steps


def solve_qp(H, f, lb, ub, lbA, ubA, nu, guess=None):
    nv = len(f)
    z = np.zeros(nv); act = C.c_uint(); obj = C.c_double(); it = C.c_int()
    g = C.c_uint(0xFFFFFFFF if guess is None else guess)
    L = lib()
    L.orc_solve_qp.restype = C.c_int
    st = L.orc_solve_qp(nv, nu, _p(f64(H)), _p(f64(f)), _p(f64(lb)), _p(f64(ub)), _p(f64(lbA)),
                        _p(f64(ubA)), C.byref(g), _p(z), C.byref(act), C.byref(obj), C.byref(it))
    return dict(status=st, z=z, active=act.value, objective=obj.value, iterations=it.value,
                working_set=g.value)
