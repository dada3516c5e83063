"""Host-side mirror of the reference's controller interface, batched over B scenarios.

Same names, call order and argument meaning as the reference
(ctor -> SetWeights -> SetOutputReference -> Initialize -> GetNextInput*):
  ControllerInterface::GetNextInput          include/controller_interface.h:46
  NerveCenter::{SetWeights,SetOutputReference,Initialize,GetNextInput}
                                             include/nerve_center.h:89-182
  InputConstraints                           include/input_constraints.h:11-27
Every method forwards to the C ABI (include/cmpc.h); all numerics run on the GPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import check, f64, lib, ptr


@dataclasses.dataclass
class InputConstraints:
    lower_bound: np.ndarray
    upper_bound: np.ndarray
    lower_rate_bound: np.ndarray
    upper_rate_bound: np.ndarray
    use_rate_constraints: bool = False  # never read by the reference either (mpc_qp_solver.cc:57-64)


@dataclasses.dataclass
class SubController:
    """One sub-controller of a general configuration: the template arguments of the reference's
    DistributedController / AugmentedLinearizedSystem (n_sub_control_inputs, ControlledOutputIndices,
    ControlInputIndices) as run-time values.  input_indices None: the own inputs are the system
    inputs that follow those of the sub-controllers before it, the others follow in system order."""
    n_inputs: int
    controlled_outputs: Sequence[int]
    input_indices: Optional[Sequence[int]] = None


@dataclasses.dataclass
class Configuration:
    """Run-time form of include/{parallel,serial}_compressors_constants.h (cmpc_config)."""
    plant: int
    controllers: Sequence[SubController]
    p: int = 100
    m: int = 2
    delays: Sequence[int] = (0, 40, 0, 40)
    n_iterations: int = 9
    Ts: float = 0.05

    def input_permutations(self):
        out, first = [], 0
        for sc in self.controllers:
            if sc.input_indices is not None:
                out.append(list(sc.input_indices))
            else:
                own = list(range(first, first + sc.n_inputs))
                out.append(own + [i for i in range(4) if i not in own])
            first += sc.n_inputs
        return out

    def to_cfg(self, batch: int) -> "capi.Config":
        cfg = capi.Config()
        cfg.plant, cfg.mode, cfg.p, cfg.m, cfg.Ts = self.plant, -1, self.p, self.m, self.Ts
        cfg.n_iterations, cfg.batch, cfg.n_disturbance_states = self.n_iterations, batch, 4
        cfg.n_controllers = len(self.controllers)
        cfg.n_sub_control_inputs = self.controllers[0].n_inputs
        for i in range(4):
            cfg.delays[i] = int(self.delays[i])
        for c, (sc, perm) in enumerate(zip(self.controllers, self.input_permutations())):
            cfg.n_sub_control_inputs_per[c] = sc.n_inputs
            cfg.n_controlled_outputs[c] = len(sc.controlled_outputs)
            for i, o in enumerate(sc.controlled_outputs):
                cfg.controlled_output_indices[c][i] = int(o)
            for i in range(4):
                cfg.control_input_indices[c][i] = int(perm[i])
        return cfg


class NerveCenter:
    """Batched NerveCenter: one cooperative / non-cooperative / centralised MPC per scenario."""

    def __init__(self, plant: int, mode: int, batch: int = 1, p: int = 100, n_solver_iterations: Optional[int] = None,
                 device: int = 0, constraints: Optional[InputConstraints] = None, cfg=None):
        if cfg is None:
            cfg = capi.default_config(plant, mode, batch)
            cfg.p = p
            if n_solver_iterations is not None:
                cfg.n_iterations = n_solver_iterations
        self.cfg = cfg
        self.plant, self.mode, self.batch, self.p, self.m = cfg.plant, mode, batch, cfg.p, cfg.m
        self.n_controllers = cfg.n_controllers
        self.n_sub_control_inputs = cfg.n_sub_control_inputs
        self.n_sub_inputs = [cfg.n_sub_control_inputs_per[c] or cfg.n_sub_control_inputs for c in range(cfg.n_controllers)]
        self.n_controlled_outputs = [cfg.n_controlled_outputs[c] for c in range(cfg.n_controllers)]
        self.controlled_output_indices = [list(cfg.controlled_output_indices[c])[: self.n_controlled_outputs[c]]
                                          for c in range(cfg.n_controllers)]
        self.control_input_indices = [list(cfg.control_input_indices[c]) for c in range(cfg.n_controllers)]
        self.nv = cfg.m * self.n_sub_control_inputs
        self.nvo = cfg.m * (4 - self.n_sub_control_inputs)
        self.n_delay_states = sum(cfg.delays)
        n, nin = C.c_int(), C.c_int()
        check(lib().cmpc_plant_dims(cfg.plant, C.byref(n), C.byref(nin)))
        self.n_states, self.n_inputs = n.value, nin.value
        self._h = C.c_void_p()
        check(lib().cmpc_create(C.byref(cfg), device, C.byref(self._h)))
        if constraints is not None:
            for c in range(self.n_controllers):
                self.SetConstraints(c, constraints)

    @classmethod
    def from_configuration(cls, conf: Configuration, batch: int = 1, device: int = 0) -> "NerveCenter":
        """A configuration outside the reference's own instantiations (other delays, move horizon,
        output partitions, up to four sub-controllers)."""
        return cls(conf.plant, -1, batch=batch, device=device, cfg=conf.to_cfg(batch))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().cmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup (nerve_center.h:98-122) -------------------------------------------------
    def SetWeights(self, uwt: np.ndarray, ywts: Sequence[np.ndarray]):
        """uwt: full 4x4 input weight; ywts: one n_y x n_y matrix per sub-controller
        (tuple overload, nerve_center.h:113-116; sub-matrix selection :225-234)."""
        uwt = f64(uwt)
        for c in range(self.n_controllers):
            idx = self.control_input_indices[c][: self.n_sub_inputs[c]]
            sub = f64(uwt[np.ix_(idx, idx)])
            check(lib().cmpc_set_weights(self._h, c, ptr(sub), ptr(f64(ywts[c]))))

    def SetOutputReference(self, y_ref: np.ndarray):
        """y_ref: (p, 4) reference for all plant outputs over the horizon (or (4,), replicated)."""
        y_ref = f64(y_ref)
        if y_ref.ndim == 1:
            y_ref = f64(np.tile(y_ref, (self.p, 1)))
        assert y_ref.shape == (self.p, 4)
        check(lib().cmpc_set_output_reference(self._h, ptr(y_ref)))

    def SetConstraints(self, ctrl: int, ic: InputConstraints):
        check(lib().cmpc_set_constraints(self._h, ctrl, ptr(f64(ic.lower_bound)), ptr(f64(ic.upper_bound)),
                                         ptr(f64(ic.lower_rate_bound)), ptr(f64(ic.upper_rate_bound))))

    def SetObserverGain(self, ctrl: int, M: np.ndarray):
        M = f64(M)
        assert M.shape == (self.n_states + 4, 4)
        check(lib().cmpc_set_observer_gain(self._h, ctrl, ptr(M)))

    def Initialize(self, x_init, u_init, u_init_full, y_init):
        B = self.batch
        bc = lambda a, k: f64(np.broadcast_to(f64(a), (B, k)))
        check(lib().cmpc_initialize(self._h, ptr(bc(x_init, self.n_states)), ptr(bc(u_init, 4)),
                                    ptr(bc(u_init_full, self.n_inputs)), ptr(bc(y_init, 4))))

    # ---- the hot path ---------------------------------------------------------------------
    def GetNextInput(self, y) -> np.ndarray:
        """y: (B, 4) host array -> u: (B, 4) (relative to the input offset), as the reference returns."""
        y = f64(np.broadcast_to(f64(y), (self.batch, 4)))
        u = np.empty((self.batch, 4))
        check(lib().cmpc_get_next_input(self._h, ptr(y), ptr(u)))
        return u

    def GetNextInputWithTiming(self, y, n_timing_iterations: int = -1):
        """NerveCenter::GetNextInputWithTiming (nerve_center.h:134-182): returns (u, time_ns), the time
        covering QP generation, the first n_timing_iterations sweeps and what follows the sweeps."""
        y = f64(np.broadcast_to(f64(y), (self.batch, 4)))
        u = np.empty((self.batch, 4))
        ns = C.c_int64()
        check(lib().cmpc_get_next_input_timed(self._h, ptr(y), ptr(u), int(n_timing_iterations), C.byref(ns)))
        return u, ns.value

    def GetNextInputRaw(self, y_host_ptr: int, u_host_ptr: int):
        """Same call on raw host addresses (e.g. pinned torch tensors): B x 4 doubles each."""
        check(lib().cmpc_get_next_input(self._h, C.c_void_p(y_host_ptr), C.c_void_p(u_host_ptr)))

    def GetNextInputDevice(self, y_dev_ptr: int, u_dev_ptr: int, stream: int = 0):
        check(lib().cmpc_get_next_input_device(self._h, C.c_void_p(y_dev_ptr), C.c_void_p(u_dev_ptr),
                                               C.c_void_p(stream)))

    def step_info(self):
        n = self.batch * self.n_controllers
        st = np.zeros(n, dtype=np.int32); act = np.zeros(n, dtype=np.uint32); obj = np.zeros(n)
        check(lib().cmpc_get_step_info(self._h, ptr(st), ptr(act), ptr(obj)))
        shp = (self.batch, self.n_controllers)
        return dict(status=st.reshape(shp), active=act.reshape(shp), objective=obj.reshape(shp))

    # ---- closed loop --------------------------------------------------------------------
    def run_closed_loop(self, x0, block_end, block_off, n_steps, want_traj=True, want_qp=True,
                        n_timing_iterations=None):
        """n_timing_iterations given: the result also holds step_ns (n_steps,), the reference's timing
        window around the control step of every record (cmpc_run_closed_loop_timed)."""
        B = self.batch
        x0 = f64(np.broadcast_to(f64(x0), (B, self.n_states)))
        block_end = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(block_end), (B, np.atleast_2d(block_end).shape[1])), dtype=np.int32)
        nb = block_end.shape[1]
        block_off = f64(np.broadcast_to(f64(block_off).reshape(-1, nb, self.n_inputs), (B, nb, self.n_inputs)))
        rec = 1 + self.n_states + 8
        traj = np.zeros((B, n_steps, rec)) if want_traj else None
        act = np.zeros((B, n_steps, self.n_controllers), dtype=np.uint32) if want_qp else None
        obj = np.zeros((B, n_steps, self.n_controllers)) if want_qp else None
        st = np.zeros((B, n_steps, self.n_controllers), dtype=np.int32) if want_qp else None
        if n_timing_iterations is None:
            check(lib().cmpc_run_closed_loop(self._h, n_steps, ptr(x0), nb, ptr(block_end), ptr(block_off),
                                             ptr(traj), ptr(act), ptr(obj), ptr(st)))
            return dict(traj=traj, active=act, objective=obj, status=st)
        ns = np.zeros(n_steps, dtype=np.int64)
        check(lib().cmpc_run_closed_loop_timed(self._h, n_steps, ptr(x0), nb, ptr(block_end), ptr(block_off),
                                               ptr(traj), ptr(act), ptr(obj), ptr(st), int(n_timing_iterations), ptr(ns)))
        return dict(traj=traj, active=act, objective=obj, status=st, step_ns=ns)

    def run_closed_loop_device(self, first_step, n_steps, total_steps, x0_ptr, n_blocks, block_end_ptr,
                               block_off_ptr, traj_ptr=0, act_ptr=0, obj_ptr=0, st_ptr=0, stream=0):
        """Device-resident closed loop (raw device addresses, e.g. torch.Tensor.data_ptr())."""
        vp = lambda v: C.c_void_p(v) if v else None
        check(lib().cmpc_run_closed_loop_device(self._h, first_step, n_steps, total_steps, vp(x0_ptr), n_blocks,
                                                vp(block_end_ptr), vp(block_off_ptr), vp(traj_ptr),
                                                vp(act_ptr), vp(obj_ptr), vp(st_ptr), vp(stream)))

    def closed_loop_start(self, x0):
        """Streaming closed loop with host I/O: place the scenarios at x0 (B, n_states)."""
        x0 = f64(np.broadcast_to(f64(x0), (self.batch, self.n_states)))
        check(lib().cmpc_closed_loop_start(self._h, ptr(x0)))

    def closed_loop_pipeline(self, on: bool = True):
        """Before the first step of a streaming run: launch the control step of the next record behind each
        plant advance, so that it runs while the caller handles the record (cmpc_closed_loop_pipeline)."""
        check(lib().cmpc_closed_loop_pipeline(self._h, 1 if on else 0))

    def closed_loop_step(self, plant_offset) -> np.ndarray:
        """One sample: plant-input offsets (B, n_inputs) in, record [t, x, u, y] (B, 1+n+8) out."""
        off = f64(np.broadcast_to(f64(plant_offset), (self.batch, self.n_inputs)))
        rec = np.empty((self.batch, 1 + self.n_states + 8))
        check(lib().cmpc_closed_loop_step(self._h, ptr(off), ptr(rec)))
        return rec

    def closed_loop_step_raw(self, offset_host_ptr: int, record_host_ptr: int):
        """Same on raw host addresses (pinned buffers)."""
        check(lib().cmpc_closed_loop_step(self._h, C.c_void_p(offset_host_ptr), C.c_void_p(record_host_ptr)))

    def set_timing(self, on: bool = True):
        check(lib().cmpc_set_timing(self._h, int(on)))

    def get_timing(self):
        n = C.c_int64(); ms = C.c_double(); ams = C.c_double()
        check(lib().cmpc_get_timing(self._h, C.byref(n), C.byref(ms), C.byref(ams)))
        return n.value, ms.value, ams.value

    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib().cmpc_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- parity hooks ---------------------------------------------------------------------
    def set_capture(self, on: bool = True):
        check(lib().cmpc_set_capture(self._h, int(on)))

    def linearization(self, ctrl: int):
        B, n = self.batch, self.n_states
        A = np.zeros((B, n, n)); Bd = np.zeros((B, n, 4)); f = np.zeros((B, n))
        check(lib().cmpc_get_linearization(self._h, ctrl, ptr(A), ptr(Bd), ptr(f)))
        return A, Bd, f

    def qp(self, ctrl: int, cross_term: bool = True):
        B = self.batch
        nv = self.cfg.m * self.n_sub_inputs[ctrl]
        nvo = self.cfg.m * (4 - self.n_sub_inputs[ctrl])
        H = np.zeros((B, nv, nv)); f = np.zeros((B, nv)); G = np.zeros((B, nv, max(nvo, 1)))
        check(lib().cmpc_get_qp(self._h, ctrl, ptr(H), ptr(f), ptr(G) if cross_term else None))
        return H, f, (G if nvo and cross_term else None)

    def prediction(self, ctrl: int):
        B, rows = self.batch, self.p * self.n_controlled_outputs[ctrl]
        Su = np.zeros((B, rows, self.nv)); Suo = np.zeros((B, rows, max(self.nvo, 1)))
        check(lib().cmpc_generate_prediction(self._h, ctrl, ptr(Su), ptr(Suo)))
        return Su, (Suo if self.nvo else None)

    def controller_state(self, ctrl: int):
        B, n = self.batch, self.n_states
        x = np.zeros((B, n)); dx = np.zeros((B, n + 4 + self.n_delay_states)); yo = np.zeros((B, 4)); uo = np.zeros((B, 4))
        check(lib().cmpc_get_controller_state(self._h, ctrl, ptr(x), ptr(dx), ptr(yo), ptr(uo)))
        return x, dx, yo, uo


def from_setup(setup, batch: int = 1, p: int = 100, device: int = 0, n_solver_iterations=None) -> NerveCenter:
    """Build a NerveCenter the way the reference driver does from a setup file."""
    ic = InputConstraints(setup.lower, setup.upper, setup.rate_lower, setup.rate_upper)
    nc = NerveCenter(setup.plant, setup.mode, batch=batch, p=p, device=device, constraints=ic,
                     n_solver_iterations=n_solver_iterations if n_solver_iterations is not None else setup.n_iterations)
    nc.SetWeights(setup.uwt, setup.ywt)
    nc.SetOutputReference(np.asarray(setup.yref, dtype=np.float64))
    return nc
