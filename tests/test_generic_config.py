"""configurations outside the reference's own instantiations (run-time replacement of
constexpr_array.h / *_constants.h / the NerveCenter sub-controller pack) on the GPU against the
generalised oracle: other per-input delays, move horizons, unequal output partitions, three
sub-controllers.  Same tolerances as the tuned path (applied inputs and objective 1e-6 relative,
identical active sets)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import CASES

RTOL_U, ATOL_U = 1e-6, 1e-9


def rel_err(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


def build(pkg, conf, s, ywts, cons, batch):
    nc = pkg.NerveCenter.from_configuration(conf, batch=batch)
    nc.SetWeights(s.uwt, ywts)
    nc.SetOutputReference(np.asarray(s.yref, dtype=np.float64))
    for c, (lo, up, rlo, rup) in enumerate(cons):
        nc.SetConstraints(c, pkg.InputConstraints(lo, up, rlo, rup))
    return nc


def general_cases(pkg, setups):
    sp, ss = setups["coop-par"], setups["coop-ser"]
    SC, CF = pkg.SubController, pkg.Configuration
    two = lambda s: [(s.lower, s.upper, s.rate_lower, s.rate_upper)] * 2
    w2 = np.diag([1.0, 420.0])
    one = lambda s, i: (s.lower[i:i + 1], s.upper[i:i + 1], s.rate_lower[i:i + 1], s.rate_upper[i:i + 1])
    return {
        "delays-0-20-0-60": (sp, CF(0, [SC(2, [0, 1, 3]), SC(2, [0, 1, 3])], delays=(0, 20, 0, 60)), sp.ywt, two(sp)),
        "delays-on-torque": (ss, CF(1, [SC(2, [0, 1, 2, 3]), SC(2, [0, 1, 2, 3])], delays=(5, 0, 0, 33), p=60), ss.ywt, two(ss)),
        "m3": (sp, CF(0, [SC(2, [0, 1, 3]), SC(2, [0, 1, 3])], m=3), sp.ywt, two(sp)),
        "m1-p30": (sp, CF(0, [SC(2, [0, 1, 3]), SC(2, [0, 1, 3])], m=1, p=30, n_iterations=4), sp.ywt, two(sp)),
        "partition-013-13": (sp, CF(0, [SC(2, [0, 1, 3]), SC(2, [1, 3])]), [sp.ywt[0], w2], two(sp)),
        "three-controllers": (sp, CF(0, [SC(1, [0, 3]), SC(1, [0, 3]), SC(2, [1, 3])], n_iterations=5), [w2, w2, w2],
                              [one(sp, 0), one(sp, 1), two(sp)[0]]),
        "four-controllers-ser": (ss, CF(1, [SC(1, [0, 1]), SC(1, [1]), SC(1, [2, 3]), SC(1, [3, 0, 1])], p=50, n_iterations=3,
                                        delays=(0, 12, 0, 40)),
                                 [np.eye(2), np.eye(1), np.eye(2), np.eye(3)], [one(ss, 0), one(ss, 1), one(ss, 0), one(ss, 1)]),
        "centralised-m2-delays": (setups["cent-par"], CF(0, [SC(4, [0, 1, 3])], n_iterations=1, delays=(0, 25, 3, 40)),
                                  setups["cent-par"].ywt, [(setups["cent-par"].lower, setups["cent-par"].upper,
                                                            setups["cent-par"].rate_lower, setups["cent-par"].rate_upper)]),
    }


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["delays-0-20-0-60", "delays-on-torque", "m3", "m1-p30", "partition-013-13",
                                  "three-controllers", "four-controllers-ser", "centralised-m2-delays"])
def test_general_configuration_matches_oracle(name, setups, pkg, gpu_lib):
    s, conf, ywts, cons = general_cases(pkg, setups)[name]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 3, 160
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = [35, 50, 65]
    nc = build(pkg, conf, s, ywts, cons, B)
    g = nc.run_closed_loop(x0, be, bo, T)
    o = ol.Oracle.from_configuration(conf, s.uwt, ywts, cons, s.yref).run_closed_loop(x0, be, bo, T, n_threads=3)
    n = len(x_def)
    assert (g["status"] == 0).all() and (o["status"] == 0).all()
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.allclose(g["traj"][:, :, 1:1 + n], o["traj"][:, :, 1:1 + n], rtol=1e-8, atol=1e-11)
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U
    assert np.abs(g["traj"][:, -1, 1 + n:5 + n]).max() > 1e-3          # the controllers did act on the disturbance
    # controller state after a few host-facing steps: the reference's state order, delay chains included
    nc2 = build(pkg, conf, s, ywts, cons, B)
    y0 = np.stack([ol.plant_output(s.plant, x) for x in x0])
    _, u_def = ol.plant_defaults(s.plant)
    nc2.Initialize(x0, np.zeros(4), u_def, y0)
    for k in range(5):
        u = nc2.GetNextInput(g["traj"][:, k, 5 + n:])
        assert np.array_equal(u, g["traj"][:, k, 1 + n:5 + n])
    x, dx, yo, uo = nc2.controller_state(0)
    assert dx.shape == (B, n + 4 + sum(conf.delays)) and np.isfinite(dx).all()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_general_path_on_the_reference_shapes(case, setups, pkg, gpu_lib):
    """The reference's own six configurations sent down the general path (CMPC_FORCE_GENERIC=1):
    same records as the oracle, and as the tuned kernels to rounding."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 2, 90
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = [30, 45]
    tuned = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    os.environ["CMPC_FORCE_GENERIC"] = "1"
    try:
        nc = pkg.from_setup(s, batch=B)
    finally:
        del os.environ["CMPC_FORCE_GENERIC"]
    g = nc.run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=2)
    n = len(x_def)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"]) and np.array_equal(g["active"], tuned["active"])
    assert np.allclose(g["traj"], tuned["traj"], rtol=1e-7, atol=1e-10)
    with pytest.raises(pkg.capi.CmpcError, match="tuned path only"):
        nc.set_capture(True)


@pytest.mark.gpu
def test_refused_general_configurations(pkg, gpu_lib):
    SC, CF = pkg.SubController, pkg.Configuration
    two = [SC(2, [0, 1, 3]), SC(2, [0, 1, 3])]
    for conf, text in [(CF(0, two, delays=(0, 1, 0, 40)), "delays must be"),
                       (CF(0, [SC(4, [0, 1, 3])], m=3), "at most 8"),
                       (CF(0, [SC(2, [0, 1, 3]), SC(1, [0])]), "add up to the four"),
                       (CF(0, [SC(2, [0, 1]), SC(2, [1, 3], input_indices=[3, 2, 0, 1])]), "follow"),
                       (CF(0, two, m=5), "move horizon"),
                       (CF(0, two, m=3, p=300), "prediction horizon")]:
        with pytest.raises(pkg.capi.CmpcError, match=text):
            pkg.NerveCenter.from_configuration(conf, batch=2)


def test_generalised_oracle_reduces_to_the_reference_configuration(setups, pkg):
    """CPU: the run-time configuration of the oracle with the reference's own constants gives the
    records of the fixed configuration bit for bit (so the golden runs pin the general code too)."""
    for case in ("coop-par", "ncoop-ser", "cent-ser"):
        s = setups[case]
        x_def, _ = ol.plant_defaults(s.plant)
        T = 80
        be = np.array([[30, T]], dtype=np.int32)
        ref = ol.Oracle(s).run_closed_loop(x_def, be, s.sim_offsets[None], T)
        conf = pkg.Configuration(s.plant, [pkg.SubController(s.n_sub_control_inputs, o) for o in s.controlled_outputs],
                                 n_iterations=s.n_iterations)
        cons = [(s.lower, s.upper, s.rate_lower, s.rate_upper)] * s.n_controllers
        r = ol.Oracle.from_configuration(conf, s.uwt, s.ywt, cons, s.yref).run_closed_loop(x_def, be, s.sim_offsets[None], T)
        assert np.array_equal(r["traj"], ref["traj"]) and np.array_equal(r["active"], ref["active"])


@pytest.mark.gpu
def test_random_general_configurations(pkg, gpu_lib):
    """25 random general configurations (tests/fuzz_parity.py run_general): plant x controller split x
    output partitions x delays x move and prediction horizons x sweep counts, against the oracle."""
    import fuzz_parity
    assert fuzz_parity.run_general(seed=11, n_cfg=25) == []


@pytest.mark.gpu
def test_general_path_through_the_other_entry_points(setups, pkg, gpu_lib):
    """The general path behind the streaming host API and the timing window (there the window is the
    whole step: one kernel does it)."""
    s, conf, ywts, cons = general_cases(pkg, setups)["three-controllers"]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 2, 40
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = [12, 20]
    ref = build(pkg, conf, s, ywts, cons, B).run_closed_loop(x0, be, bo, T)
    timed = build(pkg, conf, s, ywts, cons, B).run_closed_loop(x0, be, bo, T, n_timing_iterations=2)
    assert np.array_equal(timed["traj"], ref["traj"]) and (timed["step_ns"] > 0).all()
    nc = build(pkg, conf, s, ywts, cons, B)
    nc.closed_loop_start(x0)
    for k in range(T):
        off = np.stack([bo[b, min(int((k >= be[b]).sum()), be.shape[1] - 1)] for b in range(B)])
        assert np.array_equal(nc.closed_loop_step(off), ref["traj"][:, k]), k
    H, f, G = nc.qp(2, cross_term=False)
    assert H.shape == (B, 4, 4) and np.allclose(H, np.swapaxes(H, 1, 2)) and (np.linalg.eigvalsh(H) > 0).all()


@pytest.mark.gpu
def test_general_path_failed_qp_gives_zero_move(setups, pkg, gpu_lib):
    """mpc_qp_solver.cc:66-69 on the general path: a sub-controller whose QP cannot be solved (here:
    lower bound above upper bound) applies the zero move and reports a status, the others go on; same
    records as the oracle."""
    s, conf, ywts, cons = general_cases(pkg, setups)["three-controllers"]
    cons = list(cons)
    lo, up, rlo, rup = cons[1]
    cons[1] = (np.asarray(up) + 0.2, np.asarray(up), rlo, rup)          # empty interval for sub-controller 1
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 2, 50
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 10
    g = build(pkg, conf, s, ywts, cons, B).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle.from_configuration(conf, s.uwt, ywts, cons, s.yref).run_closed_loop(x0, be, bo, T, n_threads=2)
    n = len(x_def)
    assert (g["status"][:, :, 1] != 0).all() and (o["status"][:, :, 1] != 0).all()
    assert (g["status"][:, :, [0, 2]] == 0).all()
    assert (g["traj"][:, :, 1 + n + 1] == 0).all()                        # its input (system input 1) never moves
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
