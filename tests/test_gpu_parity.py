"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): applied inputs and QP objective within 1e-6 relative
(floor 1e-9 absolute) over a closed-loop run, identical optimal active sets.  Intermediate
quantities (Jacobians, H, f) are held to much tighter bounds because both sides compute in FP64.
"""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import CASES

pytestmark = pytest.mark.gpu

RTOL_U = 1e-6
ATOL_U = 1e-9


def rel_err(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("plant", [0, 1])
def test_plant_eval_matches_oracle(plant, pkg, gpu_lib):
    x0, u0 = ol.plant_defaults(plant)
    rng = np.random.default_rng(1)
    nq = 64
    X = x0 * (1 + 0.02 * rng.uniform(-1, 1, (nq, len(x0))))
    U = u0 + 0.05 * rng.uniform(-1, 1, (nq, len(u0)))
    idx_rec = [3, 7]
    U[:, idx_rec] = np.abs(U[:, idx_rec])         # recycle valves >= 0
    U[: nq // 4, 3] = 0.005                        # inside the smoothed dead zone (compressor.cc:154-164)
    U[nq // 4: nq // 2, 3] = 0.015
    X[0], U[0] = x0, u0
    out = pkg.capi.plant_eval(plant, X, U)
    for b in range(nq):
        A, B, C, f = ol.plant_linearize(plant, X[b], U[b])
        assert np.allclose(out["dxdt"][b], ol.plant_derivative(plant, X[b], U[b]), rtol=1e-9, atol=1e-12)
        assert np.allclose(out["y"][b], ol.plant_output(plant, X[b]), rtol=1e-13)
        assert np.allclose(out["A"][b], A, rtol=1e-9, atol=1e-12)
        assert np.allclose(out["B"][b], B, rtol=1e-11, atol=1e-15)
        assert np.allclose(out["C"][b], C, rtol=1e-13)


@pytest.mark.parametrize("plant", [0, 1])
def test_plant_integrate_matches_oracle(plant, pkg, gpu_lib):
    x0, u0 = ol.plant_defaults(plant)
    rng = np.random.default_rng(2)
    nq = 32
    X = x0 * (1 + 0.02 * rng.uniform(-1, 1, (nq, len(x0))))
    U = np.tile(u0, (nq, 1))
    U[:, 0] += 0.05 * rng.uniform(-1, 1, nq)
    U[:, -1] += -0.3 * rng.uniform(0, 1, nq)
    Xg, steps = pkg.capi.plant_integrate(plant, X, U)
    for b in range(nq):
        xo, so = ol.plant_integrate(plant, X[b], U[b])
        assert so == steps[b]
        assert np.allclose(Xg[b], xo, rtol=1e-10, atol=1e-13)


def test_straight_line_sqrt_and_division_round_like_the_standard_ones(pkg, gpu_lib):
    """The integrator's branch-free square root and division (plant_dev.cuh) against sqrt() and / on the
    device: identical bits over the magnitudes a plant state can take and far beyond; operands outside
    [1e-290, 1e290] (zero, negative roots, infinities, NaN) are flagged, which is when the integrator
    falls back to the standard operations."""
    rng = np.random.default_rng(11)
    n = 1 << 20
    a = np.exp(rng.uniform(np.log(1e-280), np.log(1e280), n))
    a[: n // 4] = rng.uniform(1e-3, 1e3, n // 4)                    # where the plant lives
    a[n // 4: n // 2] = 1.0 + rng.uniform(-1, 1, n // 4) * 1e-9      # differences of nearly equal pressures
    b = np.exp(rng.uniform(np.log(1e-140), np.log(1e140), n)) * rng.choice([-1.0, 1.0], n)
    b[: n // 4] = rng.uniform(0.1, 1e4, n // 4)
    r = pkg.capi.inrange_math(a, b)
    ok = r["flagged"] == 0
    with np.errstate(over="ignore", under="ignore"):
        q = a / b
    representable = (np.abs(q) > 1e-290) & (np.abs(q) < 1e290)
    assert ok[: n // 2].all() and ok.mean() > 0.99
    assert np.array_equal(r["sqrt_fast"][ok], r["sqrt_std"][ok])
    assert np.array_equal(r["sqrt_std"], np.sqrt(a))
    sel = ok & representable
    d = np.abs(r["div_fast"][sel] - r["div_std"][sel]) / np.abs(r["div_std"][sel])
    assert d.max() <= 2.3e-16, d.max()                               # at most one unit in the last place ...
    assert (r["div_fast"][sel] == r["div_std"][sel]).mean() > 0.999   # ... and almost never that
    edge_a = np.array([0.0, -1.0, np.inf, np.nan, 1e-300, 1e300, 4.0, 1.0, 0.0, 1e-300])
    edge_b = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 1e-300, 2.0, 3.0])
    e = pkg.capi.inrange_math(edge_a, edge_b)
    assert (e["flagged"][:6] & 1).all() and (e["flagged"][6:8] & 2).all() and (e["flagged"][9] & 2)
    assert e["flagged"][8] == 1 and e["div_fast"][8] == 0.0           # 0 / 2: fine for the division (and sqrt(0) is flagged)


@pytest.mark.parametrize("nv", [4, 8])
def test_qp_solver_matches_oracle(nv, pkg, gpu_lib):
    rng = np.random.default_rng(3)
    nq, nu = 256, nv // 2
    H = np.zeros((nq, nv, nv)); f = np.zeros((nq, nv))
    lb = np.zeros((nq, nv)); ub = np.zeros((nq, nv)); lbA = np.zeros((nq, nv)); ubA = np.zeros((nq, nv))
    for q in range(nq):
        M = rng.standard_normal((nv, nv))
        H[q] = M @ M.T + 0.3 * np.eye(nv)
        f[q] = rng.standard_normal(nv) * rng.choice([0.1, 1.0, 5.0])
        lb[q] = -rng.uniform(0.05, 1, nv); ub[q] = rng.uniform(0.05, 1, nv)
        lbA[q] = -rng.uniform(0.05, 0.5, nv); ubA[q] = rng.uniform(0.05, 0.5, nv)
    g = pkg.capi.solve_qp(H, f, lb, ub, lbA, ubA)
    n_active = 0
    for q in range(nq):
        r = ol.solve_qp(H[q], f[q], lb[q], ub[q], lbA[q], ubA[q], nu)
        assert g["status"][q] == r["status"] == 0
        assert np.allclose(g["z"][q], r["z"], rtol=1e-9, atol=1e-12)
        assert g["active"][q] == r["active"], (q, bin(g["active"][q]), bin(r["active"]))
        assert np.isclose(g["objective"][q], r["objective"], rtol=1e-9, atol=1e-12)
        n_active += bin(r["active"]).count("1")
    assert n_active > nq  # the random problems do exercise the constraints
    # warm start from the optimal working sets gives the same answers
    g2 = pkg.capi.solve_qp(H, f, lb, ub, lbA, ubA, guess=g["working_set"])
    assert np.allclose(g2["z"], g["z"], rtol=1e-12, atol=1e-14) and (g2["active"] == g["active"]).all()


def _oracle_drive(setup, x0, u_def, n_steps, p=100):
    """Run the oracle closed loop and return the measurement sequence it saw."""
    o = ol.Oracle(setup, p=p)
    out = o.run_closed_loop(x0, setup.block_end_records(n_steps=n_steps), setup.sim_offsets, n_steps)
    n = len(x0)
    return out["traj"][0][:, 5 + n:]


@pytest.mark.parametrize("case", CASES)
def test_control_step_matches_oracle(case, setups, pkg, gpu_lib):
    """Step-by-step: same y sequence into both controllers, compare QP data, inputs and state."""
    s = setups[case]
    x_def, u_def = ol.plant_defaults(s.plant)
    T = 60
    ys = _oracle_drive(s, x_def, u_def, T)
    ys[45:] *= 1.0 + 2e-3  # a measurement jump the controller has not seen coming
    nc = pkg.from_setup(s, batch=2)
    nc.set_capture(True)
    o = ol.Oracle(s)
    y0 = ol.plant_output(s.plant, x_def)
    nc.Initialize(x_def, np.zeros(4), u_def, y0)
    o.initialize(x_def, np.zeros(4), u_def, y0)
    for k in range(T):
        ug = nc.GetNextInput(ys[k])
        uo = o.get_next_input(ys[k])
        assert np.array_equal(ug[0], ug[1])
        assert rel_err(ug[0], uo, 1e-9 / RTOL_U) < RTOL_U, (k, ug[0], uo)
        info = nc.step_info()
        for c in range(nc.n_controllers):
            Hg, fg, Gg = nc.qp(c)
            Ho, fo = o.qp(c)
            assert np.allclose(Hg[0], Ho, rtol=1e-10, atol=1e-9), (k, c)
            assert np.allclose(fg[0], fo, rtol=1e-8, atol=1e-9 * np.abs(fo).max()), (k, c, fg[0], fo)
            Ag, Bg, fdg = nc.linearization(c)
            Ao, Bo, Ado, Co, fdo = o.linearization(c)
            assert np.allclose(Ag[0], Ao, rtol=1e-11, atol=1e-14)
            # the discretised B in this controller's input order: the reference keeps the undelayed
            # columns in Borig and the delayed ones in Adelay (aug_lin_sys.cc:156-173)
            assert np.allclose(Bg[0][:, [0, 2]], Bo, rtol=1e-10, atol=1e-15)
            assert np.allclose(Bg[0][:, [1, 3]], Ado, rtol=1e-10, atol=1e-15)
            assert np.allclose(fdg[0], fdo, rtol=1e-10, atol=1e-15)
            Sug, Suog = nc.prediction(c)
            Suo_, _, _, Suoo = o.prediction(c)
            assert np.allclose(Sug[0], Suo_, rtol=1e-9, atol=1e-12), (k, c)
            if Suog is not None:
                assert np.allclose(Suog[0], Suoo, rtol=1e-9, atol=1e-12)
                # cross term: f_it = f + Gx du_other  <=>  Gx = (Q Su)' Su_other
                ny = nc.n_controlled_outputs[c]
                Q = np.kron(np.eye(nc.p), np.asarray(s.ywt[c]))
                assert np.allclose(Gg[0], (Q @ Suo_).T @ Suoo, rtol=1e-9, atol=1e-9)
            xg, dxg, yog, uog = nc.controller_state(c)
            xo, dxo, yoo, uoo = o.ctrl_state(c)
            assert np.allclose(xg[0], xo, rtol=1e-12, atol=1e-14)
            assert np.allclose(dxg[0], dxo, rtol=1e-9, atol=1e-13)
            assert np.allclose(uog[0], uoo, rtol=1e-8, atol=1e-13)
            st, act, obj = o.last_qp_info(c)
            assert info["status"][0, c] == st == 0
            assert info["active"][0, c] == act, (k, c, bin(info["active"][0, c]), bin(act))
            assert np.isclose(info["objective"][0, c], obj, rtol=RTOL_U, atol=1e-12)


@pytest.mark.parametrize("case", CASES)
def test_closed_loop_matches_oracle(case, setups, pkg, gpu_lib):
    """Perturbed scenarios through the on-device closed loop vs the oracle's closed loop."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 6, 400
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = np.minimum(be[:, 0], 150 + 20 * np.arange(B))   # pull the disturbance into the window
    nc = pkg.from_setup(s, batch=B)
    g = nc.run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=4)
    n = len(x_def)
    ug, uo = g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n]
    assert rel_err(ug, uo, ATOL_U / RTOL_U) < RTOL_U
    assert rel_err(g["traj"][:, :, 1:1 + n], o["traj"][:, :, 1:1 + n], 1e-3) < 1e-8
    assert (g["status"] == 0).all() and (o["status"] == 0).all()
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U
    assert np.array_equal(g["traj"][:, :, 0], o["traj"][:, :, 0])
    assert (g["active"] != 0).any()


def test_scenario0_reproduces_reference_records(setups, golden, pkg, gpu_lib):
    """GPU closed loop, nominal scenario, against the reference's recorded coop9.dat run."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    T = 1600
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, 2, T)
    g = pkg.from_setup(s, batch=2).run_closed_loop(x0, be, bo, T)
    idx = golden["coop-par/index"]; rec = golden["coop-par/records"]
    keep = idx < T
    tr = g["traj"][0][idx[keep]]
    n = len(x_def)
    assert (np.abs(tr[:, 1:1 + n] - rec[keep, 1:1 + n]) / np.maximum(np.abs(rec[keep, 1:1 + n]), 1e-3)).max() < 1e-5
    assert np.abs(tr[:, 1 + n:5 + n] - rec[keep, 1 + n:5 + n]).max() < 5e-6


def test_constraint_stress_matches_oracle(setups, pkg, gpu_lib):
    """Tight rate limits and a large disturbance: rate rows and bounds become active, the working
    set changes often (the reference's goldens never exercise this)."""
    import copy
    s = copy.deepcopy(setups["coop-par"])
    s.rate_lower = np.array([-2e-3, -2e-3]); s.rate_upper = np.array([2e-3, 2e-3])
    s.upper = np.array([0.05, 0.2])
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 4, 300
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 40
    bo[:, 1, 8] = -0.45
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=4)
    n = len(x_def)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    rate_bits = 0xFF00
    assert (g["active"] & rate_bits).any(), "rate constraints never became active"
    assert len(np.unique(g["active"])) > 3


def test_long_horizon_p200_matches_oracle(setups, pkg, gpu_lib):
    """BASELINE configs[4] shape: cooperative-parallel with twice the prediction horizon (p = 200)."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T, p = 4, 260, 200
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 120
    g = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s, p=p).run_closed_loop(x0, be, bo, T, n_threads=4)
    n = len(x_def)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U


@pytest.mark.parametrize("case,p", [("coop-par", 41), ("coop-par", 64), ("ncoop-ser", 90), ("cent-par", 129),
                                    ("coop-ser", 150), ("cent-ser", 256)])
def test_other_horizons_match_oracle(case, p, setups, pkg, gpu_lib):
    """Horizons other than 100/200 run the kernels that read the horizon at run time: shorter and
    longer power ladders, a partial last giant-step block, both rows-per-thread settings."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 3, 160
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 60
    g = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s, p=p).run_closed_loop(x0, be, bo, T, n_threads=3)
    n = len(x_def)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U


@pytest.mark.parametrize("case", ["coop-par", "cent-ser"])
def test_custom_observer_gain_matches_oracle(case, setups, pkg, gpu_lib):
    """A gain with non-zero plant-state rows: the a-posteriori estimate then depends on the new
    measurement, so the closed loop must linearise after the plant step (with the reference's
    M = [0; I] it linearises next to it)."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 4, 200
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 40
    rng = np.random.default_rng(11)
    nc = pkg.from_setup(s, batch=B)
    o = ol.Oracle(s)
    for c in range(nc.n_controllers):
        M = np.zeros((n + 4, 4))
        M[n:, :] = np.eye(4) * 0.8
        M[:n, :] = 2e-3 * rng.standard_normal((n, 4)) * np.abs(x_def)[:, None]
        nc.SetObserverGain(c, M)
        o.set_observer_gain(c, M)
    g = nc.run_closed_loop(x0, be, bo, T)
    r = o.run_closed_loop(x0, be, bo, T, n_threads=4)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], r["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert rel_err(g["traj"][:, :, 1:1 + n], r["traj"][:, :, 1:1 + n], 1e-3) < 1e-8
    assert np.array_equal(g["active"], r["active"])
    # and it is a different controller from the default one
    d = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    assert np.abs(d["traj"][:, :, 1 + n:5 + n] - g["traj"][:, :, 1 + n:5 + n]).max() > 1e-6


def test_old_noncooperative_serial_partition_matches_oracle(setups, pkg, gpu_lib):
    """SERIAL_CTRL_NONCOOP_OLD{1,2} (serial_compressors_constants.h:47-59,103-104): three controlled
    outputs per sub-controller, {0,1,2} and {2,3,1}.  The reference instantiates it
    (distributed_controller_list.h:29-30) but ships no setup file or recorded run for it, so parity
    is against the oracle only."""
    import copy
    s = copy.deepcopy(setups["ncoop-ser"])
    s.mode = pkg.setupfile.MODE_NCOOP_OLD
    s.ywt = [np.diag([1900.0, 3.0, 100.0]), np.diag([1900.0, 3.0, 1.0])]
    assert s.controlled_outputs == [[0, 1, 2], [2, 3, 1]]
    x_def, u_def = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 4, 150
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 40
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=4)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U
    # step-level check of the prediction columns for the permuted output set {2,3,1}
    nc = pkg.from_setup(s, batch=1)
    nc.set_capture(True)
    oc = ol.Oracle(s)
    y0 = ol.plant_output(s.plant, x_def)
    nc.Initialize(x_def, np.zeros(4), u_def, y0)
    oc.initialize(x_def, np.zeros(4), u_def, y0)
    y1 = y0 * (1 + 1e-3)
    ug, uo = nc.GetNextInput(y1)[0], oc.get_next_input(y1)
    assert np.allclose(ug, uo, rtol=1e-7, atol=1e-11)
    for c in range(2):
        Hg, fg, _ = nc.qp(c)
        Ho, fo = oc.qp(c)
        assert np.allclose(Hg[0], Ho, rtol=1e-9, atol=1e-9 * np.abs(Ho).max())
        assert np.allclose(fg[0], fo, rtol=1e-8, atol=1e-9 * np.abs(fo).max())


@pytest.mark.parametrize("case,seed", [("coop-par", 1), ("cent-ser", 2), ("ncoop-ser", 3), ("coop-ser", 4)])
def test_random_weights_and_constraints_match_oracle(case, seed, setups, pkg, gpu_lib):
    """The reference's setups only ever use diagonal weights.  Here: dense symmetric positive
    definite output and input weights, shifted references, tighter and asymmetric input bounds."""
    import copy
    rng = np.random.default_rng(seed)
    s = copy.deepcopy(setups[case])

    def spd_like(m):
        m = np.asarray(m, dtype=np.float64)
        d = np.sqrt(np.diag(m))
        n = len(d)
        c = rng.uniform(-0.4, 0.4, (n, n))
        c = (c + c.T) / 2
        np.fill_diagonal(c, 1.0)
        c = c @ c.T                      # positive definite correlation-like matrix
        c /= np.sqrt(np.outer(np.diag(c), np.diag(c)))
        return c * np.outer(d, d) * rng.uniform(0.5, 2.0)

    s.ywt = [spd_like(w) for w in s.ywt]
    s.uwt = spd_like(s.uwt)
    s.yref = np.asarray(s.yref) * (1 + 2e-3 * rng.standard_normal(4))
    s.lower = np.asarray(s.lower) * rng.uniform(0.3, 1.0, len(s.lower))
    s.upper = np.asarray(s.upper) * rng.uniform(0.3, 1.0, len(s.upper))
    s.rate_lower = np.asarray(s.rate_lower) * rng.uniform(0.05, 1.0, len(s.rate_lower))
    s.rate_upper = np.asarray(s.rate_upper) * rng.uniform(0.05, 1.0, len(s.rate_upper))
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 4, 140
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 30
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=4)
    assert (g["status"] == 0).all() and (o["status"] == 0).all()
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U


@pytest.mark.parametrize("case", ["coop-par", "cent-ser"])
def test_infeasible_qp_gives_zero_move_like_the_reference(case, setups, pkg, gpu_lib):
    """MpcQpSolver::SolveQP returns zeros when qpOASES fails (mpc_qp_solver.cc:66-69).  Contradictory
    bounds on the first input make every QP infeasible: both sides must report a failed solve and
    apply no move, step after step, and keep running."""
    import copy
    s = copy.deepcopy(setups[case])
    s.lower = np.array(s.lower, dtype=np.float64); s.upper = np.array(s.upper, dtype=np.float64)
    s.lower[0], s.upper[0] = 0.2, 0.1          # lower > upper
    x_def, u_def = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 3, 25
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T, n_threads=3)
    assert (g["status"] != 0).any()
    assert np.array_equal(g["status"] != 0, o["status"] != 0)
    ug, uo = g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n]
    assert np.isfinite(g["traj"]).all()
    assert rel_err(ug, uo, ATOL_U / RTOL_U) < RTOL_U
    failed = (g["status"] != 0).all(axis=2)    # every sub-controller failed in that record
    assert (np.abs(np.diff(ug, axis=1, prepend=0.0))[failed] == 0.0).all()


@pytest.mark.parametrize("B", [1, 17, 33])
def test_ragged_batch_sizes(B, setups, pkg, gpu_lib):
    """Batch sizes that do not fill a warp, a 16-scenario plant block or a 32-pair block: every
    scenario of the ragged batch equals the same scenario run in a batch of its own size class."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    T = 60
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, 33, T)
    be[:, 0] = 20
    full = pkg.from_setup(s, batch=33).run_closed_loop(x0, be, bo, T)
    part = pkg.from_setup(s, batch=B).run_closed_loop(x0[:B], be[:B], bo[:B], T)
    for key in ("traj", "active", "objective", "status"):
        assert np.array_equal(part[key], full[key][:B]), key


def test_shortest_and_longest_horizons_and_rejected_configurations(setups, pkg, gpu_lib):
    """p = 2 (one giant-step block, three ladder stages) and p = 256 (the longest the on-chip tables
    take) against the oracle; configurations outside the built scope are refused, not approximated."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 2, 40
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 10
    for p in (2, 9, 256):
        g = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)
        o = ol.Oracle(s, p=p).run_closed_loop(x0, be, bo, T, n_threads=2)
        assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U, p
        assert np.array_equal(g["active"], o["active"]), p
    C = pkg.capi
    # (other move horizons, delays and output partitions are served by the general path since round 2:
    # tests/test_generic_config.py)
    for mutate in (lambda c: setattr(c, "p", 257), lambda c: setattr(c, "p", 1), lambda c: setattr(c, "batch", 0),
                   lambda c: setattr(c, "m", 5), lambda c: c.delays.__setitem__(1, 1),
                   lambda c: setattr(c, "n_controllers", 3), lambda c: setattr(c, "n_disturbance_states", 2),
                   lambda c: c.n_controlled_outputs.__setitem__(1, 5)):
        cfg = C.default_config(0, 1, 4)
        mutate(cfg)
        h = C.C.c_void_p()
        rc = C.lib().cmpc_create(C.C.byref(cfg), 0, C.C.byref(h))
        assert rc != 0 and not h.value, "configuration outside the scope must be refused"
        assert C.lib().cmpc_last_error()


@pytest.mark.parametrize("case,p,n_iter", [("coop-par", 100, 1), ("coop-par", 100, 4), ("ncoop-par", 20, 9),
                                            ("cent-ser", 5, 1), ("coop-ser", 39, 3), ("ncoop-ser", 40, 9),
                                            ("cent-par", 8, 2)])
def test_sweep_counts_short_horizons_and_reference_ramps(case, p, n_iter, setups, pkg, gpu_lib):
    """Sweep counts other than the setups' 1 / 9, horizons around the 40-sample delay (no delayed row
    at all, exactly one, ...) and below one giant step, with a reference that changes over the
    horizon (SetOutputReference takes p rows, nerve_center.h:119-122)."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 3, 90
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 30
    ramp = np.asarray(s.yref)[None, :] * (1 + 1e-3 * np.linspace(0, 1, p)[:, None] * np.array([1.0, -1.0, 0.5, -0.5]))
    nc = pkg.from_setup(s, batch=B, p=p, n_solver_iterations=n_iter)
    nc.SetOutputReference(ramp)
    o = ol.Oracle(s, p=p, n_iter=n_iter)
    o.set_output_reference(ramp)
    g = nc.run_closed_loop(x0, be, bo, T)
    r = o.run_closed_loop(x0, be, bo, T, n_threads=3)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], r["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], r["active"])
    assert rel_err(g["objective"], r["objective"], 1e-6) < RTOL_U
    # the ramp matters: a constant reference gives another trajectory
    c = pkg.from_setup(s, batch=B, p=p, n_solver_iterations=n_iter).run_closed_loop(x0, be, bo, T)
    assert np.abs(c["traj"][:, :, 1 + n:5 + n] - g["traj"][:, :, 1 + n:5 + n]).max() > 1e-8


def test_handles_of_different_horizons_and_shapes_coexist(setups, pkg, gpu_lib):
    """Several handles alive at once, created in an order that would break per-function launch
    limits set per handle (two run-time horizons share one kernel instantiation), stepped in turn."""
    cases = [("coop-par", 120), ("coop-par", 64), ("cent-ser", 100), ("coop-par", 250)]
    B, T = 3, 30
    ctl, ref = [], []
    for case, p in cases:
        s = setups[case]
        x_def, _ = ol.plant_defaults(s.plant)
        x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
        be[:, 0] = 10
        ref.append(pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)["traj"])
        ctl.append((pkg.from_setup(s, batch=B, p=p), x0, be, bo, len(x_def)))
    import torch
    dev = torch.device("cuda", 0)
    bufs = []
    for nc, x0, be, bo, n in ctl:
        bufs.append((torch.from_numpy(x0).to(dev), torch.from_numpy(be).to(dev), torch.from_numpy(bo).to(dev),
                     torch.zeros((B, T, 1 + n + 8), dtype=torch.float64, device=dev)))
    st = torch.cuda.current_stream().cuda_stream
    for k in range(T):           # one record of every handle in turn
        for (nc, x0, be, bo, n), (dx, dbe, dbo, dtraj) in zip(ctl, bufs):
            nc.run_closed_loop_device(k, 1, T, dx.data_ptr(), be.shape[1], dbe.data_ptr(), dbo.data_ptr(),
                                      dtraj.data_ptr(), 0, 0, 0, st)
    torch.cuda.synchronize()
    for r, (_, _, _, dtraj) in zip(ref, bufs):
        assert np.array_equal(dtraj.cpu().numpy(), r)


@pytest.mark.parametrize("case,p", [("coop-par", 200), ("coop-ser", 64), ("cent-par", 41), ("ncoop-par", 7)])
def test_parity_hooks_at_other_horizons(case, p, setups, pkg, gpu_lib):
    """GeneratePrediction / QP / linearisation hooks after a few host-facing steps at horizons
    whose table layout differs from p = 100 (pruned columns start elsewhere, or nowhere)."""
    s = setups[case]
    x_def, u_def = ol.plant_defaults(s.plant)
    nc = pkg.from_setup(s, batch=2, p=p)
    nc.set_capture(True)
    o = ol.Oracle(s, p=p)
    y0 = ol.plant_output(s.plant, x_def)
    nc.Initialize(x_def, np.zeros(4), u_def, y0)
    o.initialize(x_def, np.zeros(4), u_def, y0)
    rng = np.random.default_rng(5)
    for k in range(45):          # long enough for the delay lines to fill with non-zero moves
        y = y0 * (1 + 2e-3 * rng.standard_normal(4))
        ug, uo = nc.GetNextInput(y)[0], o.get_next_input(y)
        assert np.allclose(ug, uo, rtol=1e-6, atol=1e-10), k
    for c in range(nc.n_controllers):
        Sug, Suog = nc.prediction(c)
        Suo_, _, _, Suoo = o.prediction(c)
        assert np.allclose(Sug[0], Suo_, rtol=1e-9, atol=1e-12)
        if Suog is not None:
            assert np.allclose(Suog[0], Suoo, rtol=1e-9, atol=1e-12)
        Hg, fg, _ = nc.qp(c)
        Ho, fo = o.qp(c)
        assert np.allclose(Hg[0], Ho, rtol=1e-9, atol=1e-9 * np.abs(Ho).max())
        assert np.allclose(fg[0], fo, rtol=1e-7, atol=1e-9 * np.abs(fo).max())
        xg, dxg, _, uog = nc.controller_state(c)
        xo, dxo, _, uoo = o.ctrl_state(c)
        assert np.allclose(xg[0], xo, rtol=1e-10, atol=1e-13)
        assert np.allclose(dxg[0], dxo, rtol=1e-8, atol=1e-12)
        assert np.allclose(uog[0], uoo, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("case", ["coop-par", "ncoop-ser", "cent-par"])
def test_per_controller_constraints_and_nonzero_initial_inputs(case, setups, pkg, gpu_lib):
    """Different input constraints for the two sub-controllers and an Initialize with non-zero
    u_init (the reference's drivers always start from zero): 60 host-facing steps, which also takes
    the delay rings once around."""
    s = setups[case]
    x_def, u_def = ol.plant_defaults(s.plant)
    B = 2
    nc = pkg.from_setup(s, batch=B)
    o = ol.Oracle(s)
    lo, up = np.asarray(s.lower, dtype=np.float64), np.asarray(s.upper, dtype=np.float64)
    rlo, rup = np.asarray(s.rate_lower, dtype=np.float64), np.asarray(s.rate_upper, dtype=np.float64)
    for c in range(nc.n_controllers):
        k = 1.0 / (c + 2)
        ic = pkg.InputConstraints(lo * k, up * k, rlo * k, rup * k)
        nc.SetConstraints(c, ic)
        ol.lib().orc_set_constraints(o.h, c, ol._p(ol.f64(lo * k)), ol._p(ol.f64(up * k)), ol._p(ol.f64(rlo * k)),
                                     ol._p(ol.f64(rup * k)))
    u_init = np.array([0.01, 0.02, -0.01, 0.03])
    y0 = ol.plant_output(s.plant, x_def)
    nc.Initialize(x_def, u_init, u_def, y0)
    o.initialize(x_def, u_init, u_def, y0)
    rng = np.random.default_rng(9)
    act_seen = False
    for k in range(60):
        y = y0 * (1 + 4e-3 * rng.standard_normal(4))
        ug, uo = nc.GetNextInput(y), o.get_next_input(y)
        assert np.allclose(ug[0], uo, rtol=1e-6, atol=1e-10), k
        assert np.array_equal(ug[0], ug[1])
        info = nc.step_info()
        for c in range(nc.n_controllers):
            st, act, _ = o.last_qp_info(c)
            assert info["status"][0, c] == st == 0
            assert info["active"][0, c] == act
            act_seen |= act != 0
    assert act_seen or nc.n_controllers == 1, "the tightened constraints never became active"


def test_runaway_plant_terminates(setups, pkg, gpu_lib):
    """An unphysical input offset (three times the setup's disturbance) drives the serial plant into
    a state where the adaptive integrator takes ever smaller steps (found by fuzzing: an unbounded
    odeint loop, i.e. a kernel that never returns).  Both sides now stop an interval after 4000
    accepted steps: the run must terminate, agree until the plant runs away, and report failed QPs
    afterwards instead of hanging."""
    import time
    s = setups["cent-ser"]
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B, T = 2, 71
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 33
    bo[:, 1, :] *= 3.0
    t0 = time.time()
    g = pkg.from_setup(s, batch=B, p=60, n_solver_iterations=8).run_closed_loop(x0, be, bo, T)
    assert time.time() - t0 < 60
    o = ol.Oracle(s, p=60, n_iter=8).run_closed_loop(x0, be, bo, T, n_threads=2)
    fin = np.isfinite(o["traj"]).all(axis=2)
    assert not fin.all(), "the scenario was meant to run away"
    k_ok = int(np.argmin(fin.all(axis=0))) - 2     # records well before the first non-finite one
    assert k_ok > 33
    ug, uo = g["traj"][:, :k_ok, 1 + n:5 + n], o["traj"][:, :k_ok, 1 + n:5 + n]
    assert rel_err(ug, uo, ATOL_U / RTOL_U) < 1e-5
    assert np.isfinite(g["traj"][:, :k_ok]).all() and not np.isfinite(g["traj"]).all()
    assert (g["status"][:, -1] != 0).all()      # the controller reports failed solves once the state is lost


def test_random_configuration_sweep(pkg, gpu_lib):
    """40 random configurations (tests/fuzz_parity.py): shapes x horizons 2..256 x sweep counts x
    dense weights x scaled constraints x batch sizes x disturbances.  The sweep that found the
    unbounded integrator loop; 360 configurations were clean when it was added."""
    import fuzz_parity
    assert fuzz_parity.run(seed=7, n_cfg=40) == []


def test_closed_loop_in_pieces_and_handle_state(setups, pkg, gpu_lib):
    """The device-resident loop may be advanced in pieces (bench.py does, one record per call):
    the records are bit-identical to a single call.  A closed-loop run leaves the controller
    half a step ahead (its plant kernel has already linearised the next record), so the
    host-facing step refuses to run until the handle is initialised again."""
    import torch
    s = setups["coop-par"]
    x_def, u_def = ol.plant_defaults(s.plant)
    B, T = 5, 90
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 30
    ref = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    nc = pkg.from_setup(s, batch=B)
    dev = torch.device("cuda", 0)
    d_x0, d_be, d_bo = (torch.from_numpy(a).to(dev) for a in (x0, be, bo))
    rec = 1 + len(x_def) + 8
    d_traj = torch.zeros((B, T, rec), dtype=torch.float64, device=dev)
    d_act = torch.zeros((B, T, 2), dtype=torch.int32, device=dev)
    for first, count in ((0, 25), (25, 1), (26, 1), (27, T - 27)):
        nc.run_closed_loop_device(first, count, T, d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(),
                                  d_traj.data_ptr(), d_act.data_ptr(), 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_traj.cpu().numpy(), ref["traj"])
    assert np.array_equal(d_act.cpu().numpy().astype(np.uint32), ref["active"])
    y0 = np.tile(ol.plant_output(s.plant, x_def), (B, 1))
    with pytest.raises(pkg.capi.CmpcError, match="cmpc_initialize"):
        nc.GetNextInput(y0)
    nc.Initialize(np.tile(x_def, (B, 1)), np.zeros(4), u_def, y0)
    u = nc.GetNextInput(y0)
    assert np.isfinite(u).all()
    # ... and a controller restarted on its own has no closed loop to continue
    with pytest.raises(pkg.capi.CmpcError, match="first_step = 0"):
        nc.run_closed_loop_device(T // 2, 1, T, d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(),
                                  d_traj.data_ptr(), 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    fresh = pkg.from_setup(s, batch=B)
    with pytest.raises(pkg.capi.CmpcError, match="first_step = 0"):
        fresh.run_closed_loop_device(3, 1, T, d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(),
                                     d_traj.data_ptr(), 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(pkg.capi.CmpcError):
        fresh.run_closed_loop_device(0, T + 1, T, d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(),
                                     d_traj.data_ptr(), 0, 0, 0, torch.cuda.current_stream().cuda_stream)


def test_no_launch_reads_uninitialised_shared_memory(pkg, gpu_lib):
    """The assemble kernel does not clear its operand region: zero padding comes from stored
    results.  The test build fills shared memory with NaNs at the start of every launch; parity
    must still hold (tests/poison_check.py, run in a fresh process because the library is chosen at
    import time)."""
    import os, pathlib, subprocess, sys
    root = pathlib.Path(__file__).resolve().parent.parent
    lib = root / "compressor-mpc_b200" / "libcmpc_b200_poison.so"
    assert lib.exists(), "libcmpc_b200_poison.so is missing: run __graft_entry__.build()"
    env = dict(os.environ, CMPC_B200_LIB=str(lib))
    out = subprocess.run([sys.executable, str(root / "tests" / "poison_check.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]


def test_full_size_batch_properties(setups, golden, pkg, gpu_lib):
    """BASELINE configs[3] at full size: 4096 perturbed scenarios, closed loop over the disturbance
    onset.  Size-independent properties: every QP solved, trajectories finite and inside the input
    constraints, scenario 0 = the reference's recorded run, duplicate scenarios give identical
    trajectories, and a spread sample of scenarios agrees with the oracle."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 4096, 1300
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    x0[B - 1], be[B - 1], bo[B - 1] = x0[17], be[17], bo[17]     # a duplicate far away in the batch
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    n = len(x_def)
    u = g["traj"][:, :, 1 + n:5 + n]
    assert (g["status"] == 0).all()
    assert np.isfinite(g["traj"]).all()
    lo = np.tile(s.lower, 2); hi = np.tile(s.upper, 2)
    assert (u >= lo - 1e-9).all() and (u <= hi + 1e-9).all()
    du = np.diff(u, axis=1)
    assert (du >= np.tile(s.rate_lower, 2) - 1e-9).all() and (du <= np.tile(s.rate_upper, 2) + 1e-9).all()
    assert np.array_equal(g["traj"][B - 1], g["traj"][17]) and np.array_equal(g["active"][B - 1], g["active"][17])
    idx = golden["coop-par/index"]; rec = golden["coop-par/records"]
    keep = idx < T
    tr = g["traj"][0][idx[keep]]
    assert (np.abs(tr[:, 1:1 + n] - rec[keep, 1:1 + n]) / np.maximum(np.abs(rec[keep, 1:1 + n]), 1e-3)).max() < 1e-5
    assert np.abs(tr[:, 1 + n:5 + n] - rec[keep, 1 + n:5 + n]).max() < 5e-6
    sample = np.array([1, 2, 3, 500, 1023, 2048, 3000, 4094])
    o = ol.Oracle(s).run_closed_loop(x0[sample], be[sample], bo[sample], T, n_threads=8)
    assert rel_err(u[sample], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"][sample], o["active"])
    assert rel_err(g["objective"][sample], o["objective"], 1e-6) < RTOL_U
    # the perturbed scenarios do differ from each other
    assert np.unique(np.round(u[:, -1, 0], 9)).size > B // 2


@pytest.mark.parametrize("case", CASES)
def test_setup_workflow_writes_reference_records(case, setups, golden, pkg, gpu_lib, tmp_path):
    """setup file in, .dat records out: each of the reference's six workflows (run-all-tests.sh) on
    the GPU, against the first 400 records of the reference's own output file."""
    s = setups[case]
    folder = "parallel" if s.plant == 0 else "serial"
    n = 11 if s.plant == 0 else 10
    f = tmp_path / f"setup-{case}"
    f.write_text(pkg.setupfile.format_setup(s))
    r = pkg.workflow.run_setup(f, batch=2, out_dir=tmp_path / folder, n_records=400)
    assert [p.name for p in r["paths"]] == [s.output_filename, s.output_filename + ".s1"]
    got = pkg.workflow.parse_records(r["paths"][0].read_text(), n)
    rec = golden[f"{case}/records"][:400]
    assert got.shape == (400, 1 + n + 8 + 1)
    assert np.allclose(got[:, 0], rec[:, 0], rtol=1e-5)
    assert (np.abs(got[:, 1:1 + n] - rec[:, 1:1 + n]) / np.maximum(np.abs(rec[:, 1:1 + n]), 1e-3)).max() < 2e-5
    assert np.abs(got[:, 1 + n:5 + n] - rec[:, 1 + n:5 + n]).max() < 1e-5
    assert (got[:, -1] > 0).all()


def test_cxx_setup_driver_matches_reference_records(setups, golden, pkg, gpu_lib, tmp_path):
    """The C++ host side (reference-named facade + setup-file driver) end to end: setup file in,
    the reference's centralized-serial .dat records out."""
    import subprocess
    from conftest import ROOT
    exe = ROOT / "compressor-mpc_b200" / "cmpc_run_setup"
    assert exe.exists(), "build with __graft_entry__.build()"
    s = setups["cent-ser"]
    import copy
    s = copy.deepcopy(s)
    s.sim_t_end = np.array([10.0, 20.0])          # 400 records instead of 10 000
    (tmp_path / "serial").mkdir()
    (tmp_path / "setup-cent-ser").write_text(pkg.setupfile.format_setup(s))
    out = subprocess.run([str(exe), "setup-cent-ser", "serial", "centralized"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = pkg.workflow.parse_records((tmp_path / "serial" / "centralized.dat").read_text(), 10)
    T = got.shape[0]
    assert T == 400
    # the first 200 records are disturbance free exactly like the reference's first block
    rec = golden["cent-ser/records"][:200]
    assert (np.abs(got[:200, 1:11] - rec[:, 1:11]) / np.maximum(np.abs(rec[:, 1:11]), 1e-3)).max() < 2e-5
    assert np.abs(got[:200, 11:15] - rec[:, 11:15]).max() < 1e-5
    # after record 200 the -0.1 offset on plant input 6 (second outlet valve) acts on its outlet pressure
    assert abs(got[-1, 7] - got[199, 7]) > 1e-3


@pytest.mark.parametrize("case", CASES)
def test_full_recorded_run_on_gpu(case, setups, golden, pkg, gpu_lib):
    """All 10 000 records of each of the reference's six recorded runs (results/*/run1), closed loop
    on the GPU: through the disturbance at record 1001 to the end, within the print precision of the
    .dat files (6 significant digits)."""
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    T = int(golden[f"{case}/n_records"])
    assert T == 10000
    g = pkg.from_setup(s, batch=1).run_closed_loop(x_def, s.block_end_records(n_steps=T), s.sim_offsets, T)
    idx, rec = golden[f"{case}/index"], golden[f"{case}/records"]
    tr = g["traj"][0][idx]
    n = len(x_def)
    assert (g["status"] == 0).all()
    assert np.allclose(tr[:, 0], rec[:, 0], rtol=1e-5, atol=1e-9)
    assert (np.abs(tr[:, 1:1 + n] - rec[:, 1:1 + n]) / np.maximum(np.abs(rec[:, 1:1 + n]), 1e-3)).max() < 1e-5
    assert np.abs(tr[:, 1 + n:5 + n] - rec[:, 1 + n:5 + n]).max() < 5e-6
    assert (np.abs(tr[:, 5 + n:] - rec[:, 5 + n:]) / np.maximum(np.abs(rec[:, 5 + n:]), 1e-3)).max() < 1e-5


def test_sweep_shape_at_size(setups, pkg, gpu_lib):
    """BASELINE configs[4] shape at real occupancy: p = 200 (2 CTAs/SM, table behind the powers),
    8192 scenarios, 300 records, against the oracle on a 64-scenario sample spread over the batch."""
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T, p = 8192, 300, 200
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = np.minimum(be[:, 0], 60 + (np.arange(B) % 200))     # the disturbance arrives inside the run
    g = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)
    n = len(x_def)
    assert (g["status"] == 0).all() and np.isfinite(g["traj"]).all()
    sample = np.unique(np.concatenate([np.arange(16), np.linspace(16, B - 1, 48).astype(int)]))
    o = ol.Oracle(s, p=p).run_closed_loop(x0[sample], be[sample], bo[sample], T, n_threads=16)
    u = g["traj"][:, :, 1 + n:5 + n]
    assert rel_err(u[sample], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"][sample], o["active"])
    assert rel_err(g["objective"][sample], o["objective"], 1e-6) < RTOL_U


def test_headline_config_parity_protocol(setups, pkg, gpu_lib):
    """coop-par x 4096 scenarios.  Scenarios 0..63 over the full run
    and ALL 4096 scenarios over the first records, applied inputs / objectives / active sets against
    the oracle.  The full protocol (10 000 and 1500 records: about 7 M oracle steps, minutes of host
    time) runs with CMPC_FULL_PROTOCOL=1 (log in profiles/); the default sizes keep the suite short."""
    import os
    full = os.environ.get("CMPC_FULL_PROTOCOL") == "1"
    T_long, T_all = (10000, 1500) if full else (2000, 120)
    threads = os.cpu_count() or 1
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    n = len(x_def)
    B = 4096
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T_long)
    nc = pkg.from_setup(s, batch=64)
    g = nc.run_closed_loop(x0[:64], be[:64], bo[:64], T_long)
    o = ol.Oracle(s).run_closed_loop(x0[:64], be[:64], bo[:64], T_long, n_threads=threads)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U
    nc.close()
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T_all)
    if not full:
        be[:, 0] = np.minimum(be[:, 0], 30 + (np.arange(B) % 60))   # every scenario meets its disturbance
    g = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T_all)
    o = ol.Oracle(s).run_closed_loop(x0, be, bo, T_all, n_threads=threads)
    assert rel_err(g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n], ATOL_U / RTOL_U) < RTOL_U
    assert np.array_equal(g["active"], o["active"])
    assert rel_err(g["objective"], o["objective"], 1e-6) < RTOL_U


def test_streaming_closed_loop_with_host_buffers(setups, pkg, gpu_lib):
    """cmpc_closed_loop_start / _step (one record per call, host buffers: bench.py's e2e path) gives
    the records of cmpc_run_closed_loop bit for bit."""
    s = setups["coop-ser"]
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 6, 70
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 20 + np.arange(B)
    ref = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)["traj"]
    nc = pkg.from_setup(s, batch=B)
    with pytest.raises(pkg.capi.CmpcError, match="cmpc_closed_loop_start"):
        nc.closed_loop_step(bo[:, 0])
    nc.closed_loop_start(x0)
    for k in range(T):
        off = np.stack([bo[b, min(int((k >= be[b]).sum()), be.shape[1] - 1)] for b in range(B)])
        rec = nc.closed_loop_step(off)
        assert np.array_equal(rec, ref[:, k]), k


@pytest.mark.parametrize("pipelined", [False, True])
@pytest.mark.parametrize("case,B", [("coop-par", 37), ("cent-ser", 16), ("ncoop-ser", 5)])
def test_streaming_closed_loop_with_page_locked_buffers(setups, pkg, gpu_lib, case, B, pipelined):
    """The same with page-locked (mapped) host buffers: the plant kernel reads the offsets and writes the
    record rows itself, chunk by chunk (no copies on the stream).  Ragged last block, both plants; a
    buffer that is only 8-byte aligned goes back to the copy path.  Bit for bit in every case."""
    import torch
    s = setups[case]
    x_def, _ = ol.plant_defaults(s.plant)
    n, T = len(x_def), 50
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 15 + np.arange(B) % 20
    ref = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)["traj"]
    nc = pkg.from_setup(s, batch=B)
    nc.closed_loop_start(x0)
    nc.closed_loop_pipeline(pipelined)   # (the control step of the next record launched ahead, or not)
    off = torch.empty((B, bo.shape[2]), dtype=torch.float64).pin_memory()
    big = torch.empty((B * (1 + n + 8) + 1,), dtype=torch.float64).pin_memory()
    rec_al, rec_odd = big[:-1].view(B, 1 + n + 8), big[1:].view(B, 1 + n + 8)   # 16-byte aligned / not
    for k in range(T):
        off.copy_(torch.from_numpy(np.stack([bo[b, min(int((k >= be[b]).sum()), be.shape[1] - 1)] for b in range(B)])))
        rec = rec_al if k % 3 else rec_odd
        rec.fill_(float("nan"))
        nc.closed_loop_step_raw(off.data_ptr(), rec.data_ptr())
        assert np.array_equal(rec.numpy(), ref[:, k]), k
    if pipelined:
        with pytest.raises(pkg.capi.CmpcError, match="before its first step"):
            nc.closed_loop_pipeline(False)
        # a restart drops the control step that was launched ahead
        nc.closed_loop_start(x0)
        nc.closed_loop_step_raw(off.data_ptr(), rec_al.data_ptr())
        assert np.array_equal(rec_al.numpy()[:, :1 + n], ref[:, 0, :1 + n])


@pytest.mark.parametrize("case,B", [("coop-par", 45), ("cent-ser", 33)])
def test_get_next_input_with_page_locked_buffers(setups, pkg, gpu_lib, case, B):
    """cmpc_get_next_input with page-locked (mapped) measurement and input buffers -- the linearisation kernel
    reads the measurements itself, the solve kernel writes the inputs straight out -- gives what the copy path
    gives, bit for bit (ragged last block of the linearisation kernel; one and two sub-controllers)."""
    import torch
    s = setups[case]
    x_def, u_def = ol.plant_defaults(s.plant)
    rng = np.random.default_rng(5)
    y0 = ol.plant_output(s.plant, x_def)
    a, b = pkg.from_setup(s, batch=B), pkg.from_setup(s, batch=B)
    for nc in (a, b):
        nc.Initialize(x_def, np.zeros(4), u_def, y0)
    y_pin = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    u_pin = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    for k in range(12):
        y = y0 * (1 + 1e-3 * rng.uniform(-1, 1, (B, 4)))
        u_copy = a.GetNextInput(y)
        y_pin.copy_(torch.from_numpy(y)); u_pin.fill_(float("nan"))
        b.GetNextInputRaw(y_pin.data_ptr(), u_pin.data_ptr())
        assert np.array_equal(u_pin.numpy(), u_copy), k


def test_argument_validation_added_in_round_two(setups, pkg, gpu_lib):
    import torch
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(s.plant)
    nc = pkg.from_setup(s, batch=2)
    lib, ptr, f64 = pkg.capi.lib(), pkg.capi.ptr, pkg.capi.f64
    bad_uwt = f64([[1.0, 0.5], [0.25, 1.0]])
    assert lib.cmpc_set_weights(nc._h, 0, ptr(bad_uwt), None) == 3        # CMPC_ERR_UNSUPPORTED: non-symmetric uwt
    ok = f64([0.0, 0.0])
    assert lib.cmpc_set_constraints(nc._h, 0, ptr(f64([np.nan, 0.0])), ptr(ok), ptr(ok), ptr(ok)) == 1
    inf = f64([np.inf, np.inf])
    assert lib.cmpc_set_constraints(nc._h, 0, ptr(-inf), ptr(inf), ptr(-inf), ptr(inf)) == 0            # unbounded is fine
    # a continuation must pick up where the previous call stopped
    B, T = 2, 20
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    dev = torch.device("cuda", 0)
    d_x0, d_be, d_bo = (torch.from_numpy(a).to(dev) for a in (x0, be, bo))
    d_traj = torch.zeros((B, T, 20), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(), d_traj.data_ptr(), 0, 0, 0, st)
    nc.run_closed_loop_device(0, 5, T, *args)
    with pytest.raises(pkg.capi.CmpcError, match="continuation does not match"):
        nc.run_closed_loop_device(7, 1, T, *args)
    with pytest.raises(pkg.capi.CmpcError, match="continuation does not match"):
        nc.run_closed_loop_device(5, 1, T + 1, *args)
    nc.run_closed_loop_device(5, 15, T, *args)
    torch.cuda.synchronize()
    assert torch.isfinite(d_traj).all()
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("case", ["coop-par", "cent-ser", "ncoop-ser"])
def test_timing_window_leaves_the_results_alone(case, setups, pkg, gpu_lib):
    """n-timing-iterations (GetNextInputWithTiming, nerve_center.h:134-182): the sweeps from
    n_timing_iterations on are left out of the measured time.  On the device the solve is then split
    into timed sweeps, untimed sweeps and the finishing part; records, active sets and objectives stay
    bit-identical to the unsplit run, and every record gets a positive time of its own."""
    s = setups[case]
    x_def, u_def = ol.plant_defaults(s.plant)
    B, T = 4, 60
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 15
    ref = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T)
    means = {}
    for n_t in (-1, 0, 3, s.n_iterations):
        r = pkg.from_setup(s, batch=B).run_closed_loop(x0, be, bo, T, n_timing_iterations=n_t)
        assert np.array_equal(r["traj"], ref["traj"]), n_t
        assert np.array_equal(r["active"], ref["active"]) and np.array_equal(r["objective"], ref["objective"])
        assert r["step_ns"].shape == (T,) and (r["step_ns"] > 0).all()
        means[n_t] = r["step_ns"][5:].mean()
    assert all(1e3 < m < 5e6 for m in means.values()), means       # microseconds to milliseconds, not garbage
    # the host-facing call with a window
    nc, nc2 = pkg.from_setup(s, batch=B), pkg.from_setup(s, batch=B)
    y0 = np.stack([ol.plant_output(s.plant, x) for x in x0])
    nc.Initialize(x0, np.zeros(4), u_def, y0)
    nc2.Initialize(x0, np.zeros(4), u_def, y0)
    for k in range(6):
        y = ref["traj"][:, k, 5 + len(x_def):]
        u, ns = nc.GetNextInputWithTiming(y, 2)
        assert np.array_equal(u, nc2.GetNextInput(y)) and ns > 0
        # (the closed loop linearises inside its plant kernel, the host-facing step in lin_kernel: same
        # numbers to rounding, not to the bit)
        assert np.allclose(u, ref["traj"][:, k, 1 + len(x_def):5 + len(x_def)], rtol=1e-9, atol=1e-12)


def test_run_all_tests_workflow(setups, golden, pkg, gpu_lib, tmp_path):
    """setup/run-all-tests.sh on the GPU path (1 + 1 + n x 4 runs with the setup files rewritten per
    run), then read_timing_data.m's aggregation over what was written."""
    wf = pkg.workflow
    sdir = tmp_path / "setup"
    sdir.mkdir()
    for case in CASES:
        (sdir / f"setup-{case}").write_text(pkg.setupfile.format_setup(setups[case]))
    out = tmp_path / "results"
    written = wf.run_all_tests(sdir, out, n_records=80, n_max=2, log=lambda *_: None)
    names = sorted(str(p.relative_to(out)) for p in written)
    assert names == sorted([f"{f}/{n}.dat" for f in ("parallel", "serial")
                            for n in ("centralized", "coop1", "coop2", "ncoop1", "ncoop2")])
    for case, fname in (("coop-par", "parallel/coop2.dat"), ("ncoop-ser", "serial/ncoop1.dat"), ("cent-ser", "serial/centralized.dat")):
        n = 11 if setups[case].plant == 0 else 10
        got = wf.parse_records((out / fname).read_text(), n)
        rec = golden[f"{case}/records"][:80]
        assert got.shape == (80, 1 + n + 9)
        assert np.abs(got[:, 1 + n:5 + n] - rec[:, 1 + n:5 + n]).max() < 1e-5
        assert (got[:, -1] > 0).all()
    t = wf.read_timing_data(out, runs=("",), n_max=2)
    for folder in ("parallel", "serial"):
        assert all(t[folder][k].shape == (2,) and (t[folder][k] > 0).all() for k in ("cent", "coop", "ncoop"))
    assert (sdir / "setup-coop-par").read_text() == pkg.setupfile.format_setup(setups["coop-par"])   # inputs untouched
