"""The oracle against the reference's own recorded closed-loop runs and the
derived known-answer values.  CPU only."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import CASES

# all 10 000 records of every recorded run (about 10 s of oracle time each)
N_RECORDS = {case: 10000 for case in CASES}


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_trajectory(case, setups, golden):
    s = setups[case]
    x0, _ = ol.plant_defaults(s.plant)
    T = N_RECORDS[case]
    out = ol.Oracle(s).run_closed_loop(x0, s.block_end_records(n_steps=T), s.sim_offsets, T)
    idx = golden[case + "/index"]
    rec = golden[case + "/records"]
    keep = idx < T
    idx, rec = idx[keep], rec[keep]
    tr = out["traj"][0][idx]
    n = len(x0)
    # the reference prints 6 significant digits: half a unit of the last digit is 5e-6 relative
    assert np.allclose(tr[:, 0], rec[:, 0], rtol=1e-5, atol=1e-9)
    xerr = np.abs(tr[:, 1:1 + n] - rec[:, 1:1 + n]) / np.maximum(np.abs(rec[:, 1:1 + n]), 1e-3)
    assert xerr.max() < 1e-5, xerr.max()
    uerr = np.abs(tr[:, 1 + n:5 + n] - rec[:, 1 + n:5 + n])
    assert uerr.max() < 5e-6, uerr.max()
    yerr = np.abs(tr[:, 5 + n:] - rec[:, 5 + n:]) / np.maximum(np.abs(rec[:, 5 + n:]), 1e-3)
    assert yerr.max() < 1e-5, yerr.max()
    assert (out["status"] == 0).all()


def test_plant_known_answers():
    # SURVEY.md §8c "Plant KATs"
    x0, u0 = ol.plant_defaults(0)
    d = ol.plant_derivative(0, x0, u0)
    assert np.allclose(d[:4], [-7.7643e-4, -2.4371e-4, 6.32278e-3, -8.543159e-2], rtol=2e-5)
    assert d[4] == 0 and np.allclose(d[5:10], d[:5]) and np.isclose(d[10], 1.3178e-4, rtol=1e-4)
    A, B, C, f = ol.plant_linearize(0, x0, u0)
    assert np.isclose(A[0, 0], -1.012103191705, rtol=1e-11)
    assert np.isclose(A[0, 2], -0.986225423656, rtol=1e-11)
    assert np.isclose(A[10, 1], 0.381056435364, rtol=1e-11)
    assert np.isclose(A[10, 10], -0.876959866173, rtol=1e-11)
    Ad, Bd, Cd, fd = ol.plant_discretize(0, x0, u0)
    assert np.allclose(np.diag(Ad)[:5], [0.927176143114, 0.66823061593, 0.803799725176, 0.990645274512, 0.9048375], rtol=1e-10)
    assert np.isclose(Ad[10, 10], 0.963033852389, rtol=1e-10)
    assert np.isclose(Bd[3, 0], 6.992542420742744, rtol=1e-12)
    assert np.isclose(Bd[4, 1], 0.07445161571741146, rtol=1e-12)
    assert np.isclose(Bd[2, 0], 0.0032518630328751477, rtol=1e-11)
    assert np.allclose(fd[:4], [-4.478606600552e-05, 3.14223284584e-06, 2.706372194916e-04, -5.655450098368e-03], rtol=1e-10)
    xs, us = ol.plant_defaults(1)
    ds = ol.plant_derivative(1, xs, us)
    assert np.allclose(ds, [7.2258e-4, 5.14941e-3, 1.039337e-2, 8.974562e-2, 0, -2.61544e-3, -2.94779e-3,
                            -4.136247e-2, 8.974562e-2, 0], rtol=2e-5)


KAT = {  # SURVEY.md §8c "Derived KATs": y0, step-0 u, controller-0 H[0,:] and f
    "cent-par": ([-4.268478570310e-05, 0, -4.268478570310e-05, 0],
                 [2.001161645593e4, 6.125245052176, -0.8213685898195, -2.017125022063, 173.9281732775,
                  216.0860845385, -64.41456118887, -102.9937557434],
                 [0.929842411556, 1.177291153199, 0.929842411556, 1.177291153199, 24.778806891718,
                  57.381285631742, 24.778806891718, 57.381285631742]),
    "coop-par": ([-4.480364006781e-05, 0, -4.480364006781e-05, 0],
                 [1.901161305439e4, 6.127727358425, 173.7742694076, 216.1529016240],
                 [0.92936141508, 1.177633822584, 24.757392099168, 57.390586602239]),
    "ncoop-par": ([-4.049496251928e-05, 0, -4.049496251928e-05, 0],
                  [2.201148427777e4, 6.219420639736, 167.9139540101, 218.8759105040],
                  [0.999191842837, 1.126630089438, 28.007245427466, 55.870198300505]),
    "cent-ser": ([1.157851129557e-05, 1.121339102255e-06, 8.324812951184e-04, 0],
                 [2.001648038990e4, -6.665632398511, 1.436617614285, 45.14338609280, 393.8005916622,
                  -278.1440640793, 414.8129760405, 1957.129756361],
                 [-0.261820421431, -0.305238556501, -16.77449591566, 1.959240430923, 11.843256559576,
                  -8.102276042366, -12.348746064128, 66.984178986606]),
    "coop-ser": ([2.078009130435e-05, 6.442347123477e-07, 6.870321020548e-04, 0],
                 [4.103579985707e4, -10.88127595338, 453.7867956647, -526.9122278009],
                 [-0.916206353914, -0.49667263265, -2.029067520062, -10.776144459854]),
    "ncoop-ser": ([2.020617093483e-05, 0, 2.021753662601e-04, 0],
                  [3.001999451883e4, -4.709132708576, 93.03421157125, -266.0459771880],
                  [-0.644209643529, -0.19876948465, -13.375187514746, -2.889301480092]),
}


@pytest.mark.parametrize("case", CASES)
def test_step0_known_answers(case, setups):
    s = setups[case]
    x0, u0 = ol.plant_defaults(s.plant)
    o = ol.Oracle(s)
    y0 = ol.plant_output(s.plant, x0)
    o.initialize(x0, np.zeros(4), u0, y0)
    u = o.get_next_input(y0)
    u_kat, h_kat, f_kat = KAT[case]
    assert np.allclose(u, u_kat, rtol=1e-8, atol=1e-13)
    H, f = o.qp(0)
    assert np.allclose(H[0], h_kat, rtol=1e-9)
    assert np.allclose(f, f_kat, rtol=1e-8)


def test_qp_solver_against_bruteforce():
    """Exact active-set solve vs enumeration of all bound/rate combinations on random small QPs."""
    import itertools
    rng = np.random.default_rng(7)
    nv, nu = 4, 2
    Ain = np.eye(nv)
    for i in range(nu, nv):
        Ain[i, i - nu] = -1
    rows = np.vstack([np.eye(nv), -np.eye(nv), Ain, -Ain])
    for trial in range(40):
        M = rng.standard_normal((nv, nv))
        H = M @ M.T + 0.5 * np.eye(nv)
        f = 3 * rng.standard_normal(nv)
        lb = -rng.uniform(0.1, 1, nv); ub = rng.uniform(0.1, 1, nv)
        lbA = -rng.uniform(0.05, 0.5, nv); ubA = rng.uniform(0.05, 0.5, nv)
        rhs = np.concatenate([lb, -ub, lbA, -ubA])
        r = ol.solve_qp(H, f, lb, ub, lbA, ubA, nu)
        assert r["status"] == 0
        best, best_obj = None, np.inf
        for k in range(0, nv + 1):
            for W in itertools.combinations(range(4 * nv), k):
                N = rows[list(W)]
                if k and np.linalg.matrix_rank(N) < k:
                    continue
                K = np.block([[H, -N.T], [N, np.zeros((k, k))]])
                sol = np.linalg.solve(K, np.concatenate([-f, rhs[list(W)]]))
                z, lam = sol[:nv], sol[nv:]
                if (rows @ z - rhs < -1e-9).any() or (lam < -1e-9).any():
                    continue
                obj = 0.5 * z @ H @ z + f @ z
                if obj < best_obj:
                    best, best_obj = z, obj
        assert best is not None
        assert np.allclose(r["z"], best, atol=1e-9), (trial, r["z"], best)
        assert np.isclose(r["objective"], best_obj, rtol=1e-9, atol=1e-12)
        # warm start from the optimal working set reproduces the solution without iterations
        r2 = ol.solve_qp(H, f, lb, ub, lbA, ubA, nu, guess=r["working_set"])
        assert r2["iterations"] == 0 and np.allclose(r2["z"], r["z"], atol=1e-12)
