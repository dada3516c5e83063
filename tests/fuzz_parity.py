"""Random sweep of GPU-vs-oracle closed loops: shapes, horizons, sweep counts, dense weights,
scaled constraints, batch sizes and (sometimes) a three-fold disturbance.  Test infrastructure:
`run(seed, n)` is used by tests/test_gpu_parity.py; as a script, `python tests/fuzz_parity.py SEED N`.

Records are compared up to shortly before the oracle's first non-finite record: a three-fold
disturbance can drive the serial plant into a runaway (both sides then cap the integrator and
produce NaNs, whose pattern is not a parity statement)."""
import copy
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib as ol  # noqa: E402


def run(seed: int, n_cfg: int, verbose: bool = False):
    pkg = ge.load_package()
    setups = {k: pkg.setupfile.setup_from_dict(v)
              for k, v in json.load(open(ROOT / "tests" / "golden" / "setups.json")).items()}
    cases = list(setups)
    rng = np.random.default_rng(seed)

    def spd_like(m):
        m = np.asarray(m, dtype=np.float64)
        d = np.sqrt(np.diag(m))
        c = rng.uniform(-0.4, 0.4, (len(d), len(d)))
        c = (c + c.T) / 2
        np.fill_diagonal(c, 1.0)
        c = c @ c.T
        c /= np.sqrt(np.outer(np.diag(c), np.diag(c)))
        return c * np.outer(d, d) * rng.uniform(0.5, 2.0)

    bad = []
    for it in range(n_cfg):
        case = cases[rng.integers(len(cases))]
        s = copy.deepcopy(setups[case])
        p = int(rng.choice([rng.integers(2, 20), rng.integers(20, 60), rng.integers(60, 140), rng.integers(140, 257)]))
        n_iter = int(rng.integers(1, 10))
        if rng.random() < 0.5:
            s.ywt = [spd_like(w) for w in s.ywt]
            s.uwt = spd_like(s.uwt)
        if rng.random() < 0.5:
            for k in ("lower", "upper", "rate_lower", "rate_upper"):
                setattr(s, k, np.asarray(getattr(s, k)) * rng.uniform(0.05, 1.0, len(getattr(s, k))))
        x_def, _ = ol.plant_defaults(s.plant)
        n = len(x_def)
        B, T = int(rng.integers(1, 6)), int(rng.integers(20, 90))
        x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T, first=int(rng.integers(0, 1000)))
        be[:, 0] = rng.integers(5, T)
        if rng.random() < 0.3:
            bo[:, 1, :] *= 3.0
        cfg = dict(case=case, p=p, n_iter=n_iter, B=B, T=T)
        g = pkg.from_setup(s, batch=B, p=p, n_solver_iterations=n_iter).run_closed_loop(x0, be, bo, T)
        o = ol.Oracle(s, p=p, n_iter=n_iter).run_closed_loop(x0, be, bo, T, n_threads=4)
        fin = np.isfinite(o["traj"]).all(axis=2).all(axis=0)
        K = T if fin.all() else max(int(np.argmin(fin)) - 3, 0)
        ug, uo = g["traj"][:, :K, 1 + n:5 + n], o["traj"][:, :K, 1 + n:5 + n]
        err = float(np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3))) if K else 0.0
        ok = (err < 1e-6 and np.array_equal(g["active"][:, :K], o["active"][:, :K])
              and np.array_equal(g["status"][:, :K] != 0, o["status"][:, :K] != 0))
        if verbose:
            print(it, cfg, "K", K, "err", err, "ok", ok, flush=True)
        if not ok:
            bad.append(dict(cfg, K=K, err=err))
    return bad


def general_configs(seed: int, n_cfg: int):
    """The random general configurations of run_general, as a generator (shared with the sensitivity
    analysis below): yields (index, conf, base setup, ywts, constraints, x0, block_end, block_off, T)."""
    pkg = ge.load_package()
    setups = {k: pkg.setupfile.setup_from_dict(v)
              for k, v in json.load(open(ROOT / "tests" / "golden" / "setups.json")).items()}
    rng = np.random.default_rng(seed)
    splits = [[4], [2, 2], [1, 3], [3, 1], [1, 1, 2], [2, 1, 1], [1, 2, 1], [1, 1, 1, 1]]
    for it in range(n_cfg):
        plant = int(rng.integers(0, 2))
        base = setups["coop-par" if plant == 0 else "coop-ser"]
        nus = splits[rng.integers(len(splits))]
        m = int(rng.integers(1, 5))
        while m * max(nus) > 8:
            m -= 1
        p = int(rng.choice([rng.integers(m, 12), rng.integers(12, 70), rng.integers(70, 200)]))
        delays = [int(rng.choice([0, rng.integers(2, 61)])) for _ in range(4)]
        ctrls, ywts, cons, first = [], [], [], 0
        for nu in nus:
            ny = int(rng.integers(1, 5))
            outs = [int(v) for v in rng.permutation(4)[:ny]]
            ctrls.append(pkg.SubController(nu, outs))
            ywts.append(np.diag(rng.choice([1.0, 10.0, 420.0], ny)))
            idx = [(first + i) % 2 for i in range(nu)]          # torque / recycle bounds of the setup, by input kind
            cons.append(tuple(np.asarray(getattr(base, k))[idx] * rng.uniform(0.3, 1.0)
                              for k in ("lower", "upper", "rate_lower", "rate_upper")))
            first += nu
        conf = pkg.Configuration(plant, ctrls, p=p, m=m, delays=delays, n_iterations=int(rng.integers(1, 7)))
        x_def, _ = ol.plant_defaults(plant)
        B, T = int(rng.integers(1, 4)), int(rng.integers(30, 110))
        x0, be, bo = pkg.scenarios.make_scenarios(base, x_def, B, T, first=int(rng.integers(0, 1000)))
        be[:, 0] = rng.integers(5, T)
        yield it, conf, base, ywts, cons, x0, be, bo, T


def _healthy_prefix(o, g, T):
    healthy = np.isfinite(o["traj"]).all(axis=2).all(axis=0) & (o["status"] == 0).all(axis=2).all(axis=0) \
        & (g["status"] == 0).all(axis=2).all(axis=0)
    return T if healthy.all() else max(int(np.argmin(healthy)) - 3, 0)


def oracle_sensitivity(conf, base, ywts, cons, x0, be, bo, T, rel=1e-14):
    """How much this closed loop amplifies a perturbation of the size of rounding errors: the oracle
    against itself with the initial state moved by `rel` (relative).  Returns the largest relative
    input difference (same measure as the parity check) over the healthy prefix."""
    n = x0.shape[1]
    mk = lambda: ol.Oracle.from_configuration(conf, base.uwt, ywts, cons, base.yref)
    a = mk().run_closed_loop(x0, be, bo, T, n_threads=3)
    b = mk().run_closed_loop(x0 * (1.0 + rel), be, bo, T, n_threads=3)
    K = _healthy_prefix(a, b, T)
    ua, ub = a["traj"][:, :K, 1 + n:5 + n], b["traj"][:, :K, 1 + n:5 + n]
    return float(np.max(np.abs(ua - ub) / np.maximum(np.abs(ua), 1e-3))) if K else 0.0


def run_general(seed: int, n_cfg: int, verbose: bool = False):
    """Random GENERAL configurations on the GPU's general path against the generalised
    oracle: plant, number of sub-controllers and their input counts, controlled-output partitions,
    per-input delays (0 or 2..60), move horizon, prediction horizon, sweep count.

    Compared up to shortly before the first record that is not healthy on either side (a non-finite
    state, or a QP the exact solver gives up on: random configurations do reach numerically hopeless
    problems, e.g. an unstable linearisation raised to the 164th power makes H indefinite in floating
    point, and which of two summation orders trips first is not a parity statement; the zero move on a
    failed QP has its own test).  Tolerance: 1e-6 like everywhere else -- or, for a loop that amplifies
    rounding noise itself, 100 times what the ORACLE shows against itself when its initial state is
    moved by 1e-14 (long horizons with heavy output weights make H ill-conditioned: cond(H) times the
    1e-13 of two summation orders)."""
    pkg = ge.load_package()
    bad = []
    for it, conf, base, ywts, cons, x0, be, bo, T in general_configs(seed, n_cfg):
        n = x0.shape[1]
        B = x0.shape[0]
        nc = pkg.NerveCenter.from_configuration(conf, batch=B)
        nc.SetWeights(base.uwt, ywts)
        nc.SetOutputReference(np.asarray(base.yref, dtype=np.float64))
        for c, k in enumerate(cons):
            nc.SetConstraints(c, pkg.InputConstraints(*k))
        g = nc.run_closed_loop(x0, be, bo, T)
        o = ol.Oracle.from_configuration(conf, base.uwt, ywts, cons, base.yref).run_closed_loop(x0, be, bo, T, n_threads=3)
        K = _healthy_prefix(o, g, T)
        ug, uo = g["traj"][:, :K, 1 + n:5 + n], o["traj"][:, :K, 1 + n:5 + n]
        err = float(np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3))) if K else 0.0
        sens = None
        ok = err < 1e-6 and np.array_equal(g["active"][:, :K], o["active"][:, :K])
        if not ok:
            sens = oracle_sensitivity(conf, base, ywts, cons, x0, be, bo, T)
            ok = err < 100.0 * sens      # the loop itself turns 1e-14 into `sens`
        cfg = dict(plant=conf.plant, nus=[c.n_inputs for c in conf.controllers], m=conf.m, p=conf.p, delays=list(conf.delays),
                   outs=[c.controlled_outputs for c in conf.controllers], n_iter=conf.n_iterations, B=B, T=T)
        if verbose:
            print(it, cfg, "K", K, "err", err, "oracle self-sensitivity", sens, "ok", ok, flush=True)
        if not ok:
            per_rec = np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3), axis=(0, 2)) if K else np.zeros(1)
            prof = {int(k): float(per_rec[k]) for k in np.linspace(0, max(K - 1, 0), 8).astype(int)}
            bad.append(dict(cfg, K=K, err=err, oracle_self_sensitivity=sens, err_by_record=prof))
        nc.close()
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "general":
        b = run_general(int(sys.argv[1]), int(sys.argv[2]), verbose=True)
        print(f"{len(b)} mismatches", b)
        raise SystemExit

    b = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 40, verbose=True)
    print(f"{len(b)} mismatches", b)
