"""Replay ONE configuration of tests/fuzz_parity.py's tuned sweep (python tools/fuzz_case.py SEED INDEX) and show
where GPU and oracle part: per record, max relative input difference, active-set and status agreement."""
import copy, json, pathlib, sys
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge
import oracle_lib as ol

seed, index = int(sys.argv[1]), int(sys.argv[2])
pkg = ge.load_package()
setups = {k: pkg.setupfile.setup_from_dict(v) for k, v in json.load(open(ROOT / "tests" / "golden" / "setups.json")).items()}
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

for it in range(index + 1):   # the same draws, in the same order, as fuzz_parity.run
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
print("config", it, dict(case=case, p=p, n_iter=n_iter, B=B, T=T))
g = pkg.from_setup(s, batch=B, p=p, n_solver_iterations=n_iter).run_closed_loop(x0, be, bo, T)
o = ol.Oracle(s, p=p, n_iter=n_iter).run_closed_loop(x0, be, bo, T, n_threads=4)
fin = np.isfinite(o["traj"]).all(axis=2).all(axis=0)
print("oracle finite up to record", int(np.argmin(fin)) if not fin.all() else T)
for k in range(T):
    ug, uo = g["traj"][:, k, 1 + n:5 + n], o["traj"][:, k, 1 + n:5 + n]
    err = float(np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3)))
    print(k, "err %.2e" % err, "active equal", bool(np.array_equal(g["active"][:, k], o["active"][:, k])),
          "status gpu/oracle", g["status"][:, k].tolist(), o["status"][:, k].tolist(),
          "max|u| %.3g" % float(np.nanmax(np.abs(uo))))
