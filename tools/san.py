"""Smallest runs that touch every kernel, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool racecheck python tools/san.py
    compute-sanitizer --tool memcheck  python tools/san.py
"""
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
raw = json.loads((ROOT / "tests" / "golden" / "setups.json").read_text())
B, T = 3, 4
for case, p in (("coop-par", 100), ("coop-par", 200), ("cent-ser", 100), ("ncoop-par", 64), ("coop-ser", 41)):
    s = pkg.setupfile.setup_from_dict(raw[case])
    x_def, u_def = pkg.plant_defaults(s.plant)
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    nc = pkg.from_setup(s, batch=B, p=p)
    r = nc.run_closed_loop(x0, be, bo, T)
    r2 = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T, n_timing_iterations=1)   # split solve kernels
    assert np.array_equal(r["traj"], r2["traj"])
    nc2 = pkg.from_setup(s, batch=B, p=p)                                                          # host-facing step (lin_kernel)
    nc2.Initialize(x0, np.zeros(4), u_def, r["traj"][:, 0, -4:])
    u = nc2.GetNextInput(r["traj"][:, 0, -4:])
    print(case, p, "ok", np.isfinite(r["traj"]).all(), int(r["status"].max()), np.array_equal(u, r["traj"][:, 0, -8:-4]))
# the general configuration path
s = pkg.setupfile.setup_from_dict(raw["coop-par"])
x_def, _ = pkg.plant_defaults(0)
conf = pkg.Configuration(0, [pkg.SubController(1, [0, 3]), pkg.SubController(1, [0, 3]), pkg.SubController(2, [1, 3])],
                         m=3, p=40, delays=(0, 6, 0, 9), n_iterations=3)
nc = pkg.NerveCenter.from_configuration(conf, batch=B)
nc.SetOutputReference(np.asarray(s.yref))
nc.SetWeights(s.uwt, [np.diag([1.0, 420.0])] * 3)
for c, sl in enumerate((slice(0, 1), slice(1, 2), slice(0, 2))):
    nc.SetConstraints(c, pkg.InputConstraints(s.lower[sl], s.upper[sl], s.rate_lower[sl], s.rate_upper[sl]))
x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, 12)
r = nc.run_closed_loop(x0, be, bo, 12)
print("general", "ok", np.isfinite(r["traj"]).all(), int(r["status"].max()))
