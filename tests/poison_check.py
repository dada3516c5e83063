"""Helper for test_no_launch_reads_uninitialised_shared_memory: run as a subprocess with
CMPC_B200_LIB pointing at the NaN-poisoning test build.  Closed loops for a few shapes and horizons
against the oracle; prints OK or raises."""
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib as ol  # noqa: E402

pkg = ge.load_package()
assert "poison" in str(pkg.capi.LIB_PATH), pkg.capi.LIB_PATH
setups = json.load(open(ROOT / "tests" / "golden" / "setups.json"))
for case, p in (("coop-par", 100), ("cent-ser", 100), ("ncoop-par", 72), ("coop-ser", 200)):
    s = pkg.setupfile.setup_from_dict(setups[case])
    x_def, _ = ol.plant_defaults(s.plant)
    B, T = 3, 120
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T)
    be[:, 0] = 50
    g = pkg.from_setup(s, batch=B, p=p).run_closed_loop(x0, be, bo, T)
    o = ol.Oracle(s, p=p).run_closed_loop(x0, be, bo, T, n_threads=3)
    n = len(x_def)
    ug, uo = g["traj"][:, :, 1 + n:5 + n], o["traj"][:, :, 1 + n:5 + n]
    assert np.isfinite(g["traj"]).all(), case
    err = np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3))
    assert err < 1e-6, (case, p, err)
    assert np.array_equal(g["active"], o["active"]), case
print("OK")
