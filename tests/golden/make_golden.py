#!/usr/bin/env python
"""Build the committed golden fixtures from the reference's own recorded runs.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/setups.json      the six setup-{cent,coop,ncoop}-{par,ser} files of the
                                reference (setup/setup-*), parsed into JSON
  tests/golden/golden_traj.npz  the six distinct closed-loop trajectories of
                                results/{parallel,serial}/run1/{centralized,coop9,ncoop9}.dat
                                (t, x, u, y per record; 6 significant digits as printed by the
                                reference), sub-sampled: every record for k < 1600 (start-up
                                transient and the disturbance onset at record 1001), every 25th
                                after that.  The timing column is kept as its mean only.
Nothing at test time reads /root/reference.
"""
import importlib.util
import json
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = pathlib.Path("/root/reference")

spec = importlib.util.spec_from_file_location("setupfile", ROOT / "compressor-mpc_b200" / "setupfile.py")
setupfile = importlib.util.module_from_spec(spec)
sys.modules["setupfile"] = setupfile
spec.loader.exec_module(setupfile)

CASES = {
    "cent-par": ("setup-cent-par", 0, 0, "parallel/run1/centralized.dat"),
    "coop-par": ("setup-coop-par", 0, 1, "parallel/run1/coop9.dat"),
    "ncoop-par": ("setup-ncoop-par", 0, 2, "parallel/run1/ncoop9.dat"),
    "cent-ser": ("setup-cent-ser", 1, 0, "serial/run1/centralized.dat"),
    "coop-ser": ("setup-coop-ser", 1, 1, "serial/run1/coop9.dat"),
    "ncoop-ser": ("setup-ncoop-ser", 1, 2, "serial/run1/ncoop9.dat"),
}


def read_dat(path, n_states):
    vals = np.array(path.read_text().split(), dtype=np.float64)
    rec = 1 + n_states + 4 + 4 + 1
    assert vals.size % rec == 0, (vals.size, rec)
    return vals.reshape(-1, rec)


def keep_index(n):
    idx = np.arange(n)
    return idx[(idx < 1600) | (idx % 25 == 0)]


def main():
    setups, arrays = {}, {}
    for name, (sfile, plant, mode, dat) in CASES.items():
        s = setupfile.parse_setup((REF / "setup" / sfile).read_text(), plant, mode)
        setups[name] = setupfile.setup_to_dict(s)
        full = read_dat(REF / "results" / dat, setupfile.N_STATES[plant])
        idx = keep_index(full.shape[0])
        arrays[name + "/index"] = idx.astype(np.int32)
        arrays[name + "/records"] = full[idx, :-1]
        arrays[name + "/mean_step_ns"] = np.array(full[:, -1].mean())
        arrays[name + "/n_records"] = np.array(full.shape[0])
        print(name, full.shape, "kept", idx.size, "mean step ns", full[:, -1].mean())
    (HERE / "setups.json").write_text(json.dumps(setups, indent=1))
    np.savez_compressed(HERE / "golden_traj.npz", **arrays)
    timing_golden()


def timing_golden():
    """tests/golden/timing_golden.json: the mean of the timing column of every recorded .dat file
    (results/{parallel,serial}/run1..5/{centralized,coop1..9,ncoop1..9}.dat) and what the reference's
    read_timing_data.m makes of them (res.parallel.coop etc.: per-controller means over the runs)."""
    out = {"file_mean_ns": {}, "read_timing_data": {}}
    for folder, n_states in (("parallel", 11), ("serial", 10)):
        names = ["centralized"] + [f"coop{i}" for i in range(1, 10)] + [f"ncoop{i}" for i in range(1, 10)]
        per = {}
        for run in range(1, 6):
            for name in names:
                full = read_dat(REF / "results" / folder / f"run{run}" / f"{name}.dat", n_states)
                per[(run, name)] = float(full[:, -1].mean())
                out["file_mean_ns"][f"{folder}/run{run}/{name}.dat"] = per[(run, name)]
        # read_timing_data.m:18-62, written out independently of workflow.read_timing_data
        cent = np.mean([per[(r, "centralized")] for r in range(1, 6)])
        out["read_timing_data"][folder] = {
            "cent": [float(cent)] * 9,
            "coop": [float(np.mean([per[(r, f"coop{i}")] / 2 for r in range(1, 6)])) for i in range(1, 10)],
            "ncoop": [float(np.mean([per[(r, f"ncoop{i}")] / 2 for r in range(1, 6)])) for i in range(1, 10)]}
        print(folder, {k: [round(v) for v in vals[:3]] for k, vals in out["read_timing_data"][folder].items()})
    (HERE / "timing_golden.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
