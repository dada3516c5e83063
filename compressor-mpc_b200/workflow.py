"""The reference's `setup-{cent,coop,ncoop}-{ser,par}` workflows on the GPU path.

The reference builds six executables (tests/*-with-timing.cc) that take a setup file, run a
500 s closed-loop simulation and write one record per 50 ms sample to
`<folder-name>/<output-filename>`:

    t
    x (n_states values)
    u (4 values, relative to the default input)
    y (4 values)
    ns (wall time of GetNextInput in nanoseconds)
    <blank>

(layout reconstructed in SURVEY.md §3.1; `read_timing_data.m:18-23` reads it back as
whitespace separated numbers).  `run_setup` does the same with the batched CUDA controller and
the on-device plant; with batch > 1 the additional scenarios are the perturbed ones of
`scenarios.py` and are written to `<output-filename>.s<k>`.

    python compressor-mpc_b200/workflow.py setup-coop-par [--batch 64] [--out DIR] [--records N]
"""
from __future__ import annotations

import argparse
import pathlib
import sys
import time

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
    import __graft_entry__ as _entry
    _pkg = _entry.load_package()
    capi, scenarios, setupfile, controller = _pkg.capi, _pkg.scenarios, _pkg.setupfile, sys.modules[_pkg.__name__ + ".controller"]
else:
    from . import capi, controller, scenarios, setupfile

PLANTS = {"parallel": setupfile.PLANT_PARALLEL, "serial": setupfile.PLANT_SERIAL}
MODES = {"centralized": setupfile.MODE_CENT, "cooperative": setupfile.MODE_COOP, "noncoop": setupfile.MODE_NCOOP,
         "noncoop-old": setupfile.MODE_NCOOP_OLD}


def _fmt_row(values) -> str:
    """Eigen's default stream format: 6 significant digits, every coefficient padded to the
    width of the widest one, separated by one blank."""
    toks = ["%g" % v for v in values]
    w = max(len(t) for t in toks)
    return " ".join(t.rjust(w) for t in toks)


def format_records(traj: np.ndarray, n_states: int, step_ns) -> str:
    """traj: (T, 1 + n + 4 + 4) rows [t, x, u, y]."""
    step_ns = np.broadcast_to(np.asarray(step_ns), (traj.shape[0],))
    out = []
    for r, ns in zip(traj, step_ns):
        out.append("%g" % r[0])
        out.append(_fmt_row(r[1:1 + n_states]))
        out.append(_fmt_row(r[1 + n_states:5 + n_states]))
        out.append(_fmt_row(r[5 + n_states:]))
        out.append("%d" % int(round(ns)))
        out.append("")
    return "\n".join(out) + "\n"


def parse_records(text: str, n_states: int) -> np.ndarray:
    vals = np.array(text.split(), dtype=np.float64)
    return vals.reshape(-1, 1 + n_states + 4 + 4 + 1)


def infer_test(setup_text: str, name: str = ""):
    """(plant, mode) from the folder-name / output-filename keys, like the reference's choice of
    executable (setup/run-all-tests.sh:6-35)."""
    toks = setup_text.split()
    folder = toks[toks.index("folder-name") + 1] if "folder-name" in toks else ""
    fname = toks[toks.index("output-filename") + 1] if "output-filename" in toks else name
    plant = PLANTS["serial" if "ser" in folder else "parallel"]
    mode = MODES["noncoop"] if fname.startswith("ncoop") else MODES["cooperative"] if fname.startswith("coop") else MODES["centralized"]
    return plant, mode


def run_setup(setup_path, plant=None, mode=None, batch=1, out_dir=None, n_records=None, device=0, Ts=0.05):
    text = pathlib.Path(setup_path).read_text()
    if plant is None or mode is None:
        plant, mode = infer_test(text, pathlib.Path(setup_path).name)
    s = setupfile.parse_setup(text, plant, mode)
    ends = s.block_end_records(Ts=Ts)
    T = int(ends[-1]) if n_records is None else int(n_records)
    x_def, _ = capi.plant_defaults(plant)
    x0, be, bo = scenarios.make_scenarios(s, x_def, batch, T, Ts=Ts)
    nc = controller.from_setup(s, batch=batch, device=device)
    t0 = time.perf_counter()
    # every record carries the time of its own control step inside the reference's timing window
    # (GetNextInputWithTiming with the setup's n-timing-iterations, nerve_center.h:134-182)
    res = nc.run_closed_loop(x0, be, bo, T, want_qp=True, n_timing_iterations=s.n_timing_iterations)
    wall = time.perf_counter() - t0
    step_ns = res["step_ns"]
    out_dir = pathlib.Path(out_dir if out_dir is not None else s.folder_name)
    out_dir.mkdir(parents=True, exist_ok=True)
    n = len(x_def)
    paths = []
    for b in range(batch):
        p = out_dir / (s.output_filename if b == 0 else f"{s.output_filename}.s{b}")
        p.write_text(format_records(res["traj"][b], n, step_ns))
        paths.append(p)
    nc.close()
    return dict(paths=paths, result=res, step_ns=step_ns, ns_per_step=float(step_ns.mean()) if T else 0.0,
                wall_s=wall, setup=s)


# setup/run-all-tests.sh:6-35: the two centralised runs, then for i = 1..9 the four distributed ones
# with n-timing-iterations = i written to coop<i>.dat / ncoop<i>.dat
ALL_TESTS_CENT = [("setup-cent-par", "parallel", "centralized"), ("setup-cent-ser", "serial", "centralized")]
ALL_TESTS_DIST = [("setup-coop-par", "parallel", "cooperative", "coop"), ("setup-coop-ser", "serial", "cooperative", "coop"),
                  ("setup-ncoop-par", "parallel", "noncoop", "ncoop"), ("setup-ncoop-ser", "serial", "noncoop", "ncoop")]


def set_setup_params(text: str, n_timing_iterations: int, filename: str) -> str:
    """The two gawk edits of setup/run-all-tests.sh:39-47: the line after the n-timing-iterations key
    and the line after the output-filename key are replaced."""
    out, pending = [], None
    for line in text.splitlines():
        if pending is not None:
            out.append(pending)
            pending = None
            continue
        out.append(line)
        if "n-timing-iterations" in line:
            pending = str(n_timing_iterations)
        elif "output-filename" in line:
            pending = filename
    return "\n".join(out) + "\n"


def run_all_tests(setup_dir, out_root, n_records=None, n_max=9, batch=1, device=0, log=print):
    """setup/run-all-tests.sh on the GPU path: 1 + 1 + n_max x 4 runs.  The setup files are read from
    setup_dir (never rewritten there; the edited text goes to a scratch copy under out_root), the
    .dat files land in out_root/<folder-name>/ like the reference's.  Returns the written paths."""
    setup_dir, out_root = pathlib.Path(setup_dir), pathlib.Path(out_root)
    scratch = out_root / "setup"
    scratch.mkdir(parents=True, exist_ok=True)
    written = []

    def one(fname, plant, mode, text):
        f = scratch / fname
        f.write_text(text)
        folder = setupfile.parse_setup(text, PLANTS[plant], MODES[mode]).folder_name
        r = run_setup(f, PLANTS[plant], MODES[mode], batch=batch, out_dir=out_root / folder, n_records=n_records,
                      device=device)
        log(f"{fname} -> {r['paths'][0]}: mean step {r['ns_per_step'] / 1e3:.1f} us")
        written.append(r["paths"][0])

    for fname, plant, mode in ALL_TESTS_CENT:
        one(fname, plant, mode, (setup_dir / fname).read_text())
    for i in range(1, n_max + 1):
        log(f"Using {i} timing iterations.")
        for fname, plant, mode, prefix in ALL_TESTS_DIST:
            one(fname, plant, mode, set_setup_params((setup_dir / fname).read_text(), i, f"{prefix}{i}.dat"))
    return written


def read_timing_data(results_root, runs=("run1", "run2", "run3", "run4", "run5"), n_max=9):
    """read_timing_data.m:18-62: the mean of the last value of every record (the step time in ns) per
    file; distributed runs divided by 2 (time per sub-controller), the centralised value repeated for
    every iteration count; then the mean over the runs.  results_root holds parallel/ and serial/,
    each with one folder per run (or the .dat files directly: runs=("",)).
    Returns {"parallel": {"cent", "coop", "ncoop"}, "serial": {...}}, each an array of n_max values."""
    results_root = pathlib.Path(results_root)

    def mean_ns(path, n_states):
        vals = np.array(path.read_text().split(), dtype=np.float64)
        return vals.reshape(-1, 1 + n_states + 9)[:, -1].mean()

    res = {}
    for folder, n_states in (("parallel", 11), ("serial", 10)):
        cent = np.zeros((n_max, len(runs))); coop = np.zeros_like(cent); ncoop = np.zeros_like(cent)
        for r, run in enumerate(runs):
            d = results_root / folder / run
            cent[:, r] = mean_ns(d / "centralized.dat", n_states)
            for i in range(1, n_max + 1):
                coop[i - 1, r] = mean_ns(d / f"coop{i}.dat", n_states) / 2
                ncoop[i - 1, r] = mean_ns(d / f"ncoop{i}.dat", n_states) / 2
        res[folder] = {"cent": cent.mean(axis=1), "coop": coop.mean(axis=1), "ncoop": ncoop.mean(axis=1)}
    return res


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("setup_file", help="a setup file, or with --all the directory that holds the six of them")
    ap.add_argument("--all", action="store_true", help="setup/run-all-tests.sh: 1 + 1 + 9 x 4 runs")
    ap.add_argument("--plant", choices=sorted(PLANTS))
    ap.add_argument("--mode", choices=sorted(MODES))
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out")
    ap.add_argument("--records", type=int)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    if a.all:
        out = a.out or "."
        run_all_tests(a.setup_file, out, n_records=a.records, batch=a.batch, device=a.device)
        t = read_timing_data(out, runs=("",))
        for folder, d in t.items():
            for k, v in d.items():
                print(folder, k, " ".join("%.0f" % x for x in v), "ns")
        return
    r = run_setup(a.setup_file, PLANTS.get(a.plant) if a.plant else None, MODES.get(a.mode) if a.mode else None,
                  batch=a.batch, out_dir=a.out, n_records=a.records, device=a.device)
    print(f"wrote {len(r['paths'])} file(s), first: {r['paths'][0]}; control step {r['ns_per_step'] / 1e3:.1f} us "
          f"per batched step, wall {r['wall_s']:.2f} s")


if __name__ == "__main__":
    main()
