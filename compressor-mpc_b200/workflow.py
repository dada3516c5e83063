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
    nc.set_timing(True)
    t0 = time.perf_counter()
    res = nc.run_closed_loop(x0, be, bo, T, want_qp=True)
    wall = time.perf_counter() - t0
    n_timed, step_ms, _ = nc.get_timing()
    ns_per_step = step_ms / max(n_timed, 1) * 1e6       # mean device time of one (batched) control step
    out_dir = pathlib.Path(out_dir if out_dir is not None else s.folder_name)
    out_dir.mkdir(parents=True, exist_ok=True)
    n = len(x_def)
    paths = []
    for b in range(batch):
        p = out_dir / (s.output_filename if b == 0 else f"{s.output_filename}.s{b}")
        p.write_text(format_records(res["traj"][b], n, ns_per_step))
        paths.append(p)
    return dict(paths=paths, result=res, ns_per_step=ns_per_step, wall_s=wall, setup=s)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("setup_file")
    ap.add_argument("--plant", choices=sorted(PLANTS))
    ap.add_argument("--mode", choices=sorted(MODES))
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out")
    ap.add_argument("--records", type=int)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    r = run_setup(a.setup_file, PLANTS.get(a.plant) if a.plant else None, MODES.get(a.mode) if a.mode else None,
                  batch=a.batch, out_dir=a.out, n_records=a.records, device=a.device)
    print(f"wrote {len(r['paths'])} file(s), first: {r['paths'][0]}; control step {r['ns_per_step'] / 1e3:.1f} us "
          f"per batched step, wall {r['wall_s']:.2f} s")


if __name__ == "__main__":
    main()
