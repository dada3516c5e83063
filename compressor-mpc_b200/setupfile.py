"""Setup-file reader/writer for the `setup-{cent,coop,ncoop}-{ser,par}` workflows.

Mirrors the tokeniser of the reference's `include/read_files.h:13-81`
(`ReadString` / `ReadNumbers`): lines starting with `#`, empty lines and
all-whitespace lines are skipped; numbers are read across lines until the
requested count is reached; trailing text on the last consumed line is ignored
with a warning.  Keys are read in the fixed order the reference driver uses
(SURVEY.md §3.1): n-iterations, n-timing-iterations, folder-name,
output-filename, yref, uwt, ywt (one matrix per sub-controller), constraints-
lower/-upper/-rate-lower/-rate-upper, simulation (repeated blocks of n_inputs
offsets followed by t_end).
"""
from __future__ import annotations

import dataclasses
import sys
from typing import List

import numpy as np

PLANT_PARALLEL, PLANT_SERIAL = 0, 1
MODE_CENT, MODE_COOP, MODE_NCOOP, MODE_NCOOP_OLD = 0, 1, 2, 3

# (plant, mode) -> (n_controllers, n_sub_control_inputs, controlled outputs per controller)
# include/parallel_compressors_constants.h:70-93, include/serial_compressors_constants.h:84-109
SHAPES = {
    (PLANT_PARALLEL, MODE_CENT): (1, 4, [[0, 1, 3]]),
    (PLANT_PARALLEL, MODE_COOP): (2, 2, [[0, 1, 3], [0, 1, 3]]),
    (PLANT_PARALLEL, MODE_NCOOP): (2, 2, [[0, 3], [1, 3]]),
    (PLANT_SERIAL, MODE_CENT): (1, 4, [[0, 1, 2, 3]]),
    (PLANT_SERIAL, MODE_COOP): (2, 2, [[0, 1, 2, 3], [0, 1, 2, 3]]),
    (PLANT_SERIAL, MODE_NCOOP): (2, 2, [[0, 1], [2, 3]]),
    # SERIAL_CTRL_NONCOOP_OLD{1,2} (serial_compressors_constants.h:47-59,103-104): instantiated by the
    # reference (distributed_controller_list.h:29-30) but used by none of its main programs
    (PLANT_SERIAL, MODE_NCOOP_OLD): (2, 2, [[0, 1, 2], [2, 3, 1]]),
}
N_STATES = {PLANT_PARALLEL: 11, PLANT_SERIAL: 10}
N_INPUTS = {PLANT_PARALLEL: 9, PLANT_SERIAL: 8}


@dataclasses.dataclass
class Setup:
    plant: int
    mode: int
    n_iterations: int
    n_timing_iterations: int
    folder_name: str
    output_filename: str
    yref: np.ndarray                 # (4,)
    uwt: np.ndarray                  # (4, 4) full system
    ywt: List[np.ndarray]            # per controller (n_y, n_y)
    lower: np.ndarray                # (n_sub_control_inputs,)
    upper: np.ndarray
    rate_lower: np.ndarray
    rate_upper: np.ndarray
    sim_offsets: np.ndarray          # (n_blocks, n_inputs)
    sim_t_end: np.ndarray            # (n_blocks,)

    @property
    def n_controllers(self) -> int:
        return SHAPES[(self.plant, self.mode)][0]

    @property
    def n_sub_control_inputs(self) -> int:
        return SHAPES[(self.plant, self.mode)][1]

    @property
    def controlled_outputs(self):
        return SHAPES[(self.plant, self.mode)][2]

    def block_end_records(self, Ts: float = 0.05, n_steps: int | None = None) -> np.ndarray:
        """Index of the first record that no longer belongs to each block.

        The reference driver advances `t += Ts` and runs a block `while t < t_end`
        (t_1000 = 49.9999999999993 < 50, so block 1 holds records
        0..1000); the accumulated floating-point sum is reproduced here.
        """
        ends = []
        t, k = 0.0, 0
        for t_end in self.sim_t_end:
            while t < t_end:
                t += Ts
                k += 1
                if n_steps is not None and k >= n_steps:
                    break
            ends.append(k)
        return np.asarray(ends, dtype=np.int32)


class _Reader:
    def __init__(self, text: str):
        self.lines = text.splitlines()
        self.pos = 0

    def _next_line(self):
        while self.pos < len(self.lines):
            line = self.lines[self.pos]
            self.pos += 1
            if line.startswith("#") or not line.strip():
                continue
            return line
        return None

    def read_string(self) -> str:
        line = self._next_line()
        if line is None:
            raise RuntimeError(f"Error reading setup file at line number{self.pos}")
        toks = line.split()
        if len(toks) > 1:
            print(f'Extra text "{toks[1]}" in line {self.pos} being ignored.', file=sys.stderr)
        return toks[0]

    def read_numbers(self, n: int, throw_error: bool = True) -> List[float]:
        out: List[float] = []
        rest: List[str] = []
        while len(out) < n:
            line = self._next_line()
            if line is None:
                if throw_error:
                    raise RuntimeError(
                        f"Error reading setup file at line number{self.pos}"
                        " (probably not enough entries given)")
                break
            toks = line.split()
            rest = []
            for i, tok in enumerate(toks):
                try:
                    out.append(float(tok))
                except ValueError:
                    rest = toks[i:]
                    break
                if len(out) == n:
                    rest = toks[i + 1:]
                    break
        if rest:
            print(f'Extra text "{rest[0]}" in line {self.pos} being ignored.', file=sys.stderr)
        return out


def parse_setup(text: str, plant: int, mode: int) -> Setup:
    n_ctrl, n_sub, outs = SHAPES[(plant, mode)]
    n_in = N_INPUTS[plant]
    r = _Reader(text)

    def key(expected):
        k = r.read_string()
        if k != expected:
            raise RuntimeError(f"setup file: expected key '{expected}', found '{k}'")

    key("n-iterations"); n_it = int(r.read_numbers(1)[0])
    key("n-timing-iterations"); n_tim = int(r.read_numbers(1)[0])
    key("folder-name"); folder = r.read_string()
    key("output-filename"); fname = r.read_string()
    key("yref"); yref = np.array(r.read_numbers(4))
    key("uwt"); uwt = np.array(r.read_numbers(16)).reshape(4, 4)
    key("ywt")
    ywt = []
    for c in range(n_ctrl):
        ny = len(outs[c])
        ywt.append(np.array(r.read_numbers(ny * ny)).reshape(ny, ny))
    key("constraints-lower"); lo = np.array(r.read_numbers(n_sub))
    key("constraints-upper"); up = np.array(r.read_numbers(n_sub))
    key("constraints-rate-lower"); rlo = np.array(r.read_numbers(n_sub))
    key("constraints-rate-upper"); rup = np.array(r.read_numbers(n_sub))
    key("simulation")
    offs, tends = [], []
    while True:
        vals = r.read_numbers(n_in + 1, throw_error=False)
        if len(vals) < n_in + 1:
            break
        offs.append(vals[:n_in])
        tends.append(vals[n_in])
    if not offs:
        raise RuntimeError("setup file: no simulation block")
    return Setup(plant, mode, n_it, n_tim, folder, fname, yref, uwt, ywt, lo, up, rlo, rup,
                 np.array(offs), np.array(tends))


def format_setup(s: Setup) -> str:
    """Write a Setup back in the reference's file format."""
    def mat(m):
        return "\n".join("\t".join(repr(float(v)) for v in row) for row in np.atleast_2d(m))
    out = [f"n-iterations\n{s.n_iterations}\n", f"n-timing-iterations\n{s.n_timing_iterations}\n",
           f"folder-name\n{s.folder_name}\n", f"output-filename\n{s.output_filename}\n",
           "yref\n" + mat(s.yref) + "\n", "uwt\n" + mat(s.uwt) + "\n",
           "ywt\n" + "\n\n".join(mat(w) for w in s.ywt) + "\n",
           "constraints-lower\n" + mat(s.lower) + "\n", "constraints-upper\n" + mat(s.upper) + "\n",
           "constraints-rate-lower\n" + mat(s.rate_lower) + "\n",
           "constraints-rate-upper\n" + mat(s.rate_upper) + "\n", "simulation"]
    for off, te in zip(s.sim_offsets, s.sim_t_end):
        out.append(mat(off) + "\n" + repr(float(te)) + "\n")
    return "\n".join(out)


def setup_to_dict(s: Setup) -> dict:
    d = dataclasses.asdict(s)
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            d[k] = v.tolist()
    d["ywt"] = [np.asarray(w).tolist() for w in s.ywt]
    return d


def setup_from_dict(d: dict) -> Setup:
    d = dict(d)
    for k in ("yref", "uwt", "lower", "upper", "rate_lower", "rate_upper", "sim_offsets", "sim_t_end"):
        d[k] = np.asarray(d[k], dtype=np.float64)
    d["ywt"] = [np.asarray(w, dtype=np.float64) for w in d["ywt"]]
    return Setup(**d)
