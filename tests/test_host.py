"""CPU-only tests: setup-file reader, scenario generator, C-ABI surface, shard/gather logic."""
import ctypes
import os
import re
import pathlib
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from conftest import CASES, ROOT


def test_setup_file_roundtrip_and_tokeniser(pkg, setups):
    sf = pkg.setupfile
    for case in CASES:
        s = setups[case]
        text = sf.format_setup(s)
        # comments, blank lines and trailing text are tolerated like read_files.h:13-81
        noisy = "# header comment\n\n" + text.replace("yref\n", "yref\n# reference\n   \n", 1)
        s2 = sf.parse_setup(noisy, s.plant, s.mode)
        assert s2.n_iterations == s.n_iterations and s2.output_filename == s.output_filename
        for k in ("yref", "uwt", "lower", "upper", "rate_lower", "rate_upper", "sim_offsets", "sim_t_end"):
            assert np.array_equal(getattr(s, k), getattr(s2, k)), k
        assert all(np.array_equal(a, b) for a, b in zip(s.ywt, s2.ywt))
    with pytest.raises(RuntimeError):
        sf.parse_setup("n-iterations\n9\n", 0, 1)              # truncated file
    with pytest.raises(RuntimeError):
        sf.parse_setup(text.replace("uwt", "uwx"), s.plant, s.mode)  # wrong key


def test_block_end_records_match_reference_driver(setups):
    # records 0..1000 use block-1 offsets, 10 000 records in total
    for case in CASES:
        assert list(setups[case].block_end_records()) == [1001, 10000]


def test_scenarios_deterministic_and_shardable(pkg, setups):
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(0)
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, 64, 2000)
    assert np.array_equal(x0[0], x_def) and list(be[0]) == [1001, 2000] and bo[0, 1, 8] == -0.3
    assert np.abs(x0[:, [0, 1, 2, 3, 10]] / x_def[[0, 1, 2, 3, 10]] - 1).max() <= 0.01 + 1e-12
    assert (x0[:, 1] > x0[:, 0]).all() and (x0[:, 4] == 0).all()     # p2 > p1, recycle flow state 0
    assert (np.abs(be[:, 0] - 1001) <= 200).all() and (be[:, 1] == 2000).all()
    assert (bo[:, 1, 8] <= -0.15 + 1e-12).all() and (bo[:, 1, 8] >= -0.45 - 1e-12).all()
    # shards: rank r of 4 gets scenarios [16r, 16r+16)
    parts = [pkg.scenarios.make_scenarios(s, x_def, 16, 2000, first=16 * r) for r in range(4)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), x0)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), be)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), bo)


def test_c_abi_exports_every_declared_symbol(pkg):
    """include/cmpc.h vs the built library (no compute calls: there is no GPU here)."""
    header = (ROOT / "include" / "cmpc.h").read_text()
    declared = set(re.findall(r"\b(cmpc_[a-z0-9_]+)\s*\(", header))
    declared -= {"cmpc_handle", "cmpc_config", "cmpc_fp64_peak"}
    assert len(declared) >= 25
    lib = pkg.capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in cmpc.h but not exported"
    assert declared == set(pkg.capi.EXPORTED_SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", str(pkg.capi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cmpc_[a-z0-9_]+)", nm))
    assert declared <= exported
    # host-only entry points work without a GPU
    cfg = pkg.capi.default_config(0, 1, 4096)
    assert (cfg.p, cfg.m, cfg.n_controllers, cfg.n_sub_control_inputs, cfg.n_iterations) == (100, 2, 2, 2, 9)
    assert list(cfg.delays) == [0, 40, 0, 40] and list(cfg.control_input_indices[1]) == [2, 3, 0, 1]
    x, u = pkg.plant_defaults(1)
    assert np.array_equal(x, ol.plant_defaults(1)[0]) and np.array_equal(u, ol.plant_defaults(1)[1])
    # every (plant, mode) the library knows gives the output partition of setupfile.SHAPES
    for (plant, mode), (n_ctrl, n_sub, outs) in pkg.setupfile.SHAPES.items():
        c = pkg.capi.default_config(plant, mode, 8)
        assert (c.n_controllers, c.n_sub_control_inputs) == (n_ctrl, n_sub)
        for k in range(n_ctrl):
            assert list(c.controlled_output_indices[k])[: c.n_controlled_outputs[k]] == outs[k]
    with pytest.raises(pkg.capi.CmpcError):
        pkg.capi.default_config(0, pkg.setupfile.MODE_NCOOP_OLD, 8)   # the old partition is a serial-plant shape


def test_product_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CmpcError) as e:
        pkg.NerveCenter(0, 1, batch=4)
    assert "CMPC_ERR_CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or call anything under oracle/."""
    for path in list((ROOT / "compressor-mpc_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.suffix in {".py", ".cu", ".cuh", ".h", ".hpp"} or path.name == "Makefile":
            text = path.read_text()
            assert "oracle" not in text.lower(), f"{path} mentions the oracle"
    so = ROOT / "compressor-mpc_b200" / "libcmpc_b200.so"
    ldd = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def _gloo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import json
    import __graft_entry__ as entry
    import bench
    import oracle_lib as ol2
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    s = pkg.setupfile.setup_from_dict(json.loads((ROOT / "tests/golden/setups.json").read_text())["coop-par"])
    x_def, _ = ol2.plant_defaults(0)
    first, B = bench.sweep_shard(6, world, rank)      # the sweep's strong-scaling shard rule
    T = 25
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, B, T, first=first)
    res = ol2.Oracle(s).run_closed_loop(x0, be, bo, T)
    traj = torch.from_numpy(np.ascontiguousarray(res["traj"]))
    # bench.py's collective: the whole trajectory tensor, a chunk of records per all-gather
    full = torch.zeros((world * B, T, traj.shape[2]), dtype=torch.float64)

    def sink(k0, out):
        full[:, k0:k0 + out.shape[2]] = out.reshape(world * B, out.shape[2], out.shape[3])
    calls, n_rec = bench.chunked_all_gather(torch, dist, world, traj, 8, sink)
    # health counters over all ranks (bench.py Dist.reduce): failures summed, finiteness and-ed, time max-ed
    fails = torch.tensor([float((res["status"] != 0).sum() + rank)])
    dist.all_reduce(fails, op=dist.ReduceOp.SUM)
    finite = torch.tensor([1.0 if rank == 0 else 0.0])
    dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    t_max = torch.tensor([float(rank + 1)])
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)    # max-over-ranks timing reduction
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "scalars.npy"), np.array([t_max.item(), fails.item(), finite.item(), calls, n_rec]))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_gloo(pkg, setups, tmp_path):
    """The N > 1 host logic of bench.py on two CPU ranks: shard rule, chunked all-gather of the
    trajectory tensor in global scenario order, all-reduced health counters."""
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    s = setups["coop-par"]
    x_def, _ = ol.plant_defaults(0)
    x0, be, bo = pkg.scenarios.make_scenarios(s, x_def, 6, 25)
    ref = ol.Oracle(s).run_closed_loop(x0, be, bo, 25)["traj"]
    assert np.array_equal(got[:, :24], ref[:, :24])      # 3 chunks of 8 records
    assert (got[:, 24] == 0).all()                       # the ragged tail is not gathered
    t_max, fails, finite, calls, n_rec = np.load(tmp_path / "scalars.npy")
    assert (t_max, fails, finite, calls, n_rec) == (2.0, 1.0, 0.0, 3.0, 24.0)
    import bench
    assert [bench.sweep_shard(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 8192)]


def test_dat_record_format_roundtrip(pkg, golden):
    """The .dat writer prints records the way the reference does (6 significant digits, Eigen row
    format) and the reference's reader (whitespace separated numbers) gets them back."""
    wf = pkg.workflow
    rec = golden["coop-par/records"][:50]
    text = wf.format_records(rec, 11, 1176270)
    first = text.splitlines()[:6]
    assert first[0] == "0"
    assert first[1] == "0.916 1.145 0.152   440     0 0.916 1.145 0.152   440     0  1.12"   # as in results/parallel/run1/coop9.dat
    assert first[2] == "-4.48036e-05            0 -4.48036e-05            0"
    assert first[4] == "1176270" and first[5] == ""
    back = wf.parse_records(text, 11)
    assert back.shape == (50, 21) and np.allclose(back[:, :-1], rec, rtol=1e-5, atol=1e-12)
    assert wf.infer_test("folder-name\nserial\noutput-filename\nncoop9.dat\n") == (1, 2)
    assert wf.infer_test("folder-name\nparallel\noutput-filename\ncentralized.dat\n") == (0, 0)


def test_cxx_driver_built_and_reports_usage():
    import subprocess
    exe = ROOT / "compressor-mpc_b200" / "cmpc_run_setup"
    if not exe.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "compressor-mpc_b200")])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_read_timing_data_reproduces_reference_aggregation(pkg, tmp_path):
    """read_timing_data.m:18-62 as workflow.read_timing_data: a results tree whose files carry the
    per-file mean step times of the reference's own recorded runs (tests/golden/timing_golden.json,
    made from results/*/run1..5 by make_golden.py) aggregates to the reference's table: mean of the
    last column, distributed runs divided by 2, mean over the five runs."""
    import json
    wf = pkg.workflow
    gold = json.loads((ROOT / "tests" / "golden" / "timing_golden.json").read_text())
    assert len(gold["file_mean_ns"]) == 2 * 5 * 19
    for rel, mean_ns in gold["file_mean_ns"].items():
        folder = rel.split("/")[0]
        n = 11 if folder == "parallel" else 10
        m = int(round(mean_ns))
        traj = np.zeros((3, 1 + n + 8))
        f = tmp_path / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_text(wf.format_records(traj, n, np.array([m - 7, m, m + 7])))
    got = wf.read_timing_data(tmp_path)
    for folder in ("parallel", "serial"):
        for kind in ("cent", "coop", "ncoop"):
            want = np.array(gold["read_timing_data"][folder][kind])
            assert got[folder][kind].shape == (9,)
            assert np.abs(got[folder][kind] - want).max() <= 0.5, (folder, kind)     # rounding of the written integers
    # the reference's headline numbers: about 0.35 ms centralised, 0.36 - 0.45 ms per cooperative controller
    assert 3.4e5 < got["parallel"]["cent"][0] < 3.6e5 and 3.6e5 < got["parallel"]["coop"][0] < 3.7e5


def test_set_setup_params_edits_like_run_all_tests(pkg, setups):
    """setup/run-all-tests.sh:39-47 (two gawk edits): the value lines after the n-timing-iterations
    and output-filename keys are replaced, nothing else changes."""
    wf, sf = pkg.workflow, pkg.setupfile
    s = setups["ncoop-ser"]
    text = sf.format_setup(s)
    s2 = sf.parse_setup(wf.set_setup_params(text, 4, "ncoop4.dat"), s.plant, s.mode)
    assert (s2.n_timing_iterations, s2.output_filename, s2.n_iterations) == (4, "ncoop4.dat", s.n_iterations)
    assert np.array_equal(s2.uwt, s.uwt) and np.array_equal(s2.sim_offsets, s.sim_offsets)
    assert [t[0] for t in wf.ALL_TESTS_CENT + wf.ALL_TESTS_DIST] == [
        "setup-cent-par", "setup-cent-ser", "setup-coop-par", "setup-coop-ser", "setup-ncoop-par", "setup-ncoop-ser"]
