#!/usr/bin/env python
"""Headline benchmark: batched closed-loop MPC steps/s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)

Headline workload at every N: setup-coop-par (cooperative MPC, parallel compressors + tank, p = 100,
9 Jacobi sweeps) batched over 4096 perturbed scenarios PER GPU (BASELINE.json configs[3];
scenarios are independent, so ranks hold disjoint shards and there is no data-path
collective: weak scaling).  One "step" = one closed-loop sample of every scenario of the
batch: control step (observer, linearisation, prediction, QP build, 9 sweeps of QP solves,
a-priori update) + plant advance (actuator delay + Dormand-Prince over 50 ms).

The same JSON line also carries
  sweep      BASELINE.json configs[4]: coop-par at twice the horizon (p = 200), 65 536 scenarios IN
             TOTAL, sharded 65 536 / N per GPU (strong scaling), with the roofline of its own
             assemble_kernel instantiation and an oracle sample on every rank
  configs    BASELINE.json configs[0-2] (cent-ser, coop-ser, ncoop-par) and the other three setups
             as B = 1 runs: closed-loop rate, p50 / p99 latency of one cmpc_get_next_input call,
             next to the reference's own recorded per-step time (N = 1 only)
  gather     SURVEY.md 8(e): the trajectory tensor all-gathered over the ranks in chunks (NCCL),
             timed on its own; health counters all-reduced over the ranks
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

METRIC = "closed-loop MPC steps/sec (batched scenario-steps/s)"
UNIT = "steps/s"
# contract FP64 flops per plant-step, keyed by (case, p)
FLOPS_PER_STEP = {("coop-par", 100): 418414, ("coop-par", 200): 812014, ("cent-ser", 100): 285652,
                  ("coop-ser", 100): 515278, ("ncoop-par", 100): 280814, ("cent-par", 100): 230020}
SWEEP_SCENARIOS = 65536      # BASELINE.json configs[4]
SWEEP_P = 200
ALL_CASES = ["cent-ser", "coop-ser", "ncoop-par", "cent-par", "coop-par", "ncoop-ser"]   # configs[0-2] first


def load_setup(pkg, case):
    raw = json.loads((ROOT / "tests" / "golden" / "setups.json").read_text())
    return pkg.setupfile.setup_from_dict(raw[case])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.6 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU implementation of the path: the oracle port (the reference cannot
    be compiled here: it needs Eigen, Boost and qpOASES), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol
    setup = load_setup(pkg, args.case)
    x_def, _ = ol.plant_defaults(setup.plant)
    cores = os.cpu_count() or 1
    W, K = args.warmup, args.steps
    # a bounded sample of the 4096 scenarios, sized for about 10 s of CPU work (3 k steps/s per thread) whatever K is,
    # so that thread start-up does not decide the figure
    bs = min(4096, max(args.ref_scenarios, -(-32768 // max(K, 1))))
    x0, be, bo = pkg.scenarios.make_scenarios(setup, x_def, bs, W + K)
    o = ol.Oracle(setup, p=args.p)
    o.run_closed_loop(x0, be, bo, max(W, 1), n_threads=cores)          # warm-up
    t0 = time.perf_counter()
    o.run_closed_loop(x0, be, bo, K, n_threads=cores)
    dt = time.perf_counter() - t0
    val = bs * K / dt
    sample = f"{bs} of the 4096 scenarios x {K} closed-loop steps (control step + plant advance), {cores} host threads"
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"setup-{args.case} x {bs} scenarios (bounded sample), p={args.p}, CPU oracle port",
                       "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Dist:
    """torch.distributed over NCCL when launched under torchrun, a no-op at N = 1."""

    def __init__(self, torch, local):
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.dev = torch.device("cuda", local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, value, op):
        """all-reduce of one float over the ranks; op in {"max", "min", "sum"}"""
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                    "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def chunked_all_gather(torch, dist, world, traj, chunk, sink):
    """the trajectory tensor [B, T, rec] of every rank all-gathered `chunk` records
    at a time; sink(k0, out) sees each gathered chunk as [world, B, chunk, rec] (rank-major, i.e.
    global scenario order).  Returns (calls, records covered).  Backend-agnostic (NCCL in the bench,
    gloo in the CPU test)."""
    B, T, rec = traj.shape
    out = torch.empty((world * B, chunk, rec), dtype=traj.dtype, device=traj.device)
    calls = 0
    for k0 in range(0, T - T % chunk, chunk):
        dist.all_gather_into_tensor(out, traj[:, k0:k0 + chunk].contiguous())
        sink(k0, out.view(world, B, chunk, rec))
        calls += 1
    return calls, calls * chunk


def sweep_shard(total, world, rank):
    """Scenarios [first, first + count) of the configs[4] sweep owned by `rank`."""
    count = total // world
    return rank * count, count


def traffic_for(case, p, batch):
    """DRAM bytes per assemble_kernel launch from the ncu --set full capture of this configuration
    (profiles/traffic.json, one entry per captured configuration; None when there is none)."""
    f = ROOT / "profiles" / "traffic.json"
    if not f.exists():
        return None, None
    try:
        e = json.loads(f.read_text()).get(f"{case}:p{p}:B{batch}")
    except Exception:
        return None, None
    return (e.get("dram_bytes_per_launch"), e.get("source")) if e else (None, None)


def oracle_sample(pkg, setup, p, x0, be, bo, n_rec, d_traj, d_act, n, threads):
    """The first scenarios of this rank's shard on the CPU oracle: timing and parity of the GPU run."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol
    o = ol.Oracle(setup, p=p)
    t0 = time.perf_counter()
    ref = o.run_closed_loop(x0, be, bo, n_rec, n_threads=threads)
    dt = time.perf_counter() - t0
    bs = x0.shape[0]
    g = d_traj[:bs, :n_rec].cpu().numpy()
    uo, ug = ref["traj"][:, :, 1 + n:5 + n], g[:, :, 1 + n:5 + n]
    err = float(np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3)))
    act_equal = bool(np.array_equal(d_act[:bs, :n_rec].cpu().numpy().astype(np.uint32), ref["active"]))
    return dt, err, act_equal


def closed_loop_buffers(torch, dev, pkg, setup, x_def, B, T, first, ncz):
    x0, be, bo = pkg.scenarios.make_scenarios(setup, x_def, B, T, first=first)
    rec = 1 + len(x_def) + 8
    d = dict(x0=x0, be=be, bo=bo,
             d_x0=torch.from_numpy(x0).to(dev), d_be=torch.from_numpy(be).to(dev), d_bo=torch.from_numpy(bo).to(dev),
             traj=torch.zeros((B, T, rec), dtype=torch.float64, device=dev),
             act=torch.zeros((B, T, ncz), dtype=torch.int32, device=dev),
             obj=torch.zeros((B, T, ncz), dtype=torch.float64, device=dev),
             st=torch.zeros((B, T, ncz), dtype=torch.int32, device=dev))
    return d


def fp64_roofline(pkg, local, case, p, B, kern_ms, ctrl_ms, share, peak):
    peak_tf = max(peak["dfma_tflops"], peak["dmma_m8n8k4_tflops"])
    flops_step = FLOPS_PER_STEP.get((case, p))
    achieved_tf = flops_step * B / (kern_ms * 1e-3) / 1e12 if flops_step else None
    traffic, traffic_src = traffic_for(case, p, B)
    # "tensor": the kernel's dense algebra runs on the FP64 tensor cores (DMMA), which share the SM's
    # FP64 pipe with DFMA; the denominator is that pipe's measured peak, not the bf16 figure
    return {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 + DFMA)",
            "kernel": f"assemble_kernel<{case}, p={p}> (discretise + predict + QP assembly)",
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured live by cmpc_measure_fp64_peak (DFMA %.1f, DMMA m8n8k4 %.1f TF); "
                           "MEASURED_PEAKS.json holds no FP64 figure" % (peak["dfma_tflops"], peak["dmma_m8n8k4_tflops"]),
            "flops_per_unit": flops_step, "units_per_launch": B, "kernel_ms": kern_ms, "control_step_ms": ctrl_ms,
            "control_step_frac_of_peak": (flops_step * B / (ctrl_ms * 1e-3) / 1e12 / peak_tf) if flops_step else None,
            "kernel_share_of_step": share}


def run_sweep(args, torch, D, pkg, local, flush, peak):
    """BASELINE.json configs[4] / SURVEY.md 8d config (5): coop-par, p = 200, 65 536 scenarios in
    total, 65 536 / N per GPU.  Records are timed one by one with an L2 flush in between, like the
    headline; a second pass with the library's events gives the assemble kernel's own duration."""
    world, rank, dev = D.world, D.rank, D.dev
    total = args.sweep_scenarios
    first, B = sweep_shard(total, world, rank)
    K, W, KT = args.sweep_steps, 5, min(args.sweep_steps, 40)
    T = W + K + KT
    setup = load_setup(pkg, "coop-par")
    x_def, _ = pkg.plant_defaults(setup.plant)
    n = len(x_def)
    nc = pkg.from_setup(setup, batch=B, p=SWEEP_P, device=local)
    buf = closed_loop_buffers(torch, dev, pkg, setup, x_def, B, T, first, nc.n_controllers)
    stream = torch.cuda.current_stream().cuda_stream

    def advance(first_rec, count):
        nc.run_closed_loop_device(first_rec, count, T, buf["d_x0"].data_ptr(), buf["be"].shape[1], buf["d_be"].data_ptr(),
                                  buf["d_bo"].data_ptr(), buf["traj"].data_ptr(), buf["act"].data_ptr(),
                                  buf["obj"].data_ptr(), buf["st"].data_ptr(), stream)

    advance(0, W)
    torch.cuda.synchronize()
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    D.barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()
        ev_s[k].record()
        advance(W + k, 1)
        ev_e[k].record()
    torch.cuda.synchronize()
    D.barrier()
    per_ms = np.array([s.elapsed_time(e) for s, e in zip(ev_s, ev_e)])
    total_ms = D.reduce(per_ms.sum(), "max")
    nc.set_timing(True)
    for k in range(KT):
        flush.zero_()
        advance(W + K + k, 1)
    torch.cuda.synchronize()
    n_timed, step_ms, asm_ms = nc.get_timing()
    nc.set_timing(False)
    kern_ms = D.reduce(asm_ms / max(n_timed, 1), "max")
    ctrl_ms = D.reduce(step_ms / max(n_timed, 1), "max")
    failures = D.reduce(int((buf["st"] != 0).sum().item()), "sum")
    finite = D.reduce(1.0 if bool(torch.isfinite(buf["traj"]).all().item()) else 0.0, "min")
    bs, nrec = args.sweep_oracle_scenarios, min(T, 30)
    _, err, act_equal = oracle_sample(pkg, setup, SWEEP_P, buf["x0"][:bs], buf["be"][:bs], buf["bo"][:bs], nrec,
                                      buf["traj"], buf["act"], n, threads=2)
    err = D.reduce(err, "max")
    act_equal = D.reduce(1.0 if act_equal else 0.0, "min")
    out = {"workload": f"setup-coop-par (BASELINE configs[4]) p={SWEEP_P}, {total} scenarios in total, "
                       f"{B} per GPU x {world} GPU(s), 9 sweeps, closed loop (control step + plant advance)",
           "scaling": "strong", "value": total * K / (total_ms * 1e-3), "unit": UNIT, "steps": K, "warmup": W,
           "ms_per_step": total_ms / K, "p50_step_ms": float(np.median(per_ms)), "scenarios_total": total,
           "scenarios_per_gpu": B, "p": SWEEP_P, "l2": "flushed (256 MiB memset) between timed records",
           "roofline": fp64_roofline(pkg, local, "coop-par", SWEEP_P, B, kern_ms, ctrl_ms,
                                     kern_ms / (total_ms / K), peak),
           "health": {"qp_failures": int(failures), "finite": bool(finite),
                      "oracle_sample": f"first {bs} scenarios of every rank x {nrec} records",
                      "gpu_vs_oracle_max_rel_err_u": err, "active_sets_identical": bool(act_equal)}}
    nc.close()
    return out


def run_b1_configs(args, torch, dev, pkg, local):
    """The six setups as B = 1 runs (BASELINE.json configs[0-2] and the other three): closed-loop rate
    of the device-resident loop, latency of one host-facing cmpc_get_next_input call, the recorded
    run of the reference as the check, and the reference's own recorded time per step beside it."""
    golden = np.load(ROOT / "tests" / "golden" / "golden_traj.npz")
    out = {}
    T, NL = args.b1_records, args.b1_latency_steps
    stream = torch.cuda.current_stream().cuda_stream
    for case in ALL_CASES:
        setup = load_setup(pkg, case)
        x_def, u_def = pkg.plant_defaults(setup.plant)
        n = len(x_def)
        nc = pkg.from_setup(setup, batch=1, device=local)
        buf = closed_loop_buffers(torch, dev, pkg, setup, x_def, 1, T, 0, nc.n_controllers)

        def advance(first, count):
            nc.run_closed_loop_device(first, count, T, buf["d_x0"].data_ptr(), buf["be"].shape[1],
                                      buf["d_be"].data_ptr(), buf["d_bo"].data_ptr(), buf["traj"].data_ptr(),
                                      buf["act"].data_ptr(), buf["obj"].data_ptr(), buf["st"].data_ptr(), stream)
        advance(0, 20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        advance(20, T - 20)
        e1.record()
        torch.cuda.synchronize()
        loop_ms = e0.elapsed_time(e1) / (T - 20)
        traj = buf["traj"][0].cpu().numpy()
        idx = golden[f"{case}/index"]
        keep = idx < T
        gold = golden[f"{case}/records"][keep]
        du = float(np.max(np.abs(traj[idx[keep], 1 + n:5 + n] - gold[:, 1 + n:5 + n])))
        # host-facing step: replay the measurements of that run through cmpc_get_next_input
        y_seq = torch.from_numpy(np.ascontiguousarray(traj[:, 5 + n:9 + n])).pin_memory()
        u_host = torch.empty((1, 4), dtype=torch.float64).pin_memory()
        nc.Initialize(buf["x0"], np.zeros(4), u_def, y_seq[0].numpy())
        lat = []
        for k in range(min(T, NL + 20)):
            t0 = time.perf_counter()
            nc.GetNextInputRaw(y_seq[k].data_ptr(), u_host.data_ptr())
            t1 = time.perf_counter()
            if k >= 20:
                lat.append((t1 - t0) * 1e6)
        ref_ns = float(golden[f"{case}/mean_step_ns"])
        out[case] = {"workload": f"setup-{case}, B=1, p=100, {setup.n_iterations} sweep(s)",
                     "closed_loop_steps_per_s": 1e3 / loop_ms, "closed_loop_record_us": loop_ms * 1e3,
                     "get_next_input_p50_us": float(np.median(lat)), "get_next_input_p99_us": float(np.percentile(lat, 99)),
                     "latency_steps": len(lat), "records": T,
                     "max_abs_du_vs_reference_recorded_run": du,
                     "reference_recorded_step_us": ref_ns / 1e3,
                     "reference_recorded_on": "the reference's own results/*/run1 files (its authors' CPU, 2016)"}
        nc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--case", default="coop-par")
    ap.add_argument("--batch", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--p", type=int, default=100)
    ap.add_argument("--ref-scenarios", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=128, help="scenarios in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] sweep object")
    ap.add_argument("--sweep-scenarios", type=int, default=SWEEP_SCENARIOS, help="scenarios of the sweep IN TOTAL")
    ap.add_argument("--sweep-steps", type=int, default=200)
    ap.add_argument("--sweep-oracle-scenarios", type=int, default=8)
    ap.add_argument("--no-b1", action="store_true", help="skip the B = 1 runs of the six setups")
    ap.add_argument("--b1-records", type=int, default=600)
    ap.add_argument("--b1-latency-steps", type=int, default=400)
    ap.add_argument("--gather-chunk", type=int, default=16, help="records per all-gather call")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the control step has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    D = Dist(torch, local)
    world, rank, dev = D.world, D.rank, D.dev

    pkg = entry.load_package()
    setup = load_setup(pkg, args.case)
    x_def, u_def = pkg.plant_defaults(setup.plant)
    B, W, K, p = args.batch, args.warmup, args.steps, args.p
    T = W + 3 * K
    n, nin = len(x_def), len(u_def)
    rec = 1 + n + 8
    n_it = int(os.environ["CMPC_BENCH_NITER"]) if "CMPC_BENCH_NITER" in os.environ else None   # experiments only
    nc = pkg.from_setup(setup, batch=B, p=p, device=local, n_solver_iterations=n_it)
    ncz = nc.n_controllers
    buf = closed_loop_buffers(torch, dev, pkg, setup, x_def, B, T, rank * B, ncz)
    x0, be, bo = buf["x0"], buf["be"], buf["bo"]
    d_traj, d_act, d_st = buf["traj"], buf["act"], buf["st"]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def advance(first, count):
        nc.run_closed_loop_device(first, count, T, buf["d_x0"].data_ptr(), be.shape[1], buf["d_be"].data_ptr(),
                                  buf["d_bo"].data_ptr(), d_traj.data_ptr(), d_act.data_ptr(), buf["obj"].data_ptr(),
                                  d_st.data_ptr(), stream)

    # ---- device-resident closed loop: W warm-up steps, then exactly K timed steps -------------
    advance(0, W)
    torch.cuda.synchronize()
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = nc.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    D.barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        ev_s[k].record()
        advance(W + k, 1)
        ev_e[k].record()
    torch.cuda.synchronize()
    D.barrier()
    gpu_launches = nc.launch_count() - launches0
    per_step_ms = np.array([s.elapsed_time(e) for s, e in zip(ev_s, ev_e)])
    total_ms = D.reduce(per_step_ms.sum(), "max")
    # the same loop once more with the library's own events between its kernels: the per-kernel
    # durations behind the roofline (kept out of the timed run, whose kernels then follow each
    # other without an event in between)
    nc.set_timing(True)
    for k in range(K):
        flush.zero_()
        advance(W + K + k, 1)
    torch.cuda.synchronize()
    n_timed, step_kernel_ms, assemble_ms = nc.get_timing()
    nc.set_timing(False)
    # back-to-back variant (no L2 flush, one event pair)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    e0.record()
    advance(W + 2 * K, K)
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = D.reduce(e0.elapsed_time(e1), "max")
    value = world * B * K / (total_ms * 1e-3)

    # ---- e2e: the closed loop one record per call with HOST buffers (cmpc_closed_loop_step: this
    # sample's plant-input offsets go up from pinned memory, the record [t, x, u, y] comes back),
    # copies and the blocking wait inside the timed region.  Like for like with the CPU arm, whose
    # step is also control step + plant advance.
    blk = (np.arange(W + K)[None, :, None] >= be[:, None, :]).sum(axis=2)            # block of record k, per scenario
    blk = np.minimum(blk, be.shape[1] - 1)
    off_seq = torch.from_numpy(np.ascontiguousarray(np.take_along_axis(bo, blk[:, :, None], axis=1).transpose(1, 0, 2)))
    off_seq = off_seq.pin_memory()                                                      # (W+K, B, NIN)
    rec_host = torch.empty((B, rec), dtype=torch.float64).pin_memory()
    off_ptrs = [off_seq[k].data_ptr() for k in range(W + K)]
    rec_ptr = rec_host.data_ptr()
    ref_traj = d_traj[:, : W + K].cpu()
    nc.closed_loop_start(x0)
    e2e_s, e2e_mismatch = 0.0, 0.0
    D.barrier()
    for k in range(W + K):
        if k >= W:
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        nc.closed_loop_step_raw(off_ptrs[k], rec_ptr)   # H2D offsets, assemble/solve/advance, D2H record, sync
        if k >= W:
            e2e_s += time.perf_counter() - t0
            e2e_mismatch = max(e2e_mismatch, float((rec_host - ref_traj[:, k]).abs().max()))
    e2e_value = world * B * K / D.reduce(e2e_s, "max")
    # the same loop pipelined (cmpc_closed_loop_pipeline: the control step of record k + 1 is launched behind the
    # plant advance of record k and runs while the host handles record k), W + K calls back to back under one
    # clock: the calls overlap by design, so there is no place for an L2 flush between them -- reported next to
    # `e2e`, not instead of it
    nc.closed_loop_start(x0)
    nc.closed_loop_pipeline(True)
    pipe_mismatch = 0.0
    for k in range(W):
        nc.closed_loop_step_raw(off_ptrs[k], rec_ptr)
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        nc.closed_loop_step_raw(off_ptrs[k], rec_ptr)
    torch.cuda.synchronize()
    pipe_s = time.perf_counter() - t0
    pipe_mismatch = float((rec_host - ref_traj[:, W + K - 1]).abs().max())
    e2e_pipe_value = world * B * K / D.reduce(pipe_s, "max")
    nc.closed_loop_start(x0)
    nc.closed_loop_pipeline(False)

    # the control step alone through the reference-facing call (cmpc_get_next_input: host y -> host u),
    # replaying the measurements of the device run
    y_seq = d_traj[:, : W + K, 1 + n + 4:].permute(1, 0, 2).contiguous().cpu().pin_memory()   # (T, B, 4)
    u_seq = d_traj[:, : W + K, 1 + n: 1 + n + 4].permute(1, 0, 2).contiguous().cpu()
    u_host = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    nc.Initialize(x0, np.zeros(4), u_def, y_seq[0].numpy())
    ctl_s, max_du = 0.0, 0.0
    y_ptrs = [y_seq[k].data_ptr() for k in range(W + K)]
    u_ptr = u_host.data_ptr()
    for k in range(W + K):
        if k >= W:
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        nc.GetNextInputRaw(y_ptrs[k], u_ptr)   # H2D of y, the three kernels, D2H of u, sync
        if k >= W:
            ctl_s += time.perf_counter() - t0
            max_du = max(max_du, float((u_host - u_seq[k]).abs().max()))
    ctl_value = world * B * K / D.reduce(ctl_s, "max")
    clocks = sampler.stop() if sampler else None

    # ---- health over ALL ranks: QP failures, finiteness, an oracle sample of every shard --------
    st_bad = D.reduce(int((d_st != 0).sum().item()), "sum")
    finite = D.reduce(1.0 if bool(torch.isfinite(d_traj).all().item()) else 0.0, "min")
    cores = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        ks = min(K, 200)
        bs = min(B, max(args.cpu_sample, -(-32768 // (W + ks))))   # about 10 s of CPU work (see run_reference)
        dt, parity, act_equal = oracle_sample(pkg, setup, p, x0[:bs], be[:bs], bo[:bs], W + ks, d_traj, d_act, n, cores)
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as ol
        t1 = time.perf_counter()
        ol.Oracle(setup, p=p).run_closed_loop(x0[:1], be[:1], bo[:1], W + ks, n_threads=1)
        dt1 = time.perf_counter() - t1
        cpu_baseline = {"value": bs * (W + ks) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {bs} scenarios x {W + ks} closed-loop steps on {cores} host threads",
                        "single_thread_steps_per_s": (W + ks) / dt1,
                        "gpu_vs_oracle_max_rel_err_u": parity, "active_sets_identical": act_equal}
        sample_note = f"first {bs} scenarios x {W + ks} records (the cpu_baseline sample)"
    else:
        bs, ks = 8, min(K, 25)
        _, parity, act_equal = oracle_sample(pkg, setup, p, x0[:bs], be[:bs], bo[:bs], W + ks, d_traj, d_act, n,
                                             max(1, cores // max(world, 1)))
        sample_note = f"first {bs} scenarios of every rank x {W + ks} records"
    parity_all = D.reduce(parity, "max")
    act_all = D.reduce(1.0 if act_equal else 0.0, "min")

    # ---- NCCL: all-gather the trajectory tensor [B, T, 1+n+8] of every shard in
    # chunks of records, timed on its own; the checksum covers every record of every rank ---------
    gather = None
    checksum = float(d_traj[:, :, 1 + n: 1 + n + 4].sum().item())
    if world > 1:
        import torch.distributed as dist
        ch = args.gather_chunk
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        chunked_all_gather(torch, dist, world, d_traj[:, :ch], ch, lambda k0, out: None)   # warm-up (communicator set-up)
        torch.cuda.synchronize()
        D.barrier()
        acc = torch.zeros((), dtype=torch.float64, device=dev)

        def sink(k0, out):
            acc.add_(out[:, :, :, 1 + n: 1 + n + 4].sum())
        g0.record()
        n_calls, n_rec = chunked_all_gather(torch, dist, world, d_traj, ch, sink)
        g1.record()
        torch.cuda.synchronize()
        g_ms = D.reduce(g0.elapsed_time(g1), "max")
        g_bytes = n_calls * world * B * ch * rec * 8
        own = D.reduce(float(d_traj[:, :n_rec, 1 + n: 1 + n + 4].sum().item()), "sum")
        gather = {"collective": "ncclAllGather (torch.distributed all_gather_into_tensor)", "records": n_rec,
                  "chunk_records": ch, "calls": n_calls, "bytes_received_per_rank": g_bytes, "ms": g_ms,
                  "GB_per_s_per_rank": g_bytes / (g_ms * 1e-3) / 1e9,
                  "checksum_gathered": float(acc.item()), "checksum_sum_of_shards": own,
                  "note": "after the timed region; includes the strided pack of each chunk and a checksum of the result"}
        checksum = own

    # ---- sub-results ------------------------------------------------------------------------------
    peak = pkg.measure_fp64_peak(local)
    kern_ms = assemble_ms / max(n_timed, 1)          # dominant kernel: assemble_kernel
    ctrl_ms = step_kernel_ms / max(n_timed, 1)      # whole control step
    roofline = fp64_roofline(pkg, local, args.case, p, B, kern_ms, ctrl_ms, assemble_ms / per_step_ms.sum(), peak)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak = json.loads(peaks_file.read_text()).get("hbm_gbs") if peaks_file.exists() else 6650.0
    alg_bytes = B * (ncz * 2 * 1024 + 2 * 128 + 64 + ncz * (nc.nv * nc.nv + nc.nv + nc.nv * max(nc.nvo, 1) + 3) * 8)
    roofline["hbm"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                       "peak_gbs": hbm_peak, "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak}
    nv_, nvo_, n_iter_ = nc.nv, nc.nvo, nc.cfg.n_iterations
    nc.close()
    del buf, d_traj, d_act, d_st, ref_traj, y_seq, u_seq, off_seq
    torch.cuda.empty_cache()
    sweep = None if args.no_sweep else run_sweep(args, torch, D, pkg, local, flush, peak)
    configs = None
    if world == 1 and not args.no_b1:
        configs = run_b1_configs(args, torch, dev, pkg, local)

    if rank != 0:
        D.close()
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"setup-{args.case} (BASELINE configs[3]) x {B} perturbed scenarios per GPU, p={p}, "
                                   f"{n_iter_} sweeps, closed loop (control step + plant advance)",
                       "scenarios_per_gpu": B, "p": p, "l2": "flushed (256 MiB memset) between timed iterations",
                       "parallelism": f"scenario shards x{world}, no data-path collective"},
            "p50_step_ms": float(np.median(per_step_ms)), "p99_step_ms": float(np.percentile(per_step_ms, 99)),
            "value_back_to_back": world * B * K / (b2b_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nin * 8, "d2h_bytes_per_step": B * rec * 8,
                    "api": "cmpc_closed_loop_step (pinned host plant-input offsets -> control step + plant advance "
                           "on the device -> pinned host record [t, x, u, y]), blocking",
                    "max_abs_diff_vs_device_run": e2e_mismatch},
            "e2e_pipelined": {"value": e2e_pipe_value, "unit": UNIT, "h2d_bytes_per_step": B * nin * 8, "d2h_bytes_per_step": B * rec * 8,
                              "api": "cmpc_closed_loop_pipeline(1) + cmpc_closed_loop_step, all calls back to back under one clock "
                                     "(no L2 flush: consecutive calls overlap by design)",
                              "max_abs_diff_vs_device_run_last_record": pipe_mismatch},
            "e2e_control_step": {"value": ctl_value, "unit": UNIT, "h2d_bytes_per_step": B * 32, "d2h_bytes_per_step": B * 32,
                                 "api": "cmpc_get_next_input (host y -> host u, pinned): the control step alone, "
                                        "as the reference's ControllerInterface::GetNextInput", "max_abs_du_vs_device_run": max_du},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "health": {"qp_failures": int(st_bad), "finite": bool(finite), "u_checksum": checksum, "ranks_covered": world,
                       "oracle_sample": sample_note, "gpu_vs_oracle_max_rel_err_u": parity_all,
                       "active_sets_identical": bool(act_all)},
            "gather": gather, "sweep": sweep, "configs": configs}
    print(json.dumps(line), flush=True)
    D.close()


if __name__ == "__main__":
    main()
