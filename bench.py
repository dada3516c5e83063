#!/usr/bin/env python
"""Headline benchmark: batched closed-loop MPC steps/s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)

Workload at every N: setup-coop-par (cooperative MPC, parallel compressors + tank, p = 100,
9 Jacobi sweeps) batched over 4096 perturbed scenarios PER GPU (BASELINE.json configs[3];
scenarios are independent, so ranks hold disjoint shards and there is no data-path
collective: weak scaling).  One "step" = one closed-loop sample of every scenario of the
batch: control step (observer, linearisation, prediction, QP build, 9 sweeps of QP solves,
a-priori update) + plant advance (actuator delay + Dormand-Prince over 50 ms).
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
    bs = args.ref_scenarios
    W, K = args.warmup, args.steps
    x0, be, bo = pkg.scenarios.make_scenarios(setup, x_def, bs, W + K)
    o = ol.Oracle(setup, p=args.p)
    o.run_closed_loop(x0, be, bo, max(W, 1), n_threads=cores)          # warm-up
    t0 = time.perf_counter()
    o.run_closed_loop(x0, be, bo, K, n_threads=cores)
    dt = time.perf_counter() - t0
    val = bs * K / dt
    sample = f"{bs} of the 4096 scenarios x {K} closed-loop steps, {cores} host threads"
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"setup-{args.case} x {bs} scenarios (bounded sample), p={args.p}, CPU oracle port",
                       "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the control step has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    pkg = entry.load_package()
    setup = load_setup(pkg, args.case)
    x_def, u_def = pkg.plant_defaults(setup.plant)
    B, W, K, p = args.batch, args.warmup, args.steps, args.p
    T = W + 3 * K
    n, nin = len(x_def), len(u_def)
    rec = 1 + n + 8
    x0, be, bo = pkg.scenarios.make_scenarios(setup, x_def, B, T, first=rank * B)
    n_it = int(os.environ["CMPC_BENCH_NITER"]) if "CMPC_BENCH_NITER" in os.environ else None   # experiments only
    nc = pkg.from_setup(setup, batch=B, p=p, device=local, n_solver_iterations=n_it)
    ncz = nc.n_controllers

    d_x0 = torch.from_numpy(x0).to(dev)
    d_be = torch.from_numpy(be).to(dev)
    d_bo = torch.from_numpy(bo).to(dev)
    d_traj = torch.zeros((B, T, rec), dtype=torch.float64, device=dev)
    d_act = torch.zeros((B, T, ncz), dtype=torch.int32, device=dev)
    d_obj = torch.zeros((B, T, ncz), dtype=torch.float64, device=dev)
    d_st = torch.zeros((B, T, ncz), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def advance(first, count):
        nc.run_closed_loop_device(first, count, T, d_x0.data_ptr(), be.shape[1], d_be.data_ptr(), d_bo.data_ptr(),
                                  d_traj.data_ptr(), d_act.data_ptr(), d_obj.data_ptr(), d_st.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident closed loop: W warm-up steps, then exactly K timed steps -------------
    advance(0, W)
    torch.cuda.synchronize()
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = nc.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        ev_s[k].record()
        advance(W + k, 1)
        ev_e[k].record()
    torch.cuda.synchronize()
    barrier()
    gpu_launches = nc.launch_count() - launches0
    per_step_ms = np.array([s.elapsed_time(e) for s, e in zip(ev_s, ev_e)])
    total_ms = torch.tensor([per_step_ms.sum()], dtype=torch.float64, device=dev)
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
    barrier()
    torch.cuda.synchronize()
    e0.record()
    advance(W + 2 * K, K)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    b2b_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(b2b_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item()); b2b_ms = float(b2b_ms.item())
    value = world * B * K / (total_ms * 1e-3)

    # closed-loop health: every QP solved, trajectories finite
    st_bad = int((d_st[:, : W + 3 * K] != 0).sum().item())
    finite = bool(torch.isfinite(d_traj).all().item())

    # ---- e2e: the reference-facing call with HOST buffers, copies inside the timed region ------
    y_seq = d_traj[:, : W + K, 1 + n + 4:].permute(1, 0, 2).contiguous().cpu().pin_memory()   # (T, B, 4)
    u_seq = d_traj[:, : W + K, 1 + n: 1 + n + 4].permute(1, 0, 2).contiguous().cpu()
    u_host = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    y0 = y_seq[0].numpy()
    nc.Initialize(x0, np.zeros(4), u_def, y0)
    e2e_s, max_du = 0.0, 0.0
    y_ptrs = [y_seq[k].data_ptr() for k in range(W + K)]   # pinned host rows; no tensor indexing in the timed region
    u_ptr = u_host.data_ptr()
    barrier()
    for k in range(W + K):
        if k >= W:
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        nc.GetNextInputRaw(y_ptrs[k], u_ptr)   # H2D of y, the three kernels, D2H of u, sync
        if k >= W:
            e2e_s += time.perf_counter() - t0
            max_du = max(max_du, float((u_host - u_seq[k]).abs().max()))
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(e2e_t.item())

    # ---- NCCL: gather the last closed-loop record of every shard (the only collective) ---------
    last = d_traj[:, W + 3 * K - 1, :].contiguous()
    if world > 1:
        allrec = torch.empty((world * B, rec), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allrec, last)
    else:
        allrec = last
    checksum = float(allrec[:, 1 + n: 1 + n + 4].sum().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (control step): FP64 pipe ------------------------------
    peak = pkg.measure_fp64_peak(local)
    peak_tf = max(peak["dfma_tflops"], peak["dmma_m8n8k4_tflops"])
    flops_step = FLOPS_PER_STEP.get((args.case, p))
    kern_ms = assemble_ms / max(n_timed, 1)          # dominant kernel: assemble_kernel
    ctrl_ms = step_kernel_ms / max(n_timed, 1)      # whole control step (3 kernels)
    achieved_tf = flops_step * B / (kern_ms * 1e-3) / 1e12 if flops_step else None
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak = json.loads(peaks_file.read_text()).get("hbm_gbs") if peaks_file.exists() else 6650.0
    alg_bytes = B * (ncz * 2 * 1024 + 2 * 128 + 64 + ncz * (nc.nv * nc.nv + nc.nv + nc.nv * max(nc.nvo, 1) + 3) * 8)
    traffic = None
    prof = ROOT / "profiles" / "r01_step_kernel_traffic.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # "tensor": the kernel's dense algebra runs on the FP64 tensor cores (DMMA), which share the SM's
    # FP64 pipe with DFMA; the denominator is that pipe's measured peak, not the bf16 figure
    roofline = {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 + DFMA)",
                "kernel": "assemble_kernel (discretise + predict + QP assembly)", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                "peak_source": "measured live by cmpc_measure_fp64_peak (DFMA %.1f, DMMA m8n8k4 %.1f TF); "
                               "MEASURED_PEAKS.json holds no FP64 figure" % (peak["dfma_tflops"], peak["dmma_m8n8k4_tflops"]),
                "flops_per_unit": flops_step, "units_per_launch": B, "kernel_ms": kern_ms,
                "control_step_ms": ctrl_ms, "control_step_frac_of_peak": (flops_step * B / (ctrl_ms * 1e-3) / 1e12 / peak_tf) if flops_step else None,
                "kernel_share_of_step": assemble_ms / per_step_ms.sum(),
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak}}

    # ---- CPU baseline on this box's host cores (oracle port), bounded sample ----------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as ol
        cores = os.cpu_count() or 1
        bs, ks = args.cpu_sample, min(K, 200)
        o = ol.Oracle(setup, p=p)
        t0 = time.perf_counter()
        ref = o.run_closed_loop(x0[:bs], be[:bs], bo[:bs], W + ks, n_threads=cores)
        dt = time.perf_counter() - t0
        # the same sample doubles as a parity check of the timed GPU run
        g = d_traj[:bs, : W + ks].cpu().numpy()
        uo, ug = ref["traj"][:, :, 1 + n:5 + n], g[:, :, 1 + n:5 + n]
        parity = float(np.max(np.abs(ug - uo) / np.maximum(np.abs(uo), 1e-3)))
        act_equal = bool(np.array_equal(d_act[:bs, : W + ks].cpu().numpy().astype(np.uint32), ref["active"]))
        t1 = time.perf_counter()
        ol.Oracle(setup, p=p).run_closed_loop(x0[:1], be[:1], bo[:1], W + ks, n_threads=1)
        dt1 = time.perf_counter() - t1
        cpu_baseline = {"value": bs * (W + ks) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {bs} scenarios x {W + ks} closed-loop steps on {cores} host threads",
                        "single_thread_steps_per_s": (W + ks) / dt1,
                        "gpu_vs_oracle_max_rel_err_u": parity, "active_sets_identical": act_equal}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"setup-{args.case} (BASELINE configs[3]) x {B} perturbed scenarios per GPU, p={p}, "
                                   f"{nc.cfg.n_iterations} sweeps, closed loop (control step + plant advance)",
                       "scenarios_per_gpu": B, "p": p, "l2": "flushed (256 MiB memset) between timed iterations",
                       "parallelism": f"scenario shards x{world}, no data-path collective"},
            "p50_step_ms": float(np.median(per_step_ms)), "p99_step_ms": float(np.percentile(per_step_ms, 99)),
            "value_back_to_back": world * B * K / (b2b_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 32, "d2h_bytes_per_step": B * 32,
                    "api": "cmpc_get_next_input (host y -> host u, pinned)", "max_abs_du_vs_device_run": max_du},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": clocks, "health": {"qp_failures": st_bad, "finite": finite, "u_checksum": checksum}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
