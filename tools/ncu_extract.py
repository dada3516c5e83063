#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) as JSON: the counters DESIGN.md and bench.py quote.

    python tools/ncu_extract.py out.json rep1.ncu-rep [rep2.ncu-rep ...]
"""
import csv, io, json, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    return x * SCALE[unit] if unit in SCALE else x


def main():
    out = {}
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        H, U = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(H, r)), dict(zip(H, U))
            e = {"kernel": d["Kernel Name"], "report": rep.split("/")[-1]}
            for w in WANT:
                if w in d:
                    unit = u[w]
                    e[w] = num(d[w], unit)
                    if unit and unit not in SCALE:
                        e[w + " [unit]"] = unit
            e["duration_us"] = e.get("gpu__time_duration.sum")
            e["dram_bytes_per_launch"] = e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
            key = rep.split("/")[-1].replace(".ncu-rep", "") + ":" + d["ID"]
            out[key] = e
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    for k, e in out.items():
        print(k, e["kernel"][:60], "| %.1f us" % e["duration_us"], "| dram %.1f MB" % (e["dram_bytes_per_launch"] / 1e6),
              "| dmma %.1f%% fp64 %.1f%% lsu-wavefronts %.1f%% issue %.1f%%" % (
                  e.get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", 0),
                  e.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0),
                  e.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0),
                  e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)),
              "| regs", e.get("launch__registers_per_thread"), "| conflicts", e.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))


if __name__ == "__main__":
    main()
