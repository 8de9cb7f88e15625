"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into small JSON files under profiles/ (the metrics the
profiling recipe names: duration, DRAM bytes / throughput, instruction and pipe utilisation, occupancy, stall reasons)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_integer_pred_on.sum",
]


def summarise(rep, note):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        raise SystemExit(f"{rep}: no data")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"report": rep.split("/")[-1], "note": note, "kernel": vals[hdr.index("Kernel Name")]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            v = vals[i].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            d[k] = {"value": v, "unit": units[i]}
    return d


if __name__ == "__main__":
    rep, dst, note = sys.argv[1], sys.argv[2], " ".join(sys.argv[3:])
    json.dump(summarise(rep, note), open(dst, "w"), indent=1)
    print(dst)
