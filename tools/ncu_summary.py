"""Print the roofline-relevant metrics of an .ncu-rep (first result) as a small table.
usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    short = {h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] == "TriageCompute" else h: i
             for i, h in enumerate(hdr)}
    full = {h: i for i, h in enumerate(hdr)}
    print("## %s" % path)
    for k in KEYS:
        i = full.get(k, short.get(k))
        if i is not None:
            print("%-86s %s %s" % (k, vals[i][:90], units[i]))
    print()
