"""Box-side summary of one `ncu --set full` report: key raw metrics + hottest source lines, as text.
Usage: python profiles/summarize_rep.py <report.ncu-rep> <out.txt> [title]   (the .ncu-rep files themselves are too
large to bring back from the GPU box; these summaries are what profiles/ keeps)."""
import csv
import os
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
if len(rr) >= 3:
    Hh, U, V = rr[0], rr[1], rr[2]
    out.append("# %s: ncu --set full --clock-control none (%s)" % (title, V[Hh.index("Kernel Name")][:120] if "Kernel Name" in Hh else ""))
    for i, h in enumerate(Hh):
        if h in keep or "dmma" in h.lower():
            out.append("%-92s %-16s %s" % (h, U[i], V[i]))
src = subprocess.run([sys.executable, os.path.join(root, "profiles", "ncu_source_summary.py"), rep, "22"],
                     capture_output=True, text=True).stdout
out.append("")
out.append("# hottest source lines (warp-state samples)")
out.append(src)
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
