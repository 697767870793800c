"""Writes profiles/<tag>_*.txt from gpurun_out/*.ncu-rep + launch list.  Usage: python profiles/summarize_ncu.py <tag>
(expects gpurun_out/launches_<tag>.csv, gpurun_out/prof_fwd_<tag>.ncu-rep, gpurun_out/prof_bwd_<tag>.ncu-rep)"""
import csv, collections, subprocess, sys, os
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
out = []
rows = list(csv.reader(open(os.path.join(go, "launches_%s.csv" % tag))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    tot[r[ki][:90]] += v; cnt[r[ki][:90]] += 1
T = sum(tot.values())
out.append("# launch list (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache: compare SHARES)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    out.append("%-92s n=%3d total %9.3f ms  avg %8.3f ms  %5.1f%%" % (k, cnt[k], v / 1e6, v / 1e6 / cnt[k], 100 * v / T))
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active"]
for which in ("fwd", "bwd"):
    rep = os.path.join(go, "prof_%s_%s.ncu-rep" % (which, tag))
    if not os.path.exists(rep): continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines())); Hh, U, V = rr[0], rr[1], rr[2]
    out.append("")
    out.append("# %s kernel, ncu --set full (%s)" % (which, V[Hh.index("Kernel Name")] if "Kernel Name" in Hh else ""))
    for i, h in enumerate(Hh):
        if h in keep: out.append("%-86s %-16s %s" % (h, U[i], V[i]))
    src = subprocess.run([sys.executable, os.path.join(root, "profiles", "ncu_source_summary.py"), rep, "25"],
                         capture_output=True, text=True).stdout
    out.append("")
    out.append("# %s kernel, hottest source lines (warp-state samples)" % which)
    out.append(src)
open(os.path.join(root, "profiles", "r01_%s_ncu_summary.txt" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
