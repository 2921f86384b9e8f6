"""Summarise an ncu --set full report (raw page) into the handful of figures DESIGN.md quotes.
usage: python scripts/ncu_summary.py report.ncu-rep [evals_per_launch ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
keys = [
 ("gpu__time_duration.sum", "duration"),
 ("launch__grid_size", "grid"),
 ("launch__registers_per_thread", "regs"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64_pipe_active_pct"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
 ("smsp__inst_executed.sum", "warp_instr"),
 ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "dfma_per_cyc"),
 ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "dmul_per_cyc"),
 ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "dadd_per_cyc"),
 ("sm__cycles_elapsed.avg", "cycles"),
 ("dram__bytes_write.sum", "dram_write"),
 ("dram__bytes_read.sum", "dram_read"),
 ("sass__inst_executed_local_loads", "local_ld"),
 ("sass__inst_executed_local_stores", "local_st"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
 ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instr"),
 ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
 ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
 ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_instr"),
]
evals = [float(x) for x in sys.argv[2:]]
for n, r in enumerate(rows[2:]):
    print("== %s" % r[ix["Kernel Name"]])
    d = {}
    for k, nm in keys:
        if k in ix:
            d[nm] = r[ix[k]]
            print("  %-22s %s %s" % (nm, r[ix[k]], units[ix[k]]))
    if n < len(evals):
        ev = evals[n]
        cyc = float(d["cycles"])
        f64 = (float(d["dfma_per_cyc"]) + float(d["dmul_per_cyc"]) + float(d["dadd_per_cyc"])) * cyc
        print("  -> per evaluation: %.0f FP64 arithmetic thread-instr, %.0f warp-instr per warp-evaluation"
              % (f64 / ev, float(d["warp_instr"]) / (ev / 32)))
