"""Extract the metrics DESIGN.md / bench.py quote from an .ncu-rep into a small markdown table.
usage: python tools/ncu_summary.py <report.ncu-rep> [launch_index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
 "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
 "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "smsp__sass_average_branch_targets_threads_uniform.pct",
]
ix = {h: i for i, h in enumerate(hdr)}
for li, vals in enumerate(rows[2:]):
    if len(sys.argv) > 2 and li != int(sys.argv[2]): continue
    name = vals[ix["Kernel Name"]] if "Kernel Name" in ix else "?"
    print(f"### launch {li}: `{name}`\n\n| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in ix: print(f"| {w} | {vals[ix[w]]} | {units[ix[w]]} |")
    print()
