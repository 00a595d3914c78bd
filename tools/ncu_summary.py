#!/usr/bin/env python3
"""Key metrics of an .ncu-rep as a small CSV (one row per metric), for profiles/.
usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_static launch__shared_mem_per_block_dynamic
launch__occupancy_limit_shared_mem launch__occupancy_limit_registers launch__occupancy_limit_warps
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
smsp__sass_average_branch_targets_threads_uniform.pct
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum""".split()

rep, outp = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
with open(outp, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "value", "unit"])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for m in WANT:
            if m in hdr:
                w.writerow([name, m, r[hdr.index(m)], units[hdr.index(m)]])
print("wrote", outp)
