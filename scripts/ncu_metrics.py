"""Key metrics of the first kernel in an .ncu-rep: python scripts/ncu_metrics.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); hdr, units, row = r[0], r[1], r[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
want = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__shared_mem_per_block_dynamic sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum
dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct l1tex__throughput.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio smsp__warps_eligible.avg.per_cycle_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sector_hit_rate.pct""".split()
for w in want:
    if w in hdr:
        i = hdr.index(w); print('%-82s %s %s' % (w, row[i], units[i]))
