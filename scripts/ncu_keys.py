#!/usr/bin/env python3
"""Print the handful of ncu raw-page metrics this project tracks (one block per captured launch)."""
import csv, sys
KEYS = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum',
 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct','l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
 'smsp__average_warp_latency_per_inst_issued.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---- ' + r[hdr.index('Kernel Name')][:70])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('  %-88s %s %s' % (k, r[i], units[i]))
