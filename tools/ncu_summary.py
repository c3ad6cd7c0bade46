"""Summarise an .ncu-rep (read on the CPU box): key metrics per profiled launch."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__grid_size','launch__block_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum',
 'lts__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','sm__cycles_elapsed.avg',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
name_i = h.index('Kernel Name')
for r in rows[2:]:
    print('==', r[name_i][:90])
    for k in KEYS:
        if k in h:
            i = h.index(k); print(f'   {k} = {r[i]} {units[i]}')
    if len(sys.argv) > 2:
        for i,k in enumerate(h):
            if sys.argv[2] in k: print(f'   {k} = {r[i]} {units[i]}')
