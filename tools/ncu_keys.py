#!/usr/bin/env python3
"""Print the headline metrics of every kernel in an `ncu --page raw --csv` export."""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
hdr, units = r[0], r[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for row in r[2:]:
    print('=' * 100)
    for i, h in enumerate(hdr):
        if h in keys:
            print(f"{h} [{units[i]}] = {row[i]}")
    st = sorted(((float(row[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
