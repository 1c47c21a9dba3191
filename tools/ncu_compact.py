"""Compact summary of one kernel from `ncu -i X.ncu-rep --page raw --csv`: python tools/ncu_compact.py raw.csv
(the fixed metric list the summaries under profiles/ quote + the warp stall reasons per issue-active cycle)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, d = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.per_cycle_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for w in WANT:
    if w in col and d[col[w]] not in ('', 'n/a'):
        print('%-80s %s %s' % (w, d[col[w]], units[col[w]]))
print('\nwarp stall reasons (average warps stalled per issue-active cycle):')
pre, suf = 'smsp__average_warps_issue_stalled_', '_per_issue_active.ratio'
for h in sorted(hdr):
    if h.startswith(pre) and h.endswith(suf) and 'not_issued' not in h:
        print('  %-40s %s' % (h[len(pre):-len(suf)], d[col[h]]))
