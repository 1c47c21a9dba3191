"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [substr ...]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = sys.argv[2:] or ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
                        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct',
                        'pipe_fp64', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'inst_executed_pipe_',
                        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'bank_conflicts', 'warp_issue_stalled',
                        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
for d in data:
    print('=' * 100)
    for i, h in enumerate(hdr):
        if any(w in h for w in want) and not h.endswith('_peak_sustained') and d[i] not in ('', 'n/a'):
            if 'warp_issue_stalled' in h and not h.endswith('per_warp_active.pct'):
                continue
            print('%-95s %s %s' % (h, d[i], units[i]))
