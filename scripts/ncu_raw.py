"""Key metrics of every kernel in an ncu report: ncu -i X.ncu-rep --page raw --csv > raw.csv; python ncu_raw.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('==', r[h.index('Kernel Name')][:60])
    for w in want:
        if w in h:
            print('  ', w, r[h.index(w)], rows[1][h.index(w)])
    for i, x in enumerate(h):
        if 'issue_stalled' in x and 'per_issue_active' in x:
            try:
                if float(r[i]) > 0.3:
                    print('   stall', x.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i])
            except ValueError:
                pass
