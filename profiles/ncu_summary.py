#!/usr/bin/env python
"""Print the judged counters of every kernel in an `ncu --page raw --csv` export."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
stall = [h for h in H if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-' * 100)
    for w in WANT:
        if w in H:
            i = H.index(w)
            print("%-72s %s %s" % (w, r[i][:70], U[i]))
    st = sorted(((float(r[H.index(s)] or 0), s) for s in stall), reverse=True)[:6]
    for v, s in st:
        print("   stall %-50s %.2f" % (s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
