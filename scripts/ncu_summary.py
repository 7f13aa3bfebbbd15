"""Key metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` export (the rows quoted in profiles/*.md).
usage: python scripts/ncu_summary.py raw.csv"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum']
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d['Kernel Name'][:90], d.get('Grid Size'), d.get('Block Size'))
    for k in KEYS:
        if k in d:
            print('   %-82s %s %s' % (k, d[k], u[h.index(k)]))
    for k in h:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(d[k] or 0) >= 0.1:
            print('   stall %-76s %s' % (k.split('issue_stalled_')[1].split('_per_issue')[0], d[k]))
