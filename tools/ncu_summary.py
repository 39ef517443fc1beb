"""Summarise an `ncu --page raw --csv` export: per kernel the headline metrics and stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'sm__cycles_elapsed.max',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d['Kernel Name'], 'grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
    for w in want:
        if w in d:
            print(f"   {w:72s} {d[w]:>18s} {units[hdr.index(w)]}")
    items = [(k, float(d[k].replace(',', ''))) for k in hdr
             if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio')]
    items.sort(key=lambda x: -x[1])
    print('   stalls per issue:', ', '.join('%s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for k, v in items[:7]))
