"""Print the handful of ncu raw-page metrics that decide what bounds a kernel.  usage: ncu_summary.py raw.csv [name filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_ldgsts.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if flt and flt not in name:
        continue
    print('=====', name[:100])
    for w in want:
        if w in hdr:
            print(f'  {w:75s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}')
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:40s} {v:.2f}')
