"""Print the handful of ncu raw-page metrics used when reading a capture: python scratch/ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_ldgsts.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    d = dict(zip(hdr, r))
    for k in want:
        if k in d: print("  %-75s %s" % (k, d[k]))
    st = [(float(v.replace(',', '')), k) for k, v in d.items() if 'issue_stalled' in k and k.endswith('per_warp_active.pct') and v not in ('', 'n/a')]
    for v, k in sorted(st, reverse=True)[:7]: print("  stall %-60s %.1f" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
