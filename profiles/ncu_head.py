"""Headline metrics of one ncu capture:  python profiles/ncu_head.py <report.ncu-rep>"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__cycles_elapsed.avg']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, v = rows[0], rows[2] if len(rows) > 2 else rows[1]
for w in WANT:
    if w in h:
        print(f'{w:75s} {v[h.index(w)]}')
for i, name in enumerate(h):
    if name.startswith('smsp__average_warps_issue_stalled') and name.endswith('per_issue_active.ratio') and float(v[i] or 0) >= 0.3:
        print(f'{name:75s} {v[i]}')
