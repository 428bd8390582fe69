"""Key metrics of every kernel in an .ncu-rep (ncu --set full) as a small text table for profiles/."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('#', d['Kernel Name'])
    for k in KEYS:
        if k in d:
            print(f'{k},{units[hdr.index(k)]},{d[k]}')
