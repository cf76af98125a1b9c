"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump: one block per kernel launch with the metrics that matter here.
Usage: python tools/ncu_raw_summary.py raw.csv > profiles/summary.txt"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(f"== {d['Kernel Name'][:110]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for k in KEYS:
        if k in d:
            print(f'   {k:90s} {d[k]:>16s} {units[h.index(k)]}')
