"""Print the metrics of an `ncu --page raw --csv` dump that the profile summaries quote."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'pipe_fp64', 'pipe_tensor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'data_bank_conflicts_pipe_lsu_mem_shared', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__grid_size', 'launch__block_size', 'issue_stalled', 'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct',
        'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'smsp__inst_executed.sum ']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for line in rows[2:]:
    for h, u, v in zip(hdr, rows[1], line):
        if any(w in h for w in WANT):
            print('%-95s %-14s %s' % (h, u, v))
    print('-' * 40)
