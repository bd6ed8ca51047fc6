#!/usr/bin/env python3
"""Key raw metrics of an ncu report.  usage: python tools/ncu_raw.py X.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
for vals in rows[2:]:
    print("##", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print("%-90s %s %s" % (h, v, u))
