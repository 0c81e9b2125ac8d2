#!/usr/bin/env python
"""Extracts the headline metrics of an ncu report (ncu -i X.ncu-rep --page raw --csv piped on stdin)."""
import csv
import sys

KEEP = [
    'Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
    'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
    'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
    'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
    'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
    'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
    'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_selected',
]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEEP:
        if k in d and d[k] not in ('', 'n/a'):
            print('%-100s %s %s' % (k, d[k], units[hdr.index(k)]))
    print()
