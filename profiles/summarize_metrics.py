#!/usr/bin/env python
"""Per-launch table (time, DRAM bytes, tensor-pipe activity, L2 hit rate) from an `ncu --metrics ... --csv` log:
python profiles/summarize_metrics.py gpurun_out/ll5_metrics_r2b.csv"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:64]), {})[r[mi]] = float(r[vi].replace(',', ''))
tot = 0.0
for (i, k), m in d.items():
    t = m['gpu__time_duration.sum'] / 1e3
    tot += t
    rd, wr = m.get('dram__bytes_read.sum', 0.0), m.get('dram__bytes_write.sum', 0.0)
    print('%3s %-64s %9.1f us  rd %8.1f MB  wr %8.1f MB  tensor %5.1f%%  L2 hit %5.1f%%  %5.0f GB/s' % (
        i, k, t, rd / 1e6, wr / 1e6, m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0),
        m.get('lts__t_sector_hit_rate.pct', 0.0), (rd + wr) / t / 1e3))
print('total %.1f us over %d launches (serialised, cold caches: compare shares)' % (tot, len(d)))
