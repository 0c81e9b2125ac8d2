#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of time)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get('Metric Name') == 'gpu__time_duration.sum':
                v = float(d['Metric Value'].replace(',', ''))
                unit = d['Metric Unit']
                us = v / 1000 if unit in ('ns', 'nsecond') else v * 1000 if unit in ('ms', 'msecond') else v
                agg[d['Kernel Name'][:70]].append(us)
    tot = sum(sum(v) for v in agg.values())
    print('%-72s %5s %10s %7s' % ('kernel', 'n', 'avg_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print('%-72s %5d %10.2f %6.1f%%' % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print('total_us %.1f over %d launches' % (tot, sum(len(v) for v in agg.values())))


if __name__ == '__main__':
    main(sys.argv[1])
