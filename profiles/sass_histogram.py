#!/usr/bin/env python
"""SASS evidence: per kernel of libsert_b200.so, how many instructions of the mnemonics that prove a Blackwell-native
path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, vector reductions -> REDG...).

    python profiles/sass_histogram.py > profiles/sass_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sert_b200', 'libsert_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'REDG',
         'ATOMG', 'ATOMS', 'REDUX', 'VOTE', 'SHFL', 'FMNMX', 'MUFU', 'HMMA', 'LDG', 'STG', 'LDS', 'STS']

out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, check=True).stdout.decode()
kernels = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE).stdout.decode().strip()
        name = name.replace('(anonymous namespace)::', '').replace('sert::', '')
        name = re.sub(r'\(.*', '', name)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and name:
        op = m.group(1)
        kernels[name]['_total'] += 1
        for w in WATCH:
            if op.startswith(w):
                kernels[name][w] += 1
                break
print('SASS opcode histogram of %s (cuobjdump -sass, sm_100a)' % os.path.relpath(LIB, ROOT))
print('%-58s %6s  %s' % ('kernel', 'instr', 'watched mnemonics'))
for k, c in kernels.items():
    seen = ' '.join('%s:%d' % (w, c[w]) for w in WATCH if c[w])
    print('%-58s %6d  %s' % (k[:58], c['_total'], seen))
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print('\nlibrary totals: ' + ' '.join('%s:%d' % (w, tot[w]) for w in WATCH if tot[w]))
