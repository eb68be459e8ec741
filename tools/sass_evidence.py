#!/usr/bin/env python
"""Regenerable SASS evidence for profiles/: per kernel of libse3b200.so, the count of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2/HFMA2 =
packed math).  usage: python tools/sass_evidence.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'se3_transformer_pytorch_b200', 'libse3b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
names = ['UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'SYNCS', 'FFMA2', 'HFMA2', 'HMUL2', 'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR']
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        counts[cur]['total'] += 1
        for nme in names:
            if op == nme or (nme == 'UTCHMMA' and op.startswith('UTC') and op.endswith('MMA')):
                counts[cur][nme] += 1
print('# cuobjdump -sass se3_transformer_pytorch_b200/libse3b200.so, instruction counts per kernel (tools/sass_evidence.py)')
print('kernel,' + ','.join(['total'] + names))
for k, c in counts.items():
    print(k.replace(',', ';') + ',' + ','.join(str(c[n]) for n in ['total'] + names))
