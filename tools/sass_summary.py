"""Opcode summary of every kernel in the built library: cuobjdump -sass, counted per kernel for the mnemonics that
prove the hardware paths in use (DMMA = FP64 tensor MMA, UBLKCP / SYNCS = bulk-copy engine and mbarriers,
STTM / LDTM / UTCATOMSWS = tensor-memory stores, loads and allocation,
LDL / STL = local-memory spills).   python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'conmech_b200', 'lib', 'libconmech_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
WATCH = ['DMMA', 'UBLKCP', 'SYNCS', 'STTM', 'LDTM', 'UTCATOMSWS', 'LDL', 'STL', 'LDG', 'STG', 'LDS', 'STS', 'MEMBAR', 'BAR', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'ATOMS', 'NANOSLEEP']
cur, counts, total = None, collections.OrderedDict(), {}
for ln in out.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        if op in WATCH:
            counts[cur][op] += 1
print('# {} ({} bytes)'.format(os.path.relpath(lib, REPO), os.path.getsize(lib)))
print('# SASS instruction counts per kernel (static); sm_100a')
for k, c in counts.items():
    name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'\(cm_model_dev.*$', '', name)
    print('{:70s} total {:6d}  '.format(name[:70], total[k]) + '  '.join('{} {}'.format(op, c[op]) for op in WATCH if c[op]))
