"""Blackwell-native evidence in-tree: per object file, how many tcgen05 / TMEM / TMA instructions the SASS holds.
usage: python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt   (after progen_b200/csrc/build.sh)"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'SYNCS', 'USETMAXREG', 'FFMA2', 'FADD2', 'FMNMX3',
       'MUFU.EX2', 'MUFU.TANH', 'HMMA', 'LDGSTS', 'REDG', 'ATOMG']
print('SASS opcode counts per object (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / st,')
print('UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA = legacy mma.sync.\n')
print(f'{"object":22s}' + ''.join(f'{o:>11s}' for o in OPS) + f'{"2CTA mma":>11s}')
for obj in sorted(glob.glob(os.path.join(ROOT, 'progen_b200', 'csrc', 'build', '*.o'))):
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    row = [len(re.findall(r'\b' + re.escape(o) + r'\b', sass)) for o in OPS]
    two = len(re.findall(r'UTCHMMA\.2CTA', sass))
    print(f'{os.path.basename(obj):22s}' + ''.join(f'{c:11d}' for c in row) + f'{two:11d}')
