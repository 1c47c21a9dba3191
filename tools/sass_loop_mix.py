"""Instruction mix of the innermost loops of a kernel: python tools/sass_loop_mix.py lib.so kernel_substr"""
import collections
import re
import subprocess
import sys
lib, sub = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ins, on = [], False
for l in txt.splitlines():
    if 'Function :' in l:
        on = sub in l
    if on:
        m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r'BRA(?:\.\w+)*\s+(?:!?\w+,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    c = collections.Counter()
    for t in body:
        p = t.split()
        op = p[1] if p[0].startswith('@') else p[0]
        c[op.split('.')[0]] += 1
    fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DADD', 'DMUL', 'DSETP'))
    if fp64 >= 8:
        print('loop 0x%x-0x%x: %d instr, fp64 %d, other %d' % (lo, hi, len(body), fp64, len(body) - fp64))
        print('   ', c.most_common())
