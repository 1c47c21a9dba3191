"""Classify FP64 instructions of the hottest loop of a kernel by number of distinct register source operands."""
import collections, re, subprocess, sys
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
    m = re.search(r'BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
best = None
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    n64 = sum(1 for t in body if re.match(r'(@!?P\d+\s+)?D(FMA|ADD|MUL|SETP)', t))
    if n64 and len(body) < 400 and (best is None or n64 > best[0]):
        best = (n64, lo, hi, body)
n64, lo, hi, body = best
print('loop 0x%x-0x%x: %d instr, %d fp64' % (lo, hi, len(body), n64))
hist = collections.Counter()
for t in body:
    m = re.match(r'(@!?P\d+\s+)?(D(FMA|ADD|MUL|SETP)\S*)\s+(.*)', t)
    if not m:
        continue
    ops = [o.strip() for o in m.group(4).split(',')]
    srcs = ops[1:] if not m.group(2).startswith('DSETP') else ops[2:]
    regs = set()
    reuse = 0
    for o in srcs:
        r = re.match(r'[-|]*R(\d+)(\.reuse)?', o)
        if r and r.group(1) != 'Z':
            regs.add(r.group(1))
            reuse += 1 if r.group(2) else 0
    hist[(m.group(2).split('.')[0], len(regs))] += 1
for k in sorted(hist):
    print('  %-6s distinct reg sources=%d : %d' % (k[0], k[1], hist[k]))
tot = sum(v * (2 if k[1] <= 2 else 3) for k, v in hist.items())
print('  RF-port model: %d cycles for %d fp64 instr (2 cyc if <=2 reg sources, 3 cyc if 3)' % (tot, n64))
