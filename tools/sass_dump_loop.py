"""Print the SASS between two addresses of one kernel: python tools/sass_dump_loop.py lib.so kernel_substr 0xLO 0xHI"""
import re
import subprocess
import sys
lib, sub, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
on = False
for l in txt.splitlines():
    if 'Function :' in l:
        on = sub in l
    if on:
        m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2).strip())
