#!/usr/bin/env python3
"""tools/sass_region.py <nvdisasm -gi output> <kernel prefix> <file> <first line> <last line> — print the SASS instructions of a
kernel whose innermost-listed source frame in <file> falls inside the line range (offline instruction counting per region)."""
import re, sys
sass, prefix, fname, a, b = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + prefix)][0]
cur, last, n = None, None, 0
for l in lines[start + 1:]:
    if l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if cur is None:
            cur = []
        cur.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if cur:
            last = cur
        cur = None
        ls = [x for x in (last or []) if x[0] == fname and a <= x[1] <= b]
        if ls:
            n += 1
            if '-q' not in sys.argv:
                print(m.group(1), ls[0][1], m.group(2))
print("instructions in region:", n)
