#!/usr/bin/env python3
"""tools/ncu_lines.py — attribute the per-SASS-instruction metrics of an ncu report to CUDA source lines.
usage: ncu_lines.py <sass.csv from `ncu -i rep --page source --csv -k regex:<kernel> --launch-count 1`> <all.sass from
`nvdisasm -gi <cubin>`> <mangled-kernel-prefix> [top N]. Frames in rt_math.h / rng.h / CUDA headers are skipped so that
an instruction is charged to the statement of the algorithm it belongs to (deepest remaining inline frame)."""
import re, csv, collections, sys
sasscsv, allsass, prefix = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(allsass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + prefix)][0]
seq, group, in_group = [], [], False
for l in lines[start + 1:]:
    if l.startswith('.text.') and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not in_group:
            group = []
            in_group = True
        group.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        in_group = False
        seq.append((list(group), m.group(2)))
rows = list(csv.reader(open(sasscsv)))
hdr, data = rows[1], rows[2:]
ie, it, iss = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
SKIP = ('rt_math.h', 'rng.h')
agg = collections.defaultdict(lambda: [0, 0, 0])
data = data[:len(seq)]  # a report with several launches of the kernel lists them one after the other
for k in range(min(len(seq), len(data))):
    g, ins = seq[k]
    r = data[k]
    key = None
    for f in g:
        if f[0].endswith(('.cuh', '.cu', '.h')) and f[0] not in SKIP and not f[0].startswith(('sm_', 'device_', 'cuda', 'math_', 'crt')):
            key = f
            break
    if key is None and g:
        key = g[-1]
    a = agg[key]
    a[0] += int(r[ie]); a[1] += int(r[it]); a[2] += int(r[iss])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values()); tott = sum(a[1] for a in agg.values())
print("instructions %d, thread-instructions %d (%.2f threads/inst), stall samples %d" % (tot, tott, tott / tot, tots))
print("%-28s %8s %9s %9s" % ("file:line", "inst %", "thr/inst", "samples %"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-28s %7.1f%% %9.1f %8.1f%%" % ("%s:%d" % key if key else "?", 100 * a[0] / tot, a[1] / max(a[0], 1), 100 * a[2] / tots))
