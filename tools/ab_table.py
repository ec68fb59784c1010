#!/usr/bin/env python3
"""tools/ab_table.py FILE — best Mrays/s per (config, library variant) of a tools/gpu_ab*.sh log."""
import re, sys
cfg = var = None; res = {}
for l in open(sys.argv[1]):
    m = re.match(r'== cfg (.*?) (shipped|accel\S+|env:\S+)', l)
    if m: cfg = m.group(1); var = m.group(2).split('/')[-1]; continue
    m = re.search(r'([\d.]+) Mrays/s', l)
    if m and cfg: res.setdefault(cfg, {}).setdefault(var, []).append(float(m.group(1)))
for c in res:
    print(c, ' '.join('%s=%.0f' % (v, max(x)) for v, x in res[c].items()))
