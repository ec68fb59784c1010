#!/usr/bin/env python3
"""tools/wl_sum.py <wave log> — per-kernel throughput from an RT_WAVE_LOG file (one line per wave: n, rays, trace ms, shade ms)."""
import sys
rows = [l.split() for l in open(sys.argv[1])]
tr = sum(float(r[2]) for r in rows); sh = sum(float(r[3]) for r in rows); rays = sum(int(r[1]) for r in rows)
print("wavelog: waves %d rays %d trace %.2f ms shade %.2f ms -> trace %.1f Mrays/s shade %.1f Mrays/s" %
      (len(rows), rays, tr, sh, rays / tr / 1e3, rays / sh / 1e3))
