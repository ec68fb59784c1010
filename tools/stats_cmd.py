#!/usr/bin/env python3
"""tools/stats_cmd.py — per-ray work counters of the traversal (diagnostics library built by tools/build_variant.sh stats -DRT_STATS)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RT_LIB"] = os.environ.get("STATS_LIB") or os.path.join(ROOT, "accelerated-ray-tracer_b200", "lib", "variants", "stats.so")
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import pyrt
from bench import texture_dir
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sid = int(sys.argv[2]) if len(sys.argv) > 2 else 9
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 800
ny = int(sys.argv[4]) if len(sys.argv) > 4 else 800
sc = pyrt.Scene(sid, nx, ny, texture_dir=texture_dir())
out = (C.c_ulonglong * 12)()
pyrt.lib().rt_debug_stats(out, 1)
st = sc.render(spp=spp, rng_mode=0)
pyrt.lib().rt_debug_stats(out, 1)
names = ["node expansions", "sphere tests", "geom tests", "medium tests", "node phases (warp)", "leaf phases (warp)", "leaves queued", "-",
         "node phase: lanes finished", "node phase: lanes blocked", "node phase: lanes expanding", "-"]
print("scene %d %dx%d spp %d: %d rays, %.1f Mrays/s (stats build), bvh nodes %d" % (sid, nx, ny, spp, st.rays, st.rays / st.device_ms / 1e3, sc.info.n_bvh_nodes))
for n, v in zip(names, out):
    print("  %-22s %14d  %8.3f per ray" % (n, v, v / st.rays))
