#!/usr/bin/env python3
"""tools/prof_cmd.py — the short command profiled under ncu: a few progressive passes of the bench workload
(Book-2 final scene, 800x800, Philox mode). Usage: prof_cmd.py [spp] [scene] [nx] [ny] [grid_half]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200"))
sys.path.insert(0, ROOT)
import pyrt
from bench import texture_dir
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sid = int(sys.argv[2]) if len(sys.argv) > 2 else 9
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 800
ny = int(sys.argv[4]) if len(sys.argv) > 4 else 800
gh = int(sys.argv[5]) if len(sys.argv) > 5 else 0
sc = pyrt.Scene(sid, nx, ny, grid_half=gh, texture_dir=texture_dir())
st = sc.render(spp=spp, rng_mode=0, profile=bool(os.environ.get("RT_WAVE_LOG")))
print("prof_cmd: scene %d %dx%d spp %d: %.2f ms, %d rays, %.1f Mrays/s, %d waves, %d launches, slots %d" %
      (sid, nx, ny, spp, st.device_ms, st.rays, st.rays / st.device_ms / 1e3, st.waves, st.kernel_launches, st.n_slots))
sc.close()
