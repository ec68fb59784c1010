#!/usr/bin/env python3
"""tools/build_timing.py — wall time of rt_build_scene (host generator + reference leaf order + H2D + device BVH build) per config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import pyrt
from bench import texture_dir
pyrt.Scene(7, 64, 64).close()  # CUDA context + module load
for name, sid, nx, ny, gh in (("C1", 1, 400, 225, 0), ("C4", 9, 800, 800, 0), ("C5-10k", 1, 3840, 2160, 50), ("C5-100k", 1, 3840, 2160, 158), ("C5-1M", 1, 3840, 2160, 500)):
    ts = []
    for i in range(4):
        t0 = time.perf_counter(); sc = pyrt.Scene(sid, nx, ny, grid_half=gh, texture_dir=texture_dir()); t1 = time.perf_counter()
        dev = sc.info.bvh_build_ms; n = sc.info.n_top; sc.close()
        ts.append((t1 - t0) * 1e3)
    print("build_timing %s: %d objects, rt_build_scene %.2f ms (best of 4; device BVH build %.2f ms)" % (name, n, min(ts), dev))
