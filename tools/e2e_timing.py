#!/usr/bin/env python3
"""tools/e2e_timing.py — where the end-to-end (C ABI, host buffers) time of one bench step goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import numpy as np
import pyrt
from bench import texture_dir
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pyrt.Scene(7, 64, 64).close()  # CUDA context + module load
host = np.empty((800, 800, 3), dtype=np.float32)
for it in range(3):
    t0 = time.perf_counter(); sc = pyrt.Scene(9, 800, 800, texture_dir=texture_dir())
    t1 = time.perf_counter(); st = sc.render(spp=spp, rng_mode=0)
    t2 = time.perf_counter(); st2 = sc.render(spp=spp, rng_mode=0)
    t3 = time.perf_counter(); pyrt._check(pyrt.lib().rt_readback(sc._h, host.ctypes.data, None, None))
    t4 = time.perf_counter(); sc.close()
    t5 = time.perf_counter()
    print("build %.1f ms | render#1 %.1f ms (device %.1f) | render#2 %.1f ms (device %.1f) | readback %.1f ms | destroy %.1f ms" %
          ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st.device_ms, (t3 - t2) * 1e3, st2.device_ms, (t4 - t3) * 1e3, (t5 - t4) * 1e3))
