#!/usr/bin/env python3
"""tools/c1_timing.py — BASELINE config C1 exactly (Book-1 scene, 400x225, 10 spp, Philox mode): device time of repeated renders of one
resident scene (min / median of 30) and the end-to-end time of build + render + readback + destroy (median of 10)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import pyrt
sc = pyrt.Scene(1, 400, 225)
ms = []
for i in range(35):
    st = sc.render(spp=10, rng_mode=0)
    if i >= 5: ms.append(st.device_ms)
rays = st.rays
build_ms = sc.info.bvh_build_ms
sc.close()
e2e = []
for i in range(12):
    t0 = time.perf_counter()
    s2 = pyrt.Scene(1, 400, 225)
    s2.render(spp=10, rng_mode=0)
    fb = s2.framebuffer()
    s2.close()
    if i >= 2: e2e.append((time.perf_counter() - t0) * 1e3)
ms, e2e = np.array(ms), np.array(e2e)
print("c1_timing: %d rays; device ms min %.3f median %.3f -> %.0f / %.0f Mrays/s; bvh build %.3f ms; e2e (build + render + readback + destroy) median %.3f ms -> %.0f Mrays/s"
      % (rays, ms.min(), np.median(ms), rays / ms.min() / 1e3, rays / np.median(ms) / 1e3, build_ms, np.median(e2e), rays / np.median(e2e) / 1e3))
