#!/usr/bin/env python3
"""tools/all_configs.py — run ON THE GPU BOX: throughput of every BASELINE config (C1..C5) in Philox mode, one JSON line each."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import pyrt
from bench import texture_dir
CFG = [("C1", 1, 400, 225, 10, 0), ("C1 at the scene function's own size", 1, 1200, 600, 100, 0), ("C2", 7, 600, 600, 1000, 0),
       ("C3", 8, 600, 600, 1000, 0), ("C4 (one 1000-spp pass)", 9, 800, 800, 1000, 0), ("C5 10k spheres", 1, 3840, 2160, 64, 50),
       ("C5 100k spheres", 1, 3840, 2160, 64, 158), ("C5 1M spheres", 1, 3840, 2160, 64, 500)]
for name, sid, nx, ny, spp, gh in CFG:
    sc = pyrt.Scene(sid, nx, ny, grid_half=gh, texture_dir=texture_dir())
    sc.render(spp=min(spp, 16))  # warm-up
    st = sc.render(spp=spp)
    print(json.dumps({"config": name, "scene": sid, "nx": nx, "ny": ny, "spp": spp, "objects": sc.info.n_top, "bvh_nodes": sc.info.n_bvh_nodes,
                      "bvh_build_ms": round(sc.info.bvh_build_ms, 3), "device_ms": round(st.device_ms, 3), "rays": int(st.rays),
                      "rays_per_sample": round(st.rays / st.samples, 4), "mrays_per_s": round(st.rays / st.device_ms / 1e3, 1),
                      "msamples_per_s": round(st.samples / st.device_ms / 1e3, 1), "waves": st.waves, "launches": st.kernel_launches}), flush=True)
    sc.close()
