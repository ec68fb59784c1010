#!/usr/bin/env python3
"""tools/render_png.py <scene id> <spp> <out.png> [nx ny] — render a scene (Philox mode) and store it as an 8-bit PNG
(clamped; written with zlib only). Row 0 of the framebuffer is the bottom scanline, like the reference's."""
import os, struct, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import numpy as np
import pyrt
from bench import texture_dir
sid, spp, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
nx, ny = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
sc = pyrt.Scene(sid, nx, ny, texture_dir=texture_dir())
st = sc.render(spp=spp)
img = np.clip(pyrt.to_8bit(sc.framebuffer()), 0, 255).astype(np.uint8)[::-1]
h, w = img.shape[:2]
raw = b"".join(b"\x00" + img[j].tobytes() for j in range(h))
def chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
open(out, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))
print("%s: scene %d %dx%d %d spp, %.1f ms, %.0f Mrays/s" % (out, sid, w, h, spp, st.device_ms, st.rays / st.device_ms / 1e3))
