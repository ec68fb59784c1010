#!/usr/bin/env python3
"""tools/pack_goldens.py — pack the raw dumps of the reference harnesses (ref_gpu / ref_cpu:
<prefix>.sd / .ids / .fb) into one compressed .npz per run and delete the raw files.

usage: pack_goldens.py <dir> [--keep]
The .npz keys: sd (raw bytes of the SD file, uint8), ids_obj/ids_t/ids_mat/ids_bvh_t, fb (float32,
ny x nx x 3, row j=0 is the BOTTOM scanline like the reference's fb), nx, ny, ns.
"""
import sys, os, glob
import numpy as np

def pack(prefix, keep=False):
    out = {}
    raw = []
    p = prefix + ".sd"
    if os.path.exists(p):
        out["sd"] = np.fromfile(p, dtype=np.uint8); raw.append(p)
    p = prefix + ".ids"
    if os.path.exists(p):
        a = np.fromfile(p, dtype=np.int32)
        nx, ny = int(a[0]), int(a[1]); n = nx * ny
        body = a[2:]
        out["ids_obj"] = body[0:n].reshape(ny, nx).astype(np.int32)
        out["ids_t"] = body[n:2 * n].view(np.float32).reshape(ny, nx)
        out["ids_mat"] = body[2 * n:3 * n].reshape(ny, nx).astype(np.int32)
        out["ids_bvh_t"] = body[3 * n:4 * n].view(np.float32).reshape(ny, nx)
        out["nx"], out["ny"] = nx, ny
        raw.append(p)
    p = prefix + ".fb"
    if os.path.exists(p):
        a = np.fromfile(p, dtype=np.int32, count=3)
        nx, ny, ns = int(a[0]), int(a[1]), int(a[2])
        fb = np.fromfile(p, dtype=np.float32, offset=12).reshape(ny, nx, 3)
        out["fb"] = fb; out["nx"], out["ny"], out["ns"] = nx, ny, ns
        raw.append(p)
        # <name>_dsK: keep only the KxK box-downsampled LINEAR radiance (gamma 2.2 undone, averaged): the reference at
        # nx x ny x ns is, per downsampled pixel, a K*K*ns-sample estimate of the same pixel footprint integral
        m = __import__("re").search(r"_ds(\d+)$", prefix)
        if m:
            K = int(m.group(1))
            lin = np.maximum(fb, 0).astype(np.float64) ** 2.2
            out["lin_ds"] = lin.reshape(ny // K, K, nx // K, K, 3).mean(axis=(1, 3)).astype(np.float32)
            out["ds"] = K
            del out["fb"]
    if out:
        np.savez_compressed(prefix + ".npz", **out)
        if not keep:
            for r in raw: os.remove(r)
    return bool(out)

if __name__ == "__main__":
    d = sys.argv[1]; keep = "--keep" in sys.argv
    prefixes = sorted({os.path.splitext(f)[0] for f in glob.glob(os.path.join(d, "*")) if f.endswith((".sd", ".ids", ".fb"))})
    for p in prefixes:
        pack(p, keep)
    tot = sum(os.path.getsize(f) for f in glob.glob(os.path.join(d, "*")))
    print("packed %d runs, dir size %.1f MiB" % (len(prefixes), tot / 2**20))
