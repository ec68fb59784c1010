#!/usr/bin/env python3
"""tools/gpu_check.py — run ON THE GPU BOX. First-contact parity + speed report of the CUDA path
against the reference dumps in gpurun_out/ref (or tests/golden)."""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200"))
import pyrt

GOLD = os.environ.get("RT_GOLDEN", os.path.join(ROOT, "tests", "golden", "ref_gpu"))
TEX = os.path.join(ROOT, "oracle", "_ref", "textures")

def gold(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return d, (pyrt.SD(d["sd"].tobytes()) if "sd" in d else None)

def check(name, sid, nx, ny, spp):
    d, ref = gold(name)
    out = {"case": name}
    t0 = time.time()
    sc = pyrt.Scene(sid, nx, ny, texture_dir=TEX)
    out["build_s"] = round(time.time() - t0, 3)
    out["bvh_nodes"] = sc.info.n_bvh_nodes; out["bvh_ms"] = round(sc.info.bvh_build_ms, 3)
    mine, rank = sc.export()
    inv = np.argsort(rank)
    bad = sum(1 for pos in range(len(ref.top)) if ref.obj_key(int(ref.top[pos])) != mine.obj_key(int(mine.top[inv[pos]])))
    cam_bad = [n for n in pyrt.CAM_DT.names if not np.array_equal(np.atleast_1d(ref.cam[n]).view(np.uint8), np.atleast_1d(mine.cam[n]).view(np.uint8))]
    out["sd_obj_mismatch"] = bad; out["sd_cam_mismatch"] = cam_bad
    # reference-RNG render + AOV
    st = sc.render(spp=spp, rng_mode=1, aov=True)
    fb = sc.framebuffer(); obj, mat, t = sc.aov()
    out["ref_mode_ms"] = round(st.device_ms, 3); out["rays"] = int(st.rays); out["waves"] = st.waves
    g_obj = d["ids_obj"]; g_t = d["ids_t"]
    mine_pos = np.where(obj >= 0, rank[np.maximum(obj, 0)], -1)
    out["id_mismatch"] = int((mine_pos != g_obj).sum())
    out["t_mismatch"] = int((t.view(np.uint32) != g_t.view(np.uint32)).sum())
    # material classes
    mk_ref = {k: ref.mat_key(k) for k in np.unique(d["ids_mat"]) if k >= 0}
    mk_mine = {k: mine.mat_key(k) for k in np.unique(mat) if k >= 0}
    mm = 0
    for k in np.unique(mat):
        sel = mat == k
        gm = np.unique(d["ids_mat"][sel])
        for g in gm:
            if (k < 0) != (g < 0) or (k >= 0 and mk_mine[k] != mk_ref[g]):
                mm += int((d["ids_mat"][sel] == g).sum())
    out["mat_mismatch"] = mm
    gfb = d["fb"]
    out["fb_bit_mismatch_px"] = int((fb.view(np.uint32) != gfb.view(np.uint32)).any(axis=2).sum())
    a8, b8 = pyrt.to_8bit(fb), pyrt.to_8bit(gfb)
    out["fb_8bit_mismatch_px"] = int((a8 != b8).any(axis=2).sum())
    out["fb_maxabs"] = float(np.abs(fb - gfb).max())
    out["npix"] = nx * ny
    if out["fb_bit_mismatch_px"]:
        jj, ii = np.nonzero((fb.view(np.uint32) != gfb.view(np.uint32)).any(axis=2))
        out["first_bad_px"] = [[int(ii[k]), int(jj[k]), [float(x) for x in fb[jj[k], ii[k]]], [float(x) for x in gfb[jj[k], ii[k]]]] for k in range(min(3, len(ii)))]
    # philox render, same spp: statistics
    st2 = sc.render(spp=spp, rng_mode=0)
    fb2 = sc.framebuffer()
    out["philox_ms"] = round(st2.device_ms, 3); out["philox_rays_per_sample"] = round(st2.rays / max(st2.samples, 1), 4)
    out["ref_rays_per_sample"] = round(st.rays / max(st.samples, 1), 4)
    out["philox_mean_abs_diff"] = float(np.abs(fb2.mean(axis=(0, 1)) - gfb.mean(axis=(0, 1))).max())
    sc.close()
    return out

def speed(sid, nx, ny, spp, reps=2):
    sc = pyrt.Scene(sid, nx, ny, texture_dir=TEX)
    res = []
    for r in range(reps + 1):
        st = sc.render(spp=spp, rng_mode=0)
        res.append((st.device_ms, st.rays))
    ms, rays = res[-1]
    best = min(x[0] for x in res[1:])
    o = {"speed": pyrt.SCENE_NAMES[sid], "nx": nx, "ny": ny, "spp": spp, "ms_best": round(best, 3), "rays": int(rays),
         "mrays_per_s": round(rays / best / 1e3, 1), "waves": st.waves, "launches": st.kernel_launches, "slots": st.n_slots,
         "bvh_nodes": sc.info.n_bvh_nodes}
    sc.close()
    return o

if __name__ == "__main__":
    cases = [("c1_400x225_10", 1, 400, 225, 10), ("c2_300x300_16", 7, 300, 300, 16), ("c3_300x300_16", 8, 300, 300, 16),
             ("c4_400x400_16", 9, 400, 400, 16), ("s2_300x150_8", 2, 300, 150, 8), ("s3_300x150_8", 3, 300, 150, 8),
             ("s4_300x150_8", 4, 300, 150, 8), ("s5_300x150_8", 5, 300, 150, 8), ("s6_300x150_8", 6, 300, 150, 8),
             ("s10_300x150_8", 10, 300, 150, 8)]
    only = sys.argv[1:] 
    for c in cases:
        if only and c[0] not in only and "all" not in only: continue
        try:
            print(json.dumps(check(*c)), flush=True)
        except Exception as e:
            print(json.dumps({"case": c[0], "error": repr(e)}), flush=True)
    if not only or "speed" in only or "all" in only:
        for sid, nx, ny, spp in [(1, 1200, 600, 100), (7, 600, 600, 200), (8, 600, 600, 200), (9, 800, 800, 100)]:
            try:
                print(json.dumps(speed(sid, nx, ny, spp)), flush=True)
            except Exception as e:
                print(json.dumps({"speed": sid, "error": repr(e)}), flush=True)
