#!/usr/bin/env python3
"""tools/dist_check.py — run under torchrun on >= 2 GPUs: the one-process-per-GPU paths of pyrt.dist against a 1-GPU
render of the same image. Tile split must be bit-identical (both RNG modes); spp split equal to float rounding."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-ray-tracer_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import pyrt
from pyrt import dist as rdist
from bench import texture_dir

world, rank, local = rdist.init_process_group("nccl")
torch.cuda.set_device(local)
ok = True
for sid, nx, ny, spp in [(1, 400, 225, 10), (9, 200, 200, 32)]:
    sc = pyrt.Scene(sid, nx, ny, texture_dir=texture_dir(), device=local)
    for rng_mode in (1, 0):
        st, fb_tile = rdist.render_distributed(sc, spp, split_mode=0, rng_mode=rng_mode)
        fb_spp = None
        if rng_mode == 0:
            st2, fb_spp = rdist.render_distributed(sc, spp, split_mode=1, rng_mode=0)
        if rank == 0:
            sc.render(spp=spp, rng_mode=rng_mode)
            whole = sc.framebuffer()
            same = np.array_equal(whole.view(np.uint32), fb_tile.view(np.uint32))
            msg = "scene %d rng %d world %d: tile split bit-identical to 1 GPU: %s" % (sid, rng_mode, world, same)
            ok &= same
            if fb_spp is not None:
                err = float(np.abs(fb_spp - whole).max())
                msg += "; spp split max abs diff %.2e" % err
                ok &= err < 1e-4
            if sid == 1 and rng_mode == 1:
                g = np.load(os.path.join(ROOT, "tests", "golden", "ref_gpu", "c1_400x225_10.npz"))
                same8 = np.array_equal(pyrt.to_8bit(fb_tile), pyrt.to_8bit(g["fb"]))
                msg += "; 8-bit identical to the reference CUDA build: %s" % same8
                ok &= same8
            print(msg, flush=True)
    # dynamic tile queue across the ranks (shared counter in the process group's store): bit-identical whoever renders what
    def share(c, n, sc=sc):
        st = sc.render(spp=spp, rng_mode=0, rank=c, world=n, split_mode=0)
        return rdist.fb_tensor(sc).view(st.rows_local, st.nx, 3).clone()
    full, mine = rdist.render_dynamic(share, 4 * world + 1, ny, nx, device=torch.device("cuda", local))
    counts = [None] * world
    torch.distributed.all_gather_object(counts, len(mine))
    if rank == 0:
        sc.render(spp=spp, rng_mode=0)
        same = np.array_equal(sc.framebuffer().view(np.uint32), full.cpu().numpy().view(np.uint32))
        print("scene %d world %d: dynamic tile queue (%d chunks, per rank %s) bit-identical to 1 GPU: %s" % (sid, world, 4 * world + 1, counts, same), flush=True)
        ok &= same
    sc.close()
if rank == 0:
    print("dist_check:", "OK" if ok else "FAILED", flush=True)
torch.distributed.barrier(device_ids=[local])
torch.distributed.destroy_process_group()
sys.exit(0 if ok else 1)
