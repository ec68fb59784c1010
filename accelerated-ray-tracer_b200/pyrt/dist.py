"""pyrt.dist — one process per GPU: how the render path shards across the GPUs of one box.

The reference is single-GPU (no device selection code at all). Pixels and samples are
independent, so the path shards with the scene + BVH REPLICATED on every GPU and no data-path
exchange until the image is assembled:

  * tile split  (split_mode 0): rank r owns scanlines j = r (mod world). Disjoint pixels; the shares
    are gathered to rank 0. Bit-identical to a 1-GPU render in both RNG modes (every pixel's stream
    is a function of (seed, pixel, sample) only). The only split possible in reference-RNG mode.
  * spp split   (split_mode 1): every rank renders all pixels with a contiguous share of the sample
    numbers into its linear-radiance sum buffer; ONE reduce(sum) to rank 0 (NCCL over NVLink/NVSwitch
    on GPUs), then 1/ns and gamma AFTER the reduce (the reference applies gamma after averaging,
    main.cu:128-131).

torch.distributed is plumbing only (process group, the reduce/gather of library-owned buffers); the
pure partition arithmetic below has no torch dependency and is what the gloo CPU tests pin.
"""
import os

import numpy as np


# ---- partition arithmetic (mirrors rt_render in csrc/rt_host.cu) ----
def rows_of_rank(ny, rank, world):
    """Tile split: the image scanlines owned by `rank` (j = rank, rank + world, ...)."""
    return list(range(rank, ny, world))


def rows_local(ny, rank, world):
    return (ny - rank + world - 1) // world if rank < ny else 0


def sample_share(spp_total, rank, world):
    """Spp split: [base, end) sample numbers rendered by `rank`."""
    return (spp_total * rank) // world, (spp_total * (rank + 1)) // world


def assemble_rows(parts, ny):
    """Interleave tile-split shares (rank r owns scanlines r, r+world, ...) into a full (ny, ...) array."""
    world = len(parts)
    out = np.empty((ny,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


# ---- process group ----
def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Join the torchrun rendezvous if there is one. Returns (world, rank, local_rank)."""
    world, rank, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
                dist.init_process_group(backend, device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend)
    return world, rank, local


def reduce_sum_to_root(t, root=0):
    """Spp split: sum the ranks' linear-radiance buffers into rank `root`, IN PLACE on the library's accumulation
    buffer (rt_accum_device_ptr: one allocation per scene, recycled through the library's block cache, so NCCL sees
    the same few addresses step after step - no staging copies)."""
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return t
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    return t


def gather_rows_to_root(t_local, ny, root=0):
    """Tile split: gather the ranks' scanline shares to rank `root` and interleave them.
    t_local: (rows_local, nx, C) tensor. Returns the full (ny, nx, C) tensor on root, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return t_local
    world, rank = dist.get_world_size(), dist.get_rank()
    rmax = rows_local(ny, 0, world)
    pad = torch.zeros((rmax,) + tuple(t_local.shape[1:]), dtype=t_local.dtype, device=t_local.device)
    pad[: t_local.shape[0]] = t_local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == root else None
    dist.gather(pad, bufs, dst=root)
    if rank != root:
        return None
    out = torch.empty((ny,) + tuple(t_local.shape[1:]), dtype=t_local.dtype, device=t_local.device)
    for r in range(world):
        n = rows_local(ny, r, world)
        out[r::world] = bufs[r][:n]
    return out


_queue_epoch = [0]


def next_chunk(key, n_chunks):
    """Cross-process dynamic tile queue: an atomic counter in the process group's key-value store (TCPStore.add is
    atomic); returns the next unclaimed chunk id or None when all n_chunks are taken."""
    import torch.distributed as dist
    store = dist.distributed_c10d._get_default_store()
    c = int(store.add(key, 1)) - 1
    return c if c < n_chunks else None


def render_dynamic(scene_render, n_chunks, ny, nx, device=None):
    """Dynamic tile queue across the ranks of the process group. The image is cut into n_chunks tile shares
    (share c = scanlines j = c mod n_chunks); every rank pulls share ids from a shared counter and calls
    scene_render(c, n_chunks) -> (rows_c, nx, 3) tensor until none is left, writing the rows into its own zero-initialised
    full image; one reduce-sum to rank 0 then merges them (every row is written by exactly one rank and x + 0 is exact, so
    the result is bit-identical to a single-GPU render). Returns (full image on rank 0 else None, chunk ids this rank rendered)."""
    import torch
    import torch.distributed as dist
    multi = dist.is_initialized() and dist.get_world_size() > 1
    _queue_epoch[0] += 1
    key = "rt_tile_queue_%d" % _queue_epoch[0]
    full = torch.zeros((ny, nx, 3), dtype=torch.float32, device=device)
    mine = []
    c = 0
    while True:
        c = next_chunk(key, n_chunks) if multi else (len(mine) if len(mine) < n_chunks else None)
        if c is None:
            break
        rows = scene_render(c, n_chunks)
        full[c::n_chunks] = rows.to(full.device).view(-1, nx, 3)
        mine.append(c)
    if multi:
        dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
        if dist.get_rank() != 0:
            return None, mine
    return full, mine


# ---- device-side helpers (library-owned buffers as torch tensors, zero copy) ----
def accum_tensor(scene):
    """This rank's linear per-pixel radiance sums (rows_local*nx*3 floats) as a CUDA tensor view."""
    import torch
    from . import DevicePtrView
    p, n = scene.accum_ptr()
    return torch.as_tensor(DevicePtrView(p, n), device="cuda")


def fb_tensor(scene):
    import torch
    from . import DevicePtrView
    p, n = scene.fb_ptr()
    return torch.as_tensor(DevicePtrView(p, n), device="cuda")


def render_distributed(scene, spp, split_mode=1, rng_mode=0, **kw):
    """One distributed render: every rank renders its share, rank 0 ends up with the full image.
    Returns (stats, fb) where fb is the full float32 (ny, nx, 3) image on rank 0 (numpy) and None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    st = scene.render(spp=spp, rng_mode=rng_mode, rank=rank, world=world, split_mode=split_mode, **kw)
    if world == 1:
        return st, scene.framebuffer()
    if split_mode == 1:
        acc = accum_tensor(scene)
        reduce_sum_to_root(acc)
        torch.cuda.synchronize()
        if rank != 0:
            return st, None
        scene.resolve(total_spp=spp)
        return st, scene.framebuffer()
    fb = fb_tensor(scene).view(st.rows_local, st.nx, 3)
    full = gather_rows_to_root(fb, scene.ny)
    return st, (full.cpu().numpy() if full is not None else None)
