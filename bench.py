#!/usr/bin/env python3
"""bench.py — the headline benchmark of the render path: Mrays/s (all bounces, device-timed) on the
Book-2 final scene (create_world_final, main.cu:498-562) at 800x800, on N B200s of one box.

  python bench.py [--gpus N] [--steps K] [--warmup W]                  (N > 1: launched by torchrun)
  python bench.py --impl reference ...      the reference's own render code on the box's host cores

A *step* is one progressive pass of --spp-per-step samples per pixel (default 1000) over the whole
800x800 image; steps use disjoint sample numbers, so the default K = 10 steps ARE the reference's
10000-spp final_scene() (main.cu:1179). At N > 1 the SAME job is split N ways (strong scaling, the
shape BASELINE.json configs[3] names: a fixed 800x800 x 10000-spp render tile/spp-split across 1/2/4/8
GPUs): --split spp (default) gives every GPU 1/N of the pass's sample numbers and sums the
linear-radiance buffers on rank 0 with one NCCL reduce per step; --split tile gives every GPU its
interleaved scanlines and gathers them. Both exchanges are inside the timed region. --scaling weak
keeps the per-GPU work fixed instead (every GPU renders --spp-per-step samples per step). After the
timed region rank 0 renders the last step's sample set alone and compares it with the image the N
ranks produced ("parity_check": tile split bit-identical, spp split to float-sum rounding).

value    whole-job Mrays/s with the scene resident in HBM: rays of all ranks / (max-over-ranks CUDA-event
         span of the K steps, barrier + synchronize on both sides).
e2e      the same metric through the C ABI with host buffers: every step is rt_build_scene (host generator,
         H2D of scene + texture, device BVH build) + rt_render + rt_readback of the float framebuffer to host
         + rt_destroy — what the reference's final_scene() does between process start and the PPM loop.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "accelerated-ray-tracer_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

METRIC = "Mrays/s (all bounces, device-timed) on Book-2 final scene"
UNIT = "Mrays/s"
SCENE_ID, NX, NY = 9, 800, 800
GRID_HALF = 0
# --config: the other BASELINE configs as bench lines for profiles/ (the driver's line is always c4)
CONFIGS = {"c4": (9, 800, 800, 0, 1000, "Book-2 final scene"),
           "c5-10k": (1, 3840, 2160, 50, 64, "C5 scale-up, 10 004 spheres 3840x2160"),
           "c5-100k": (1, 3840, 2160, 158, 64, "C5 scale-up, 99 860 spheres 3840x2160"),
           "c5-1m": (1, 3840, 2160, 500, 64, "C5 scale-up, 1 000 004 spheres 3840x2160")}
# SURVEY.md §8(d) / BASELINE.md §3: work per ray of the reference algorithm on C4 (fixed normaliser)
C4_FLOP_PER_RAY = 1685.0
C4_BYTES_PER_RAY = 2535.0
# k_trace's share of it: BVH nodes + primitives + media (1715 + 245 + 368 + 47) + ray read 32 + hit write 16
C4_TRACE_BYTES_PER_RAY = 2423.0
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "ref_cpu")
REF_GPU = os.path.join(ROOT, "baseline", "_ref", "ref_gpu")
TEXDIRS = [os.path.join(ROOT, "oracle", "_ref", "textures"), os.path.join(ROOT, "tests", "golden", "textures")]


def texture_dir():
    for d in TEXDIRS:
        if os.path.exists(os.path.join(d, "earthmap.ppm")):
            return d
    raise SystemExit("bench: earthmap.ppm not found (run __graft_entry__.build() where /root/reference exists)")


MICROBENCH = os.path.join(PKG, "lib", "rt_microbench")


def ncu_capture():
    """Per-ray figures of k_trace / k_shade from the latest committed `ncu --set full` capture (profiles/*_traffic.json,
    written by tools/summarize_profile.py): DRAM / L2 / L1 bytes, FP32 flop and warp instructions of one launch divided by
    the rays of that launch, plus issue utilisation and lanes per instruction."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    for f in files[::-1]:
        d = json.load(open(f))
        kt = d.get("k_trace") or d.get("k_trace<0>")
        ks = d.get("k_shade<0>")
        if kt and kt.get("rays") and kt.get("warp_inst"):
            return kt, ks, os.path.basename(f)
    return None, None, None


def microbench():
    """Measured denominators (the FP32 and L2 peaks 'must be microbenchmarked'): tools/microbench.cu, run live
    on this GPU when the binary is there (about 3 s), else the committed result of the last run on the same GPU model."""
    if os.path.exists(MICROBENCH):
        try:
            r = subprocess.run([MICROBENCH], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
            for line in r.stdout.splitlines():
                if line.startswith("{"):
                    d = json.loads(line)
                    d["source"] = "tools/microbench.cu run live on this GPU"
                    return d
        except Exception:  # noqa: BLE001
            pass
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_microbench.json")))
    if files:
        d = json.load(open(files[-1]))
        d["source"] = "profiles/" + os.path.basename(files[-1])
        return d
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi-equivalent clock/throttle sampling (NVML) during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def run_json(cmd, timeout):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    for line in r.stdout.splitlines()[::-1]:
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("no JSON from %s: rc=%d %s" % (cmd[0], r.returncode, r.stderr[-500:]))


def oracle_port_run(nx, ny, ns, threads):
    """Fallback when oracle/_ref/ref_cpu is not there: the CPU restatement (oracle/rt_oracle.cpp, kind "port") renders the
    same scene description; this is the one other place bench.py may execute oracle/."""
    import numpy as np
    os.environ["OMP_NUM_THREADS"] = str(threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import pyrt
    sd, _ = pyrt.export_host(SCENE_ID, nx, ny, texture_dir=texture_dir())
    o = oracle_py.Oracle(sd.raw.tobytes(), [oracle_py.load_ppm(os.path.join(texture_dir(), "earthmap.ppm"))])
    t0 = time.perf_counter()
    fb, rays = o.render(nx, ny, ns)
    ms = (time.perf_counter() - t0) * 1e3
    return {"render_ms_mean": ms, "rays": rays, "rays_per_sample": rays / float(nx * ny * ns), "mrays_per_s": rays / ms / 1e3,
            "kind": "port"}


def cpu_reference_run(nx, ny, ns, threads, count=1):
    """The reference's own render()/color() (compiled for the host behind oracle/shim) on this box's cores."""
    if not os.path.exists(REF_CPU):
        return oracle_port_run(nx, ny, ns, threads)
    cmd = [REF_CPU, "--scene", str(SCENE_ID), "--nx", str(nx), "--ny", str(ny), "--ns", str(ns), "--reps", "1",
           "--count", str(count), "--threads", str(threads), "--textures", texture_dir()]
    return run_json(cmd, 900)


def cpu_sample_plan(threads, target_s):
    """Bounded sample of the workload: the full 800x800 image at the spp that takes ~target_s on this box."""
    probe = cpu_reference_run(200, 200, 1, threads)
    per_sample_ms = probe["render_ms_mean"] / (200 * 200)
    ns = int(max(1, min(64, round(target_s * 1e3 / (per_sample_ms * NX * NY)))))
    return ns, probe["rays_per_sample"]


def reference_arm(args):
    """--impl reference: rank 0 only. kind 'reference' = /root/reference's render path compiled unmodified for
    the host (oracle/build_ref.sh); the reference ships no CPU path of its own."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind = "reference" if os.path.exists(REF_CPU) else "port"
    threads = os.cpu_count() or 1
    ns, _ = cpu_sample_plan(threads, 2.0)
    res = []
    for i in range(args.warmup + args.steps):
        o = cpu_reference_run(NX, NY, ns, threads)
        if i >= args.warmup:
            res.append(o)
    rays = sum(o["rays"] for o in res)
    ms = sum(o["render_ms_mean"] for o in res)
    v = rays / ms / 1e3
    sample = "%dx%d x %d spp per step (of the 10000-spp config; throughput is linear in spp)" % (NX, NY, ns)
    line = {"metric": METRIC, "value": round(v, 4), "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / max(len(res), 1), 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Book-2 final scene (create_world_final) 800x800, depth 50; " + sample,
                       "scene_id": SCENE_ID, "nx": NX, "ny": NY, "spp_per_step": ns},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference render()/color() compiled unmodified for the host behind oracle/shim (g++ -O2 -fopenmp, "
                    "glibc libm); the reference's CUDA build on this GPU is reported by the default arm as reference_cuda_sm100"}
    print(json.dumps(line), flush=True)
    return 0


def reference_cuda_run(ns=8):
    """The reference's own CUDA build recompiled for sm_100 (baseline/_ref/ref_gpu), same scene, on this GPU."""
    if not os.path.exists(REF_GPU):
        return None
    try:
        o = run_json([REF_GPU, "--scene", str(SCENE_ID), "--nx", str(NX), "--ny", str(NY), "--ns", str(ns), "--reps", "1",
                      "--count", "1", "--textures", texture_dir()], 600)
        return {"value": o["mrays_per_s"], "unit": UNIT, "msamples_per_s": o["msamples_per_s"],
                "sample": "%dx%d x %d spp, render_init+render device-timed" % (NX, NY, ns),
                "rays_per_sample": o["rays_per_sample"], "scene_build_ms": o["build_ms"],
                "what": "reference render kernel (main.cu:107-133) compiled unmodified, nvcc -arch=sm_100 -O3"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=1000)
    ap.add_argument("--split", default="spp", choices=["spp", "tile"],
                    help="N > 1: spp split (default: every GPU renders all pixels for its own --spp-per-step samples, one NCCL "
                         "reduce per step) or tile split (every GPU renders its interleaved scanlines for N x --spp-per-step "
                         "samples, one NCCL gather per step); per-GPU work is the same in both")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong (default) = the fixed 800x800 x spp-per-step pass split N ways; weak = every GPU "
                         "renders spp-per-step samples per step")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS),
                    help="workload: c4 (default, BASELINE's headline config) or a C5 scale-up scene (BASELINE configs[4]; "
                         "sets --spp-per-step 64 unless given; no reference arms: the reference cannot build these scenes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-microbench", action="store_true", help="do not run tools/microbench.cu (measured FP32 / issue / L2 peaks)")
    ap.add_argument("--profile-in-timed", action="store_true",
                    help="record CUDA events around every launch INSIDE the timed region (costs ~7%%); default: the "
                         "per-kernel durations come from one extra, untimed, profiled step of the same workload")
    args = ap.parse_args()
    global SCENE_ID, NX, NY, GRID_HALF, METRIC
    if args.config != "c4":
        SCENE_ID, NX, NY, GRID_HALF, spp_default, what = CONFIGS[args.config]
        METRIC = "Mrays/s (all bounces, device-timed) on " + what
        if "--spp-per-step" not in sys.argv:
            args.spp_per_step = spp_default
        args.no_cpu_baseline = args.no_reference_cuda = True
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import pyrt
    from pyrt import dist as rdist

    if not torch.cuda.is_available():
        raise SystemExit("bench: no CUDA device (the render path has no CPU fallback)")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on STDOUT, next to the one JSON line
    world, rank, local = rdist.init_process_group("nccl")
    if world != args.gpus:
        raise SystemExit("bench: --gpus %d but WORLD_SIZE=%d (launch N>1 with torchrun)" % (args.gpus, world))
    torch.cuda.set_device(local)
    import torch.distributed as tdist
    dev = torch.device("cuda", local)
    S, K, W = args.spp_per_step, args.steps, args.warmup
    tex = texture_dir()

    def barrier():
        if world > 1:
            tdist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    sc = pyrt.Scene(SCENE_ID, NX, NY, grid_half=GRID_HALF, texture_dir=tex, device=local)
    sc_info_n_top, sc_info_nodes, sc_info_bvh_ms = sc.info.n_top, sc.info.n_bvh_nodes, sc.info.bvh_build_ms
    weak = args.scaling == "weak" and world > 1
    S_step = S * world if weak else S          # samples per pixel of one step, all ranks together
    n_steps_all = W + K + 1                     # warm-up + timed + the profiled / parity step
    spp_all = S_step * n_steps_all
    tile = args.split == "tile" and world > 1

    def step(i, profile=False, alone=False):
        """Step i renders sample numbers [i * S_step, (i + 1) * S_step) of every pixel, split over the ranks
        (alone: this rank renders the whole step by itself - the parity check)."""
        w_, r_ = (1, 0) if alone else (world, rank)
        if tile and not alone:
            # tile split: rank r owns scanlines j = r (mod N) and renders all S_step samples of step i for them
            st = sc.render(spp=S_step, rng_mode=0, split_mode=0, rank=r_, world=w_, seed=1984 + i, profile=profile)
            full = rdist.gather_rows_to_root(rdist.fb_tensor(sc).view(st.rows_local, st.nx, 3), NY)  # NCCL gather: the image on rank 0
            torch.cuda.current_stream().synchronize()
            return st, full
        if tile:
            st = sc.render(spp=S_step, rng_mode=0, split_mode=0, rank=0, world=1, seed=1984 + i, profile=profile)
            return st, rdist.fb_tensor(sc).view(st.rows_local, st.nx, 3)
        # spp split: the pass's sample numbers are cut into world contiguous shares
        st = sc.render(spp=spp_all, rng_mode=0, split_mode=1, rank=i * w_ + r_, world=n_steps_all * w_, profile=profile)
        acc = rdist.accum_tensor(sc)
        if w_ > 1:
            rdist.reduce_sum_to_root(acc)  # NCCL reduce over NVLink, in place on the library's accumulation buffer
            torch.cuda.current_stream().synchronize()  # the next pass overwrites the accumulation buffer
        return st, acc

    for i in range(W):
        step(i)
    result = None
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rays = 0
    launches = 0
    kernel_ms = 0.0
    trace_ms = shade_ms = 0.0
    waves = 0
    for i in range(W, W + K):
        st, result = step(i, profile=args.profile_in_timed)
        rays += st.rays
        launches += st.kernel_launches
        kernel_ms += st.device_ms
        trace_ms += st.trace_ms
        shade_ms += st.shade_ms
        waves += st.profiled_waves
    e1.record()
    barrier()
    ck = clocks.stop()
    span_ms = e0.elapsed_time(e1)
    prof_rays = rays
    # ---- parity of the multi-GPU result: rank 0 renders the last timed step's samples alone ----
    parity = None
    if world > 1 and not args.no_parity_check:
        got = result.clone() if rank == 0 else None
        barrier()
        if rank == 0:
            _, want = step(W + K - 1, alone=True)
            torch.cuda.synchronize()
            a, b = got.float().flatten(), want.float().flatten()
            diff = (a - b).abs()
            if tile:
                ok = bool(torch.equal(a, b))
                parity = {"ok": ok, "split": "tile", "rule": "bit-identical to the 1-rank render of the same samples",
                          "pixels_differing": int((diff.view(-1, 3) > 0).any(dim=1).sum())}
            else:
                tol = 2e-6 * b.abs() + 1e-6
                ok = bool((diff <= tol).all())
                parity = {"ok": ok, "split": "spp", "rule": "|sum_N - sum_1| <= 2e-6*|sum_1| + 1e-6 per channel (float add order)",
                          "max_abs_diff": float(diff.max()), "max_rel_diff": float((diff / (b.abs() + 1e-6)).max())}
        barrier()
    if not args.profile_in_timed:  # one extra (untimed) profiled step for the per-kernel durations
        st, _ = step(W + K, profile=True, alone=True)
        trace_ms, shade_ms, waves, prof_rays = st.trace_ms, st.shade_ms, st.profiled_waves, st.rays
    t = torch.tensor([span_ms, float(rays), float(launches), kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        tdist.all_reduce(tmax, op=tdist.ReduceOp.MAX)
        tsum = t.clone()
        tdist.all_reduce(tsum, op=tdist.ReduceOp.SUM)
        span_ms, rays_all, launches_all = float(tmax[0]), float(tsum[1]), int(tsum[2])
    else:
        rays_all, launches_all = float(rays), launches
    value = rays_all / span_ms / 1e3
    samples_all = float(NX) * NY * S_step * K

    # ---- e2e: C ABI with host buffers, every step builds, renders, reads back ----
    e2e = None
    if not args.no_e2e:
        host_fb = np.empty((NY, NX, 3), dtype=np.float32)
        h2d = d2h = 0
        e2e_rays = 0
        Ke = K
        barrier()
        t0 = time.perf_counter()
        bd = {"build": 0.0, "render_wall": 0.0, "render_device": 0.0, "reduce_readback": 0.0, "destroy": 0.0}
        for i in range(Ke):
            ta = time.perf_counter()
            s2 = pyrt.Scene(SCENE_ID, NX, NY, grid_half=GRID_HALF, texture_dir=tex, device=local)
            tb = time.perf_counter()
            st = s2.render(spp=spp_all, rng_mode=0, split_mode=1, rank=(W + i) * world + rank, world=n_steps_all * world)
            tc = time.perf_counter()
            if world > 1:
                rdist.reduce_sum_to_root(rdist.accum_tensor(s2))
                torch.cuda.synchronize()
            if rank == 0:
                s2.resolve(total_spp=S_step)
                pyrt._check(pyrt.lib().rt_readback(s2._h, host_fb.ctypes.data, None, None))
                d2h += host_fb.nbytes
            td = time.perf_counter()
            h2d += int(s2.info.h2d_bytes)
            e2e_rays += st.rays
            s2.close()
            te = time.perf_counter()
            bd["build"] += (tb - ta) * 1e3; bd["render_wall"] += (tc - tb) * 1e3; bd["render_device"] += st.device_ms
            bd["reduce_readback"] += (td - tc) * 1e3; bd["destroy"] += (te - td) * 1e3
        barrier()
        dt = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([dt, float(e2e_rays)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            tdist.all_reduce(tm, op=tdist.ReduceOp.MAX)
            ts = t.clone()
            tdist.all_reduce(ts, op=tdist.ReduceOp.SUM)
            dt, e2e_rays = float(tm[0]), float(ts[1])
        e2e = {"value": round(e2e_rays / dt / 1e3, 2), "unit": UNIT, "h2d_bytes_per_step": h2d // Ke,
               "d2h_bytes_per_step": d2h // Ke if rank == 0 else 0, "ms_per_step": round(dt / Ke, 3),
               "breakdown_ms_per_step": {k: round(v / Ke, 2) for k, v in bd.items()},
               "what": "per step: rt_build_scene (host generator + H2D scene/texture + device BVH build) + rt_render + "
                       "rt_readback(float fb -> host) + rt_destroy; wall clock, max over ranks"}

    if rank == 0:
        hbm, peak_src, sm_max = peaks()
        # ---- roofline of the dominant kernel, k_trace (70% of the kernel time): isolated launch duration measured live
        # with CUDA events on the render stream (one profiled step, one pool) x per-ray figures of the committed ncu capture
        n_launch = max(waves, 1)
        rays_per_launch = prof_rays / n_launch
        trace_s = trace_ms / n_launch / 1e3 if trace_ms > 0 else None
        shade_s = shade_ms / n_launch / 1e3 if shade_ms > 0 else None
        kt, ks, ncu_src = ncu_capture()
        if args.config != "c4":
            kt = ks = None  # the committed capture's per-ray figures are C4's (for C5 at 1 M spheres see profiles/*c5*)
        mb = microbench() if not args.no_microbench else None
        trace_rays_per_s = rays_per_launch / trace_s if trace_s else None
        roof = {"bound": "issue", "kernel": "k_trace", "achieved": None, "peak": None, "unit": "Gwarp-inst/s", "frac": None, "traffic": None}
        if kt and trace_rays_per_s:
            per = lambda k, d=kt: (d[k] / d["rays"]) if d.get(k) else None  # noqa: E731
            issue_peak = mb["issue_ginst_s"] if mb else 148 * 4 * sm_max / 1e3
            ach = per("warp_inst") * trace_rays_per_s / 1e9
            roof.update({
                "achieved": round(ach, 1), "peak": round(issue_peak, 1), "frac": round(ach / issue_peak, 4),
                "traffic": round(per("dram_bytes") * rays_per_launch),
                "what": "k_trace is bound by instruction issue, not by a memory level or the FP32 lanes: achieved = warp "
                        "instructions per ray (ncu capture) x rays per second of the isolated launch (CUDA events, this run); "
                        "peak = measured issue rate of the whole GPU (1 warp instruction per scheduler per clock)",
                "peak_source": mb["source"] if mb else "148 SMs x 4 schedulers x sm_max_mhz (data sheet)",
                "ncu_source": "profiles/" + ncu_src,
                "warp_inst_per_ray": round(per("warp_inst"), 1), "lanes_per_instruction": kt.get("threads_per_instruction"),
                "useful_lane_frac": round(ach / issue_peak * kt["threads_per_instruction"] / 32.0, 4),
                "rays_per_launch": round(rays_per_launch, 1), "launch_ms": round(trace_ms / n_launch, 5),
                "mrays_per_s_isolated": round(trace_rays_per_s / 1e6, 1),
                "share_of_kernel_time": {"k_trace": round(trace_ms / max(trace_ms + shade_ms, 1e-9), 4),
                                         "k_shade": round(shade_ms / max(trace_ms + shade_ms, 1e-9), 4)},
                # the other roofs of the same kernel, each achieved / MEASURED peak (all well below 1: none of them binds)
                "fp32": {"achieved_tflops": round(per("flop") * trace_rays_per_s / 1e12, 2),
                         "peak_tflops": mb["fp32_fma_tflops"] if mb else round(148 * 128 * 2 * sm_max / 1e6, 1),
                         "flop_per_ray": round(per("flop"), 1)},
                "l1": {"achieved_gbs": round(per("l1_bytes") * trace_rays_per_s / 1e9, 1), "peak_gbs": mb["l1_read_gbs"] if mb else None,
                       "bytes_per_ray": round(per("l1_bytes"), 1), "hit_pct": kt.get("l1_hit_pct")},
                "l2": {"achieved_gbs": round(per("l2_bytes") * trace_rays_per_s / 1e9, 1), "peak_gbs": mb["l2_read_gbs"] if mb else None,
                       "bytes_per_ray": round(per("l2_bytes"), 1), "hit_pct": kt.get("l2_hit_pct")},
                "hbm": {"achieved_gbs": round(per("dram_bytes") * trace_rays_per_s / 1e9, 1), "peak_gbs": hbm, "peak_source": peak_src,
                        "bytes_per_ray": round(per("dram_bytes"), 1),
                        "algorithmic_bytes_per_ray": 40, "algorithmic": "ray read 32 B + hit write 8 B (the scene is cache-resident)"},
            })
            for k in ("fp32", "l1", "l2", "hbm"):
                a_, p_ = (roof[k].get("achieved_tflops") or roof[k].get("achieved_gbs")), (roof[k].get("peak_tflops") or roof[k].get("peak_gbs"))
                roof[k]["frac"] = round(a_ / p_, 4) if a_ and p_ else None
            if ks and shade_s:
                sr = rays_per_launch / shade_s
                roof["k_shade"] = {"bound": "hbm latency (gathers in queue order)", "mrays_per_s_isolated": round(sr / 1e6, 1),
                                   "hbm_achieved_gbs": round(ks["dram_bytes"] / ks["rays"] * sr / 1e9, 1), "hbm_peak_gbs": hbm,
                                   "hbm_frac": round(ks["dram_bytes"] / ks["rays"] * sr / 1e9 / hbm, 4),
                                   "dram_bytes_per_ray": round(ks["dram_bytes"] / ks["rays"], 1), "algorithmic_bytes_per_ray": 140,
                                   "algorithmic": "read 4 x float4 state + float2 hit + int queue entry = 76 B, write 4 x float4 = 64 B"}
            # labelled extra (round-1 figure): what the REFERENCE algorithm would have to move per ray if nothing were cached
            roof["hbm_equivalent_of_reference_algorithm"] = {
                "bytes_per_ray": C4_TRACE_BYTES_PER_RAY, "gbs": round(C4_TRACE_BYTES_PER_RAY * trace_rays_per_s / 1e9, 1),
                "note": "SURVEY.md 8d normaliser; NOT a roof of this kernel (the scene is L1/L2-resident), kept for comparison with round 1"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(span_ms / K, 3), "higher_is_better": True, "scaling": "weak" if weak else "strong",
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("Book-2 final scene (create_world_final + earthmap) 800x800, depth 50, %d spp per step "
                                    "%s; default 10 steps = the 10000-spp config" if args.config == "c4" else
                                    CONFIGS[args.config][5] + " (create_world_bouncing with the grid widened, main.cu:140-141, 160-244), "
                                    "depth 50, %d spp per step %s") %
                                   (S_step, "(%d per GPU)" % S if weak else "split over the %d GPU(s)" % world),
                       "name": args.config, "objects": int(sc_info_n_top), "bvh_nodes": int(sc_info_nodes), "bvh_build_ms": round(sc_info_bvh_ms, 3),
                       "scene_id": SCENE_ID, "nx": NX, "ny": NY, "spp_per_step": S_step, "spp_total": S_step * K,
                       "max_depth": 50, "rng": "philox4x32-10", "parallelism": "%s-split x%d, scene replicated" % (args.split if world > 1 else "spp", world),
                       "l2": "inputs larger than L2: %.0f MB of path state (%d paths in flight x 104 B: two ping-pong sets of 48 B + an 8 B hit "
                             "record) streamed every wave" % (st.n_slots * 104 / 1e6, st.n_slots)},
            "msamples_per_s": round(samples_all / span_ms / 1e3, 3),
            "rays": int(rays_all), "rays_per_sample": round(rays_all / samples_all, 4),
            "kernel_ms_per_step": round(kernel_ms / K, 3),
            "gpu_launches": launches_all,
            "clocks": ck,
            "parity_check": parity,
            "e2e": e2e,
            "roofline": roof,
            "microbench": mb,
        }
        if world == 1 and not args.no_reference_cuda:
            sc.close()
            line["reference_cuda_sm100"] = reference_cuda_run()
            rc = line["reference_cuda_sm100"]
            if rc and "value" in rc and rc["value"]:
                line["speedup_vs_reference_cuda_sm100"] = round(value / rc["value"], 1)
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                ns, _ = cpu_sample_plan(threads, 12.0)
                o = cpu_reference_run(NX, NY, ns, threads, count=1)
                line["cpu_baseline"] = {"value": round(o["mrays_per_s"], 4), "unit": UNIT, "cores": threads,
                                        "kind": o.get("kind", "reference"),
                                        "sample": "%dx%d x %d spp of the same scene (reference render() compiled for the "
                                                  "host behind oracle/shim)" % (NX, NY, ns)}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.barrier(device_ids=[local])
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
