// rt_kernels.cuh — the wavefront render kernels (device only).
//
// The reference renders with ONE megakernel: a thread per pixel loops ns samples x <= 50 bounces
// through virtual calls and device recursion (render/color, main.cu:44-133). Here the same
// integrator runs as waves over a pool of paths whose state STREAMS through two ping-pong sets of
// SoA arrays in HBM/L2 (no persistent slots, no gathers through an indirection table):
//
//   k_init     the first n_slots camera samples (render_init + main.cu:119-123); reference-RNG mode: XORWOW seeding
//   k_trace    pure closest-hit kernel (main.cu:57; bvh.cuh:95). A WARP owns RT_RANGE consecutive rays of the wave's
//              dense layout and runs them through the 4-wide BVH 32 at a time, in warp-wide node / leaf / media phases
//              (rt_intersect.cuh); when all 32 are done the lanes store their hits and take the range's next 32 rays.
//              No block-level barrier anywhere. At the end the warp bins its rays by the material class of
//              the hit into shade queues (one atomic per class per warp, spread over RT_NSUB sub-queues).
//   k_shade    one material class per warp, reading the path state in queue order (neighbouring entries were written
//              by neighbouring trace lanes): miss/background, emission, scatter, throughput (main.cu:58-83). A finished
//              sample is added to its pixel and the lane takes the NEXT camera sample right away (path regeneration,
//              main.cu:119-124); the new state is written densely at the thread's own position of the other
//              ping-pong set, which is the layout the next k_trace walks. A block walks RT_SHADE_ITEMS chunks of
//              queue positions with the next chunk's loads in flight, and a warp reserves the work items of all its
//              miss / light lanes with one atomic.
//   k_finish   the tail of a job: the last few ten thousand paths are run to their end in one kernel
//   k_accumulate / k_resolve   sums -> linear accumulation buffer -> 1/ns, gamma     (main.cu:128-132)
//   k_aov      primary-hit object/material id + t for the centre ray of every pixel
//
// Philox mode (production): a path is a worker. Work item w = sample * n_local_pixels + local_pixel is
// handed out by one 64-bit counter (one atomic per warp that finished samples), so every lane stays
// busy until the whole job is done no matter how unevenly path lengths are spread over the image (on
// the Book-2 final scene a pixel-bound scheme ran at 11% mean occupancy). A finished sample is
// added to its pixel with 64-bit FIXED-POINT atomics (2^-32 resolution): integer sums do not depend on
// the order of arrival, so the image is bit-reproducible and identical under any tile split.
// Reference-RNG mode (validation): a path IS a pixel and consumes that pixel's XORWOW stream in exactly
// the reference's order, samples one after the other, summed in float like `col += color(...)`.
#pragma once
#include "rt_shade.cuh"

namespace rt {

#ifndef RT_NSUB
#define RT_NSUB 4                       // sub-queues per material class: spreads the per-warp queue atomics over 4 addresses
#endif
#define RT_NQ (Q_COUNT * RT_NSUB)       // shade queues of a wave (<= 32: k_shade scans them with one warp scan)
#ifndef RT_RANGE
#define RT_RANGE 64                     // consecutive rays of the dense layout a trace warp owns (A/B on C4: 32 -4%, 128 -2%)
#endif
#define RT_RANGE_K (RT_RANGE / 32)

struct PathArrays {
  float4* ray_o[2];  // origin.xyz, time                                    [2]: ping-pong sets, k_shade reads [parity], writes [parity ^ 1]
  float4* ray_d[2];  // direction.xyz, local pixel (int bits; < 0 = hole: no path at this position)
  float4* thr[2];    // throughput.rgb, bounce | sample number << 8 (int bits)
                     // (no radiance in the state: of the reference's materials only diffuse_light emits and it never scatters
                     //  (material.cuh:174-190), and a miss ends the path, so `radiance += throughput * emitted` (main.cu:71) is
                     //  non-zero only in the event that ends the sample - it is computed and added to the pixel there)
  float2* hit;       // k_trace -> k_shade: t, packed hit (object | face << 25 | class << 28; -1 miss; -2 hole)
  float4* col;       // reference-RNG mode: float sum of the pixel's finished samples (per local pixel)
  uint32_t* rng;     // reference-RNG mode: 6 words per local pixel, SoA [6][n_slots]
  unsigned long long* acc64;  // Philox mode: per local pixel 3 x fixed-point (2^-32) radiance sums
  unsigned long long* acc64_odd;  // adaptive sampling: the same sums over the ODD sample numbers only (error estimate), else null
  unsigned long long* next_work;  // Philox mode: next work item to hand out (shared by all pools)
  int* queues;       // [RT_NQ][subcap] positions of the dense layout, by (class, sub-queue)
  int subcap;
};

struct WaveCounters {
  int n_queue[2][RT_NQ];       // shade queue fill, by wave parity (their sum = rays traced in that wave)
  int order_len;               // length of the dense layout the next k_trace walks
  unsigned int overflow;       // traversal stack overflow flag (must stay 0)
  unsigned int nonfinite;      // Philox mode: samples dropped because their radiance was inf/NaN/out of fixed-point range
  unsigned int pad;
  unsigned long long rays;     // closest-hit queries issued (= the reference's bounce-loop iterations)
};

struct RenderParams {
  int nx, ny;            // full image
  float inv_nx;          // 1.0f / nx
  int rows_local;        // scanlines owned by this rank (tile split: j = lr * world + rank)
  int rank, world;
  int n_slots;           // paths in flight (reference-RNG mode: one per local pixel)
  long long work_total;  // rows_local * nx * sample_count work items (Philox mode)
  int sample_base;       // first sample number of this rank's share (spp split), else 0
  int sample_count;      // samples per pixel this rank renders
  int max_depth;         // 50
  float tmin;            // 0.001f
  V3 background; int gradient;
  unsigned long long seed;  // 1984
  const int* active;     // adaptive passes: the local pixels that still take samples (null: all rows_local * nx of them)
  int n_active;
};

enum RngMode : int { RNG_PHILOX = 0, RNG_REFERENCE = 1 };
#define RT_HIT_MISS (-1)
#define RT_HIT_HOLE (-2)

struct SlotInfo { int lpix, i, j, pix; };
RT_D SlotInfo pixel_info(const RenderParams& P, int lpix) {
  SlotInfo s;
  s.lpix = lpix;
  // lpix / nx without the ~25-instruction integer divide: float estimate (off by at most one for any image with
  // fewer than 2^22 rows) plus one correction step
  int lr = __float2int_rz(__fmul_rz(__int2float_rz(lpix), P.inv_nx));
  int i = lpix - lr * P.nx;
  if (i < 0) { --lr; i += P.nx; } else if (i >= P.nx) { ++lr; i -= P.nx; }
  s.i = i;
  s.j = lr * P.world + P.rank;
  s.pix = s.j * P.nx + s.i;  // the reference's pixel_index (main.cu:115)
  return s;
}

template <int MODE> struct RngOf;
template <> struct RngOf<RNG_PHILOX> { typedef Philox type; };
template <> struct RngOf<RNG_REFERENCE> { typedef Xorwow type; };

RT_D void rng_load(Philox& g, const PathArrays&, const RenderParams& P, int, int pix, int sample, int stage) {
  g.init(P.seed, (uint32_t)pix, (uint32_t)sample, (uint32_t)stage);
}
RT_D void rng_load(Xorwow& g, const PathArrays& A, const RenderParams& P, int lpix, int, int, int) {
  const int n = P.n_slots;
  g.d = A.rng[lpix]; g.v0 = A.rng[n + lpix]; g.v1 = A.rng[2 * n + lpix]; g.v2 = A.rng[3 * n + lpix];
  g.v3 = A.rng[4 * n + lpix]; g.v4 = A.rng[5 * n + lpix];
}
RT_D void rng_store(const Philox&, const PathArrays&, const RenderParams&, int) {}
RT_D void rng_store(const Xorwow& g, const PathArrays& A, const RenderParams& P, int lpix) {
  const int n = P.n_slots;
  A.rng[lpix] = g.d; A.rng[n + lpix] = g.v0; A.rng[2 * n + lpix] = g.v1; A.rng[3 * n + lpix] = g.v2;
  A.rng[4 * n + lpix] = g.v3; A.rng[5 * n + lpix] = g.v4;
}

// New camera sample (main.cu:121-123): jitter, lens, shutter time; throughput 1, radiance 0. Written at position
// `at` of ping-pong set `pp`; ray_d.w carries the local pixel, thr.w bounce 0 and the sample number.
template <class RNG>
RT_D void start_sample(const DScene& S, const RenderParams& P, const PathArrays& A, int pp, int at, const SlotInfo& si, int sample, RNG& g) {
  const float u = fdiv(fadd((float)si.i, g.uniform()), (float)P.nx);
  const float v = fdiv(fadd((float)si.j, g.uniform()), (float)P.ny);
  const Ray r = camera_get_ray(S.cam, u, v, g);
  A.ray_o[pp][at] = make_float4(r.o.x, r.o.y, r.o.z, r.tm);
  A.ray_d[pp][at] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(si.lpix));
  A.thr[pp][at] = make_float4(1.f, 1.f, 1.f, __int_as_float(sample << 8));
}
RT_D void write_hole(const PathArrays& A, int pp, int at) { A.ray_d[pp][at] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); }

#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif
#ifndef RT_SHADE_MINB
#define RT_SHADE_MINB 3   // resident blocks per SM the shade kernel is compiled for (register cap 65536 / (256 * MINB))
#endif
#ifndef RT_TBLOCK
#define RT_TBLOCK 64      // k_trace block size: no block-level synchronisation, so small blocks retire warp by warp
#endif
#define RT_TWARPS (RT_TBLOCK / 32)

#define RT_FIXED_ONE 4294967296.0f  /* 2^32 */
#define RT_FIXED_MAX 1048576.0f     /* samples at or above 2^20 are counted as dropped: 2^11 of them still fit the Q31.32 sum */
RT_D void fixed_add(unsigned long long* acc, float v) {
  atomicAdd(acc, (unsigned long long)__float2ll_rn(v * RT_FIXED_ONE));  // two's complement: negative values add correctly
}
RT_D void work_to_pixel_sample(const RenderParams& P, unsigned long long w, int& lpix, int& sample) {
  const unsigned n_px = (unsigned)(P.active ? P.n_active : P.rows_local * P.nx);
  if (P.work_total <= 0xFFFFFFFFll) {  // the usual case: 32-bit divide
    const unsigned w32 = (unsigned)w;
    const unsigned s = w32 / n_px;
    lpix = (int)(w32 - s * n_px);
    sample = P.sample_base + (int)s;
  } else {
    const unsigned long long s = w / (unsigned long long)n_px;
    lpix = (int)(w - s * n_px);
    sample = P.sample_base + (int)s;
  }
  if (P.active) lpix = P.active[lpix];
}

// The first wave: position = slot. Philox mode: work item = work_base + slot (the host starts the work counter behind them);
// reference-RNG mode: slot = local pixel, its stream is seeded (render_init, main.cu:104) and draws its first sample.
template <int MODE>
__global__ void __launch_bounds__(RT_BLOCK) k_init(DScene S, RenderParams P, PathArrays A, int work_base) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.n_slots) return;
  if constexpr (MODE == RNG_REFERENCE) {
    const SlotInfo si = pixel_info(P, slot);
    A.col[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    Xorwow g;
    g.init((unsigned long long)(long long)(1984 + si.pix));
    start_sample(S, P, A, 0, slot, si, P.sample_base, g);
    rng_store(g, A, P, slot);
  } else {
    int lpix, sample;
    work_to_pixel_sample(P, (unsigned long long)(work_base + slot), lpix, sample);
    const SlotInfo si = pixel_info(P, lpix);
    Philox g;
    rng_load(g, A, P, lpix, si.pix, sample, 0);
    start_sample(S, P, A, 0, slot, si, sample, g);
  }
}

#ifndef RT_TRACE_MINB
#define RT_TRACE_MINB (2048 / RT_TBLOCK / 2)  // resident k_trace blocks per SM the kernel is compiled for: 32 warps, 64 registers
#endif
#ifndef RT_REFILL_MIN
#define RT_REFILL_MIN 32    // finished lanes that trigger a refill from the warp's range (A/B on C4: 8 -9%, 16 -6%: refilled lanes
                            // start at the root while the others are deep in the tree, which desynchronises the warp-wide phases).
                            // Round 2 also tried PERSISTENT warps that refill from chunks of the dense layout taken from a
                            // counter (or by a fixed stride, next rays prefetched) with two chunk slots in flight: node phases
                            // per ray 0.375 -> 0.288 (stats build) and 10% fewer warp instructions (ncu r02o), but C4 3.71 /
                            // 3.48 Grays/s against 4.04: every partial refill runs the refill, media-phase and binning code at
                            // 8..16 of 32 lanes, and issue utilisation fell from 66% to 55% (long-scoreboard stalls 3.0 -> 4.6).
#endif
template <bool CULL>
RT_D void trace_ranges(const DScene& S, float tmin, const float4* ray_o, const float4* ray_d, float2* hit, int* queues, int subcap,
                       WaveCounters* C, int parity) {
  typedef TravT<CULL> Trav;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int order_len = C->order_len;
  const int base = wid * RT_RANGE;
  if (base >= order_len) return;
  const int end = min(base + RT_RANGE, order_len);
  int next = base;  // warp-uniform: first ray of the range nobody has taken yet
  int gid = -1;     // the ray this lane is tracing
  Trav T;
  RT_TRAV_ARRAYS(m, CULL);
  T.reset();
  bool flush = false;  // warp-uniform: a deferred media list is nearly full
  while (true) {
#if RT_REFILL_MIN >= 32
    // The warp refills when ALL its lanes are done, and that is exactly when no lane can expand a node and no lane has a leaf
    // pending: the two ballots the phase choice needs anyway say so (no ballot of finished lanes, no range test per turn).
    const bool can = T.can_expand();
    const unsigned mexp = __ballot_sync(0xFFFFFFFFu, can);
    const unsigned mleaf = __ballot_sync(0xFFFFFFFFu, T.nl > 0);
    const bool refill = (mexp | mleaf) == 0u;
    const bool fin = refill;
    const unsigned mfin = 0xFFFFFFFFu;
    // media phase: the deferred media, before the hits are stored (or right away when a list is nearly full)
    if (flush || refill) {
      if (flush || __any_sync(0xFFFFFFFFu, T.nm > 0)) {
        T.media_phase(S, m_mq_tlp, m_mq_tn, true);
        flush = false;
        T.revalidate(m_stack);
        continue;
      }
    }
#else
    const bool fin = T.finished();
    const unsigned mfin = __ballot_sync(0xFFFFFFFFu, fin);
    const bool refill = mfin == 0xFFFFFFFFu || (next < end && __popc(mfin) >= RT_REFILL_MIN);
    // media phase: the finished lanes' deferred media, before their hits are stored (or everybody's when a list is nearly full)
    if (flush || (refill && __any_sync(0xFFFFFFFFu, fin && T.nm > 0))) {
      T.media_phase(S, m_mq_tlp, m_mq_tn, flush || fin);
      flush = false;
      T.revalidate(m_stack);
      continue;
    }
#endif
    if (refill) {
      // ---------------- refill: finished lanes store their hit and take the next rays of the range ----------------
      if (fin) {
        if (gid >= 0) {
          const Hit h = T.best.resolve(S);  // RT_HIT_MISS = -1 or the packed hit word
          hit[gid] = make_float2(h.t, __int_as_float(h.tlp));
          gid = -1;
        }
        const int g = next + __popc(mfin & lt);
        if (g < end) {
          const float4 d = ray_d[g], o = ray_o[g];  // both at once: one round trip (a hole's origin is allocated, never used).
                                                    // Prefetching the range's second half here: -2 % (A/B r02zz)
          if (__float_as_int(d.w) >= 0) {
            Ray r; r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
            T.begin(r, tmin, FLT_MAX);
            gid = g;
          } else {
            hit[g] = make_float2(0.f, __int_as_float(RT_HIT_HOLE));
          }
        }
      }
      if (next >= end) break;  // everyone finished and the range is drained
      next += __popc(mfin);
      continue;
    }
#if RT_REFILL_MIN < 32
    const bool can = T.can_expand();
    const unsigned mexp = __ballot_sync(0xFFFFFFFFu, can);
    const unsigned mleaf = __ballot_sync(0xFFFFFFFFu, T.nl > 0);
#endif
    const unsigned mwait = __ballot_sync(0xFFFFFFFFu, !T.have() && T.nl > 0);  // traversal done, leaves pending
    if (Trav::pick_node_phase(mexp, mleaf, mwait)) {
      if (lane == 0) RT_COUNT(4, 1);
#ifdef RT_STATS
      { const unsigned mblk = __ballot_sync(0xFFFFFFFFu, T.have() && !can); const unsigned mfin_s = __ballot_sync(0xFFFFFFFFu, T.finished()); if (lane == 0) { atomicAdd(&g_stats2[0], (unsigned long long)__popc(mfin_s)); atomicAdd(&g_stats2[1], (unsigned long long)__popc(mblk)); atomicAdd(&g_stats2[2], (unsigned long long)__popc(mexp)); } }
#endif
      if (can) T.node_step(S, &C->overflow, RT_TRAV_ARGS(m));
    } else {
      if (lane == 0) RT_COUNT(5, 1);
      flush = T.leaf_phase(S, m_lq_ref, m_lq_tlp, m_lq_tn, m_mq_tlp, m_mq_tn);
      T.revalidate(m_stack);
    }
  }
#if RT_RANGE <= 64
  // ---- bin the range by material class: ONE atomic per class per warp, positions by ballot ----
  __syncwarp();
  int qk[RT_RANGE_K];
  int mycnt = 0;  // lane c < Q_COUNT: rays of class c in this range
#pragma unroll
  for (int k = 0; k < RT_RANGE_K; ++k) {
    const int g = base + 32 * k + lane;
    int q = -1;
    if (g < end) {
      const int packed = __float_as_int(hit[g].y);
      if (packed != RT_HIT_HOLE) q = packed < 0 ? (int)Q_MISS : ((packed >> 28) & 7);
    }
    qk[k] = q;
#pragma unroll
    for (int c = 0; c < Q_COUNT; ++c) {
      const int n = __popc(__ballot_sync(0xFFFFFFFFu, q == c));
      if (lane == c) mycnt += n;
    }
  }
  const int sub = wid & (RT_NSUB - 1);
  int mybase = 0;
  if (lane < Q_COUNT && mycnt > 0) mybase = atomicAdd(&C->n_queue[parity][lane * RT_NSUB + sub], mycnt);
#pragma unroll
  for (int c = 0; c < Q_COUNT; ++c) {
    int off = __shfl_sync(0xFFFFFFFFu, mybase, c);
    int* qp = queues + (size_t)(c * RT_NSUB + sub) * subcap;
#pragma unroll
    for (int k = 0; k < RT_RANGE_K; ++k) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, qk[k] == c);
      if (qk[k] == c) qp[off + __popc(m & lt)] = base + 32 * k + lane;
      off += __popc(m);
    }
  }
#else  // long ranges: the same in two rolled passes over the range's hits (just written by this warp: L1 / L2 hits)
  __syncwarp();
  int mycnt = 0;  // lane c < Q_COUNT: rays of class c in this range
#pragma unroll 1
  for (int g0 = base; g0 < end; g0 += 32) {
    const int g = g0 + lane;
    int q = -1;
    if (g < end) {
      const int packed = __float_as_int(hit[g].y);
      if (packed != RT_HIT_HOLE) q = packed < 0 ? (int)Q_MISS : ((packed >> 28) & 7);
    }
#pragma unroll
    for (int c = 0; c < Q_COUNT; ++c) {
      const int n = __popc(__ballot_sync(0xFFFFFFFFu, q == c));
      if (lane == c) mycnt += n;
    }
  }
  const int sub = wid & (RT_NSUB - 1);
  int mybase = 0;  // lane c: next free position of class c's sub-queue for this warp
  if (lane < Q_COUNT && mycnt > 0) mybase = atomicAdd(&C->n_queue[parity][lane * RT_NSUB + sub], mycnt);
#pragma unroll 1
  for (int g0 = base; g0 < end; g0 += 32) {
    const int g = g0 + lane;
    int q = -1;
    if (g < end) {
      const int packed = __float_as_int(hit[g].y);
      if (packed != RT_HIT_HOLE) q = packed < 0 ? (int)Q_MISS : ((packed >> 28) & 7);
    }
#pragma unroll
    for (int c = 0; c < Q_COUNT; ++c) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, q == c);
      const int off = __shfl_sync(0xFFFFFFFFu, mybase, c);
      if (q == c) queues[(size_t)(c * RT_NSUB + sub) * subcap + off + __popc(m & lt)] = g;
      if (lane == c) mybase += __popc(m);
    }
  }
#endif
}
__global__ void __launch_bounds__(RT_TBLOCK, RT_TRACE_MINB) k_trace(DScene S, float tmin, const float4* ray_o, const float4* ray_d,
                                                                    float2* hit, int* queues, int subcap, WaveCounters* C, int parity) {
  trace_ranges<RT_CULL_DEFAULT>(S, tmin, ray_o, ray_d, hit, queues, subcap, C, parity);
}
#ifndef RT_CULL_MIN_NODES
#define RT_CULL_MIN_NODES 16  // C2 / C3 (3 nodes): 6286 / 5947 Mrays/s with the check, 6351 / 6133 without; every other config has >= 233 nodes
#endif
// The same for scenes of a handful of BVH nodes, without the stack entries' distances (nothing to drop there; rt_host.cu picks)
__global__ void __launch_bounds__(RT_TBLOCK, RT_TRACE_MINB) k_trace_small(DScene S, float tmin, const float4* ray_o, const float4* ray_d,
                                                                          float2* hit, int* queues, int subcap, WaveCounters* C, int parity) {
  trace_ranges<false>(S, tmin, ray_o, ray_d, hit, queues, subcap, C, parity);
}

// One event of a path (main.cu:57-83): the hit `hh` of ray r is shaded - miss / background, emission, scatter, throughput.
// Returns true when the sample has ended; its radiance is then already added to the pixel (reference-RNG mode: the float
// sum `col += color(...)`, main.cu:124; Philox mode: fixed-point atomics). Otherwise r / thr / bounce hold the next ray.
// q = shade class of the hit (Q_MISS for a miss). Used by k_shade (one class per warp) and by k_finish (tail of a job).
template <int MODE, class RNG>
RT_D bool shade_event(const DScene& S, const RenderParams& P, const PathArrays& A, WaveCounters* C, int q, float2 hh, Ray& r, V3& thr,
                      int& bounce, int lpix, int sample, RNG& g) {
  V3 rad = v3(0.f, 0.f, 0.f);
  bool sample_done;
  Ray nxt; nxt.o = r.o; nxt.d = r.d; nxt.tm = r.tm;
  if (q == Q_MISS) {
    // main.cu:58-68
    V3 bg = P.background;
    if (P.gradient) {
      const float uy = fdiv(r.d.y, vlen(r.d));
      const float t = fmul(0.5f, fadd(uy, 1.0f));
      const float omt = fsub(1.0f, t);
      bg = v3(ffma(t, 0.5f, omt), ffma(t, 0.7f, omt), fadd(t, omt));
    }
    rad = v3(ffma(thr.x, bg.x, rad.x), ffma(thr.y, bg.y, rad.y), ffma(thr.z, bg.z, rad.z));
    sample_done = true;
  } else {
    const int packed = __float_as_int(hh.y);
    const int tlp = packed & (int)RT_TLP_MASK, face = (packed >> 25) & 7;
    const DTlp T = S.tlp[tlp];
    const DMat m = S.mats[T.mat];
    Rec rec;
    if (ref_type(T.ref) == G_MEDIUM) {  // constant_medium.cuh:58-62
      rec.t = hh.x;
      rec.p = vmad(hh.x, r.d, r.o);
      rec.n = v3(1, 0, 0);
      rec.u = rec.v = 0.f;
    } else {
      geom_hit<true>(S, T.ref, r, P.tmin, FLT_MAX, m.needs_uv != 0, rec, face);
    }
    if (q == Q_LIGHT) {  // main.cu:71: radiance += throughput * emitted
      const V3 e = material_emitted(S, m, rec);
      rad = v3(ffma(thr.x, e.x, rad.x), ffma(thr.y, e.y, rad.y), ffma(thr.z, e.z, rad.z));
    }
    V3 att;
    const bool scattered = material_scatter(S, m, r, rec, g, att, nxt);  // main.cu:76
    if (scattered) {
      thr = vmul(thr, att);  // main.cu:82
      ++bounce;
    }
    sample_done = !scattered || bounce >= P.max_depth;
  }
  if (sample_done) {
    if constexpr (MODE == RNG_REFERENCE) {
      // render: col += color(...) (main.cu:124)
      float4 c = A.col[lpix];
      c.x = fadd(c.x, rad.x); c.y = fadd(c.y, rad.y); c.z = fadd(c.z, rad.z);
      A.col[lpix] = c;
    } else {
      // non-finite samples (and samples too large for the fixed-point sum) are dropped and counted
      if (fabsf(rad.x) < RT_FIXED_MAX && fabsf(rad.y) < RT_FIXED_MAX && fabsf(rad.z) < RT_FIXED_MAX) {
        unsigned long long* acc = A.acc64 + 3 * (size_t)lpix;
        fixed_add(acc + 0, rad.x); fixed_add(acc + 1, rad.y); fixed_add(acc + 2, rad.z);
        if (A.acc64_odd && (sample & 1)) {
          unsigned long long* ao = A.acc64_odd + 3 * (size_t)lpix;
          fixed_add(ao + 0, rad.x); fixed_add(ao + 1, rad.y); fixed_add(ao + 2, rad.z);
        }
      } else {
        atomicAdd(&C->nonfinite, 1u);
      }
    }
    return true;
  }
  r = nxt;
  return false;
}

#ifndef RT_SHADE_ITEMS
#define RT_SHADE_ITEMS 8   // consecutive chunks of RT_BLOCK queue positions a k_shade block walks (software pipeline, see below)
#endif
RT_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// k_shade is bound by the LATENCY of its gathers (queue entry -> path state -> scene records: ncu r02h, 60% of the stall
// samples on the long scoreboard, 36% of the HBM bandwidth with 19 resident warps per SM), so a block walks RT_SHADE_ITEMS
// chunks and keeps the next chunk's memory traffic in flight under the current chunk's shading: the next queue entry is
// loaded into a register one chunk ahead, and as soon as it has arrived the four state records it points to are
// prefetched into L2 (no destination registers).
template <int MODE>
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINB) k_shade(DScene S, RenderParams P, PathArrays A, WaveCounters* C, int parity) {
  const int lane = threadIdx.x & 31;
  // warp -> (queue, position): the queues are laid end to end, each padded to a whole warp; one warp scan finds the warp's queue
  const int cnt = lane < RT_NQ ? C->n_queue[parity][lane] : 0;
  const int padded = (cnt + 31) & ~31;
  int incl = padded;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
  const int excl = incl - padded;
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int rays = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane < RT_NQ) C->n_queue[parity ^ 1][lane] = 0;  // the next wave's trace fills these
    if (lane == 0) { C->rays += (unsigned long long)rays; C->order_len = total; }  // the next k_trace walks this wave's layout
  }
  const int po = parity ^ 1;
  // queue entry of position g (a whole warp asks for its 32 consecutive positions): -1 = padding between two queues
  auto entry_of = [&](int g, int& qi_out) -> int {
    const int ws = g - lane;
    const unsigned mq = __ballot_sync(0xFFFFFFFFu, lane < RT_NQ && excl <= ws);
    const int qi = 31 - __clz(mq);
    const int qbase = __shfl_sync(0xFFFFFFFFu, excl, qi), qcount = __shfl_sync(0xFFFFFFFFu, cnt, qi);
    qi_out = qi;
    const int pos = g - qbase;
    return pos < qcount ? A.queues[(size_t)qi * A.subcap + pos] : -1;
  };
  int gid = (blockIdx.x * RT_SHADE_ITEMS) * RT_BLOCK + threadIdx.x;
  if (gid - lane >= total) return;
  // Miss and light warps end every sample they hold (main.cu:58-68; lights do not scatter, material.cuh:174-178), so how many
  // new work items this warp will need for them is known from the queue layout alone. They are reserved NOW, for all the
  // warp's chunks with ONE atomic on the work counter: its round trip overlaps the first state loads, and the counter - one
  // address for the whole GPU - sees 1/RT_SHADE_ITEMS of the atomics (one per warp and chunk was 17% of k_shade's stall
  // samples, ncu r02h).
  unsigned long long w_early = 0;  // this warp's reservation: items w_early .. are handed to its miss / light lanes in order
#ifndef RT_NO_EARLY_BLOCK
  if constexpr (MODE == RNG_PHILOX) {
    int n_early = 0;
#pragma unroll 1
    for (int rep = 0, ws = gid - lane; rep < RT_SHADE_ITEMS && ws < total; ++rep, ws += RT_BLOCK) {
      const unsigned mq = __ballot_sync(0xFFFFFFFFu, lane < RT_NQ && excl <= ws);
      const int qq = 31 - __clz(mq), cls = qq / RT_NSUB;
      const int qbase = __shfl_sync(0xFFFFFFFFu, excl, qq), qcount = __shfl_sync(0xFFFFFFFFu, cnt, qq);
      if (cls == Q_MISS || cls == Q_LIGHT) n_early += max(0, min(32, qcount - (ws - qbase)));
    }
    if (n_early > 0) {
      if (lane == 0) w_early = atomicAdd(A.next_work, (unsigned long long)n_early);
      w_early = __shfl_sync(0xFFFFFFFFu, w_early, 0);  // (one atomic per BLOCK through shared memory: no further gain, A/B r02u)
    }
  }
#endif
  int qi = 0;
  int idx = entry_of(gid, qi);
#pragma unroll 1
  for (int rep = 0; rep < RT_SHADE_ITEMS; ++rep, gid += RT_BLOCK) {
    if (gid - lane >= total) break;  // warp-uniform
    const int q = qi / RT_NSUB;  // material class: warp-uniform
    const bool live = idx >= 0;  // else: warp padding between two queues
    // next chunk: its queue entry starts its way up now
    int qi_next = 0, idx_next = -1;
    const bool more = rep + 1 < RT_SHADE_ITEMS && gid + RT_BLOCK - lane < total;  // warp-uniform
    if (more) idx_next = entry_of(gid + RT_BLOCK, qi_next);
    bool need = false;     // the sample ended: take the next one
    typename RngOf<MODE>::type g;
    int lpix = 0, sample = 0;
    const bool early = MODE == RNG_PHILOX && (q == Q_MISS || q == Q_LIGHT);  // this chunk's lanes take their items from the reservation
    unsigned long long w_mine = 0;
    if (early) {
      const unsigned mlive = __ballot_sync(0xFFFFFFFFu, live);
#ifdef RT_NO_EARLY_BLOCK
      const int leader = __ffs(mlive) - 1;
      if (lane == leader) w_early = atomicAdd(A.next_work, (unsigned long long)__popc(mlive));
      w_early = __shfl_sync(0xFFFFFFFFu, w_early, leader);
      w_mine = w_early + (unsigned long long)__popc(mlive & ((1u << lane) - 1u));
#else
      w_mine = w_early + (unsigned long long)__popc(mlive & ((1u << lane) - 1u));
      w_early += (unsigned long long)__popc(mlive);
#endif
    }
    if (live) {
      const float4 o = A.ray_o[parity][idx], d = A.ray_d[parity][idx];
      const float4 thr4 = A.thr[parity][idx];
      const float2 hh = A.hit[idx];
      lpix = __float_as_int(d.w);
      if (idx_next >= 0) {  // the state loads above have arrived, so has the next queue entry: prefetch what it points to
        prefetch_l2(A.ray_o[parity] + idx_next); prefetch_l2(A.ray_d[parity] + idx_next);
        prefetch_l2(A.thr[parity] + idx_next); prefetch_l2(A.hit + idx_next);
      }
      const SlotInfo si = pixel_info(P, lpix);
      Ray r; r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
      V3 thr = v3(thr4.x, thr4.y, thr4.z);
      int bounce = __float_as_int(thr4.w) & 255;
      sample = __float_as_int(thr4.w) >> 8;
      rng_load(g, A, P, lpix, si.pix, sample, bounce + 1);
      if (shade_event<MODE>(S, P, A, C, q, hh, r, thr, bounce, lpix, sample, g)) {
        need = true;
      } else {
        A.ray_o[po][gid] = make_float4(r.o.x, r.o.y, r.o.z, r.tm);
        A.ray_d[po][gid] = make_float4(r.d.x, r.d.y, r.d.z, d.w);
        A.thr[po][gid] = make_float4(thr.x, thr.y, thr.z, __int_as_float(bounce | (sample << 8)));
      }
    } else {
      write_hole(A, po, gid);
      if (idx_next >= 0) {
        prefetch_l2(A.ray_o[parity] + idx_next); prefetch_l2(A.ray_d[parity] + idx_next);
        prefetch_l2(A.thr[parity] + idx_next); prefetch_l2(A.hit + idx_next);
      }
    }
    // ---- path regeneration (main.cu:119-123): a lane whose sample has ended takes the next one ----
    if constexpr (MODE == RNG_PHILOX) {
      const unsigned mneed = __ballot_sync(0xFFFFFFFFu, need);
      if (mneed) {
        unsigned long long w = w_mine;
        if (!early) {
          const int leader = __ffs(mneed) - 1;
          if (lane == leader) w = atomicAdd(A.next_work, (unsigned long long)__popc(mneed));  // one atomic per warp
          w = __shfl_sync(0xFFFFFFFFu, w, leader) + (unsigned long long)__popc(mneed & ((1u << lane) - 1u));
        }
        if (need) {
          if (w < (unsigned long long)P.work_total) {
            work_to_pixel_sample(P, w, lpix, sample);
            const SlotInfo si = pixel_info(P, lpix);
            rng_load(g, A, P, lpix, si.pix, sample, 0);
            start_sample(S, P, A, po, gid, si, sample, g);
          } else {
            write_hole(A, po, gid);
          }
        }
      }
    } else {
      if (need) {
        if (sample + 1 < P.sample_base + P.sample_count) {
          const SlotInfo si = pixel_info(P, lpix);
          start_sample(S, P, A, po, gid, si, sample + 1, g);  // the pixel's stream carries on
        } else {
          write_hole(A, po, gid);
        }
      }
      if (live) rng_store(g, A, P, lpix);
    }
    idx = idx_next; qi = qi_next;
  }
}

// The tail of a job (Philox mode). Once the work counter has run dry the waves shrink bounce after bounce, and a wave of a
// few thousand rays costs what its slowest ray costs (one warp, ~25 us of dependent loads per closest-hit query) plus two
// launches - the Book-1 scene at 400x225x10 spent 45 of its 56 waves like that. Here every lane KEEPS its path: trace
// (one-ray-per-lane traversal) and shade in a loop until the path ends. Same events, same Philox keys (pixel, sample,
// bounce), fixed-point sums: the image does not change by a bit, whichever wave the hand-over happens at.
// spread_log2: one path per 2^spread_log2 lanes. A warp's bounce takes as long as its SLOWEST lane's closest-hit query, and at
// the end of a job latency is all that is left, so the host spreads the paths over as many warps as the GPU holds at once.
__global__ void __launch_bounds__(128) k_finish(DScene S, RenderParams P, PathArrays A, WaveCounters* C, int parity, int spread_log2) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int n = C->order_len;
  if (((tid - lane) >> spread_log2) >= n) return;
  const int gid = (tid & ((1 << spread_log2) - 1)) == 0 ? (tid >> spread_log2) : n;  // position of the dense layout, or none
  bool active = false;
  Ray r; r.o = v3(0, 0, 0); r.d = v3(1, 1, 1); r.tm = 0.f;
  V3 thr = v3(1, 1, 1);
  int bounce = 0, sample = 0, lpix = 0;
  if (gid < n) {
    const float4 d = A.ray_d[parity][gid];
    if (__float_as_int(d.w) >= 0) {
      const float4 o = A.ray_o[parity][gid], t4 = A.thr[parity][gid];
      r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
      thr = v3(t4.x, t4.y, t4.z);
      bounce = __float_as_int(t4.w) & 255; sample = __float_as_int(t4.w) >> 8;
      lpix = __float_as_int(d.w);
      active = true;
    }
  }
  unsigned long long rays = 0;
  while (__ballot_sync(0xFFFFFFFFu, active)) {
    const Hit h = closest_hit(S, r, active, P.tmin, FLT_MAX, &C->overflow);
    if (active) {
      ++rays;
      const int q = h.tlp < 0 ? (int)Q_MISS : tlp_class(h.tlp);
      const float2 hh = make_float2(h.t, __int_as_float(h.tlp));
      const SlotInfo si = pixel_info(P, lpix);
      Philox g;
      rng_load(g, A, P, lpix, si.pix, sample, bounce + 1);
      if (shade_event<RNG_PHILOX>(S, P, A, C, q, hh, r, thr, bounce, lpix, sample, g)) {
        // path regeneration, lane by lane (the counter is dry or nearly so by now)
        const unsigned long long w = atomicAdd(A.next_work, 1ull);
        if (w < (unsigned long long)P.work_total) {
          work_to_pixel_sample(P, w, lpix, sample);
          const SlotInfo s2 = pixel_info(P, lpix);
          rng_load(g, A, P, lpix, s2.pix, sample, 0);
          const float u = fdiv(fadd((float)s2.i, g.uniform()), (float)P.nx);
          const float v = fdiv(fadd((float)s2.j, g.uniform()), (float)P.ny);
          r = camera_get_ray(S.cam, u, v, g);
          thr = v3(1.f, 1.f, 1.f); bounce = 0;
        } else {
          active = false;
        }
      }
    }
  }
  rays = __reduce_add_sync(0xFFFFFFFFu, (unsigned)rays);
  if (lane == 0 && rays) atomicAdd(&C->rays, rays);
  // (the wave counters are left as they are: rt_render uploads fresh ones before the next job, and a reset here would race
  //  with blocks of this grid that have not read order_len yet)
}

RT_D float apply_gamma(float c, float gamma) {  // main.cu:37-42
  if (gamma == 1.0f) return c;
  const float inv = frcp(gamma);
  return powf(fmaxf(c, 0.0f), inv);
}

// accum: linear sum over this rank's samples per LOCAL pixel (float3); the unit the NCCL reduce sums.
// add != 0: progressive pass, accum += this pass (rt_render_params.accumulate).
template <int MODE>
__global__ void k_accumulate(RenderParams P, PathArrays A, float* accum, int add) {
  const int lpix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npl = P.rows_local * P.nx;
  if (lpix >= npl) return;
  float x, y, z;
  if constexpr (MODE == RNG_REFERENCE) {
    const float4 c = A.col[lpix];
    x = c.x; y = c.y; z = c.z;
  } else {
    const double k = 1.0 / 4294967296.0;
    x = (float)((double)(long long)A.acc64[3 * (size_t)lpix + 0] * k);
    y = (float)((double)(long long)A.acc64[3 * (size_t)lpix + 1] * k);
    z = (float)((double)(long long)A.acc64[3 * (size_t)lpix + 2] * k);
  }
  if (add) { x = fadd(accum[3 * lpix + 0], x); y = fadd(accum[3 * lpix + 1], y); z = fadd(accum[3 * lpix + 2], z); }
  accum[3 * lpix + 0] = x; accum[3 * lpix + 1] = y; accum[3 * lpix + 2] = z;
}

// ---- adaptive sampling ----
// Error estimate of a TILE of tile x tile local pixels from two half-buffers (all samples / odd sample numbers):
// mean over the tile's pixels of |I_even - I_odd|_1 / sqrt(|I_all|_1 + 1e-3)  (Dammertz et al. 2010).
__global__ void k_tile_error(int nx, int rows, int tile, int tiles_x, const unsigned long long* acc, const unsigned long long* acc_odd,
                             const int* n_even, const int* n_odd, float* err) {
  const int t = blockIdx.x, tx = t % tiles_x, ty = t / tiles_x;
  const double k = 1.0 / 4294967296.0;
  const int ne = n_even[t], no = n_odd[t];
  float sum = 0.f; int cnt = 0;
  for (int q = threadIdx.x; q < tile * tile; q += blockDim.x) {
    const int i = tx * tile + q % tile, j = ty * tile + q / tile;
    if (i >= nx || j >= rows) continue;
    const size_t p = 3 * ((size_t)j * nx + i);
    float d = 0.f, a = 0.f;
    for (int c = 0; c < 3; ++c) {
      const double all = (double)(long long)acc[p + c] * k, odd = (double)(long long)acc_odd[p + c] * k;
      const double io = no > 0 ? odd / no : 0.0, ie = ne > 0 ? (all - odd) / ne : 0.0;
      d += (float)fabs(ie - io);
      a += (float)fmax(all / (double)max(ne + no, 1), 0.0);
    }
    sum += d / sqrtf(a + 1e-3f);
    ++cnt;
  }
  __shared__ float s_sum[256]; __shared__ int s_cnt[256];
  s_sum[threadIdx.x] = sum; s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) { s_sum[threadIdx.x] += s_sum[threadIdx.x + w]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) err[t] = s_cnt[0] > 0 ? s_sum[0] / (float)s_cnt[0] : 0.f;
}
// accum (linear SUM over a pixel's samples) -> per-pixel MEAN: every tile has its own sample count
__global__ void k_normalize_tiles(int nx, int rows, int tile, int tiles_x, const int* n_even, const int* n_odd, float* accum) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nx * rows) return;
  const int i = p % nx, j = p / nx, t = (j / tile) * tiles_x + i / tile;
  const float k = frcp((float)max(n_even[t] + n_odd[t], 1));
  accum[3 * p + 0] = fmul(accum[3 * p + 0], k); accum[3 * p + 1] = fmul(accum[3 * p + 1], k); accum[3 * p + 2] = fmul(accum[3 * p + 2], k);
}

// fb = gamma(accum / ns) (main.cu:128-132). accum/fb are indexed by local pixel.
__global__ void k_resolve(int n_pix, int ns, float gamma, const float* accum, float* fb) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const float k = frcp((float)ns);  // col /= float(ns): k = 1.0/t, then three multiplies (vec3.cuh:145-153)
  fb[3 * p + 0] = apply_gamma(fmul(accum[3 * p + 0], k), gamma);
  fb[3 * p + 1] = apply_gamma(fmul(accum[3 * p + 1], k), gamma);
  fb[3 * p + 2] = apply_gamma(fmul(accum[3 * p + 2], k), gamma);
}

// Primary-hit AOV: centre ray of the pixel (no jitter, no lens offset, time0).
__global__ void k_aov(DScene S, RenderParams P, int* obj, int* mat, float* tout, WaveCounters* C) {
  const int lpix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npl = P.rows_local * P.nx;
  if ((lpix & ~31) >= npl) return;
  const bool active = lpix < npl;
  Ray r; r.o = v3(0, 0, 0); r.d = v3(1, 1, 1); r.tm = 0.f;
  if (active) {
    const int lr = lpix / P.nx, i = lpix - lr * P.nx, j = lr * P.world + P.rank;
    const float s = fdiv(fadd((float)i, 0.5f), (float)P.nx), t = fdiv(fadd((float)j, 0.5f), (float)P.ny);
    r.o = S.cam.origin;
    r.d = vsub(vmad(t, S.cam.vertical, vmad(s, S.cam.horizontal, S.cam.llc)), S.cam.origin);
    r.tm = (float)S.cam.time0;
  }
  const Hit h = closest_hit(S, r, active, P.tmin, FLT_MAX, &C->overflow);
  if (!active) return;
  obj[lpix] = h.tlp >= 0 ? tlp_index(h.tlp) : -1;
  mat[lpix] = h.tlp >= 0 ? S.tlp[tlp_index(h.tlp)].mat : -1;
  tout[lpix] = h.tlp >= 0 ? h.t : 0.f;
}

}  // namespace rt
