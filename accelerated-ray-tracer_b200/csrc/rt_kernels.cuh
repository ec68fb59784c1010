// rt_kernels.cuh — the wavefront render kernels (device only).
//
// The reference renders with ONE megakernel: a thread per pixel loops ns samples x <= 50 bounces
// through virtual calls and device recursion (render/color, main.cu:44-133). Here the same
// integrator runs as waves over a pool of path slots kept in SoA arrays in HBM/L2:
//
//   k_init     slot state: "needs a sample"; reference-RNG mode: XORWOW seeding      (render_init, main.cu:96-105)
//   k_trace    one thread per SLOT (coalesced state access, no ray lists): a slot whose sample has ended
//              takes its next camera sample here (path regeneration, main.cu:119-123), then the closest
//              hit over the 4-wide BVH; paths are binned by the hit material class into shade queues
//              (per-block aggregation, one global atomic per class per block)        (main.cu:57; bvh.cuh:95)
//   k_shade    one material class per warp: miss/background, emission, scatter, throughput; a finished
//              sample is added to its pixel and the slot is flagged "needs a sample"  (main.cu:58-83, 124)
//   k_accumulate / k_resolve   sums -> linear accumulation buffer -> 1/ns, gamma     (main.cu:128-132)
//   k_aov      primary-hit object/material id + t for the centre ray of every pixel
//
// Philox mode (production): a slot is a worker. Work item w = sample * n_local_pixels + local_pixel is
// handed out by one 64-bit counter (one atomic per block, at the head of k_trace), so every slot stays
// busy until the whole job is done no matter how unevenly path lengths are spread over the image (on
// the Book-2 final scene a pixel-bound scheme ran at 11% mean slot occupancy). A finished sample is
// added to its pixel with 64-bit FIXED-POINT atomics (2^-32 resolution): integer sums do not depend on
// the order of arrival, so the image is bit-reproducible and identical under any tile split.
// Reference-RNG mode (validation): a slot IS a pixel and consumes that pixel's XORWOW stream in exactly
// the reference's order, samples one after the other, summed in float like `col += color(...)`.
#pragma once
#include "rt_shade.cuh"

namespace rt {

struct PathArrays {
  float4* ray_o;   // origin.xyz, time
  float4* ray_d;   // direction.xyz, local pixel (int bits)
  float2* hit;     // t, (box face << 28 | top-level object index) as int bits; -1 = miss
  float4* thr;     // throughput.rgb, bounce (int bits); bounce = SLOT_NEEDS_SAMPLE / SLOT_DEAD are slot states
  float4* rad;     // radiance.rgb of the current sample, sample number (int bits)
  float4* col;     // reference-RNG mode: float sum of the pixel's finished samples
  uint32_t* rng;   // reference-RNG mode: 6 words per slot, SoA [6][n_slots]
  unsigned long long* acc64;  // Philox mode: per local pixel 3 x fixed-point (2^-32) radiance sums
  int* order;      // k_trace thread -> slot: the previous wave's shade-queue layout (-1 = hole), see k_shade
  unsigned long long* next_work;  // Philox mode: next work item to hand out (shared by all slot pools)
};

struct WaveCounters {
  int n_queue[2][Q_COUNT + 1]; // shade queue fill, by wave parity (their sum = rays traced in that wave)
  int order_len;               // entries of PathArrays::order in use
  unsigned long long rays;     // closest-hit queries issued (= the reference's bounce-loop iterations)
  unsigned long long samples;  // finished samples
  unsigned int overflow;       // traversal stack overflow flag (must stay 0)
  unsigned int nonfinite;      // Philox mode: samples dropped because their radiance was inf/NaN
};

struct RenderParams {
  int nx, ny;            // full image
  float inv_nx;          // 1.0f / nx
  int rows_local;        // scanlines owned by this rank (tile split: j = lr * world + rank)
  int rank, world;
  int n_slots;           // path slots in flight (reference-RNG mode: one per local pixel)
  long long work_total;  // rows_local * nx * sample_count work items (Philox mode)
  int sample_base;       // first sample number of this rank's share (spp split), else 0
  int sample_count;      // samples per pixel this rank renders
  int max_depth;         // 50
  float tmin;            // 0.001f
  V3 background; int gradient;
  unsigned long long seed;  // 1984
};

enum RngMode : int { RNG_PHILOX = 0, RNG_REFERENCE = 1 };
enum SlotState : int { SLOT_NEEDS_SAMPLE = -1, SLOT_DEAD = -2 };

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

template <int MODE>
RT_D void rng_load(typename RngOf<MODE>::type& g, const PathArrays& A, const RenderParams& P, int slot, int pix, int sample, int stage);
template <>
RT_D void rng_load<RNG_PHILOX>(Philox& g, const PathArrays&, const RenderParams& P, int, int pix, int sample, int stage) {
  g.init(P.seed, (uint32_t)pix, (uint32_t)sample, (uint32_t)stage);
}
template <>
RT_D void rng_load<RNG_REFERENCE>(Xorwow& g, const PathArrays& A, const RenderParams& P, int slot, int, int, int) {
  const int n = P.n_slots;
  g.d = A.rng[slot]; g.v0 = A.rng[n + slot]; g.v1 = A.rng[2 * n + slot]; g.v2 = A.rng[3 * n + slot];
  g.v3 = A.rng[4 * n + slot]; g.v4 = A.rng[5 * n + slot];
}
RT_D void rng_store(const Philox&, const PathArrays&, const RenderParams&, int) {}
RT_D void rng_store(const Xorwow& g, const PathArrays& A, const RenderParams& P, int slot) {
  const int n = P.n_slots;
  A.rng[slot] = g.d; A.rng[n + slot] = g.v0; A.rng[2 * n + slot] = g.v1; A.rng[3 * n + slot] = g.v2;
  A.rng[4 * n + slot] = g.v3; A.rng[5 * n + slot] = g.v4;
}

// New camera sample for a slot (main.cu:121-123): jitter, lens, shutter time; throughput 1, radiance 0.
template <class RNG>
RT_D Ray start_sample(const DScene& S, const RenderParams& P, const PathArrays& A, int slot, const SlotInfo& si, int sample, RNG& g) {
  // ray_d.w carries the slot's local pixel, rad.w its sample number
  const float u = fdiv(fadd((float)si.i, g.uniform()), (float)P.nx);
  const float v = fdiv(fadd((float)si.j, g.uniform()), (float)P.ny);
  const Ray r = camera_get_ray(S.cam, u, v, g);
  A.ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, r.tm);
  A.ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(si.lpix));
  A.thr[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float(0));
  A.rad[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(sample));
  return r;
}

// Block-aggregated reservation: EVERY thread of the block calls it; threads with `pred` get consecutive
// positions starting at a base taken with ONE global atomic per block (a wave of 1 Mi rays appends to a single
// counter: one atomic per warp serialised 32 Ki same-address atomics in L2 and cost ~40% of k_shade, profiles/r01).
#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif
#ifndef RT_SHADE_MINB
#define RT_SHADE_MINB 3   // resident blocks per SM the shade kernel is compiled for (register cap 65536 / (256 * MINB))
#endif
#define RT_WARPS (RT_BLOCK / 32)
#ifndef RT_TBLOCK
#define RT_TBLOCK 256     // k_trace block size
#endif
#define RT_TWARPS (RT_TBLOCK / 32)
template <int NW, class T>
RT_D T block_reserve(T* counter, bool pred, int* s_cnt /*[NW]*/, T* s_base) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_cnt[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
    *s_base = tot > 0 ? atomicAdd(counter, (T)tot) : (T)0;
  }
  __syncthreads();
  const T pos = *s_base + (T)(s_cnt[warp] + __popc(m & ((1u << lane) - 1u)));
  __syncthreads();  // s_cnt / s_base may be reused by the next call
  return pos;
}

#define RT_FIXED_ONE 4294967296.0f  /* 2^32 */
RT_D void fixed_add(unsigned long long* acc, float v) {
  atomicAdd(acc, (unsigned long long)__float2ll_rn(v * RT_FIXED_ONE));  // two's complement: negative values add correctly
}
RT_D void work_to_pixel_sample(const RenderParams& P, unsigned long long w, int& lpix, int& sample) {
  if (P.work_total <= 0xFFFFFFFFll) {  // the usual case: 32-bit divide
    const unsigned npl = (unsigned)(P.rows_local * P.nx), w32 = (unsigned)w;
    const unsigned s = w32 / npl;
    lpix = (int)(w32 - s * npl);
    sample = P.sample_base + (int)s;
  } else {
    const unsigned long long npl = (unsigned long long)(P.rows_local * P.nx);
    const unsigned long long s = w / npl;
    lpix = (int)(w - s * npl);
    sample = P.sample_base + (int)s;
  }
}

// Slot state before the first wave: every slot needs a sample. Reference-RNG mode also seeds the pixel's stream
// (render_init, main.cu:104) and zeroes its float sum; rad.w = the sample number BEFORE the first one.
template <int MODE>
__global__ void __launch_bounds__(RT_BLOCK) k_init(RenderParams P, PathArrays A) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.n_slots) return;
  A.order[slot] = slot;  // first wave: identity (the host sets order_len = n_slots)
  A.thr[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float((int)SLOT_NEEDS_SAMPLE));
  A.rad[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(P.sample_base - 1));
  if constexpr (MODE == RNG_REFERENCE) {
    const SlotInfo si = pixel_info(P, slot);  // slot == local pixel
    A.col[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    Xorwow g;
    g.init((unsigned long long)(long long)(1984 + si.pix));
    rng_store(g, A, P, slot);
  }
}

#ifndef RT_TRACE_MINB
#define RT_TRACE_MINB 4     // resident k_trace blocks per SM the kernel is compiled for (64 registers)
#endif
template <int MODE>
__global__ void __launch_bounds__(RT_TBLOCK, RT_TRACE_MINB) k_trace(DScene S, RenderParams P, PathArrays A, int* __restrict__ queues,
                                                     WaveCounters* C, int parity) {
  __shared__ int s_cnt[RT_TWARPS];
  __shared__ unsigned long long s_wbase;
  __shared__ int s_q[RT_TWARPS][Q_COUNT];
  __shared__ int s_qbase[Q_COUNT];
  // Thread -> slot through the previous wave's queue layout: paths that hit the same material class sit next to
  // each other, and the primary rays regenerated behind the miss / light queues come out in pixel order, which
  // keeps warps far more coherent than slot order (measured: 0.32 ms vs 0.53 ms per 1 Mi-ray wave on C4).
  // No early exit: closest_hit is warp-wide, the binning block-wide.
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int slot = -1;
  if (gid < C->order_len) slot = A.order[gid];
  const bool in_range = slot >= 0;
  int state = SLOT_DEAD;
  if (in_range) state = __float_as_int(A.thr[slot].w);
  bool active = state >= 0;
  Ray r; r.o = v3(0, 0, 0); r.d = v3(1, 1, 1); r.tm = 0.f;
  // ---- path regeneration (main.cu:119-123): a slot whose sample has ended takes the next one ----
  const bool need = state == SLOT_NEEDS_SAMPLE;
  if constexpr (MODE == RNG_PHILOX) {
    const unsigned long long w = block_reserve<RT_TWARPS>(A.next_work, need, s_cnt, &s_wbase);
    if (need) {
      if (w < (unsigned long long)P.work_total) {
        int lpix, sample;
        work_to_pixel_sample(P, w, lpix, sample);
        const SlotInfo si = pixel_info(P, lpix);
        Philox g;
        rng_load<MODE>(g, A, P, slot, si.pix, sample, 0);
        r = start_sample(S, P, A, slot, si, sample, g);
        active = true;
      } else {
        A.thr[slot].w = __int_as_float((int)SLOT_DEAD);
      }
    }
  } else {
    if (need) {
      const int sample = __float_as_int(A.rad[slot].w) + 1;
      if (sample < P.sample_base + P.sample_count) {
        const SlotInfo si = pixel_info(P, slot);
        Xorwow g;
        rng_load<MODE>(g, A, P, slot, si.pix, sample, 0);
        r = start_sample(S, P, A, slot, si, sample, g);
        rng_store(g, A, P, slot);
        active = true;
      } else {
        A.thr[slot].w = __int_as_float((int)SLOT_DEAD);
      }
    }
  }
  if (active && !need) {
    const float4 o = A.ray_o[slot], d = A.ray_d[slot];
    r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
  }
  // ---- closest hit (main.cu:57) ----
  const Hit h = closest_hit(S, r, active, P.tmin, FLT_MAX, &C->overflow);
  int q = -1;
  if (active) {
    A.hit[slot] = make_float2(h.t, __int_as_float(h.tlp < 0 ? -1 : (h.tlp | (h.face << 28))));
    q = h.tlp < 0 ? (int)Q_MISS : S.tlp[h.tlp].queue;
  }
  // ---- bin by material class: per-warp counts per class in shared memory, ONE global atomic per class per block ----
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < Q_COUNT) s_q[warp][lane] = 0;
  __syncwarp();
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, q);
  if (q >= 0 && lane == __ffs(peers) - 1) s_q[warp][q] = __popc(peers);
  __syncthreads();
  if (threadIdx.x < Q_COUNT) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < RT_TWARPS; ++w) { const int c = s_q[w][threadIdx.x]; s_q[w][threadIdx.x] = tot; tot += c; }
    s_qbase[threadIdx.x] = tot > 0 ? atomicAdd(&C->n_queue[parity][threadIdx.x], tot) : 0;
  }
  __syncthreads();
  if (q >= 0) queues[(size_t)q * P.n_slots + s_qbase[q] + s_q[warp][q] + __popc(peers & ((1u << lane) - 1u))] = slot;
}

template <int MODE>
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINB) k_shade(DScene S, RenderParams P, PathArrays A, const int* __restrict__ queues,
                                                                   WaveCounters* C, int parity) {
  // thread -> (queue, position): queues are laid end to end, each padded to a whole warp
  int cnt[Q_COUNT];
  int total = 0, rays = 0;
#pragma unroll
  for (int k = 0; k < Q_COUNT; ++k) { cnt[k] = C->n_queue[parity][k]; total += (cnt[k] + 31) & ~31; rays += cnt[k]; }
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < Q_COUNT; ++k) C->n_queue[parity ^ 1][k] = 0;  // next wave's trace fills these
    C->rays += (unsigned long long)rays;
    C->order_len = total;  // the next k_trace walks this wave's queue layout
  }
  if (gid >= total) return;
  int q = 0, base = 0, qcount = cnt[0];
#pragma unroll
  for (int k = 0; k < Q_COUNT - 1; ++k) {
    const int padded = (cnt[k] + 31) & ~31;
    if (q == k && gid >= base + padded) { base += padded; q = k + 1; qcount = cnt[k + 1]; }
  }
  const int pos = gid - base;
  if (pos >= qcount) { A.order[gid] = -1; return; }  // warp padding between two queues
  const int slot = queues[(size_t)q * P.n_slots + pos];
  A.order[gid] = slot;
  const float4 o = A.ray_o[slot], d = A.ray_d[slot];
  const float4 thr4 = A.thr[slot], rad4 = A.rad[slot];
  const SlotInfo si = pixel_info(P, __float_as_int(d.w));
  Ray r; r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
  V3 thr = v3(thr4.x, thr4.y, thr4.z), rad = v3(rad4.x, rad4.y, rad4.z);
  int bounce = __float_as_int(thr4.w);
  const int sample = __float_as_int(rad4.w);
  typename RngOf<MODE>::type g;
  rng_load<MODE>(g, A, P, slot, si.pix, sample, bounce + 1);
  bool sample_done;
  Ray next; next.o = r.o; next.d = r.d; next.tm = r.tm;
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
    const float2 hh = A.hit[slot];
    const int packed = __float_as_int(hh.y);
    const int tlp = packed & 0x0FFFFFFF, face = packed >> 28;
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
    const bool scattered = material_scatter(S, m, r, rec, g, att, next);  // main.cu:76
    if (scattered) {
      thr = vmul(thr, att);  // main.cu:82
      ++bounce;
    }
    sample_done = !scattered || bounce >= P.max_depth;
  }
  if (sample_done) {
    if constexpr (MODE == RNG_REFERENCE) {
      // render: col += color(...) (main.cu:124)
      float4 c = A.col[slot];
      c.x = fadd(c.x, rad.x); c.y = fadd(c.y, rad.y); c.z = fadd(c.z, rad.z);
      A.col[slot] = c;
    } else {
      if (isfinite(rad.x) && isfinite(rad.y) && isfinite(rad.z)) {
        unsigned long long* acc = A.acc64 + 3 * (size_t)si.lpix;
        fixed_add(acc + 0, rad.x); fixed_add(acc + 1, rad.y); fixed_add(acc + 2, rad.z);
      } else {
        atomicAdd(&C->nonfinite, 1u);
      }
    }
    A.thr[slot].w = __int_as_float((int)SLOT_NEEDS_SAMPLE);  // k_trace of the next wave regenerates the slot
  } else {
    A.ray_o[slot] = make_float4(next.o.x, next.o.y, next.o.z, next.tm);
    A.ray_d[slot] = make_float4(next.d.x, next.d.y, next.d.z, d.w);
    A.thr[slot] = make_float4(thr.x, thr.y, thr.z, __int_as_float(bounce));
    A.rad[slot] = make_float4(rad.x, rad.y, rad.z, rad4.w);
  }
  rng_store(g, A, P, slot);
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
  obj[lpix] = h.tlp;
  mat[lpix] = h.tlp >= 0 ? S.tlp[h.tlp].mat : -1;
  tout[lpix] = h.tlp >= 0 ? h.t : 0.f;
}

}  // namespace rt
