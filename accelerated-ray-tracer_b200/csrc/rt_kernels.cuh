// rt_kernels.cuh — the wavefront render kernels (device only).
//
// The reference renders with ONE megakernel: a thread per pixel loops ns samples x <= 50 bounces
// through virtual calls and device recursion (render/color, main.cu:44-133). Here the same
// integrator runs as waves over a pool of path slots kept in SoA arrays in HBM/L2:
//
//   k_start    every slot draws its first camera ray                         (main.cu:119-123)
//   k_trace    closest hit of every live ray over the 4-wide BVH; paths are binned by the hit
//              material class into shade queues (warp-aggregated atomics)   (main.cu:57; bvh.cuh:95)
//   k_shade    one material class per warp: miss/background, emission, scatter, throughput;
//              finished samples are folded into the slot's sum and the slot draws its next
//              camera ray in place (path regeneration); survivors are compacted into the
//              next wave's ray list                                          (main.cu:58-83, 119-125)
//   k_resolve  per pixel: sum the slot partials, scale by 1/ns, gamma        (main.cu:128-132)
//   k_aov      primary-hit object/material id + t for the centre ray of every pixel
//
// Philox mode (production): a slot is a worker. Work item w = sample * n_local_pixels + local_pixel is
// handed out by one 64-bit counter (warp-aggregated), so every slot stays busy until the whole job is
// done no matter how unevenly path lengths are spread over the image (on the Book-2 final scene the
// pixel-bound scheme ran at 11% mean slot occupancy). A finished sample is added to its pixel with
// 64-bit FIXED-POINT atomics (2^-32 resolution): integer sums do not depend on the order of arrival,
// so the image is bit-reproducible and identical under any tile split.
// Reference-RNG mode (validation): a slot IS a pixel and consumes that pixel's XORWOW stream in exactly
// the reference's order, samples one after the other, summed in float like `col += color(...)`.
#pragma once
#include "rt_shade.cuh"

namespace rt {

struct PathArrays {
  float4* ray_o;   // origin.xyz, time
  float4* ray_d;   // direction.xyz, -
  float2* hit;     // t, (box face << 28 | top-level object index) as int bits; -1 = miss
  float4* thr;     // throughput.rgb, bounce (int bits)
  float4* rad;     // radiance.rgb of the current sample, sample number (int bits)
  float4* col;     // reference-RNG mode: float sum of the pixel's finished samples
  uint32_t* rng;   // reference-RNG mode: 6 words per slot, SoA [6][n_slots]
  unsigned long long* acc64;  // Philox mode: per local pixel 3 x fixed-point (2^-32) radiance sums
};

struct WaveCounters {
  int n_active[2];             // live rays in list[parity]
  int n_queue[2][Q_COUNT + 2]; // shade queue fill, by wave parity
  unsigned long long rays;     // closest-hit queries issued (= the reference's bounce-loop iterations)
  unsigned long long samples;  // finished samples
  unsigned int overflow;       // traversal stack overflow flag (must stay 0)
  unsigned int nonfinite;      // Philox mode: samples dropped because their radiance was inf/NaN
  unsigned long long next_work; // Philox mode: next work item to hand out
};

struct RenderParams {
  int nx, ny;            // full image
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

struct SlotInfo { int lpix, i, j, pix; };
RT_D SlotInfo pixel_info(const RenderParams& P, int lpix) {
  SlotInfo s;
  s.lpix = lpix;
  const int lr = s.lpix / P.nx;
  s.i = s.lpix - lr * P.nx;
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
RT_D void start_sample(const DScene& S, const RenderParams& P, const PathArrays& A, int slot, const SlotInfo& si, int sample, RNG& g) {
  // ray_d.w carries the slot's local pixel, rad.w its sample number
  const float u = fdiv(fadd((float)si.i, g.uniform()), (float)P.nx);
  const float v = fdiv(fadd((float)si.j, g.uniform()), (float)P.ny);
  Ray r = camera_get_ray(S.cam, u, v, g);
  A.ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, r.tm);
  A.ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(si.lpix));
  A.thr[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float(0));
  A.rad[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(sample));
}

// Warp-aggregated append: lanes with `pred` get consecutive positions in a global list.
RT_D int warp_append(int* counter, bool pred) {
  const unsigned m = __ballot_sync(__activemask(), pred);
  if (!pred) return -1;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

// Warp-aggregated grab of consecutive work items: lanes with `pred` get w, w+1, ... in lane order.
RT_D unsigned long long warp_grab(unsigned long long* counter, bool pred) {
  const unsigned m = __ballot_sync(__activemask(), pred);
  if (!pred) return ~0ull;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}
#define RT_FIXED_ONE 4294967296.0f  /* 2^32 */
RT_D void fixed_add(unsigned long long* acc, float v) {
  atomicAdd(acc, (unsigned long long)__float2ll_rn(v * RT_FIXED_ONE));  // two's complement: negative values add correctly
}
RT_D void work_to_pixel_sample(const RenderParams& P, unsigned long long w, int& lpix, int& sample) {
  const unsigned long long npl = (unsigned long long)(P.rows_local * P.nx);
  const unsigned long long s = w / npl;
  lpix = (int)(w - s * npl);
  sample = P.sample_base + (int)s;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_start(DScene S, RenderParams P, PathArrays A, int* list0, WaveCounters* C) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.n_slots) return;
  typename RngOf<MODE>::type g;
  bool live;
  if constexpr (MODE == RNG_REFERENCE) {
    const SlotInfo si = pixel_info(P, slot);  // slot == local pixel
    A.col[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    g.init((unsigned long long)(long long)(1984 + si.pix));  // render_init, main.cu:104
    live = P.sample_count > 0;
    if (live) start_sample(S, P, A, slot, si, P.sample_base, g);
    rng_store(g, A, P, slot);
  } else {
    live = (long long)slot < P.work_total;  // the first n_slots work items; the counter starts behind them
    if (live) {
      int lpix, sample;
      work_to_pixel_sample(P, (unsigned long long)slot, lpix, sample);
      const SlotInfo si = pixel_info(P, lpix);
      rng_load<MODE>(g, A, P, slot, si.pix, sample, 0);
      start_sample(S, P, A, slot, si, sample, g);
    }
  }
  const int pos = warp_append(&C->n_active[0], live);
  if (live) list0[pos] = slot;
}

__global__ void __launch_bounds__(128) k_trace(DScene S, RenderParams P, PathArrays A, const int* __restrict__ list,
                                               int* __restrict__ queues, WaveCounters* C, int parity) {
  const int n = C->n_active[parity];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->n_active[parity ^ 1] = 0;  // the shade kernel of this wave appends here
    C->rays += (unsigned long long)n;
  }
  if ((gid & ~31) >= n) return;  // whole warps only: closest_hit is a warp-wide routine
  const bool active = gid < n;
  const int slot = active ? list[gid] : 0;
  Ray r; r.o = v3(0, 0, 0); r.d = v3(1, 1, 1); r.tm = 0.f;
  if (active) {
    const float4 o = A.ray_o[slot], d = A.ray_d[slot];
    r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
  }
  const Hit h = closest_hit(S, r, active, P.tmin, FLT_MAX, &C->overflow);
  int q = -1;
  if (active) {
    A.hit[slot] = make_float2(h.t, __int_as_float(h.tlp < 0 ? -1 : (h.tlp | (h.face << 28))));
    q = h.tlp < 0 ? (int)Q_MISS : S.tlp[h.tlp].queue;
  }
  // bin by material class: one warp-aggregated atomic per class present in the warp
  const unsigned peers = __match_any_sync(__activemask(), q);
  if (active) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&C->n_queue[parity][q], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    queues[(size_t)q * P.n_slots + base + __popc(peers & ((1u << lane) - 1u))] = slot;
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_shade(DScene S, RenderParams P, PathArrays A, const int* __restrict__ queues,
                                               int* __restrict__ next_list, WaveCounters* C, int parity) {
  // thread -> (queue, position): queues are laid end to end, each padded to a whole warp
  int cnt[Q_COUNT];
  int total = 0;
#pragma unroll
  for (int k = 0; k < Q_COUNT; ++k) { cnt[k] = C->n_queue[parity][k]; total += (cnt[k] + 31) & ~31; }
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < Q_COUNT; ++k) C->n_queue[parity ^ 1][k] = 0;  // next wave's trace fills these
  }
  if (gid >= total) return;
  int q = 0, base = 0, qcount = cnt[0];
#pragma unroll
  for (int k = 0; k < Q_COUNT - 1; ++k) {
    const int padded = (cnt[k] + 31) & ~31;
    if (q == k && gid >= base + padded) { base += padded; q = k + 1; qcount = cnt[k + 1]; }
  }
  const int pos = gid - base;
  const bool valid = pos < qcount;
  bool alive = false, want_work = false;
  int slot = -1;
  if (valid) {
    slot = queues[(size_t)q * P.n_slots + pos];
    const float4 o = A.ray_o[slot], d = A.ray_d[slot];
    SlotInfo si = pixel_info(P, __float_as_int(d.w));
    float4 thr4 = A.thr[slot], rad4 = A.rad[slot];
    Ray r; r.o = v3(o.x, o.y, o.z); r.d = v3(d.x, d.y, d.z); r.tm = o.w;
    V3 thr = v3(thr4.x, thr4.y, thr4.z), rad = v3(rad4.x, rad4.y, rad4.z);
    int bounce = __float_as_int(thr4.w);
    int sample = __float_as_int(rad4.w);
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
    if constexpr (MODE == RNG_REFERENCE) {
      if (sample_done) {
        // render: col += color(...) (main.cu:124), then the next sample of this pixel, if any
        float4 c = A.col[slot];
        c.x = fadd(c.x, rad.x); c.y = fadd(c.y, rad.y); c.z = fadd(c.z, rad.z);
        A.col[slot] = c;
        ++sample;
        if (sample < P.sample_base + P.sample_count) {
          start_sample(S, P, A, slot, si, sample, g);
          alive = true;
        }
      }
    } else {
      if (sample_done) {
        if (isfinite(rad.x) && isfinite(rad.y) && isfinite(rad.z)) {
          unsigned long long* acc = A.acc64 + 3 * (size_t)si.lpix;
          fixed_add(acc + 0, rad.x); fixed_add(acc + 1, rad.y); fixed_add(acc + 2, rad.z);
        } else {
          atomicAdd(&C->nonfinite, 1u);
        }
      }
    }
    if (!sample_done) {
      A.ray_o[slot] = make_float4(next.o.x, next.o.y, next.o.z, next.tm);
      A.ray_d[slot] = make_float4(next.d.x, next.d.y, next.d.z, d.w);
      A.thr[slot] = make_float4(thr.x, thr.y, thr.z, __int_as_float(bounce));
      A.rad[slot] = make_float4(rad.x, rad.y, rad.z, __int_as_float(sample));
      alive = true;
    }
    rng_store(g, A, P, slot);
    if constexpr (MODE == RNG_PHILOX) want_work = sample_done;
  }
  if constexpr (MODE == RNG_PHILOX) {
    // path regeneration: finished lanes take the next work items (consecutive pixels of one sample number)
    const unsigned long long w = warp_grab(&C->next_work, want_work);
    if (want_work && w < (unsigned long long)P.work_total) {
      int lpix, sample;
      work_to_pixel_sample(P, w, lpix, sample);
      const SlotInfo si = pixel_info(P, lpix);
      Philox g;
      rng_load<MODE>(g, A, P, slot, si.pix, sample, 0);
      start_sample(S, P, A, slot, si, sample, g);
      alive = true;
    }
  }
  const int np = warp_append(&C->n_active[parity ^ 1], alive);
  if (alive) next_list[np] = slot;
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
