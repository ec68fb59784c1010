// rng.h — the two random-number generators of the render path, host + device.
//
//  * Xorwow: cuRAND's XORWOW exactly as the reference uses it — curand_init(seed, 0, 0) (no
//    skip-ahead), curand(), curand_uniform() — restated from CUDA 12.9 curand_kernel.h:772-874 and
//    curand_uniform.h:69-72. "Reference-RNG mode" seeds one per pixel with 1984 + pixel_index
//    (main.cu:104) and one per scene with 1984 (main.cu:92). Only the 6 live words are kept
//    (the reference stores a 48-byte curandState per pixel; the Box-Muller fields are never used).
//  * Philox4x32-10 (Salmon et al., SC'11; same round/key constants as cuRAND's Philox and
//    Random123): counter-based, so a sample's stream is a pure function of (seed, pixel, sample,
//    bounce, draw index) and needs no per-pixel state in HBM.
//
// Uniform mapping is cuRAND's for both: x * 2^-32 + 2^-33, in (0, 1].
#pragma once
#include "rt_math.h"

namespace rt {

RT_HD float u32_to_uniform(uint32_t x) {
  // curand_uniform.h:71: x * CURAND_2POW32_INV + (CURAND_2POW32_INV/2.0f), fused in the reference SASS;
  // the multiplier is a power of two, so fused and unfused agree.
  return ffma((float)x, 2.3283064e-10f, 1.1641532e-10f);
}

struct Xorwow {
  uint32_t d, v0, v1, v2, v3, v4;

  RT_HD void init(unsigned long long seed) {  // curand_init(seed, 0, 0, &s)
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    d = 6615241u + t1 + t0;
    v0 = 123456789u + t0;
    v1 = 362436069u ^ t0;
    v2 = 521288629u + t1;
    v3 = 88675123u ^ t1;
    v4 = 5783321u + t0;
  }
  RT_HD uint32_t next() {  // curand()
    uint32_t t = (v0 ^ (v0 >> 2));
    v0 = v1; v1 = v2; v2 = v3; v3 = v4;
    v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    return v4 + d;
  }
  RT_HD float uniform() { return u32_to_uniform(next()); }  // curand_uniform()
};

// Philox4x32-10
struct Philox {
  uint32_t key0, key1;      // seed
  uint32_t c0, c1, c2, c3;  // counter: (draw block, pixel, sample, bounce/stage)
  uint32_t out[4];
  int have;

  RT_HD static void round_(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#if defined(__CUDA_ARCH__)
    uint32_t hi0 = __umulhi(M0, x0), hi1 = __umulhi(M1, x2);
#else
    uint32_t hi0 = (uint32_t)(((unsigned long long)M0 * x0) >> 32), hi1 = (uint32_t)(((unsigned long long)M1 * x2) >> 32);
#endif
    uint32_t lo0 = M0 * x0, lo1 = M1 * x2;
    uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
  }
  RT_HD static void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* o) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round_(c0, c1, c2, c3, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
  }
  RT_HD void init(unsigned long long seed, uint32_t pixel, uint32_t sample, uint32_t stage) {
    key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
    c0 = 0; c1 = pixel; c2 = sample; c3 = stage;
    have = 0;
  }
  RT_HD uint32_t next() {
    if (have == 0) { block(c0, c1, c2, c3, key0, key1, out); ++c0; have = 4; }
    uint32_t r = out[4 - have];
    --have;
    return r;
  }
  RT_HD float uniform() { return u32_to_uniform(next()); }
  // One whole block = four fresh 32-bit words (drops whatever was left of the previous block).
  RT_HD void next4(uint32_t* o) { block(c0, c1, c2, c3, key0, key1, o); ++c0; have = 0; }
};

}  // namespace rt
