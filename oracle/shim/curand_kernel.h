// oracle/shim/curand_kernel.h — TEST INFRASTRUCTURE: host restatement of the cuRAND XORWOW device API
// the reference calls (curand_init / curand / curand_uniform). cuRAND 12.9's own versions are
// `static __forceinline__ __device__` (curand_kernel.h:60-62) and cannot be called on the host.
// Follows /usr/local/cuda/include/curand_kernel.h:772-874 (seed scramble; subsequence = offset = 0
// means no skip-ahead) and curand_uniform.h:69-72 (x * 2^-32 + 2^-33, in (0, 1]).
#pragma once
#include "cuda_runtime.h"
struct curandStateXORWOW {
  unsigned int d, v[5];
  int boxmuller_flag, boxmuller_flag_double;
  float boxmuller_extra;
  double boxmuller_extra_double;
};
typedef curandStateXORWOW curandState;
static inline void curand_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset,
                               curandState* s) {
  if (subsequence != 0 || offset != 0) { fprintf(stderr, "shim: skip-ahead not implemented\n"); abort(); }
  unsigned int s0 = ((unsigned int)seed) ^ 0xaad26b49u;
  unsigned int s1 = (unsigned int)(seed >> 32) ^ 0xf7dcefddu;
  unsigned int t0 = 1099087573u * s0;
  unsigned int t1 = 2591861531u * s1;
  s->d = 6615241u + t1 + t0;
  s->v[0] = 123456789u + t0;
  s->v[1] = 362436069u ^ t0;
  s->v[2] = 521288629u + t1;
  s->v[3] = 88675123u ^ t1;
  s->v[4] = 5783321u + t0;
  s->boxmuller_flag = s->boxmuller_flag_double = 0;
  s->boxmuller_extra = 0.f;
  s->boxmuller_extra_double = 0.;
}
static inline unsigned int curand(curandState* s) {
  unsigned int t = (s->v[0] ^ (s->v[0] >> 2));
  s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
  s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
  s->d += 362437u;
  return s->v[4] + s->d;
}
static inline float curand_uniform(curandState* s) {
  return curand(s) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
