// oracle/shim/cuda_runtime.h — TEST INFRASTRUCTURE. Lets the reference's CUDA headers and the device
// part of main.cu compile with plain g++ so the reference's own arithmetic can run on host cores.
// Nothing here implements rendering: it only maps CUDA spellings to host C++.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <iostream>
#include <math.h>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline

// CUDA built-in coordinates: one "thread" per call, set by the harness before each kernel call.
struct rh_dim3 { unsigned x = 0, y = 0, z = 0; };
extern thread_local rh_dim3 threadIdx, blockIdx, blockDim;

// device intrinsics used by the reference
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
#define __sinf(x) sinf(x)  // texture.cuh:69 (fast-math intrinsic on the GPU; glibc already declares __sinf)
using std::min;
using std::max;

// image_io.h:24-46 uses the runtime API on the host
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1 };
static inline cudaError_t cudaMalloc(void* pp, size_t n) { *(void**)pp = malloc(n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaDeviceReset() { return 0; }
