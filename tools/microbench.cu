// tools/microbench.cu — measured denominators for the roofline of the render kernels (the FP32-FMA and L2
// peaks "must be microbenchmarked"; MEASURED_PEAKS.json only has HBM copy bandwidth and bf16 tensor throughput).
//
//   fp32_fma_tflops      8 independent FFMA chains per thread, all SMs, 2 flop per lane-FMA
//   issue_ginst_s        warp-instructions issued per second: independent FFMA (fma pipe) interleaved with LOP3/IADD3 (alu pipe)
//   l2_read_gbs          float4 grid-stride reads of a 48 MiB buffer (resident in the 126 MB L2 after the first pass)
//   l1_read_gbs          float4 reads of a 64 KiB window per block (L1-resident)
//   atomic_same_addr_gs  returning atomicAdd on ONE address from every warp (what a shared queue counter costs)
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rt_microbench microbench.cu ; prints one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// 4 FFMA + 4 integer ops per group, all independent chains: one instruction per scheduler per clock if the issue port is the limit
__global__ void __launch_bounds__(256) k_issue(float* out, int iters, float a, float b, unsigned m) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  unsigned y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, a, b); y0 = (y0 ^ m) + 0x9E3779B9u;
      x1 = fmaf(x1, a, b); y1 = (y1 ^ m) + 0x9E3779B9u;
      x2 = fmaf(x2, a, b); y2 = (y2 ^ m) + 0x9E3779B9u;
      x3 = fmaf(x3, a, b); y3 = (y3 ^ m) + 0x9E3779B9u;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + (float)(y0 + y1 + y2 + y3);
}

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ buf, size_t n, int passes, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float4 v = buf[i];
      acc += v.x + v.y + v.z + v.w;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_read_l1(const float4* __restrict__ buf, int window, int passes, float* out) {
  float acc = 0.f;
  const float4* w = buf + (size_t)blockIdx.x * window;
  for (int p = 0; p < passes; ++p)
    for (int i = threadIdx.x; i < window; i += blockDim.x) {
      const float4 v = w[i];
      acc += v.x + v.y + v.z + v.w;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_atomic(unsigned long long* ctr, int iters, unsigned long long* out) {
  unsigned long long s = 0;
  for (int i = 0; i < iters; ++i)
    if ((threadIdx.x & 31) == 0) s += atomicAdd(ctr, 32ull);
  if ((threadIdx.x & 31) == 0) out[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = s;
}

template <class F> static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  float* out; CK(cudaMalloc(&out, (size_t)sms * 64 * 256 * sizeof(float)));
  // FP32 FMA
  const int grid = sms * 8, iters = 4096;
  const float ms_fma = time_ms([&] { k_fma<<<grid, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
  const double fma_flops = (double)grid * 256 * iters * 16 * 8 * 2;
  // issue
  const float ms_issue = time_ms([&] { k_issue<<<grid, 256>>>(out, iters, 1.0001f, 0.5f, 0x5bd1e995u); }, 5);
  const double issue_insts = (double)grid * 8 /*warps*/ * iters * 16 * 12;  // per group: 4 FFMA + 4 LOP3 + 4 IADD (xor, add do not fuse)
  // L2 read bandwidth
  const size_t bytes = (size_t)48 << 20, n4 = bytes / 16;
  float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
  const int passes = 20;
  const float ms_l2 = time_ms([&] { k_read<<<sms * 8, 256>>>(buf, n4, passes, out); }, 5);
  // L1 read bandwidth: 64 KiB window per block, 2 blocks per SM
  const int window = 4096;  // float4s = 64 KiB
  const int l1_blocks = sms * 2, l1_passes = 400;
  const float ms_l1 = time_ms([&] { k_read_l1<<<l1_blocks, 256>>>(buf, window, l1_passes, out); }, 5);
  // HBM read (buffer far larger than L2)
  const size_t big = (size_t)2 << 30;
  float4* bbuf = nullptr;
  float ms_hbm = 0.f;
  if (cudaMalloc(&bbuf, big) == cudaSuccess) {
    CK(cudaMemset(bbuf, 0, big));
    ms_hbm = time_ms([&] { k_read<<<sms * 16, 256>>>(bbuf, big / 16, 1, out); }, 5);
    cudaFree(bbuf);
  } else cudaGetLastError();
  // same-address returning atomics, one per warp
  unsigned long long *ctr, *aout;
  CK(cudaMalloc(&ctr, 8)); CK(cudaMemset(ctr, 0, 8)); CK(cudaMalloc(&aout, (size_t)sms * 8 * 8 * 8));
  const int a_iters = 256;
  const float ms_at = time_ms([&] { k_atomic<<<sms * 8, 256>>>(ctr, a_iters, aout); }, 5);
  const double n_atomics = (double)sms * 8 * 8 * a_iters;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f, \"fp32_fma_tflops\": %.2f, \"fp32_fma_tflops_datasheet\": %.2f, "
         "\"issue_ginst_s\": %.1f, \"issue_ginst_s_datasheet\": %.1f, \"l2_read_gbs\": %.1f, \"l1_read_gbs\": %.1f, \"hbm_read_gbs\": %.1f, "
         "\"atomic_same_addr_gs\": %.3f, \"how\": \"tools/microbench.cu: best of 5, CUDA events\"}\n",
         prop.name, sms, clock_khz / 1e3, fma_flops / ms_fma / 1e9, sms * 128 * 2 * (clock_khz / 1e6) / 1e3,
         issue_insts / ms_issue / 1e6, sms * 4 * (clock_khz / 1e6), (double)bytes * passes / ms_l2 / 1e6,
         (double)l1_blocks * window * 16.0 * l1_passes / ms_l1 / 1e6, ms_hbm > 0 ? (double)big / ms_hbm / 1e6 : 0.0,
         n_atomics / ms_at / 1e6);
  return 0;
}
