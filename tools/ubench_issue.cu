// Does an FFMA2 (64 FMAs on a 32-lane pipe: two pipe cycles) also hold the ISSUE port for two cycles?
// Kernel 0: FFMA2 only (8 independent chains); kernel 1: one independent LOP3/IADD3 (ALU pipe) per FFMA2;
// kernel 2: two per FFMA2; kernel 3: one LDS.64 per 4 FFMA2 + one ALU op per FFMA2.  1 or 2 warps per scheduler.
// If the ALU instructions ride in the second cycle of the FFMA2s, kernel 1 costs what kernel 0 costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue tools/ubench_issue.cu && ./ubench_issue
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct __align__(8) pf2 { float x, y; };
__device__ __forceinline__ pf2 fma2(pf2 a, pf2 b, pf2 c) {
  pf2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed, uint32_t s0) {
  __shared__ float2 sm[256];
  sm[threadIdx.x] = make_float2(seed, seed * 2.f);
  __syncthreads();
  pf2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = pf2{seed + i, seed - i};
  const pf2 x{0.999f, 0.998f}, y{1e-3f, 2e-3f};
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) u[i] = s0 + i * 77u + threadIdx.x;
  float2 l = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = fma2(a[i], x, y);
        if (MODE >= 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(s0));
        if (MODE == 2) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(u[(i + 3) & 7]));
      }
      if (MODE == 3) {
        float2 t;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + r) & 255])));
        l.x += t.x;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + r + 9) & 255])));
        l.y += t.y;
      }
    }
  }
  float s = l.x + l.y;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + (float)u[i];
  if (s == 123.456f) out[0] = s;
}

int main() {
  float* d;
  cudaMalloc(&d, 64);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int threads : {128, 256}) {
    for (int mode = 0; mode < 4; ++mode) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, threads>>>(d, iters, 0.5f, 12345u);
        if (mode == 1) k<1><<<sms, threads>>>(d, iters, 0.5f, 12345u);
        if (mode == 2) k<2><<<sms, threads>>>(d, iters, 0.5f, 12345u);
        if (mode == 3) k<3><<<sms, threads>>>(d, iters, 0.5f, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double ffma2_per_warp = 64.0 * iters;
      const double cycles = best * 1e-3 * khz * 1e3;
      printf("{\"warps_per_scheduler\": %d, \"mode\": %d, \"ms\": %.4f, \"cycles_per_ffma2_per_scheduler\": %.3f}\n", threads / 128, mode,
             best, cycles / (ffma2_per_warp * (threads / 128)));
    }
  }
  return 0;
}
