// Microbenchmark of candidate inner loops for the drift contraction acc[b][j] += x[b][k] * Qs[k][j]
// at the headline shape (per SM: 28 trajectories x 2 quadratures x 72 columns, 72 k, one barrier per step).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_drift tools/ubench_drift.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef unsigned long long u64;
struct __align__(8) pf2 { float x, y; };
__device__ __forceinline__ pf2 fma2(pf2 a, pf2 b, pf2 c){ pf2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)), "l"(reinterpret_cast<u64&>(c))); return d; }
constexpr int NK = 72, NC = 72;
__constant__ float2 cqd[NK * NC];   // duplicated (q,q): 41 KB
__constant__ float cq[NK * NC];     // natural: 20 KB

// V0: the round-1a kernel's loop: thread tile (4 traj x 2 quad) x 4 cols, LDS.128 x4 + 16 FFMA2 per k
__global__ void __launch_bounds__(128, 1) v0(float* out, int steps, const float* qg){
  extern __shared__ __align__(16) float sm[];
  float4* Qd = (float4*)sm;                 // [72][2][18]
  float* X = sm + 72 * 2 * 18 * 4;          // [2][72][96]
  const int tid = threadIdx.x, rg = tid % 7, cg = tid / 7; const bool act = cg < 18; const int cgc = act ? cg : 0;
  for (int i = tid; i < 72 * 36; i += 128) Qd[i] = make_float4(qg[i % 977], qg[i % 977], qg[(i + 1) % 977], qg[(i + 1) % 977]);
  for (int i = tid; i < 2 * 72 * 96; i += 128) X[i] = 0.001f * (i % 13);
  __syncthreads();
  pf2 st[2][2][4];
  for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) st[q][p][j] = {0.01f * tid, 0.02f};
  for (int t = 0; t < steps; ++t) {
    const int buf = t & 1;
    pf2 acc[2][2][4];
    for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) acc[q][p][j] = {0.f, 0.f};
    const float4* qp = Qd + cgc; const float* xp = X + buf * 72 * 96 + rg * 4;
    for (int kc = 0; kc < 18; ++kc) {
      const float* xrow = xp + ((kc * 28) & 31);
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 qa = qp[0], qb = qp[18];
        pf2 qd[4] = {{qa.x, qa.y}, {qa.z, qa.w}, {qb.x, qb.y}, {qb.z, qb.w}};
        pf2 xv[2][2];
        #pragma unroll
        for (int q = 0; q < 2; q++) { float4 x4 = *(const float4*)(xrow + q * 28); xv[q][0] = {x4.x, x4.y}; xv[q][1] = {x4.z, x4.w}; }
        #pragma unroll
        for (int q = 0; q < 2; q++)
          #pragma unroll
          for (int p = 0; p < 2; p++)
            #pragma unroll
            for (int j = 0; j < 4; j++) acc[q][p][j] = fma2(xv[q][p], qd[j], acc[q][p][j]);
        qp += 36; xrow += 96;
      }
      xp += 4 * 96;
    }
    if (act) {
      #pragma unroll
      for (int j = 0; j < 4; j++)
        #pragma unroll
        for (int q = 0; q < 2; q++) {
          st[q][0][j] = fma2(acc[q][0][j], {1e-3f, 1e-3f}, st[q][0][j]); st[q][1][j] = fma2(acc[q][1][j], {1e-3f, 1e-3f}, st[q][1][j]);
          float* dst = X + ((buf ^ 1) * 72 + 4 * cgc + j) * 96 + ((cgc * 28) & 31) + q * 28 + rg * 4;
          *(float4*)dst = make_float4(st[q][0][j].x, st[q][0][j].y, st[q][1][j].x, st[q][1][j].y);
        }
    }
    __syncthreads();
  }
  float s = 0; for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) s += st[q][p][j].x + st[q][p][j].y;
  out[blockIdx.x * 128 + tid] = s;
}

// V1..V3: lanes = trajectories (pair = (c,s) of one trajectory), warp w owns columns [18w, 18w+18),
// Q operand is warp-uniform and comes from the constant bank; X is one LDS.64 per k.
template <int MODE>
__global__ void __launch_bounds__(128, 1) vu(float* out, int steps){
  extern __shared__ __align__(16) float sm[];
  float2* X = (float2*)sm;                  // [2][72][32] (c,s) pairs
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * 72 * 32; i += 128) X[i] = make_float2(0.001f * (i % 13), 0.002f);
  __syncthreads();
  pf2 st[18];
  for (int j = 0; j < 18; j++) st[j] = {0.01f * lane, 0.02f};
  const int col0 = w * 18;
  for (int t = 0; t < steps; ++t) {
    const int buf = t & 1;
    const float2* xp = X + buf * 72 * 32 + lane;
    if (MODE == 2) {
      float a0[18], a1[18];
      #pragma unroll
      for (int j = 0; j < 18; j++) a0[j] = a1[j] = 0.f;
      #pragma unroll 2
      for (int k = 0; k < NK; ++k) {
        const float2 xv = xp[k * 32];
        #pragma unroll
        for (int j = 0; j < 18; j++) { const float qq = cq[k * NC + col0 + j]; a0[j] = fmaf(xv.x, qq, a0[j]); a1[j] = fmaf(xv.y, qq, a1[j]); }
      }
      #pragma unroll
      for (int j = 0; j < 18; j++) st[j] = fma2({a0[j], a1[j]}, {1e-3f, 1e-3f}, st[j]);
    } else {
      pf2 acc[18];
      #pragma unroll
      for (int j = 0; j < 18; j++) acc[j] = {0.f, 0.f};
      #pragma unroll 2
      for (int k = 0; k < NK; ++k) {
        const float2 xv = xp[k * 32];
        const pf2 x = {xv.x, xv.y};
        #pragma unroll
        for (int j = 0; j < 18; j++) {
          pf2 qp;
          if (MODE == 1) { const float2 qq = cqd[k * NC + col0 + j]; qp = {qq.x, qq.y}; }
          else { const float qq = cq[k * NC + col0 + j]; qp = {qq, qq}; }
          acc[j] = fma2(x, qp, acc[j]);
        }
      }
      #pragma unroll
      for (int j = 0; j < 18; j++) st[j] = fma2(acc[j], {1e-3f, 1e-3f}, st[j]);
    }
    #pragma unroll
    for (int j = 0; j < 18; j++) X[((buf ^ 1) * 72 + col0 + j) * 32 + lane] = make_float2(st[j].x, st[j].y);
    __syncthreads();
  }
  float s = 0; for (int j = 0; j < 18; j++) s += st[j].x + st[j].y;
  out[blockIdx.x * 128 + threadIdx.x] = s;
}


// ---- TMEM helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot){
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t base){
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d){
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&r)[16]){
  uint32_t u[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(u[0]),"=r"(u[1]),"=r"(u[2]),"=r"(u[3]),"=r"(u[4]),"=r"(u[5]),"=r"(u[6]),"=r"(u[7]),"=r"(u[8]),"=r"(u[9]),"=r"(u[10]),"=r"(u[11]),"=r"(u[12]),"=r"(u[13]),"=r"(u[14]),"=r"(u[15]) : "r"(addr));
  #pragma unroll
  for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_wait_ld(){ asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st(){ asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// V4 / V5 / V6.  QSRC: 0 = natural Q from smem (LDS.128) + MOV dup, 1 = Q slice in TMEM (tcgen05.ld x16 per 4 k) + MOV dup.
// SPLIT: 0 = thread tile (4 traj x 2 quad) x 4 cols, 4 warps ; 1 = c / s split over 8 warps, tile 4 traj x 4 cols.
template <int QSRC, int SPLIT>
__global__ void __launch_bounds__(256, 1) vt(float* out, int steps, const float* qg){
  extern __shared__ __align__(16) float sm[];
  __shared__ uint32_t tslot;
  float4* Qn = (float4*)sm;                 // [72][18] natural
  float* X = sm + 72 * 18 * 4;              // [2][72][96]
  constexpr int NQ = SPLIT ? 1 : 2;
  const int tid = threadIdx.x, half = SPLIT ? (tid >> 7) : 0, t7 = tid & 127;
  const int rg = t7 % 7, cg = t7 / 7; const bool act = cg < 18; const int cgc = act ? cg : 0;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 72 * 18; i += blockDim.x) Qn[i] = make_float4(qg[i % 977], qg[(i + 1) % 977], qg[(i + 2) % 977], qg[(i + 3) % 977]);
  for (int i = tid; i < 2 * 72 * 96; i += blockDim.x) X[i] = 0.001f * (i % 13);
  uint32_t tbase = 0;
  if (QSRC == 1) {
    if (warp == 0) tmem_alloc512(&tslot);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tbase = tslot;
  }
  __syncthreads();
  const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  if (QSRC == 1 && (SPLIT == 0 || half == 0)) {
    for (int k = 0; k < 72; ++k) { float4 q4 = Qn[k * 18 + cgc]; tmem_st4(tlane + 4 * k, q4.x, q4.y, q4.z, q4.w); }
    tmem_wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pf2 st[NQ][2][4];
  for (int q = 0; q < NQ; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) st[q][p][j] = {0.01f * tid, 0.02f};
  for (int t = 0; t < steps; ++t) {
    const int buf = t & 1;
    pf2 acc[NQ][2][4];
    for (int q = 0; q < NQ; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) acc[q][p][j] = {0.f, 0.f};
    const float4* qp = Qn + cgc; const float* xp = X + buf * 72 * 96 + rg * 4 + half * 28;
    float qn[16], qnx[16];
    if (QSRC == 1) { tmem_ld16(tlane, qnx); }
    for (int kc = 0; kc < 18; ++kc) {
      const float* xrow = xp + ((kc * 28) & 31);
      if (QSRC == 1) {
        tmem_wait_ld();
        #pragma unroll
        for (int i = 0; i < 16; i++) qn[i] = qnx[i];
        if (kc + 1 < 18) tmem_ld16(tlane + 16 * (kc + 1), qnx);
      }
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float q0, q1, q2, q3;
        if (QSRC == 1) { q0 = qn[4 * kk]; q1 = qn[4 * kk + 1]; q2 = qn[4 * kk + 2]; q3 = qn[4 * kk + 3]; }
        else { float4 qa = qp[0]; q0 = qa.x; q1 = qa.y; q2 = qa.z; q3 = qa.w; }
        pf2 qd[4] = {{q0, q0}, {q1, q1}, {q2, q2}, {q3, q3}};
        pf2 xv[NQ][2];
        #pragma unroll
        for (int q = 0; q < NQ; q++) { float4 x4 = *(const float4*)(xrow + q * 28); xv[q][0] = {x4.x, x4.y}; xv[q][1] = {x4.z, x4.w}; }
        #pragma unroll
        for (int q = 0; q < NQ; q++)
          #pragma unroll
          for (int p = 0; p < 2; p++)
            #pragma unroll
            for (int j = 0; j < 4; j++) acc[q][p][j] = fma2(xv[q][p], qd[j], acc[q][p][j]);
        qp += 18; xrow += 96;
      }
      xp += 4 * 96;
    }
    if (act) {
      #pragma unroll
      for (int j = 0; j < 4; j++)
        #pragma unroll
        for (int q = 0; q < NQ; q++) {
          st[q][0][j] = fma2(acc[q][0][j], {1e-3f, 1e-3f}, st[q][0][j]); st[q][1][j] = fma2(acc[q][1][j], {1e-3f, 1e-3f}, st[q][1][j]);
          float* dst = X + ((buf ^ 1) * 72 + 4 * cgc + j) * 96 + ((cgc * 28) & 31) + (q + half) * 28 + rg * 4;
          *(float4*)dst = make_float4(st[q][0][j].x, st[q][0][j].y, st[q][1][j].x, st[q][1][j].y);
        }
    }
    __syncthreads();
  }
  float s = 0; for (int q = 0; q < NQ; q++) for (int p = 0; p < 2; p++) for (int j = 0; j < 4; j++) s += st[q][p][j].x + st[q][p][j].y;
  out[blockIdx.x * 256 + tid] = s;
  if (QSRC == 1) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_free512(tbase);
  }
}

// V8: raw tcgen05.ld bandwidth (x16 per thread, all warps), no math.
__global__ void __launch_bounds__(256, 1) vld(float* out, int iters){
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc512(&tslot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot, tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  for (int k = 0; k < 128; ++k) tmem_st4(tlane + 4 * k, 1.f, 2.f, 3.f, 4.f);
  tmem_wait_st();
  float s = 0.f;
  for (int it = 0; it < iters; ++it) {
    float a[16], b[16];
    tmem_ld16(tlane + ((it * 32) & 255), a);
    tmem_ld16(tlane + ((it * 32 + 16) & 255), b);
    tmem_wait_ld();
    s += a[0] + b[15] + a[7];
  }
  out[blockIdx.x * 256 + threadIdx.x] = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_free512(tbase);
}

int main(int argc, char** argv){
  int steps = argc > 1 ? atoi(argv[1]) : 1500;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *out, *qg; cudaMalloc(&out, sms * 256 * 4); cudaMalloc(&qg, 4096);
  float h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1e-3f * ((i * 37) % 101 - 50);
  cudaMemcpy(qg, h, 4096, cudaMemcpyHostToDevice);
  static float hq[NK * NC]; static float2 hqd[NK * NC];
  for (int i = 0; i < NK * NC; i++) { hq[i] = 1e-3f * ((i * 37) % 101 - 50); hqd[i] = make_float2(hq[i], hq[i]); }
  cudaMemcpyToSymbol(cq, hq, sizeof(hq)); cudaMemcpyToSymbol(cqd, hqd, sizeof(hqd));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const size_t sm0 = (72 * 36 * 4 + 2 * 72 * 96) * 4, sm1 = 2 * 72 * 32 * 8;
  cudaFuncSetAttribute(v0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0);
  const size_t sm2 = (72 * 18 * 4 + 2 * 72 * 96) * 4;
  cudaFuncSetAttribute(vt<0,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  cudaFuncSetAttribute(vt<1,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  cudaFuncSetAttribute(vt<0,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  cudaFuncSetAttribute(vt<1,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  for (int v = 0; v < 8; ++v) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) v0<<<sms, 128, sm0>>>(out, steps, qg);
      if (v == 1) vu<1><<<sms, 128, sm1>>>(out, steps);
      if (v == 2) vu<2><<<sms, 128, sm1>>>(out, steps);
      if (v == 3) vu<3><<<sms, 128, sm1>>>(out, steps);
      if (v == 4) vt<0, 0><<<sms, 128, sm2>>>(out, steps, qg);
      if (v == 5) vt<1, 0><<<sms, 128, sm2>>>(out, steps, qg);
      if (v == 6) vt<0, 1><<<sms, 256, sm2>>>(out, steps, qg);
      if (v == 7) vt<1, 1><<<sms, 256, sm2>>>(out, steps, qg);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    // useful FMAs per SM per step: v0: 126 thr * 32 * 72 ; vu: 112 lanes(28x4) * 18 * 2 * 72 (count all 32 lanes as issued)
    const double fma_per_step = (v == 0 || v >= 4) ? 126.0 * 32 * 72 : 128.0 * 36 * 72;
    printf("{\"variant\": %d, \"ms\": %.4f, \"us_per_step\": %.4f, \"fma_lanes_per_clk_per_sm_at_1965MHz\": %.1f, \"err\": \"%s\"}\n", v, best,
           best * 1e3 / steps, fma_per_step / (best * 1e-3 / steps * 1.965e9), cudaGetErrorString(e));
  }
  for (int nw = 4; nw <= 8; nw += 4) {
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0); vld<<<sms, nw * 32, 0>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double bytes = (double)iters * 2 * 16 * 4 * nw * 32;
    printf("{\"variant\": \"ldtm_x16\", \"warps\": %d, \"ms\": %.4f, \"bytes_per_clk_per_sm_at_1965MHz\": %.1f, \"err\": \"%s\"}\n", nw, best,
           bytes / (best * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
