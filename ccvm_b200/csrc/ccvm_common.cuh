// Shared device helpers for the CCVM B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ccvm {

enum : int { SOLVER_DL = 0, SOLVER_MF = 1, SOLVER_LV = 2, SOLVER_PLV = 3 };

constexpr int SCHED_W = 8;  // floats per iteration in the schedule table
// schedule slots (meaning depends on the solver, see build_schedule_kernel)
enum : int { SC_A = 0, SC_P1 = 1, SC_P2 = 2, SC_N1 = 3, SC_N2 = 4, SC_IB1 = 5, SC_IB2 = 6 };

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) -------------------------
// One pf2 holds the same variable of two neighbouring trajectories (x: even, y: odd).
struct __align__(8) pf2 {
  float x, y;
};
typedef unsigned long long u64;

__device__ __forceinline__ u64& as_u64(pf2& a) { return reinterpret_cast<u64&>(a); }
__device__ __forceinline__ const u64& as_u64(const pf2& a) { return reinterpret_cast<const u64&>(a); }
__device__ __forceinline__ pf2 pk(float a, float b) {
  pf2 r;
  r.x = a;
  r.y = b;
  return r;
}
__device__ __forceinline__ pf2 dup(float a) { return pk(a, a); }
__device__ __forceinline__ pf2 fma2(const pf2& a, const pf2& b, const pf2& c) {
  pf2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return d;
}
__device__ __forceinline__ pf2 mul2(const pf2& a, const pf2& b) {
  pf2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
__device__ __forceinline__ pf2 add2(const pf2& a, const pf2& b) {
  pf2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
// torch.clamp propagates NaN (a diverged trajectory must stay NaN); fminf/fmaxf
// would silently replace it by a bound, so use the NaN-propagating FMNMX forms.
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
}
__device__ __forceinline__ pf2 clamp2(const pf2& a, float lo, float hi) {
  return pk(clampf(a.x, lo, hi), clampf(a.y, lo, hi));
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ pf2 sqrt2(const pf2& a) { return pk(fast_sqrt(a.x), fast_sqrt(a.y)); }
__device__ __forceinline__ pf2 div2(const pf2& a, const pf2& b) {
  return pk(__fdividef(a.x, b.x), __fdividef(a.y, b.y));
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: no per-thread state ------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// two uniforms in (0,1] -> two standard normals (Box-Muller on the MUFU pipe).  u1 >= 2^-33 is never
// denormal, so the flush-to-zero lg2 needs none of the range fix-up __log2f carries (3 instructions
// per pair); the angle comes out of its conversion FFMA already in radians.
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = fmaf(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float th = fmaf(__uint2float_rn(b), 1.4629180792671596e-09f, 7.314590396335798e-10f);  // 2 pi u2
  const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));  // sqrt(-2 ln u1)
  float sn, cs;
  __sincosf(th, &sn, &cs);
  n0 = r * cs;
  n1 = r * sn;
}

// ---- launch parameters of the persistent SDE kernel ------------------------------------
struct SdeParams {
  const float* q;          // [n][n]
  const float* v;          // [n]
  const float* drift_s_vec;  // optional [n]: per-variable S used inside the drift
  const float* clamp_s_vec;  // optional [n]: per-variable S used by the box clamp
  const float* sched;      // [T][SCHED_W]
  const float* noise;      // replay tensor or nullptr
  float* out0;
  float* out1;
  float* out2;
  float* samples;          // evolution buffer or nullptr
  long long noise_batch;   // replay: trajectory extent of the noise tensor
  long long traj_base;     // global index of trajectory 0 of this launch
  int n, batch, iterations;
  int rg;                  // trajectory groups per CTA (trajectories per CTA = rg * TB)
  int cg;                  // column groups = ceil(n / 4)
  int xs;                  // floats per k-row of the staged state panel
  int use_tma;             // stage raw Q with a bulk async copy
  int evolution_step, num_samples;
  float drift_s, clamp_s;  // scalar S for drift / clamp when the vectors are null
  float a_half;            // (upper - lower) / 2
  float b_half;            // (upper + lower) / 2
  float dt, fs, g2, sig, dtfs;
  float beta1, beta2, omb1, omb2, adam_alpha;  // omb = 1 - beta, rounded from fp64
  int add_assign, beta2_is_one;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
  uint32_t pin_mask;       // always 0 (a run-time zero the compiler cannot fold; see sde_kernel_tmem.cuh, PIPE)
};

}  // namespace ccvm
