// Shared device helpers for the CCVM B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ccvm {

enum : int { SOLVER_DL = 0, SOLVER_MF = 1, SOLVER_LV = 2, SOLVER_PLV = 3 };

template <int SOLVER>
struct SolverTraits {
  static constexpr int K = (SOLVER == SOLVER_DL) ? 2 : 1;        // contraction inputs per trajectory
  static constexpr int NSTATE = (SOLVER == SOLVER_LV || SOLVER == SOLVER_PLV) ? 1 : 2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

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

// ---- the engine's noise ---------------------------------------------------------------------
// Two generators, both keyed by (seed, offset) and by GLOBAL trajectory indices, so results never
// depend on how a batch is split over CTAs, launches or GPUs:
//
//  * counter mode (the tcgen05 kernels, whose threads walk hundreds of column groups per iteration):
//    one Philox4x32-10 call per (trajectory, iteration, column group, quadrature) -> 4 normals;
//  * stream mode (the tiled SIMT kernels, CCVM_SIMT_RNG == 1): a thread owns the pair of trajectories
//    (2p, 2p+1) and one column group for the whole run, so it carries ONE xoshiro128+ state
//    (Blackman & Vigna 2018; recommended by its authors for floating-point generation, which only
//    consumes the high bits) seeded by a Philox4x32-10 call on (pair, column group) and draws its
//    2 x K x 4 normals per iteration from it in a fixed order.  Per 4 normals that is ~32
//    add / shift / xor instructions on the ALU pipe instead of 20 IMAD.WIDE on the FMA pipe -- the
//    pipe the drift contraction saturates (ncu, profiles/r1t_ncu_sde_dl_adam_final.txt: IMAD.WIDE was
//    6.5 % of the instructions, 16 % of the stall samples and ~4 FMA-pipe cycles apiece).
//
// Both end in the same Box-Muller transform.  ccvm_dump_noise replays either generator, so the
// normals of any production solve can be handed to the CPU oracle.
#ifndef CCVM_SIMT_RNG
#define CCVM_SIMT_RNG 1
#endif
// (One stream per TRAJECTORY -- two independent chains per thread, 8 state registers, no alignment
// rule for traj_base -- was measured and dropped: no gain at N = 70, DL + Adam 5 % slower,
// profiles/r2b_quick_bench_n70_per_trajectory_streams.jsonl.)
constexpr uint32_t NOISE_DOMAIN = 0xCC5DE200u;  // keeps solver streams apart from torch's own Philox use of the same seed

__device__ __forceinline__ void noise_normals4(uint32_t k0, uint32_t k1, uint32_t off_lo, unsigned long long gb,
                                               uint32_t t, uint32_t cg, uint32_t qi, float& n0, float& n1, float& n2,
                                               float& n3) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)gb, t, cg | (qi << 24) | ((uint32_t)(gb >> 32) << 25), off_lo),
                                make_uint2(k0, k1 ^ NOISE_DOMAIN));
  box_muller(r.x, r.y, n0, n1);
  box_muller(r.z, r.w, n2, n3);
}

struct NoiseStream {
  uint32_t s0, s1, s2, s3;
};
// stream of (global trajectory pair, column group)
__device__ __forceinline__ NoiseStream stream_init(uint32_t k0, uint32_t k1, uint32_t off_lo, unsigned long long pair,
                                                   uint32_t cg) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)pair, (uint32_t)(pair >> 32), cg | 0x80000000u, off_lo),
                                make_uint2(k0, k1 ^ NOISE_DOMAIN));
  NoiseStream s = {r.x, r.y, r.z, r.w};
  if ((r.x | r.y | r.z | r.w) == 0u) s.s0 = 1u;  // the one state xoshiro cannot leave
  return s;
}
__device__ __forceinline__ uint32_t stream_next(NoiseStream& s) {
  const uint32_t result = s.s0 + s.s3;
  const uint32_t t = __funnelshift_l(0u, s.s1, 9);  // s1 << 9 as a funnel shift: stays on the ALU pipe (no IMAD.SHL)
  s.s2 ^= s.s0;
  s.s3 ^= s.s1;
  s.s1 ^= s.s2;
  s.s0 ^= s.s3;
  s.s2 ^= t;
  s.s3 = __funnelshift_l(s.s3, s.s3, 11);
  return result;
}
__device__ __forceinline__ void stream_normals4(NoiseStream& s, float& n0, float& n1, float& n2, float& n3) {
  const uint32_t a = stream_next(s), b = stream_next(s), c = stream_next(s), d = stream_next(s);
  box_muller(a, b, n0, n1);
  box_muller(c, d, n2, n3);
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

// ---- per-iteration schedules --------------------------------------------------------------
// The reference evaluates its per-iteration schedules (pump ramp, noise-ratio decay,
// measurement-strength decay, Adam bias corrections) as fp64 host scalars
// (dl_solver.py:523-527,704,715; mf_solver.py:550-559; pumped_langevin_solver.py:278-283).
// They are evaluated once per launch, in fp64, on the device, and rounded to fp32 per use: by every
// CTA of a single-instance launch in its prologue (FusedTail::sched_inline, no extra launch), or by
// build_schedule_batch_kernel for a batched launch.
struct SchedArgs {
  int solver, adam, iterations, flag;
  double pump, dt, noise_ratio, j, fs, g, beta1, beta2;
};

__device__ __forceinline__ void schedule_row(const SchedArgs& a, int i, float* __restrict__ out) {
  if (i >= a.iterations) return;
  const double t = (double)(i + 1), T = (double)a.iterations;
  const double rate = a.flag ? t / T : 1.0;
  const double decay = exp(-t / T * 3.0);
  float r[SCHED_W] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (a.solver == SOLVER_DL) {
    const double ratio = (a.noise_ratio - 1.0) * decay + 1.0;
    const double p = a.pump * rate;  // == pump*(i+1)/T when the flag is set, else pump
    r[SC_A] = (float)(a.adam ? a.dt : a.dt * a.fs * (0.5 + rate));
    r[SC_P1] = (float)(a.dt * (-1.0 + p));
    r[SC_P2] = (float)(a.dt * (-1.0 - p));
    r[SC_N1] = (float)(2.0 * a.g * sqrt(a.dt) * ratio);
    r[SC_N2] = (float)(2.0 * a.g * sqrt(a.dt) / ratio);
  } else if (a.solver == SOLVER_MF) {
    const double ji = a.j * decay;
    r[SC_A] = (float)(sqrt(1.0 / (4.0 * ji)) / sqrt(a.dt));
    r[SC_P1] = (float)(a.pump * rate);
    r[SC_P2] = (float)ji;
    r[SC_N1] = (float)(sqrt(ji) / sqrt(a.dt));
    r[SC_N2] = (float)(1.0 + ji);
  } else if (a.solver == SOLVER_PLV) {
    r[SC_P1] = (float)(a.dt * (a.pump * rate - 1.0));
  }
  if (a.adam) {
    // beta^t as exp(t ln beta): two fp64 exps instead of two pows (the table is built inside the persistent
    // kernel's prologue now); the difference, ~1e-14 relative at t = 1500, vanishes in the rounding to fp32
    r[SC_IB1] = (float)(1.0 / (1.0 - exp(t * log(a.beta1))));
    r[SC_IB2] = a.beta2 == 1.0 ? 0.f : (float)(1.0 / (1.0 - exp(t * log(a.beta2))));
  }
  float4* o = reinterpret_cast<float4*>(out + (size_t)i * SCHED_W);
  o[0] = make_float4(r[0], r[1], r[2], r[3]);
  o[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// ---- tail of Solver.__call__ (epilogue.cuh) ---------------------------------------------------
struct EpiParams {
  const float* q;
  const float* v;
  const float* state;
  const float* m1vec;
  const float* m2vec;
  const float* scaled_by_ptr;  // optional device scalar overriding scaled_by (a scaling factor still on the device)
  float* pv;
  float* energy;
  int n, batch, ld, q_in_smem;
  int map1, map2, pp, pp_iters;
  float m1s, m1o, m2s, m2o, step, lo, hi, scaled_by;
  int wpt, cpl;  // tiled body: warps per tile of 8 trajectories, columns per lane (1, 2 or 4)
};

// result block of ccvm_solution_stats (36 bytes) ...
struct StatsOut {
  float best;
  int arg_best;
  int counts[7];
};
// ... and of a fused launch (ccvm_solve_fused): the same block followed by the device-measured
// duration of the two phases (max over CTAs, nanoseconds of %globaltimer)
struct FusedOut {
  StatsOut stats;
  uint32_t ctas;
  unsigned long long loop_ns, tail_ns;
};
// cross-CTA accumulators of a fused launch (zeroed by the host before the launch)
struct StatsAccum {
  unsigned long long key;   // (orderable(-E) << 32) | (0xffffffff - trajectory): atomicMax = best, ties to the lowest index
  int counts[7];
  int saw_nan;
  unsigned int done;        // CTAs that have merged; the last one writes FusedOut
  unsigned int pad;
  unsigned long long loop_ns, tail_ns;
};

// What a persistent SDE kernel does besides the loop, so that one Solver.__call__ is ONE launch:
// the schedule table in its prologue and -- on the CTA's own trajectories -- the change of
// variables, the post-processor, the BoxQP energy and the solution statistics after the loop.
struct FusedTail {
  int sched_inline;        // 1: evaluate the schedule table into sched_scratch[cta][T][SCHED_W] in the prologue
  int epilogue;            // 1: run `epi` on trajectories [cta range) after the loop
  int stats;               // 1: merge best / argmin / success counters into `accum`, last CTA writes `out`
  unsigned int total_ctas;
  float optimal;
  SchedArgs sa;
  float* sched_scratch;
  EpiParams epi;
  StatsAccum* accum;
  FusedOut* out;
};

}  // namespace ccvm
