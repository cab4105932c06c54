// Persistent Euler-Maruyama kernel, tensor-core variant for SMALL n (n <= 192) and large single batches:
// the drift contraction of every iteration runs on the 5th-generation tensor cores (tcgen05, FP16-split operands),
// the SIMT pipes only draw the noise and apply the solver's elementwise step.
// Reference loops: dl_solver.py:468-769, mf_solver.py:493-764, langevin_solver.py:368-561,
// pumped_langevin_solver.py:232-449 (the einsum "bi,ij->bj" + the elementwise SDE step).
//
// Why (profiles/r2x_ncu_bench_kernel.txt): the register-tile kernel of sde_kernel_tmem.cuh
// is ISSUE-bound -- an FFMA2 holds its scheduler for two cycles and the contraction is 55 % of all
// instructions (636 of 1154 per warp-iteration for DL + Adam at N = 70), so no code shape gets it far
// beyond ~0.5 of the FP32 FMA peak.  Here the contraction leaves the SIMT pipes altogether:
//
//   D[v][b] = sum_k Qs[k][v] x_b[k]        (v: variable = MMA row, b: trajectory = MMA column)
//
//   * operands are FP16 PAIRS: x sigma = hi + lo, Qs tau = hi + lo with hi = fp16(.), lo = fp16(. - hi) and sigma, tau
//     powers of two that put the largest magnitude of each operand near 2^13 (x: the clamp bound S, or 2^5
//     headroom above it for the unclamped DL amplitudes, saturating); D = Alo.Bhi + Ahi.Blo + Ahi.Bhi keeps
//     ~22 bits of every product like 3xTF32 does, but an FP16 MMA contracts K = 16 per instruction: half the
//     instructions (every tcgen05.mma costs ~20 cycles at these tiny N, measured, and the loop was bound by them);
//   * A = Qs^T (hi | lo) -- CONSTANT for the whole run -- is written ONCE into tensor memory and stays there
//     (tcgen05.mma with the A operand in TMEM: no shared-memory read of the 128 x K operand per MMA);
//   * B = the contraction input of 16 trajectories per warpgroup (hi, lo; K-major, no swizzle, 144-byte
//     K stride so that a warp's stores are conflict-free), rewritten by the update threads every iteration:
//     a few KB of shared memory;
//   * per warpgroup and iteration one elected lane of warp 8 issues 3 x K/16 tcgen05.mma kind::f16 (M = 128,
//     N = 16 or 32) and commits to an mbarrier;
//   * thread (warpgroup g, TMEM lane m) owns ONE variable for the 14-16 trajectories of its warpgroup: it
//     draws their noise and the drift-independent part of the step while the MMAs run, reads its row of D
//     with one tcgen05.ld, rescales it and adds the affine term h_v, finishes the step, splits the new contraction
//     input into (hi, lo) and stores it into B; an mbarrier (128 arrivals) hands the tile back to the issuer.
//     Two warpgroups per CTA run out of phase on the same four schedulers.
//
// Variables are dealt to the four TMEM lane quadrants round-robin (v = 4 lane + quadrant), so that the four
// warps of a warpgroup carry the same load; 70 variables occupy 18 lanes of every warp.  Variables 128 ... 191 are the
// rows of a SECOND M tile (its own A in tensor memory, its own accumulators; twice the MMAs per iteration).
// Noise: one xoshiro128+ stream per (global trajectory pair, variable), seeded by Philox4x32-10 (ccvm_common.cuh);
// a Box-Muller pair serves the two trajectories of the pair.  ccvm_dump_noise reproduces it.
#pragma once
#include "ccvm_common.cuh"
#include "epilogue.cuh"
#include "sde_kernel_tc.cuh"
#include "sde_launch.h"

namespace ccvm {

#ifndef CCVM_MMA_ISSUERS
#define CCVM_MMA_ISSUERS 1
#endif
constexpr int MMA_ISSUERS = CCVM_MMA_ISSUERS;   // 1: one warp serves both warpgroups; 2: issuer warp g serves warpgroup g
constexpr int MMA_UW = 8;                                  // update warps: two warpgroups of four
constexpr int MMA_THREADS = 32 * (MMA_UW + MMA_ISSUERS);   // the update warps, then the MMA issuer(s)
constexpr int MMA_N_MAX = 192;            // two M tiles of 128 rows; 64 D + 192 A columns of tensor memory per M tile
constexpr int MMA_LBO = 144;              // bytes between the two 16-byte K chunks of a core matrix pair (128 + 16:
                                          // 32 consecutive K positions land in different banks)
constexpr int MMA_D_COLS = 64;            // TMEM columns per M tile for the two warpgroups' accumulators (32 each)
constexpr uint32_t MMA_STREAM_TAG = 0x40000000u;      // keeps these noise streams apart from the column-group streams

// Shared-memory layout by the number of M tiles: the one-tile kernels (n <= 128) are sized for K <= 128 (71 KB per CTA instead
// of 107 KB).  Measured neutral for the loop time (+-1 %, profiles/r2zz_one_tile_layout_and_fence.txt); what did cost the
// one-tile kernels 7-14 % after the two-tile extension was the point where warpgroup 0 releases warpgroup 1 (LATE below).
// Tuning switches of that measurement: -DCCVM_MMA_WIDE_LAYOUT=1 sizes the one-tile kernels like the two-tile ones,
// -DCCVM_MMA_LATE_FENCE=1 moves the tcgen05 fence + phase arrival of every kernel behind the redistribution stores.
#ifndef CCVM_MMA_WIDE_LAYOUT
#define CCVM_MMA_WIDE_LAYOUT 0
#endif
#ifndef CCVM_MMA_LATE_FENCE
#define CCVM_MMA_LATE_FENCE 0
#endif
// one-tile kernels that take the late arrival anyway (bit = 2 solver + adam): DL settles 3-8 % faster with it (n = 70 ... 128),
// DL-adam 7 % slower, the others within 1-3 % either way
#ifndef CCVM_MMA_LATE_MASK
#define CCVM_MMA_LATE_MASK 0x01
#endif
template <int MT_>
struct MmaLayout {
  static constexpr int MT = CCVM_MMA_WIDE_LAYOUT ? 2 : MT_;
  static constexpr int KD_MAX = MT == 1 ? 128 : 192;   // K extent: n rounded up to 16 (one FP16 MMA contracts 16)
  static constexpr int NV = 128 * MT;                  // entries of the per-variable tables alpha_v, h_v
  // bytes between 8-row groups (compile-time: immediates); the 16 extra bytes shift every row group by four banks: a
  // warp's staging store covers rows 0 ... 12 of ~5 K positions, and with a stride of 0 mod 128 bytes rows 8-12 fell on
  // the banks of rows 0-4 (2.0 wavefronts per store; now 1.0-1.1)
  static constexpr int SBO = (KD_MAX / 8) * MMA_LBO + 16;
  static constexpr int TILE_BYTES = 8 * SBO;   // one warpgroup's B tile: rows [hi | lo], up to 2 x 32 (DL: 16 c + 16 s)
  // per warp and quadrature: float2 slots of the row -> item redistribution (32 MT variables per quadrant x 8 pairs,
  // capped by 32 lanes x 11 items)
  static constexpr int SCRATCH_ITEMS = MT == 1 ? 256 : 384;
  static constexpr size_t SMEM = 2 * NV * sizeof(float) + 128 + 2 * (size_t)TILE_BYTES + (size_t)MMA_UW * 2 * SCRATCH_ITEMS * 8;
};

inline size_t mma_loop_smem_bytes(int mt) { return mt == 2 ? MmaLayout<2>::SMEM : MmaLayout<1>::SMEM; }

// D[tmem] (+)= A[tmem] . B[smem]^T, A = 128 x 16 FP16 in tensor memory (lane = row, 32-bit column = two k)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle; core matrices of 8 rows x 16 bytes
template <int SBO>
__device__ __forceinline__ uint64_t umma_desc_k_none(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);   // start address  [0,14)
  d |= (uint64_t)(MMA_LBO >> 4) << 16;           // leading byte offset: next 16-byte K chunk
  d |= (uint64_t)(SBO >> 4) << 32;               // stride byte offset: next group of 8 rows
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  return d;                                      // layout type 0: SWIZZLE_NONE
}
__device__ __forceinline__ void tmem_ld_row16(uint32_t addr, float (&r)[16]) { tmem_ld16(addr, r); }
__device__ __forceinline__ void tmem_ld_row32(uint32_t addr, float (&r)[32]) {
  uint32_t u[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(addr));
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}

// FP16 split of a scaled value pair: hi = fp16(x) (saturating: an unclamped amplitude beyond the headroom must not
// become Inf), lo = fp16(x - hi); both halves of a register: .x in the low 16 bits
__device__ __forceinline__ uint32_t f16x2_sat(float lo_half, float hi_half) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
__device__ __forceinline__ float f16_lo_to_f32(uint32_t h2) {
  float r;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(r) : "r"(h2));
  return r;
}
__device__ __forceinline__ float f16_hi_to_f32(uint32_t h2) {
  float r;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(r) : "r"(h2));
  return r;
}
// largest power of two p with v p < 2^(e + 1)  (v > 0 finite; 1 for v = 0)
__device__ __forceinline__ float pow2_scale(float v, int e) {
  if (!(v > 0.f)) return 1.f;
  const int ev = (int)((__float_as_uint(v) >> 23) & 0xffu) - 127;
  int k = e - ev;
  k = k < -60 ? -60 : k > 60 ? 60 : k;
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

// Adam transform of one gradient pair (dl_solver.py:699-727 and siblings): adam_tile4 of sde_kernel_tmem.cuh
// for one element pair, the per-iteration scalars hoisted by the caller
struct AdamConsts {
  pf2 b1, b2, eps, ca, sa, aa;
};
__device__ __forceinline__ AdamConsts adam_consts(const SdeParams& p, float ib1, float ib2) {
  const bool b2one = p.beta2_is_one != 0;
  AdamConsts c;
  c.b1 = dup(p.beta1);
  c.b2 = dup(b2one ? 0.f : p.beta2);
  c.eps = dup(b2one ? 1.f : 1e-8f);
  c.ca = dup(p.adam_alpha * p.omb1 * ib1);
  c.sa = dup(b2one ? 0.f : fast_sqrt(p.omb2 * ib2));
  c.aa = dup(p.add_assign ? 1.f : 0.f);
  return c;
}
__device__ __forceinline__ pf2 adam_pair(const pf2 gr, pf2& m, pf2& v, const AdamConsts& c) {
  m = fma2(m, c.b1, gr);
  v = fma2(v, c.b2, mul2(gr, gr));
  const pf2 den = fma2(sqrt2(v), c.sa, c.eps);
  const pf2 u = mul2(m, pk(fast_rcp(den.x), fast_rcp(den.y)));
  return fma2(u, c.ca, mul2(gr, c.aa));
}

// one Box-Muller pair from the thread's stream: the same variable of the two trajectories of a pair
__device__ __forceinline__ pf2 stream_normal_pair(NoiseStream& s) {
  const uint32_t a = stream_next(s), b = stream_next(s);
  pf2 w;
  box_muller(a, b, w.x, w.y);
  return w;
}

// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

#ifdef CCVM_MMA_TRACE
// development aid: clock64 stamps of CTA 0 (first thread of warpgroup g: slots 8 g + 0..5, issuer serving g: 8 g + 6, 7)
// for iterations 64 .. 95
__device__ long long g_mma_trace[32 * 16];
#define MMA_STAMP(cond, t, slot)                                                          \
  if ((cond) && blockIdx.x == 0 && (t) >= 64 && (t) < 96) g_mma_trace[((t) - 64) * 16 + (slot)] = clock64();
#else
#define MMA_STAMP(cond, t, slot)
#endif

struct MmaLaunch {
  int kd;      // K extent of the contraction: n rounded up to a multiple of 16
  int tcols;   // TMEM columns to allocate (power of two >= MT (MMA_D_COLS + kd)); MT = 1 for n <= 128, else 2: template argument
  int nbp;     // trajectory pairs per warpgroup (<= 8): a CTA advances 4 nbp trajectories
  int stagger; // warpgroup 1 starts half an iteration after warpgroup 0
};

// Variables are dealt to the four TMEM lane quadrants round-robin, M tile by M tile: v = 128 m + 4 i + q is row 32 q + i of
// M tile m (lane i of quadrant q), has the LOCAL index li = 32 m + i in its quadrant and sits at K position koff(q) + li
// (quadrant by quadrant, no holes: a warp's stores hit consecutive K positions).  n <= 128: one M tile, li = i.
// One M tile (MT = 1): v = 4 i + q, the plain expressions the n <= 128 kernels were tuned with.
__device__ __forceinline__ int mma_qcount_tile(int n, int q, int m) {   // rows of quadrant q in M tile m
  const int r = n - 128 * m;
  return r <= 0 ? 0 : r >= 128 ? 32 : (r - q + 3) >> 2;
}
template <int MT>
__device__ __forceinline__ int mma_qcount(int n, int q) {
  if constexpr (MT == 1) return (n - q + 3) >> 2;
  else return mma_qcount_tile(n, q, 0) + mma_qcount_tile(n, q, 1);
}
template <int MT>
__device__ __forceinline__ int mma_koff(int n, int q) {
  int o = 0;
  for (int i = 0; i < q; ++i) o += mma_qcount<MT>(n, i);
  return o;
}
// (a second M tile exists only when the first is full: its local indices start at 32)
template <int MT>
__device__ __forceinline__ int mma_var(int li, int q) {
  if constexpr (MT == 1) return 4 * li + q;
  else return li < 32 ? 4 * li + q : 128 + 4 * (li - 32) + q;
}

// IPL: (variable, trajectory pair) items per lane.  A warp reads the rows of its TMEM lane quadrant (one variable per
// lane, 18 of 32 lanes at n = 70) and deals the values out again through shared memory, so that EVERY lane owns
// IPL items for the whole run (n = 70, 7 pairs: 126 items on 32 x 4 slots) -- the elementwise work, two thirds of it
// the noise, is the bound of this kernel and would otherwise run at 56 % lane occupancy.
template <int SOLVER, bool ADAM, int IPL, int MT>
__global__ void __launch_bounds__(MMA_THREADS, 1)
    sde_mma_kernel(const SdeParams p, const MmaLaunch L, const FusedTail f) {
  constexpr int K = SolverTraits<SOLVER>::K;
  using LY = MmaLayout<MT>;
  constexpr int MMA_KD_MAX = LY::KD_MAX, MMA_SBO = LY::SBO, MMA_TILE_BYTES = LY::TILE_BYTES;
  constexpr int MMA_SCRATCH_ITEMS = LY::SCRATCH_ITEMS, NV = LY::NV;
  constexpr int NR = 16 * K;   // B rows per half (hi | lo) and warpgroup (DL: c rows 0-15, s rows 16-31)
  // tcgen05 fence + phase arrival behind the redistribution stores (else right after the accumulator read)
  constexpr bool LATE = MT > 1 || CCVM_MMA_LATE_FENCE || ((CCVM_MMA_LATE_MASK >> (2 * SOLVER + (ADAM ? 1 : 0))) & 1) != 0;
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) unsigned long long bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned int s_max[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.n, T = p.iterations, KD = L.kd, NBP = L.nbp;
  const int cta = blockIdx.x;
  float* av = smem;                                                    // [NV] alpha_v
  float* hv = smem + NV;                                               // [NV] affine term h_v
  const uint32_t tiles = (smem_u32(smem) + 2 * NV * 4 + 127u) & ~127u; // [warpgroup] B tiles, rows [hi | lo]
  uint8_t* tiles_g = reinterpret_cast<uint8_t*>(smem) + (tiles - smem_u32(smem));
  float2* scratch = reinterpret_cast<float2*>(tiles_g + 2 * MMA_TILE_BYTES);   // [warp][K][MMA_SCRATCH_ITEMS]
  const uint32_t bar0 = smem_u32(bars);
  auto ready_bar = [&](int g) { return bar0 + 8u * g; };       // B tile of warpgroup g written (128 arrivals)
  auto done_bar = [&](int g) { return bar0 + 8u * (2 + g); };  // accumulator of warpgroup g complete (tcgen05.commit)
  const uint32_t phase_bar = bar0 + 8u * 4;                    // L.stagger: warpgroup 0 has read its first accumulator

  // ------------------------------------------------------------------ prologue
  const unsigned long long t_start = f.stats ? global_timer_ns() : 0ull;
  if (tid == 0) {
    mbar_init(ready_bar(0), 128);
    mbar_init(ready_bar(1), 128);
    mbar_init(done_bar(0), 1);
    mbar_init(done_bar(1), 1);
    mbar_init(phase_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_max[0] = s_max[1] = 0u;
  }
  if (warp == MMA_UW) tmem_alloc(&tmem_slot, L.tcols);
  const float* sched = p.sched;
  if (f.sched_inline) {
    float* mine = f.sched_scratch + (size_t)cta * p.iterations * SCHED_W;
    build_schedule_cta(f.sa, mine);
    sched = mine;
  }
  for (int j = tid; j < NV; j += MMA_THREADS)
    av[j] = j < N ? p.a_half / (p.drift_s_vec ? p.drift_s_vec[j] : p.drift_s) : 0.f;
  for (int i = tid; i < 2 * MMA_TILE_BYTES / 16; i += MMA_THREADS)
    reinterpret_cast<float4*>(tiles_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  {
    // operand scales: max |Qs| and max clamp bound (positive floats order like their bit patterns)
    float mq = 0.f, ms = 0.f;
    for (int idx = tid; idx < N * N; idx += MMA_THREADS) {
      const int k = idx / N, j = idx - k * N;
      mq = fmaxf(mq, fabsf(av[k] * av[j] * p.q[idx]));
    }
    for (int j = tid; j < N; j += MMA_THREADS) ms = fmaxf(ms, fabsf(p.clamp_s_vec ? p.clamp_s_vec[j] : p.clamp_s));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
      ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
    }
    if (lane == 0) {
      atomicMax(&s_max[0], __float_as_uint(mq));
      atomicMax(&s_max[1], __float_as_uint(ms));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  // Qs tau < 2^14; x sigma < 2^14 for the clamped inputs (Langevin, PumpedLangevin, MF: |x| <= S), S sigma < 2^9 for
  // the unclamped DL amplitudes (FP16 saturates at 65504: 2^7 S, far outside the dynamics)
  const float tau = pow2_scale(__uint_as_float(s_max[0]), 13);
  const float sigma = pow2_scale(__uint_as_float(s_max[1]), SOLVER == SOLVER_DL ? 8 : 13);
  const float unscale = 1.f / (tau * sigma);   // exact: powers of two

  // tensor memory: [accumulators: (warpgroup g, M tile m) at 32 (g MT + m)] [A of M tile m: hi | lo, KD / 2 columns each]
  const uint32_t a_base = tbase + (uint32_t)(MMA_D_COLS * MT);
  if (warp < 4) {
    // warpgroup 0 writes A = tau Qs^T (hi | lo) into tensor memory: lane 32 q + i of M tile m = row of variable
    // v = 128 m + 4 i + q, 32-bit column c = K positions 2 c (low half), 2 c + 1; K position P <-> input variable k
    for (int m = 0; m < MT; ++m) {
      const int v = 128 * m + 4 * lane + warp;
      const bool valid = v < N;
      float aj = 0.f;
      if (valid) {
        float cs = 0.f;
        for (int i = 0; i < N; ++i) cs += p.q[i * N + v];
        aj = av[v];
        hv[v] = -aj * (p.b_half * cs + p.v[v]);
      }
      const int o1 = mma_koff<MT>(N, 1), o2 = mma_koff<MT>(N, 2), o3 = mma_koff<MT>(N, 3);
      const uint32_t tl = a_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * KD);
      for (int c4 = 0; c4 < KD / 8; ++c4) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float val[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int P = 8 * c4 + 2 * e + u;
            val[u] = 0.f;
            if (valid && P < N) {
              const int q = P >= o3 ? 3 : P >= o2 ? 2 : P >= o1 ? 1 : 0;
              const int k = mma_var<MT>(P - (q == 3 ? o3 : q == 2 ? o2 : q == 1 ? o1 : 0), q);
              val[u] = (-av[k] * aj * p.q[k * N + v]) * tau;
            }
          }
          hi[e] = f16x2_sat(val[0], val[1]);
          lo[e] = f16x2_sat(val[0] - f16_lo_to_f32(hi[e]), val[1] - f16_hi_to_f32(hi[e]));
        }
        tmem_st4(tl + 4 * c4, __uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3]));
        tmem_st4(tl + KD / 2 + 4 * c4, __uint_as_float(lo[0]), __uint_as_float(lo[1]), __uint_as_float(lo[2]),
                 __uint_as_float(lo[3]));
      }
    }
    tmem_wait_st();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long per_cta = 4 * NBP;
  if (warp >= MMA_UW) {
    // ================================================================ MMA issuer
    // The whole warp stays converged and ONE elected lane issues (elect.sync): in a divergent `if (lane == 0)`
    // ptxas wraps every tcgen05.mma into an ELECT / BRA.U.ANY waterfall (~7 instructions and a branch per MMA:
    // measured ~35 cycles each).  Per k-step of 16:  D += Alo . Bhi^T + Ahi . Blo^T + Ahi . Bhi^T  (small terms first).
    // (Contracting [Bhi | Blo] against Ahi in ONE instruction of twice the N was measured: same tensor time -- it is
    // proportional to the columns contracted, ~28 cycles per 16 columns and k-step -- and more work for the threads.)
    // instruction descriptor: D = F32, A = B = F16, K-major, M = 128, N = NR
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int KS = KD / 16;
    auto issue = [&](int g) {
      if (elect_one()) {
        const uint64_t b_hi = umma_desc_k_none<MMA_SBO>(tiles + g * MMA_TILE_BYTES);
        const uint64_t b_lo = umma_desc_k_none<MMA_SBO>(tiles + g * MMA_TILE_BYTES + (NR / 8) * MMA_SBO);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t d_tmem = tbase + (uint32_t)(32 * (g * MT + m));
          const uint32_t a_hi = a_base + (uint32_t)(m * KD), a_lo = a_hi + KD / 2;
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t adv = (uint64_t)(ks * (2 * MMA_LBO >> 4));   // two 16-byte K chunks per MMA
            umma_f16_ts(d_tmem, a_lo + 8 * ks, b_hi + adv, idesc, ks != 0);
            umma_f16_ts(d_tmem, a_hi + 8 * ks, b_lo + adv, idesc, 1u);
            umma_f16_ts(d_tmem, a_hi + 8 * ks, b_hi + adv, idesc, 1u);
          }
        }
        umma_commit(done_bar(g));
      }
      __syncwarp();
    };
    // The warpgroups are served in STRICT ALTERNATION with blocking waits: mbarrier.try_wait suspends the warp (no issue
    // slots taken from the update warps of its scheduler) and wakes within tens of cycles of the last arrival.  Polling
    // both barriers with mbarrier.test_wait + __nanosleep ("whoever is ready first") left a ready tile waiting ~300
    // cycles for the issuer: 3-8 % of every loop (profiles/r2z_issuer_protocol.txt).
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (MMA_ISSUERS == 2 && g != warp - MMA_UW) continue;
        mbar_wait(ready_bar(g), (uint32_t)(t & 1));
        __syncwarp();
        tc_fence_after();
        MMA_STAMP(lane == 0, t, 8 * g + 6)
        issue(g);
        MMA_STAMP(lane == 0, t, 8 * g + 7)
      }
    }
  } else {
    // ================================================================ update warpgroups
    const int g = warp >> 2, w4 = warp & 3;
    const int cnt = mma_qcount<MT>(N, w4), koff = mma_koff<MT>(N, w4);   // this quadrant's variables mma_var(li, w4), li < cnt
    const int n_items = cnt * NBP;                                // (variable, pair) items: e = li * NBP + pair
    int cnt_tile[MT];                                             // ... of which in M tile m (lanes of this quadrant)
#pragma unroll
    for (int m = 0; m < MT; ++m) cnt_tile[m] = MT == 1 ? cnt : mma_qcount_tile(N, w4, m);
    const long long b0 = (long long)cta * per_cta + (long long)g * 2 * NBP;   // first trajectory of the warpgroup
    const uint32_t tl = tbase + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(32 * g * MT);
    uint8_t* tile = tiles_g + (size_t)g * MMA_TILE_BYTES;
    float2* sw = scratch + (size_t)warp * 2 * MMA_SCRATCH_ITEMS;
    const pf2 usc = dup(unscale), sg2 = dup(sigma);

    // lane L owns items e = 32 j + L (consecutive lanes read consecutive slots of the redistribution buffer)
    int vj[IPL], prj[IPL];
    bool okj[IPL];
    uint8_t* xbj[IPL];
    float hj[IPL], scj[IPL];
    pf2 st0[IPL], st1[IPL];                  // c | mu, s | sigma   (x: trajectory 2 pair, y: 2 pair + 1)
    pf2 m0[IPL], v0[IPL], m1[IPL], v1[IPL];  // Adam moments
    pf2 W[IPL];                              // MF: noise of the current measurement
    pf2 meas[IPL];
    NoiseStream rs[IPL];
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi ^ p.off_hi);
#pragma unroll
    for (int j = 0; j < IPL; ++j) {
      const int e = 32 * j + lane;
      okj[j] = e < n_items;
      const int i = okj[j] ? e / NBP : 0;
      prj[j] = okj[j] ? e - i * NBP : 0;
      vj[j] = mma_var<MT>(i, w4);
      // K position in the B tile; idle slots run the same code on zeros and store where it cannot matter: just past
      // the K extent (the row groups are laid out for KD = MMA_KD_MAX, no MMA reads there), or for KD = MMA_KD_MAX at a K position
      // whose A column is zero (n <= P < KD; the value is finite: saturating conversion, clamped or cubic-saturated
      // dynamics) -- no predicate, no branch around the stores
      const int P = okj[j] ? koff + i : (KD < MMA_KD_MAX ? KD : N);
      const int row = 2 * prj[j];
      xbj[j] = tile + (P >> 3) * MMA_LBO + (P & 7) * 2 + (row >> 3) * MMA_SBO + (row & 7) * 16;
      hj[j] = okj[j] ? hv[vj[j]] : 0.f;
      scj[j] = okj[j] ? (p.clamp_s_vec ? p.clamp_s_vec[vj[j]] : p.clamp_s) : 0.f;
      st0[j] = dup(0.f);
      st1[j] = dup(SOLVER == SOLVER_MF ? 0.5f : 0.f);
      m0[j] = v0[j] = m1[j] = v1[j] = dup(0.f);
      W[j] = meas[j] = dup(0.f);
      rs[j] = stream_init(key.x, key.y, p.off_lo, (unsigned long long)(p.traj_base + b0 + 2 * prj[j]) >> 1,
                          (uint32_t)(okj[j] ? vj[j] : 0) | MMA_STREAM_TAG);
    }
    // contraction input of an item's trajectory pair (rows 2 pair, 2 pair + 1 of the hi half; `quad` = 1: the s rows):
    // scaled, split into FP16 (hi, lo) and stored; the lo half of the tile starts NR rows further down
    auto stage2 = [&](uint8_t* base, int quad, const pf2 x) {
      const pf2 xs = mul2(x, sg2);
      const uint32_t hi = f16x2_sat(xs.x, xs.y);
      const uint32_t lo = f16x2_sat(xs.x - f16_lo_to_f32(hi), xs.y - f16_hi_to_f32(hi));
      uint8_t* dst = base + quad * 2 * MMA_SBO;
      *reinterpret_cast<unsigned short*>(dst) = (unsigned short)hi;
      *reinterpret_cast<unsigned short*>(dst + 16) = (unsigned short)(hi >> 16);
      *reinterpret_cast<unsigned short*>(dst + (NR / 8) * MMA_SBO) = (unsigned short)lo;
      *reinterpret_cast<unsigned short*>(dst + (NR / 8) * MMA_SBO + 16) = (unsigned short)(lo >> 16);
    };
    const float4* sched4 = reinterpret_cast<const float4*>(sched);
    float4 sa = sched4[0], sb = sched4[1];
    if constexpr (SOLVER == SOLVER_MF) {
      // measurement of iteration 0 (mf_solver.py:551-554): mu = 0
#pragma unroll
      for (int j = 0; j < IPL; ++j) {
        W[j] = stream_normal_pair(rs[j]);
        meas[j] = clamp2(fma2(dup(sa.x), W[j], st0[j]), -scj[j], scj[j]);
        stage2(xbj[j], 0, meas[j]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // L.stagger: the two warpgroups start HALF AN ITERATION apart (warpgroup 1 waits until warpgroup 0 has read its first
    // accumulator).  The phase between them is not pinned by anything once the loop runs: where the MMA chain has slack
    // (DL, the Adam tiles, MF at n > 96) it keeps whatever offset the start left, and with both warpgroups released
    // together one runs right behind the other -- both in their MUFU-heavy noise phase at once, 6-10 % slower, and
    // WHICH tiles fall into that state changes with unrelated code changes (profiles/r2z_issuer_protocol.txt).
    // Starting half an iteration apart puts every tile in the good state; the loops bound by the MMA chain settle out of
    // phase by themselves and do not care.
    if (L.stagger && g == 1) mbar_wait(phase_bar, 0u);
    mbar_arrive(ready_bar(g));

    // ---------------------------------------------------------------- main loop
    for (int t = 0; t < T; ++t) {
      const float4 ca = sa, cb = sb;
      if (t + 1 < T) {
        sa = sched4[2 * (t + 1)];
        sb = sched4[2 * (t + 1) + 1];
      }
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 0)
      // ---- everything that does not depend on the drift, while the tensor core contracts
      if constexpr (SOLVER == SOLVER_DL) {
        const pf2 d1 = dup(ca.y), d2 = dup(ca.z), n1 = dup(ca.w), n2 = dup(cb.x);
        const pf2 mdt = dup(-p.dt), half = dup(0.5f);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          const pf2 wc = stream_normal_pair(rs[j]), ws = stream_normal_pair(rs[j]);
          const pf2 c = st0[j], s = st1[j];
          const pf2 r2 = fma2(c, c, mul2(s, s));
          const pf2 rt = sqrt2(add2(r2, half));
          const pf2 uc = fma2(r2, mdt, d1), us = fma2(r2, mdt, d2);
          st0[j] = add2(c, fma2(c, uc, mul2(mul2(rt, n1), wc)));
          st1[j] = add2(s, fma2(s, us, mul2(mul2(rt, n2), ws)));
        }
      } else if constexpr (SOLVER == SOLVER_MF) {
        // mf_solver.py:158-233 with the constants folded (as in sde_kernel_tmem.cuh); sigma does not see the drift
        const pf2 pr = dup(ca.y), sj = dup(ca.w), opj = dup(cb.x), m2j = dup(-2.f * ca.z);
        const pf2 dtp = dup(p.dt), mhalf = dup(-0.5f);
        const pf2 ng2 = dup(-p.g2), n3g2 = dup(-3.f * p.g2), p2g2 = dup(2.f * p.g2), two = dup(2.f);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          const pf2 mu = st0[j], sg = st1[j];
          const pf2 mm = mul2(mu, mu);
          const pf2 a1 = fma2(mm, ng2, pr);
          const pf2 sh = add2(sg, mhalf);
          st0[j] = fma2(dtp, fma2(mul2(sh, W[j]), sj, mul2(a1, mu)), mu);
          const pf2 a3 = fma2(mm, n3g2, pr);
          const pf2 t3 = fma2(mm, p2g2, opj);
          const pf2 inner = fma2(mul2(sh, sh), m2j, t3);
          st1[j] = fma2(dtp, fma2(mul2(a3, sg), two, inner), sg);
          W[j] = stream_normal_pair(rs[j]);   // the measurement noise of iteration t + 1
        }
      } else {
        const pf2 sig = dup(p.sig), mdt = dup(-p.dt), d1 = dup(ca.y);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          const pf2 w = stream_normal_pair(rs[j]);
          const pf2 c = st0[j];
          pf2 pre = fma2(sig, w, c);
          if constexpr (SOLVER == SOLVER_PLV) pre = fma2(c, fma2(mul2(c, c), mdt, d1), pre);
          st0[j] = pre;
        }
      }
      AdamConsts ac;
      if constexpr (ADAM) ac = adam_consts(p, cb.y, cb.z);

      // ---- the drift of this iteration: this lane's ROW of D, dealt out to the lanes' items through the warp's
      //      redistribution buffer
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 1)
      mbar_wait(done_bar(g), (uint32_t)(t & 1));
      tc_fence_after();
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 2)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float d[NR];
        if constexpr (K == 2) tmem_ld_row32(tl + 32 * m, d);
        else tmem_ld_row16(tl + 32 * m, d);
        tmem_wait_ld();
        if constexpr (!LATE) {   // (before the redistribution stores: the order the one-tile loops were tuned in)
          tc_fence_before();
          if (L.stagger && g == 0 && t == 0) mbar_arrive(phase_bar);
        }
        if (lane < cnt_tile[m]) {
#pragma unroll
          for (int pr = 0; pr < 8; ++pr) {
            if (pr < NBP) {
#pragma unroll
              for (int h = 0; h < K; ++h)
                sw[h * MMA_SCRATCH_ITEMS + (32 * m + lane) * NBP + pr] =
                    make_float2(d[16 * h + 2 * pr], d[16 * h + 2 * pr + 1]);
            }
          }
        }
      }
      if constexpr (LATE) {
        tc_fence_before();
        if (L.stagger && g == 0 && t == 0) mbar_arrive(phase_bar);
      }
      __syncwarp();
      pf2 gq[K][IPL];
#pragma unroll
      for (int j = 0; j < IPL; ++j)
#pragma unroll
        for (int h = 0; h < K; ++h) {
          const float2 t2 = sw[h * MMA_SCRATCH_ITEMS + 32 * j + lane];
          gq[h][j] = fma2(pk(t2.x, t2.y), usc, dup(hj[j]));
        }
      __syncwarp();
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 3)

      // ---- finish the step and publish the next contraction input
      if constexpr (SOLVER == SOLVER_DL) {
        const pf2 gain = dup(ca.x);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          pf2 gc = gq[0][j], gs = gq[K - 1][j];
          if constexpr (ADAM) {
            gc = adam_pair(gc, m0[j], v0[j], ac);
            gs = adam_pair(gs, m1[j], v1[j], ac);
          }
          st0[j] = fma2(gain, gc, st0[j]);
          st1[j] = fma2(gain, gs, st1[j]);
          stage2(xbj[j], 0, st0[j]);
          stage2(xbj[j], 1, st1[j]);
        }
      } else if constexpr (SOLVER == SOLVER_MF) {
        const pf2 fs = dup(p.fs), dtp = dup(p.dt);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          pf2 gr = mul2(fs, gq[0][j]);
          if constexpr (ADAM) gr = adam_pair(gr, m0[j], v0[j], ac);
          st0[j] = fma2(dtp, gr, st0[j]);
          if (t + 1 < T) {
            meas[j] = clamp2(fma2(dup(sa.x), W[j], st0[j]), -scj[j], scj[j]);
            stage2(xbj[j], 0, meas[j]);
          }
        }
      } else {
        const pf2 dtfs = dup(p.dtfs);
#pragma unroll
        for (int j = 0; j < IPL; ++j) {
          pf2 gr = gq[0][j];
          if constexpr (ADAM) gr = adam_pair(gr, m0[j], v0[j], ac);
          st0[j] = clamp2(fma2(dtfs, gr, st0[j]), -scj[j], scj[j]);
          stage2(xbj[j], 0, st0[j]);
        }
      }
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 4)
      if (t + 1 < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(ready_bar(g));
      }
      MMA_STAMP((tid & 127) == 0, t, 8 * g + 5)
    }

    // ---------------------------------------------------------------- results
#pragma unroll
    for (int j = 0; j < IPL; ++j) {
      if (!okj[j]) continue;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long b = b0 + 2 * prj[j] + i;
        if (b >= p.batch) continue;
        const size_t o = (size_t)b * N + vj[j];
        const float x0 = i ? st0[j].y : st0[j].x, x1 = i ? st1[j].y : st1[j].x;
        if constexpr (SOLVER == SOLVER_DL) {
          p.out0[o] = clampf(x0, -scj[j], scj[j]);
          p.out1[o] = x1;
        } else if constexpr (SOLVER == SOLVER_MF) {
          p.out0[o] = x0;
          p.out1[o] = i ? meas[j].y : meas[j].x;
          p.out2[o] = x1;
        } else {
          p.out0[o] = x0;
        }
      }
    }
  }

  // ------------------------------------------------------------------ fused tail of Solver.__call__
  // (the code of sde_kernel_tmem.cuh's tail: change of variables -> post-processor -> energy -> statistics)
  if (f.epilogue) {
    __syncthreads();
    const unsigned long long t_loop = f.stats ? global_timer_ns() : 0ull;
    const long long b_begin = (long long)cta * per_cta;
    const long long b_end = b_begin + per_cta < p.batch ? b_begin + per_cta : p.batch;
    epilogue_run(f.epi, smem, b_begin, b_end, 0, 1);
    if (f.stats) {
      __syncthreads();
      StatsPartial sp;
      if (stats_block_reduce(f.epi.energy, b_begin, b_end, f.optimal, sp))
        stats_merge(sp, f.accum, f.out, f.total_ctas, t_loop - t_start, global_timer_ns() - t_loop);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_UW) tmem_free(tbase, L.tcols);
}

}  // namespace ccvm
