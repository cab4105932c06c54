// Persistent Euler-Maruyama kernel, tensor-core variant for SMALL n (n <= 128) and large single batches:
// the drift contraction of every iteration runs on the 5th-generation tensor cores (tcgen05, 3xTF32),
// the SIMT pipes only draw the noise and apply the solver's elementwise step.
// Reference loops: dl_solver.py:468-769, mf_solver.py:493-764, langevin_solver.py:368-561,
// pumped_langevin_solver.py:232-449 (the einsum "bi,ij->bj" + the elementwise SDE step).
//
// Why (profiles/r2x_ncu_bench_kernel.txt): the register-tile kernel of sde_kernel_tmem.cuh
// is ISSUE-bound -- an FFMA2 holds its scheduler for two cycles and the contraction is 55 % of all
// instructions (636 of 1154 per warp-iteration for DL + Adam at N = 70), so no code shape gets it far
// beyond ~0.5 of the FP32 FMA peak.  Here the contraction leaves the SIMT pipes altogether:
//
//   D[v][b] = sum_k Qs[k][v] x_b[k] + h_v        (v: variable = MMA row, b: trajectory = MMA column)
//
//   * A = Qs^T (hi, lo) with the affine term h as one extra K column -- CONSTANT for the whole run -- is written
//     ONCE into tensor memory and stays there (tcgen05.mma with the A operand in TMEM: no shared-memory read of
//     the 128 x K operand per MMA, which is what a 16-column MMA would otherwise be bound by);
//   * B = the contraction input of 16 trajectories per warpgroup (hi, lo; K-major, no swizzle, 144-byte
//     K stride so that a warp's stores are conflict-free), rewritten by the update threads every iteration:
//     a few KB of shared memory;
//   * per warpgroup and iteration one lane of warp 8 issues  D = Alo.Bhi + Ahi.Blo + Ahi.Bhi  (3 x K/8
//     tcgen05.mma kind::tf32, M = 128, N = 16 or 32) and commits to an mbarrier;
//   * thread (warpgroup g, TMEM lane m) owns ONE variable for the 14-16 trajectories of its warpgroup: it
//     draws their noise and the drift-independent part of the step while the MMAs run, reads its row of D
//     with one tcgen05.ld, finishes the step, splits the new contraction input into (hi, lo) and stores it
//     into B; an mbarrier (128 arrivals) hands the tile back to the issuer.  Two warpgroups per CTA run out
//     of phase on the same four schedulers.
//
// Variables are dealt to the four TMEM lane quadrants round-robin (v = 4 lane + quadrant), so that the four
// warps of a warpgroup carry the same load; 70 variables occupy 18 lanes of every warp.
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
constexpr int MMA_ISSUERS = CCVM_MMA_ISSUERS;   // 1: warp 8 serves both warpgroups; 2: warp 8 + g serves warpgroup g
constexpr int MMA_THREADS = 256 + 32 * MMA_ISSUERS;   // warps 0-3, 4-7: two update warpgroups; then the MMA issuer(s)
constexpr int MMA_KD_MAX = 136;           // K extent: n + 1 (affine column), rounded up to 8
constexpr int MMA_LBO = 144;              // bytes between the two 16-byte K chunks of a core matrix pair (128 + 16:
                                          // 32 consecutive K positions land in 32 different banks)
constexpr int MMA_SBO = (MMA_KD_MAX / 4) * MMA_LBO;   // bytes between 8-row groups (compile-time: immediates)
constexpr int MMA_TILE_BYTES = 4 * MMA_SBO;           // one B tile: up to 32 rows (DL: 16 c rows + 16 s rows)
#ifndef CCVM_MMA_NACC
#define CCVM_MMA_NACC 1
#endif
constexpr int MMA_NACC = CCVM_MMA_NACC;   // independent accumulators per warpgroup: consecutive MMAs of a chain into ONE
                                          // accumulator serialise on its read-modify-write (~21 cycles per 16-column
                                          // MMA measured, against a floor of 8), so the MMAs rotate over NACC of them
                                          // and the update thread adds the partial sums
constexpr int MMA_D_COLS = 2 * 3 * 32;    // TMEM columns [0, 192): [warpgroup][accumulator][32]
constexpr uint32_t MMA_STREAM_TAG = 0x40000000u;      // keeps these noise streams apart from the column-group streams

__host__ __device__ inline size_t mma_loop_smem_bytes() { return 128 * sizeof(float) + 128 + 4 * (size_t)MMA_TILE_BYTES; }

// D[tmem] (+)= A[tmem] . B[smem]^T, A = 128 x 8 TF32 in tensor memory (lane = row, column = k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle; core matrices of 8 rows x 16 bytes
__device__ __forceinline__ uint64_t umma_desc_k_none(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);   // start address  [0,14)
  d |= (uint64_t)(MMA_LBO >> 4) << 16;           // leading byte offset: next 16-byte K chunk
  d |= (uint64_t)(MMA_SBO >> 4) << 32;           // stride byte offset: next group of 8 rows
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

// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
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
// development aid: clock64 stamps of CTA 0 (warpgroup 0 thread 0: slots 0-5, issuer: 6-7) for iterations 64 .. 95
__device__ long long g_mma_trace[32 * 8];
#define MMA_STAMP(cond, t, slot)                                                          \
  if ((cond) && blockIdx.x == 0 && (t) >= 64 && (t) < 96) g_mma_trace[((t) - 64) * 8 + (slot)] = clock64();
#else
#define MMA_STAMP(cond, t, slot)
#endif

struct MmaLaunch {
  int kd;      // K extent of the contraction: n + 1 rounded up to a multiple of 8
  int tcols;   // TMEM columns to allocate (power of two >= MMA_D_COLS + 2 kd)
};

// Variables are dealt to the four TMEM lane quadrants round-robin: v = 4 i + q sits in lane i of quadrant q and at
// K position koff(q) + i (quadrant by quadrant, no holes: a warp's stores hit consecutive K positions)
__device__ __forceinline__ int mma_qcount(int n, int q) { return (n - q + 3) >> 2; }
__device__ __forceinline__ int mma_koff(int n, int q) {
  int o = 0;
  for (int i = 0; i < q; ++i) o += mma_qcount(n, i);
  return o;
}

// NBP: trajectory pairs per warpgroup (7 or 8; a CTA advances 4 NBP trajectories)
template <int SOLVER, bool ADAM, int NBP>
__global__ void __launch_bounds__(MMA_THREADS, 1)
    sde_mma_kernel(const SdeParams p, const MmaLaunch L, const FusedTail f) {
  constexpr int K = SolverTraits<SOLVER>::K;
  constexpr int NR = 16 * K;   // B rows = D columns per warpgroup (DL: c rows 0-15, s rows 16-31)
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.n, T = p.iterations, KD = L.kd;
  const int cta = blockIdx.x;
  float* av = smem;                                                    // [128] alpha_v
  const uint32_t tiles = (smem_u32(smem) + 128 * 4 + 127u) & ~127u;    // [warpgroup][hi | lo] B tiles
  uint8_t* tiles_g = reinterpret_cast<uint8_t*>(smem) + (tiles - smem_u32(smem));
  const uint32_t bar0 = smem_u32(bars);
  auto ready_bar = [&](int g) { return bar0 + 8u * g; };       // B tile of warpgroup g written (128 arrivals)
  auto done_bar = [&](int g) { return bar0 + 8u * (2 + g); };  // accumulator of warpgroup g complete (tcgen05.commit)

  // ------------------------------------------------------------------ prologue
  const unsigned long long t_start = f.stats ? global_timer_ns() : 0ull;
  if (tid == 0) {
    mbar_init(ready_bar(0), 128);
    mbar_init(ready_bar(1), 128);
    mbar_init(done_bar(0), 1);
    mbar_init(done_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(&tmem_slot, L.tcols);
  const float* sched = p.sched;
  if (f.sched_inline) {
    float* mine = f.sched_scratch + (size_t)cta * p.iterations * SCHED_W;
    build_schedule_cta(f.sa, mine);
    sched = mine;
  }
  for (int j = tid; j < 128; j += MMA_THREADS)
    av[j] = j < N ? p.a_half / (p.drift_s_vec ? p.drift_s_vec[j] : p.drift_s) : 0.f;
  for (int i = tid; i < 4 * MMA_TILE_BYTES / 16; i += MMA_THREADS)
    reinterpret_cast<float4*>(tiles_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const int PB = N;   // K position of the affine column

  if (warp < 4) {
    // warpgroup 0 writes A = Qs^T (hi | lo) into tensor memory: lane m = row of variable v = 4 (m % 32) + m / 32,
    // column = K position P of input variable k = 4 (P % QW) + P / QW, column PB = h_v
    const int v = 4 * lane + warp;
    const bool valid = v < N;
    float h = 0.f, aj = 0.f;
    if (valid) {
      float cs = 0.f;
      for (int i = 0; i < N; ++i) cs += p.q[i * N + v];
      aj = av[v];
      h = -aj * (p.b_half * cs + p.v[v]);
    }
    const int o1 = mma_koff(N, 1), o2 = mma_koff(N, 2), o3 = mma_koff(N, 3);
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16) + MMA_D_COLS;
    for (int c4 = 0; c4 < KD / 4; ++c4) {
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int P = 4 * c4 + e;
        float val = 0.f;
        if (valid) {
          if (P < PB) {
            const int q = P >= o3 ? 3 : P >= o2 ? 2 : P >= o1 ? 1 : 0;
            const int k = 4 * (P - (q == 3 ? o3 : q == 2 ? o2 : q == 1 ? o1 : 0)) + q;
            val = -av[k] * aj * p.q[k * N + v];
          } else if (P == PB) {
            val = h;
          }
        }
        hi[e] = tf32_rna(val);
        lo[e] = val - hi[e];
      }
      tmem_st4(tl + 4 * c4, hi[0], hi[1], hi[2], hi[3]);
      tmem_st4(tl + KD + 4 * c4, lo[0], lo[1], lo[2], lo[3]);
    }
    tmem_wait_st();
  } else if (warp < 8) {
    // warpgroup 1: the constant 1 that meets the affine column (hi tiles of both warpgroups, every row)
    const int r = tid - 128;
    if (r < 2 * NR) {
      const int g = r / NR, row = r % NR;
      *reinterpret_cast<float*>(tiles_g + (size_t)g * 2 * MMA_TILE_BYTES + (row >> 3) * MMA_SBO + (row & 7) * 16 +
                                (PB >> 2) * MMA_LBO + (PB & 3) * 4) = 1.f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long per_cta = 4 * NBP;
  if (warp >= 8) {
    // ================================================================ MMA issuer
    // The whole warp stays converged and ONE elected lane issues (elect.sync): in a divergent `if (lane == 0)`
    // ptxas wraps every tcgen05.mma into an ELECT / BRA.U.ANY waterfall (~7 instructions and a branch per MMA:
    // measured ~35 cycles each, 2 x 30 MMAs per iteration through one thread = half of the iteration).
    // instruction descriptor: D = F32, A = B = TF32, K-major, M = 128, N = NR
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int KS = KD / 8;
    const uint32_t a_hi = tbase + MMA_D_COLS, a_lo = a_hi + KD;
    // whichever warpgroup has its tile ready is served first (non-blocking tests: the two run out of phase)
    int it[2] = {0, 0};
    uint32_t spins = 0;
    long long spin_start = 0;
    while (it[0] < T || it[1] < T) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (it[g] >= T) continue;
        if (MMA_ISSUERS == 2 && g != warp - 8) {
          it[g] = T;
          continue;
        }
        if (!__all_sync(0xffffffffu, mbar_test(ready_bar(g), (uint32_t)(it[g] & 1)))) continue;
        spins = 0;
        tc_fence_after();
        MMA_STAMP(g == 0 && lane == 0, it[g], 6)
        if (elect_one()) {
          const uint32_t d_tmem = tbase + g * (3 * 32);
          const uint64_t b_hi = umma_desc_k_none(tiles + g * 2 * MMA_TILE_BYTES);
          const uint64_t b_lo = umma_desc_k_none(tiles + g * 2 * MMA_TILE_BYTES + MMA_TILE_BYTES);
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t adv = (uint64_t)(ks * (2 * MMA_LBO >> 4));   // two 16-byte K chunks per MMA
            // MMA number 3 ks + i goes to accumulator (3 ks + i) % NACC; the first one into each overwrites
            umma_tf32_ts(d_tmem + 32 * ((3 * ks + 0) % MMA_NACC), a_lo + 8 * ks, b_hi + adv, idesc, 3 * ks + 0 >= MMA_NACC);
            umma_tf32_ts(d_tmem + 32 * ((3 * ks + 1) % MMA_NACC), a_hi + 8 * ks, b_lo + adv, idesc, 3 * ks + 1 >= MMA_NACC);
            umma_tf32_ts(d_tmem + 32 * ((3 * ks + 2) % MMA_NACC), a_hi + 8 * ks, b_hi + adv, idesc, 3 * ks + 2 >= MMA_NACC);
          }
          umma_commit(done_bar(g));
        }
        __syncwarp();
        MMA_STAMP(g == 0 && lane == 0, it[g], 7)
        ++it[g];
      }
      if ((++spins & 0xfffff) == 0) {   // a protocol bug must trap, not hang the device
        const long long now = clock64();
        if (spin_start == 0) spin_start = now;
        else if (now - spin_start > 8000000000ll) __trap();
      } else if (spins == 1) {
        spin_start = 0;
      }
    }
  } else {
    // ================================================================ update warpgroups
    const int g = warp >> 2, w4 = warp & 3;
    const int v = 4 * lane + w4;                       // this thread's variable
    const bool valid = v < N;
    const int P = mma_koff(N, w4) + lane;              // its K position in the B tiles
    const long long b0 = (long long)cta * per_cta + (long long)g * 2 * NBP;   // first trajectory of the warpgroup
    const uint32_t tl = tbase + ((uint32_t)(w4 * 32) << 16) + g * (3 * 32);
    uint8_t* xb = tiles_g + (size_t)g * 2 * MMA_TILE_BYTES + (P >> 2) * MMA_LBO + (P & 3) * 4;
    const float sc = valid ? (p.clamp_s_vec ? p.clamp_s_vec[v] : p.clamp_s) : 0.f;

    pf2 st0[NBP], st1[NBP];                  // c | mu, s | sigma   (x: trajectory 2 ip, y: 2 ip + 1)
    pf2 m0[NBP], v0[NBP], m1[NBP], v1[NBP];  // Adam moments
    pf2 W[NBP];                              // MF: noise of the current measurement
    pf2 meas[NBP];
    NoiseStream rs[NBP];
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi ^ p.off_hi);
#pragma unroll
    for (int ip = 0; ip < NBP; ++ip) {
      st0[ip] = dup(0.f);
      st1[ip] = dup(SOLVER == SOLVER_MF ? 0.5f : 0.f);
      m0[ip] = v0[ip] = m1[ip] = v1[ip] = dup(0.f);
      W[ip] = meas[ip] = dup(0.f);
      rs[ip] = stream_init(key.x, key.y, p.off_lo, (unsigned long long)(p.traj_base + b0 + 2 * ip) >> 1,
                           (uint32_t)(valid ? v : 0) | MMA_STREAM_TAG);
    }
    // contraction input of row `row` (trajectory index inside the warpgroup, + 16 for the s quadrature)
    auto stage = [&](int row, float x) {
      const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);   // exact TF32 head, lo = the rest
      uint8_t* dst = xb + (row >> 3) * MMA_SBO + (row & 7) * 16;
      *reinterpret_cast<float*>(dst) = hi;
      *reinterpret_cast<float*>(dst + MMA_TILE_BYTES) = x - hi;
    };
    const float4* sched4 = reinterpret_cast<const float4*>(sched);
    float4 sa = sched4[0], sb = sched4[1];
    if constexpr (SOLVER == SOLVER_MF) {
      // measurement of iteration 0 (mf_solver.py:551-554): mu = 0
#pragma unroll
      for (int ip = 0; ip < NBP; ++ip) {
        W[ip] = stream_normal_pair(rs[ip]);
        meas[ip] = clamp2(fma2(dup(sa.x), W[ip], st0[ip]), -sc, sc);
        if (valid) {
          stage(2 * ip, meas[ip].x);
          stage(2 * ip + 1, meas[ip].y);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(ready_bar(g));

    // ---------------------------------------------------------------- main loop
    for (int t = 0; t < T; ++t) {
      const float4 ca = sa, cb = sb;
      if (t + 1 < T) {
        sa = sched4[2 * (t + 1)];
        sb = sched4[2 * (t + 1) + 1];
      }
      MMA_STAMP(tid == 0, t, 0)
      // ---- everything that does not depend on the drift, while the tensor core contracts
      if constexpr (SOLVER == SOLVER_DL) {
        const pf2 d1 = dup(ca.y), d2 = dup(ca.z), n1 = dup(ca.w), n2 = dup(cb.x);
        const pf2 mdt = dup(-p.dt), half = dup(0.5f);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          const pf2 wc = stream_normal_pair(rs[ip]), ws = stream_normal_pair(rs[ip]);
          const pf2 c = st0[ip], s = st1[ip];
          const pf2 r2 = fma2(c, c, mul2(s, s));
          const pf2 rt = sqrt2(add2(r2, half));
          const pf2 uc = fma2(r2, mdt, d1), us = fma2(r2, mdt, d2);
          st0[ip] = add2(c, fma2(c, uc, mul2(mul2(rt, n1), wc)));
          st1[ip] = add2(s, fma2(s, us, mul2(mul2(rt, n2), ws)));
        }
      } else if constexpr (SOLVER == SOLVER_MF) {
        // mf_solver.py:158-233 with the constants folded (as in sde_kernel_tmem.cuh); sigma does not see the drift
        const pf2 pr = dup(ca.y), sj = dup(ca.w), opj = dup(cb.x), m2j = dup(-2.f * ca.z);
        const pf2 dtp = dup(p.dt), mhalf = dup(-0.5f);
        const pf2 ng2 = dup(-p.g2), n3g2 = dup(-3.f * p.g2), p2g2 = dup(2.f * p.g2), two = dup(2.f);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          const pf2 mu = st0[ip], sg = st1[ip];
          const pf2 mm = mul2(mu, mu);
          const pf2 a1 = fma2(mm, ng2, pr);
          const pf2 sh = add2(sg, mhalf);
          st0[ip] = fma2(dtp, fma2(mul2(sh, W[ip]), sj, mul2(a1, mu)), mu);
          const pf2 a3 = fma2(mm, n3g2, pr);
          const pf2 t3 = fma2(mm, p2g2, opj);
          const pf2 inner = fma2(mul2(sh, sh), m2j, t3);
          st1[ip] = fma2(dtp, fma2(mul2(a3, sg), two, inner), sg);
          W[ip] = stream_normal_pair(rs[ip]);   // the measurement noise of iteration t + 1
        }
      } else {
        const pf2 sig = dup(p.sig), mdt = dup(-p.dt), d1 = dup(ca.y);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          const pf2 w = stream_normal_pair(rs[ip]);
          const pf2 c = st0[ip];
          pf2 pre = fma2(sig, w, c);
          if constexpr (SOLVER == SOLVER_PLV) pre = fma2(c, fma2(mul2(c, c), mdt, d1), pre);
          st0[ip] = pre;
        }
      }
      AdamConsts ac;
      if constexpr (ADAM) ac = adam_consts(p, cb.y, cb.z);

      // ---- the drift of this iteration
      MMA_STAMP(tid == 0, t, 1)
      mbar_wait(done_bar(g), (uint32_t)(t & 1));
      tc_fence_after();
      MMA_STAMP(tid == 0, t, 2)
      float d[NR];
      if constexpr (K == 2) tmem_ld_row32(tl, d);
      else tmem_ld_row16(tl, d);
      if constexpr (MMA_NACC > 1) {
        float e[NR], e2[NR];
        if constexpr (K == 2) tmem_ld_row32(tl + 32, e);
        else tmem_ld_row16(tl + 32, e);
        if constexpr (MMA_NACC > 2) {
          if constexpr (K == 2) tmem_ld_row32(tl + 64, e2);
          else tmem_ld_row16(tl + 64, e2);
        }
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 2 * NBP; i += 2) {
#pragma unroll
          for (int h = 0; h < K; ++h) {
            pf2 sum = add2(pk(d[16 * h + i], d[16 * h + i + 1]), pk(e[16 * h + i], e[16 * h + i + 1]));
            if constexpr (MMA_NACC > 2) sum = add2(sum, pk(e2[16 * h + i], e2[16 * h + i + 1]));
            d[16 * h + i] = sum.x;
            d[16 * h + i + 1] = sum.y;
          }
        }
      } else {
        tmem_wait_ld();
      }
      tc_fence_before();
      MMA_STAMP(tid == 0, t, 3)

      // ---- finish the step and publish the next contraction input
      if constexpr (SOLVER == SOLVER_DL) {
        const pf2 gain = dup(ca.x);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          pf2 gc = pk(d[2 * ip], d[2 * ip + 1]), gs = pk(d[16 + 2 * ip], d[16 + 2 * ip + 1]);
          if constexpr (ADAM) {
            gc = adam_pair(gc, m0[ip], v0[ip], ac);
            gs = adam_pair(gs, m1[ip], v1[ip], ac);
          }
          st0[ip] = fma2(gain, gc, st0[ip]);
          st1[ip] = fma2(gain, gs, st1[ip]);
          if (valid) {
            stage(2 * ip, st0[ip].x);
            stage(2 * ip + 1, st0[ip].y);
            stage(16 + 2 * ip, st1[ip].x);
            stage(16 + 2 * ip + 1, st1[ip].y);
          }
        }
      } else if constexpr (SOLVER == SOLVER_MF) {
        const pf2 fs = dup(p.fs), dtp = dup(p.dt);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          pf2 gr = mul2(fs, pk(d[2 * ip], d[2 * ip + 1]));
          if constexpr (ADAM) gr = adam_pair(gr, m0[ip], v0[ip], ac);
          st0[ip] = fma2(dtp, gr, st0[ip]);
          if (t + 1 < T) {
            meas[ip] = clamp2(fma2(dup(sa.x), W[ip], st0[ip]), -sc, sc);
            if (valid) {
              stage(2 * ip, meas[ip].x);
              stage(2 * ip + 1, meas[ip].y);
            }
          }
        }
      } else {
        const pf2 dtfs = dup(p.dtfs);
#pragma unroll
        for (int ip = 0; ip < NBP; ++ip) {
          pf2 gr = pk(d[2 * ip], d[2 * ip + 1]);
          if constexpr (ADAM) gr = adam_pair(gr, m0[ip], v0[ip], ac);
          st0[ip] = clamp2(fma2(dtfs, gr, st0[ip]), -sc, sc);
          if (valid) {
            stage(2 * ip, st0[ip].x);
            stage(2 * ip + 1, st0[ip].y);
          }
        }
      }
      MMA_STAMP(tid == 0, t, 4)
      if (t + 1 < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(ready_bar(g));
      }
      MMA_STAMP(tid == 0, t, 5)
    }

    // ---------------------------------------------------------------- results
    if (valid) {
#pragma unroll
      for (int ip = 0; ip < NBP; ++ip)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const long long b = b0 + 2 * ip + i;
          if (b >= p.batch) continue;
          const size_t o = (size_t)b * N + v;
          const float x0 = i ? st0[ip].y : st0[ip].x, x1 = i ? st1[ip].y : st1[ip].x;
          if constexpr (SOLVER == SOLVER_DL) {
            p.out0[o] = clampf(x0, -sc, sc);
            p.out1[o] = x1;
          } else if constexpr (SOLVER == SOLVER_MF) {
            p.out0[o] = x0;
            p.out1[o] = i ? meas[ip].y : meas[ip].x;
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
  if (warp == 8) tmem_free(tbase, L.tcols);
}

}  // namespace ccvm
