// Persistent Euler-Maruyama kernel, TMEM variant (n <= 128): the production path at the
// benchmark sizes: one launch runs ALL iterations of one Solver._solve / Solver._solve_adam call.
//
// Why (ncu, profiles/r1_ncu_sde_dl_adam_v1.txt): with the scaled matrix Qs in shared memory the
// loop was bound by shared-memory wavefronts (every LDS.128 costs 4 wavefronts no matter how much
// of it is a broadcast; 64 wavefronts per k per SM against 32 FMA-pipe cycles).  Qs is constant
// for the whole run and each thread only ever needs the 4 columns of its own tile, so:
//
//   * every thread keeps ITS slice Qs[0..NP)[4 cols] in its own TMEM lane (4*NP <= 512 columns,
//     written once with tcgen05.st) and streams it back with one tcgen05.ld.x16 per 4 k --
//     TMEM delivers > 800 B/clk/SM (tools/ubench_drift.cu) and does not touch the LSU;
//   * FFMA2 takes a scalar broadcast operand (SASS `Rq.F32`), so the natural (non-duplicated)
//     Qs feeds the packed FMA directly:  acc(b0,b1 ; j) += x(b0,b1 ; k) * Qs[k][j];
//   * only the contraction input X goes through shared memory: one LDS.128 (DL: c0,c1,s0,s1) or
//     LDS.64 per thread per k, from a double-buffered panel X[buf][k][row-group][2K];
//   * a thread owns 2 trajectories x K quadratures x 4 variables of every state array;
//   * a CTA runs NG independent trajectory groups (<= 128 threads each, one warp per SM
//     sub-partition and group), each with its own named barrier: the groups drift out of phase, so
//     one group's FMA/LSU-bound contraction overlaps the other's ALU/MUFU-bound noise + update.
//     Threads t and t+128 share a TMEM lane and therefore the same column group.
//
// Variants of the one tile body (template parameters of sde_tile_body):
//   QSRC  where the thread's Q slice lives: TMEM (n <= 128), TMEM + shared-memory tail (n <= 256),
//         streamed from L2 (fallback);
//   PIPE  production mode: Philox noise generated inside the contraction, compile-time panel stride;
//         off for noise replay (validation) and for too few column groups;
//   CGC   column-group count compiled in (N = 20, 30, ..., 70, the reference's benchmarking sizes):
//         fully unrolled contraction, immediate addresses, noise quanta scheduled by ptxas;
//   and per-tile choices made by measurement (HOIST, VSMEM, DENSE_TILE, unroll factors) -- at two
//   warps per scheduler the static schedule decides +-5 % per tile, see DESIGN.md section 4.
#pragma once
#include <type_traits>

#include "ccvm_common.cuh"
#include "epilogue.cuh"
#include "sde_launch.h"

// Per-tile code-shape choices made by measurement (bit = 2 * solver + adam; solvers DL, MF, LV, PLV;
// profiles/r2f_tuning_masks_n70.txt):
//   CCVM_HOIST_MASK  tiles whose drift-independent update math is evaluated inside the contraction
//   CCVM_UNPIN_MASK  tiles whose noise quanta are left unpinned in the compile-time variants at CG = 15, 18
//                    (all but PumpedLangevin + Adam: pinned 0.45 vs unpinned 0.40 of FP32 peak at N = 70)
// Measured and dropped in round 2 (profiles/r2m_tuning_x32_kp2_padstage.txt): tcgen05.ld.x32 for the K = 1
// tiles (half the LDTM / R2UR / wait instructions) and two k rows per panel row for n <= 128 (one LDS.128
// per two k) -- within +-3 % either way, no consistent sign over the loops and sizes.
//   CCVM_KTAIL_MASK  tiles whose N = 30 / 50 / 70 variants leave out the two padding rows of the last chunk
//                    (+1-2 % for DL, MF, Langevin(+Adam), PumpedLangevin; -1-4 % for the other Adam tiles:
//                    profiles/r2t_ktail_n30_50_70.txt)
#ifndef CCVM_HOIST_MASK
#define CCVM_HOIST_MASK 0xA3
#endif
#ifndef CCVM_KTAIL_MASK
#define CCVM_KTAIL_MASK 0x75
#endif
#ifndef CCVM_UNPIN_MASK
#define CCVM_UNPIN_MASK 0x7F
#endif

namespace ccvm {

constexpr int HYB_TMEM_CHUNKS = 32;  // chunks of 4 rows held in TMEM (128 rows x 4 columns = 512 TMEM columns)
constexpr int HYB_LD = 256;          // floats per row of the shared-memory tail

// Row stride (floats) of the state panel in the TMEM + PIPE kernels: a compile-time constant so
// that every LDS of the contraction is [base + immediate] (the run-time stride cost one IMAD on the
// FMA pipe per load).  68 = 64 + 4: rows 4 apart land 16 banks apart, which keeps the (rare) STS
// of the panel at <= 3-way conflicts without the per-row rotation of the generic layout.
constexpr int TMEM_PIPE_XS = 68;
// Same for the hybrid kernel: n > 128 leaves at most 3 trajectory pairs per 128-lane group
// (RW * RG <= 12 floats per row), and the panel has to share the SM with the 128 KB tail of Qs.
// There the K = 1 solvers pack TWO k rows into one panel row ((k even: b0,b1), (k odd: b0,b1)): one
// LDS.128 then feeds 8 FFMA2 exactly like a DL row (c0,c1,s0,s1) does, instead of one LDS.64 per
// 4 FFMA2 (n = 250: Langevin 46 -> 52 %, MF 44 -> 47 % of FP32 peak).  Measured neutral-to-worse for
// n <= 128, where up to 25 pairs share a row and the wide load costs up to 4 wavefronts: not used there.
constexpr int HYB_PIPE_XS = 20;

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(addr));
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// schedule table of one launch, evaluated by the calling CTA (fp64; see SchedArgs)
static __device__ __noinline__ void build_schedule_cta(const SchedArgs& a, float* out) {
  for (int i = threadIdx.x; i < a.iterations; i += blockDim.x) schedule_row(a, i, out);
}

// Adam transform on a 4-variable tile of trajectory pairs (dl_solver.py:699-727 and siblings).
// The moments are kept pre-divided by (1 - beta):  ms = m / (1 - beta1),  vs = v / (1 - beta2), so
//   m_hat = ms (1 - beta1) ib1,   sqrt(v_hat) = sqrt(vs) sqrt((1 - beta2) ib2)
// and the per-iteration scalars fold into two constants: 6 packed operations + 4 MUFU per pair of
// elements instead of 11 + 4 (the update phase is latency-bound, every instruction off the chain
// counts).  Same quantities as the reference up to FP32 rounding of the refactored products.
__device__ __forceinline__ void adam_tile4(pf2 (&g)[4], pf2 (&m)[4], pf2 (&v)[4], const SdeParams& p, float ib1,
                                           float ib2) {
  // The run-time options are folded into scalars so that the tile is ONE straight-line block (a
  // branch per element kept ptxas from interleaving the eight dependent MUFU chains):
  //   beta2 == 1 (dl_solver.py:707-713: update = alpha m_hat): v <- g^2, den = 0 sqrt(v) + 1 = 1;
  //   add_assign False: the gradient enters with weight 0.
  const bool b2one = p.beta2_is_one != 0;
  const pf2 b1 = dup(p.beta1), b2 = dup(b2one ? 0.f : p.beta2), eps = dup(b2one ? 1.f : 1e-8f);
  const pf2 ca = dup(p.adam_alpha * p.omb1 * ib1);
  const pf2 sa = dup(b2one ? 0.f : fast_sqrt(p.omb2 * ib2));
  const pf2 aa = dup(p.add_assign ? 1.f : 0.f);
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const pf2 gr = g[jj];
    m[jj] = fma2(m[jj], b1, gr);
    v[jj] = fma2(v[jj], b2, mul2(gr, gr));
    const pf2 den = fma2(sqrt2(v[jj]), sa, eps);
    const pf2 u = mul2(m[jj], pk(fast_rcp(den.x), fast_rcp(den.y)));
    g[jj] = fma2(u, ca, mul2(gr, aa));
  }
}

// (A variant that gave the c and s quadratures of a DL tile to different threads -- 16 warps per SM
// instead of 8 -- was measured and dropped: 43 % more instructions for a loop that is already
// FMA-pipe-bound, profiles/r1_ncu_sde_dl_tmem_v3.txt.)
// The body is shared by the single-problem kernel and the batched (many instances per launch)
// kernel: `cta` is the CTA's index inside ITS problem; threads beyond ng*gt (batched launches use
// one block size for a whole bucket of problems) only take part in the CTA-wide barriers.
//
// PIPE (Philox mode, CG > 4*K): the noise of an iteration does not depend on the state, so its
// Philox rounds and Box-Muller transforms are issued INSIDE the drift contraction, one Philox call
// (four normals) per pair of Q chunks, instead of as a separate phase after it.  The contraction
// is a pure FFMA2 stream (every FFMA2 holds the FMA pipe for two cycles and leaves every other
// issue slot free); the ALU / MUFU work and the dependent IMAD.WIDE chain of the generator fill
// those slots, and the latency-bound phase between contraction and barrier shrinks to the SDE
// update itself (profiles/r1_ncu_sde_dl_tmem_v3.txt: FMA pipe 73 % busy, 27 % bubbles, before).
// KTAIL (compile-time column-group variants only): real rows in the LAST chunk of four k -- 2 for N = 30, 50,
// 70, whose padding rows 4 CG - 2, 4 CG - 1 (zero rows of Qs) are then neither loaded nor multiplied; 0: all four.
template <int SOLVER, bool ADAM, int QSRC, bool PIPE, int CGC = 0, int KTAIL = 0>
__device__ __forceinline__ void sde_tile_body(const SdeParams& p, const TmemLaunch& L, const FusedTail& f, const int cta,
                                              float* smem, uint32_t* tmem_slot_p) {
  constexpr int K = SolverTraits<SOLVER>::K;
  constexpr int KT = K;      // quadratures handled by one thread
  constexpr bool SPLIT = false;
  constexpr int RW = 2 * K;  // floats per (k, trajectory pair): (b0,b1) or (c0,c1,s0,s1)

  uint32_t& tmem_slot = *tmem_slot_p;
  const int tid = threadIdx.x;
  constexpr bool HAS_TMEM = QSRC != QSRC_GMEM;
  // DL + Adam is the one tile at the edge of the register file (2 quadratures x (state, m, v) x 4
  // columns: ~245 registers): without the relief below ptxas schedules it measurably better (3.69 vs
  // 3.90 ms at N = 70) with the padding-column noise masked and the contraction tail left inside the
  // loop -- the two simplifications every other tile gains 5-13 % from (DENSE_TILE, kept for the
  // hybrid / streamed / replay variants of this tile).
  constexpr bool DL_ADAM = SOLVER == SOLVER_DL && ADAM;
  // DL + Adam (production kernel, n <= 128): the Adam second moments are parked in shared memory
  // between iterations (4 LDS.128 + 4 STS.128 per thread and iteration, conflict-free).  That takes
  // the tile from 245 to 240 live registers at its peak and, more to the point, lets ptxas keep the
  // peeled / unmasked / hoisted code shape of the other loops: 3.73 -> 3.59 ms at N = 70.  Each of
  // the three changes alone, or any two of them, made this tile slower (3.82-4.04 ms); parking the
  // first moments as well gave the gain back (3.74 ms).
  constexpr bool VSMEM = DL_ADAM && PIPE && QSRC == QSRC_TMEM;
  constexpr bool DENSE_TILE = DL_ADAM && !VSMEM;
  // Tiles whose drift-independent update math is evaluated inside the contraction (`precompute`):
  // chosen by measurement at N = 70 / 128 / 250 -- DL 3.23 -> 3.15 ms, Langevin + Adam 2.12 -> 2.02,
  // PumpedLangevin + Adam 2.17 -> 2.10; neutral or slower for the others (ptxas gives up FFMA2
  // overlap elsewhere), which keep the whole step after the contraction.
  constexpr int TILE_BIT = 1 << (SOLVER * 2 + (ADAM ? 1 : 0));
  constexpr bool HOIST = PIPE && (((CCVM_HOIST_MASK & TILE_BIT) != 0 && !DL_ADAM) || (VSMEM && (CCVM_HOIST_MASK & TILE_BIT) != 0));
  // compile-time panel stride (0: run time)
  constexpr int XSC = !PIPE ? 0
                      : QSRC == QSRC_TMEM ? TMEM_PIPE_XS
                      : QSRC == QSRC_HYB ? HYB_PIPE_XS : 0;
  constexpr int KP = (XSC != 0 && KT == 1 && QSRC == QSRC_HYB) ? 2 : 1;  // k rows per panel row (see HYB_PIPE_XS)
  const int N = p.n, CG = CGC ? CGC : p.cg, NP = 4 * CG, RG = L.rg, XS = XSC ? XSC : L.xs, T = p.iterations;
  const int PR = NP / KP;                            // panel rows per buffer
  const bool idle = tid >= L.ng * L.gt;
  const int grp = idle ? 0 : tid / L.gt, lg = tid - grp * L.gt;
  const int half = SPLIT ? (lg >> 7) : 0;   // SPLIT: 0 = c thread, 1 = s thread
  const int l = SPLIT ? (lg & 127) : lg;    // lane in group == TMEM lane
  const int warp = tid >> 5;

  float* hv = smem;                                   // [NP]
  float* av = hv + NP;                                // [NP]
  float* X = av + NP + (size_t)grp * 2 * PR * XS;     // this group's [2][PR][XS] panel
  float* qtail = av + NP + (size_t)L.ng * 2 * PR * XS;  // QSRC_HYB: Qs rows 128 .. NP-1, [NP - 128][HYB_LD]
  // VSMEM: Adam second moments of the DL tile, [4 slots][256 threads] float4 (conflict-free), after the panels
  float4* vsm = reinterpret_cast<float4*>(av + NP + (size_t)L.ng * 2 * PR * XS) + (tid & 255);

  // ------------------------------------------------------------------ prologue
  const unsigned long long t_start = f.stats ? global_timer_ns() : 0ull;
  if (HAS_TMEM && warp == 0) tmem_alloc(&tmem_slot, L.tcols);
  // single-instance launches evaluate the schedule table here (every CTA its own copy: T rows of 32 B,
  // fp64, ~6 rows per thread) instead of in a kernel of its own; published by the barriers below
  const float* sched = p.sched;
  if (f.sched_inline) {
    float* mine = f.sched_scratch + (size_t)cta * p.iterations * SCHED_W;
    build_schedule_cta(f.sa, mine);
    sched = mine;
  }
  for (int j = tid; j < NP; j += blockDim.x) {
    float a = 0.f;
    if (j < N) a = p.a_half / (p.drift_s_vec ? p.drift_s_vec[j] : p.drift_s);
    av[j] = a;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = HAS_TMEM ? tmem_slot : 0u;
  for (int j = tid; j < NP; j += blockDim.x) {
    float h = 0.f;
    if (j < N) {
      float cs = 0.f;
      for (int i = 0; i < N; ++i) cs += p.q[i * N + j];
      h = -av[j] * (p.b_half * cs + p.v[j]);
    }
    hv[j] = h;
  }
  if (!idle)
    for (int i = lg; i < 2 * PR * XS; i += L.gt) X[i] = 0.f;
  if constexpr (VSMEM) {
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) vsm[sl * 256] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  const int rg = l % RG, cg = l / RG;
  const bool active = cg < CG;
  const int cgc = active ? cg : 0;
  const int j0 = 4 * cgc;
  const uint32_t tlane = tbase + ((uint32_t)((l >> 5) * 32) << 16);  // this warp's TMEM lane quadrant
  if constexpr (QSRC == QSRC_HYB) {
    for (int idx = tid; idx < (NP - 4 * HYB_TMEM_CHUNKS) * HYB_LD; idx += blockDim.x) {
      const int k = 4 * HYB_TMEM_CHUNKS + idx / HYB_LD, j = idx % HYB_LD;
      qtail[idx] = (k < N && j < N) ? -av[k] * av[j] * p.q[k * N + j] : 0.f;
    }
  }
  if (HAS_TMEM && !idle && grp == 0 && half == 0) {
    // every lane stores its own copy of the 4 columns it contracts against
    const int ktm = QSRC == QSRC_HYB ? 4 * HYB_TMEM_CHUNKS : NP;
    for (int k = 0; k < ktm; ++k) {
      float qv[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj;
        qv[jj] = (k < N && j < N) ? -av[k] * av[j] * p.q[k * N + j] : 0.f;
      }
      tmem_st4(tlane + 4 * k, qv[0], qv[1], qv[2], qv[3]);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (!idle) {
  // ------------------------------------------------------------------ thread tile
  const long long gb0 = ((long long)cta * L.ng + grp) * (2 * RG) + 2 * rg;  // first of 2 trajectories
  float hreg[4], sclamp[4];
  bool colok[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    colok[jj] = active && (j0 + jj < N);
    hreg[jj] = hv[j0 + jj];
    sclamp[jj] = colok[jj] ? (p.clamp_s_vec ? p.clamp_s_vec[j0 + jj] : p.clamp_s) : 0.f;
  }

  pf2 st[2][4];             // st[0] = c | mu, st[1] = s | sigma  (x: trajectory gb0, y: gb0+1)
  pf2 am[KT][4], avv[KT][4];  // Adam moments of the tracked arrays
  pf2 W[KT][4];               // noise of the current iteration
  pf2 Wn[KT][4];              // PIPE + MF: noise of the NEXT iteration (its measurement), drawn during the drift
  pf2 meas[4];              // MF: clamped measurement
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    st[0][jj] = dup(0.f);
    st[1][jj] = dup(SOLVER == SOLVER_MF ? 0.5f : 0.f);
    meas[jj] = dup(0.f);
#pragma unroll
    for (int q = 0; q < KT; ++q) {
      am[q][jj] = dup(0.f);
      avv[q][jj] = dup(0.f);
      W[q][jj] = dup(0.f);
      Wn[q][jj] = dup(0.f);
    }
  }

  const uint2 key = make_uint2(p.seed_lo, p.seed_hi ^ p.off_hi);
#if CCVM_SIMT_RNG
  // stream mode: one xoshiro128+ state per thread = per (global trajectory pair, column group); the
  // quanta of an iteration are drawn in the fixed order (quadrature, trajectory of the pair)
  NoiseStream rs = stream_init(key.x, key.y, p.off_lo, (unsigned long long)(p.traj_base + gb0) >> 1, (uint32_t)cgc);
#endif

  // one noise quantum: the four columns of (quadrature q, trajectory i) at iteration t
  auto quantum = [&](pf2 (&Wd)[KT][4], int q, int i, int t) {
    float n0, n1, n2, n3;
#if CCVM_SIMT_RNG
    (void)t;
    stream_normals4(rs, n0, n1, n2, n3);
#else
    const unsigned long long gb = (unsigned long long)(p.traj_base + gb0 + i);
    const uint32_t qi = (uint32_t)(q + half);  // which quadrature's stream
    noise_normals4(key.x, key.y, p.off_lo, gb, (uint32_t)t, (uint32_t)cgc, qi, n0, n1, n2, n3);
#endif
    // Padding columns (j >= n) draw noise like any other -- masking it cost a SEL per normal; their
    // state stays finite (see `stage`) and is never written out.
    const float nn[4] = {n0, n1, n2, n3};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float w = (DENSE_TILE && !colok[jj]) ? 0.f : nn[jj];
      if (i) Wd[q][jj].y = w; else Wd[q][jj].x = w;
    }
  };

  auto draw = [&](int t) {
    if (PIPE || p.noise == nullptr) {
#pragma unroll
      for (int q = 0; q < KT; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i) quantum(W, q, i, t);
    } else {
#pragma unroll
      for (int q = 0; q < KT; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float w = 0.f;
            const long long b = gb0 + i;
            if (colok[jj] && b < p.batch)
              w = p.noise[(((size_t)t * K + (q + half)) * N + (j0 + jj)) * (size_t)p.noise_batch +
                          (size_t)(p.traj_base + b)];
            if (i) W[q][jj].y = w; else W[q][jj].x = w;
          }
    }
  };

  // stores the tile's contraction input for the next iteration: X[buf][k = j0+jj][rg][2K]
  const int xoff = (cgc * RW * RG) & L.xmask;
  auto stage = [&](int buf, const pf2 (&a)[4], const pf2 (&b)[4]) {
    if (!active) return;
    if constexpr (KP == 2) {
      // the tile's rows j0 .. j0+3 are two k pairs: two STS.128
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        float* dst = X + ((size_t)buf * PR + (j0 >> 1) + jp) * XS + 4 * rg;
        *reinterpret_cast<float4*>(dst) = make_float4(a[2 * jp].x, a[2 * jp].y, a[2 * jp + 1].x, a[2 * jp + 1].y);
      }
      return;
    }
    // Padding columns (j >= n) are staged like any other: their state only ever meets the zero rows of Qs.
    // That needs it to stay FINITE (Inf * 0 would poison the pair): it sees no drift, only the solver's own
    // saturating terms (DL, MF: the cubic -- a step size that blows it up blows up the real columns as well)
    // or a zero clamp (Langevin, PumpedLangevin).  Staging it behind a predicate was measured: 1-5 % slower
    // on every loop (profiles/r2m_tuning_x32_kp2_padstage.txt).
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float* dst = X + ((size_t)buf * NP + (j0 + jj)) * XS + xoff + RW * rg + 2 * half;
      if constexpr (K == 2 && !SPLIT) *reinterpret_cast<float4*>(dst) = make_float4(a[jj].x, a[jj].y, b[jj].x, b[jj].y);
      else *reinterpret_cast<float2*>(dst) = make_float2(a[jj].x, a[jj].y);
    }
  };

  const int bar_id = 1 + grp, bar_n = L.gt;
  // plain loads: the table may have been written by this CTA (sched_inline)
  const float4* sched4 = reinterpret_cast<const float4*>(sched);
  float4 sa = sched4[0], sb = sched4[1];

  if constexpr (SOLVER == SOLVER_MF) {
    draw(0);  // measurement of iteration 0 (mf_solver.py:551-554): mu = 0
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      meas[jj] = clamp2(fma2(dup(sa.x), W[0][jj], st[0][jj]), -sclamp[jj], sclamp[jj]);
    group_barrier(bar_id, bar_n);  // zero fill of the panel done
    stage(0, meas, meas);
  }
  group_barrier(bar_id, bar_n);

  // ------------------------------------------------------------------ main loop
  if ((grp & 1) && L.phase_ns > 0) __nanosleep(L.phase_ns);
  for (int t = 0; t < T; ++t) {
    const int buf = t & 1;
    const float4 ca = sa, cb = sb;
    if (t + 1 < T) {
      sa = sched4[2 * (t + 1)];
      sb = sched4[2 * (t + 1) + 1];
    }

    // ---- drift contraction: acc = h + X . Qs   (Qs from TMEM, X from shared memory)
    pf2 acc[KT][4];
#pragma unroll
    for (int q = 0; q < KT; ++q)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[q][jj] = dup(hreg[jj]);
    // PIPE: everything of the SDE step that does not depend on the drift is evaluated INSIDE the
    // contraction (called right after the unrolled noise chunks, i.e. in their straight-line block,
    // where ptxas interleaves it with the FFMA2 stream): the state arrays then hold the step's
    // drift-free part and what is left after the last FFMA2 is one FFMA2 per element pair (plus
    // Adam, which needs the finished gradient) -- the latency-bound tail between contraction and
    // barrier shrinks to almost nothing, and the noise registers die before the run-time loop.
    auto precompute = [&]() {
      if constexpr (SOLVER == SOLVER_DL) {
        const pf2 d1 = dup(ca.y), d2 = dup(ca.z), n1 = dup(ca.w), n2 = dup(cb.x);
        const pf2 mdt = dup(-p.dt), half = dup(0.5f);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 c = st[0][jj], s = st[1][jj];
          const pf2 r2 = fma2(c, c, mul2(s, s));
          const pf2 rt = sqrt2(add2(r2, half));
          const pf2 uc = fma2(r2, mdt, d1), us = fma2(r2, mdt, d2);
          st[0][jj] = add2(c, fma2(c, uc, mul2(mul2(rt, n1), W[0][jj])));
          st[1][jj] = add2(s, fma2(s, us, mul2(mul2(rt, n2), W[1][jj])));
        }
      } else if constexpr (SOLVER == SOLVER_MF) {
        const pf2 pr = dup(ca.y), sj = dup(ca.w), opj = dup(cb.x), m2j = dup(-2.f * ca.z);
        const pf2 g2 = dup(p.g2), dtp = dup(p.dt), mhalf = dup(-0.5f);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 mu = st[0][jj], sg = st[1][jj];
          const pf2 g2m2 = mul2(mul2(mu, mu), g2);
          const pf2 a1 = fma2(g2m2, dup(-1.f), pr);
          const pf2 sh = add2(sg, mhalf);
          const pf2 diff = mul2(mul2(sh, sj), W[0][jj]);
          st[0][jj] = fma2(dtp, fma2(a1, mu, diff), mu);
          const pf2 a3 = fma2(g2m2, dup(-3.f), pr);
          const pf2 t1 = mul2(mul2(a3, sg), dup(2.f));
          const pf2 t2 = mul2(mul2(sh, sh), m2j);
          const pf2 t3 = fma2(g2m2, dup(2.f), opj);
          st[1][jj] = fma2(dtp, add2(add2(t1, t2), t3), sg);
        }
      } else {
        const pf2 sig = dup(p.sig), mdt = dup(-p.dt), d1 = dup(ca.y);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 c = st[0][jj];
          pf2 pre = fma2(sig, W[0][jj], c);
          if constexpr (SOLVER == SOLVER_PLV) pre = fma2(c, fma2(mul2(c, c), mdt, d1), pre);
          st[0][jj] = pre;
        }
      }
    };
    if constexpr (XSC != 0) {
      // TMEM + PIPE: the Q chunk AND the four state rows it meets are both fetched one chunk ahead
      // into ping-pong registers (tcgen05.ld / LDS in flight under the previous chunk's FFMA2s);
      // all shared-memory addresses are xp + immediate, xp advances once per chunk pair.
      // one panel row of a pair: (c0,c1,s0,s1), (k even b0,b1, k odd b0,b1), or (b0,b1)
      typedef typename std::conditional<KT * KP == 2, float4, float2>::type XV;
      constexpr int XR = 4 / KP;     // panel rows per chunk of four k
      constexpr int ROWB = XSC * 4;  // bytes per panel row
      const char* xp = reinterpret_cast<const char*>(X + (size_t)buf * PR * XSC + RW * KP * rg);
      float qa[16], qb[16];
      XV xa[XR], xb[XR];
      // `row` counts k rows (0, 4, 8) relative to xp
      auto load_x = [&](int row, XV (&dst)[XR]) {
#pragma unroll
        for (int kk = 0; kk < XR; ++kk) dst[kk] = *reinterpret_cast<const XV*>(xp + (row / KP + kk) * ROWB);
      };
      auto contract = [&](const float (&qq)[16], const XV (&xx)[XR]) -> uint32_t {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          pf2 xv[KT];
          if constexpr (KT == 2) {
            xv[0] = pk(xx[kk].x, xx[kk].y);
            xv[1] = pk(xx[kk].z, xx[kk].w);
          } else {
            if constexpr (KP == 2) xv[0] = (kk & 1) ? pk(xx[kk / 2].z, xx[kk / 2].w) : pk(xx[kk / 2].x, xx[kk / 2].y);
            else xv[0] = pk(xx[kk].x, xx[kk].y);
          }
#pragma unroll
          for (int q = 0; q < KT; ++q)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[q][jj] = fma2(xv[q], dup(qq[4 * kk + jj]), acc[q][jj]);
        }
        return __float_as_uint(xx[XR - 1].x);
      };
      // the last chunk of the contraction without its padding rows (KTAIL)
      constexpr int LASTK = (KTAIL > 0 && KP == 1 && CGC != 0) ? KTAIL : 4;
      auto load_x_last = [&](int row, XV (&dst)[XR]) {
#pragma unroll
        for (int kk = 0; kk < (LASTK < XR ? LASTK : XR); ++kk) dst[kk] = *reinterpret_cast<const XV*>(xp + (row / KP + kk) * ROWB);
      };
      auto contract_last = [&](const float (&qq)[16], const XV (&xx)[XR]) {
        if constexpr (LASTK == 4) {
          contract(qq, xx);
        } else {
#pragma unroll
          for (int kk = 0; kk < LASTK; ++kk) {
            pf2 xv[KT];
            xv[0] = pk(xx[kk].x, xx[kk].y);
            if constexpr (KT == 2) xv[1] = pk(xx[kk].z, xx[kk].w);
#pragma unroll
            for (int q = 0; q < KT; ++q)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) acc[q][jj] = fma2(xv[q], dup(qq[4 * kk + jj]), acc[q][jj]);
          }
        }
      };
      // Small compile-time column-group counts (CG = 5, 8 <-> N = 20, 30): the contraction is too short
      // to hide a noise quantum per chunk pair -- pinned there, the quanta ran one after the other and
      // their ~250-cycle dependent chains WERE the iteration.  All of them start at the top instead,
      // unpinned: ptxas runs the chains side by side under the (fully unrolled) contraction.
      // The same holds, measured, for every larger compile-time variant (N = 40 ... 70: DL + Adam 3.57 ->
      // 3.39 ms at N = 70 and +12-15 % at N = 40 ... 60, Langevin +4-15 %, its Adam variant +6 %) except
      // MF at CG = 15, 18, which keeps one pinned quantum per chunk pair (MF + Adam loses 12 % unpinned).
      constexpr bool SMALLCG = CGC != 0 && (CGC <= 13 || (CCVM_UNPIN_MASK & TILE_BIT) != 0);
      constexpr int NQ = SMALLCG ? 2 : 2 * KT;
      const int tn = SOLVER == SOLVER_MF ? t + 1 : t;
      tmem_ld16(tlane, qa);
      load_x(0, xa);
      int kc = 0;
      if constexpr (SMALLCG) {
#pragma unroll
        for (int q = 0; q < KT; ++q)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if constexpr (SOLVER == SOLVER_MF) quantum(Wn, q, i, tn);
            else quantum(W, q, i, tn);
          }
      }
#pragma unroll
      for (int u = 0; u < NQ; ++u) {  // CG > 2*NQ: chunks 0 .. 2*NQ exist
        tmem_wait_ld();
        tmem_ld16(tlane + 16 * (2 * u + 1), qb);
        load_x(4, xb);
        // p.pin_mask is 0 at run time: the generator's counter formally depends on a state value
        // consumed in THIS pair of chunks, which keeps ptxas from hoisting all of the noise work
        // to the top of the iteration (it did) and spreads it over the contraction instead.
        const uint32_t pin = contract(qa, xa) & p.pin_mask;
        tmem_wait_ld();
        tmem_ld16(tlane + 16 * (2 * u + 2), qa);
        load_x(8, xa);
        contract(qb, xb);
        xp += (8 / KP) * ROWB;
        if constexpr (!SMALLCG) {
#if CCVM_SIMT_RNG
          rs.s0 ^= pin;  // (0) the formal dependence that pins this quantum to this pair of chunks
#endif
          if constexpr (SOLVER == SOLVER_MF) quantum(Wn, u >> 1, u & 1, tn ^ (int)pin);
          else quantum(W, u >> 1, u & 1, tn ^ (int)pin);
        }
      }
      if constexpr (HOIST) precompute();
      kc = 2 * NQ;
      if constexpr (QSRC == QSRC_HYB) {
        // chunks 2*NQ .. 31 from TMEM; the last prefetch of this part already comes from the tail
        const char* qp = reinterpret_cast<const char*>(qtail + j0);
        auto lds_q = [&](int row, float (&dst)[16]) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 v = *reinterpret_cast<const float4*>(qp + (row + kk) * (HYB_LD * 4));
            dst[4 * kk] = v.x; dst[4 * kk + 1] = v.y; dst[4 * kk + 2] = v.z; dst[4 * kk + 3] = v.w;
          }
        };
#pragma unroll((SOLVER == SOLVER_DL && !ADAM) ? 1 : 4)
        for (; kc + 2 < HYB_TMEM_CHUNKS; kc += 2) {
          tmem_wait_ld();
          tmem_ld16(tlane + 16 * (kc + 1), qb);
          load_x(4, xb);
          contract(qa, xa);
          tmem_wait_ld();
          tmem_ld16(tlane + 16 * (kc + 2), qa);
          load_x(8, xa);
          contract(qb, xb);
          xp += (8 / KP) * ROWB;
        }
        {  // kc == 30: chunk 31 is the last one in TMEM, chunk 32 the first of the tail (CG > 32)
          tmem_wait_ld();
          tmem_ld16(tlane + 16 * (kc + 1), qb);
          load_x(4, xb);
          contract(qa, xa);
          tmem_wait_ld();
          lds_q(0, qa);
          load_x(8, xa);
          contract(qb, xb);
          xp += (8 / KP) * ROWB;
          kc += 2;
        }
#pragma unroll((SOLVER == SOLVER_DL && !ADAM) ? 1 : 4)
        for (; kc + 2 < CG; kc += 2) {  // steady state; the last one or two chunks are peeled off
          lds_q(4, qb);
          load_x(4, xb);
          contract(qa, xa);
          lds_q(8, qa);
          load_x(8, xa);
          contract(qb, xb);
          xp += (8 / KP) * ROWB;
          qp += 8 * (HYB_LD * 4);
        }
        if (kc + 2 == CG) {
          lds_q(4, qb);
          load_x(4, xb);
          contract(qa, xa);
          contract(qb, xb);
        } else {
          contract(qa, xa);
        }
      } else {
        // steady state: both prefetches unconditional; the last one or two chunks are peeled off
        if constexpr (DENSE_TILE) {
          for (; kc + 2 <= CG; kc += 2) {
            tmem_wait_ld();
            tmem_ld16(tlane + 16 * (kc + 1), qb);
            load_x(4, xb);
            contract(qa, xa);
            tmem_wait_ld();
            if (kc + 2 < CG) {
              tmem_ld16(tlane + 16 * (kc + 2), qa);
              load_x(8, xa);
            }
            contract(qb, xb);
            xp += (8 / KP) * ROWB;
          }
          if (kc < CG) {
            tmem_wait_ld();
            contract(qa, xa);
          }
        } else {
        // run-time column-group count: four chunk pairs per trip for every tile but plain DL
        // (fewer branches and pointer updates: +2-6 % at N = 70..128; DL itself loses 2-5 % to it)
#pragma unroll(CGC ? 32 : (SOLVER == SOLVER_DL && !ADAM) ? 1 : 4)
        for (; kc + 2 < CG; kc += 2) {
          tmem_wait_ld();
          tmem_ld16(tlane + 16 * (kc + 1), qb);
          load_x(4, xb);
          contract(qa, xa);
          tmem_wait_ld();
          tmem_ld16(tlane + 16 * (kc + 2), qa);
          load_x(8, xa);
          contract(qb, xb);
          xp += (8 / KP) * ROWB;
        }
        tmem_wait_ld();
        if (kc + 2 == CG) {
          tmem_ld16(tlane + 16 * (kc + 1), qb);
          load_x_last(4, xb);
          contract(qa, xa);
          tmem_wait_ld();
          contract_last(qb, xb);
        } else {
          contract_last(qa, xa);
        }
        }
      }
    } else {
      // four k's against one 16-column TMEM chunk; two chunk buffers ping-pong so that the next
      // tcgen05.ld is in flight while the current chunk is consumed (no register copies)
      const float* xrow = X + (size_t)buf * NP * XS + RW * rg + 2 * half;
      int off = 0;
      // returns the bits of the last state value it loaded (scheduling pin of the PIPE variant)
      auto contract4 = [&](const float (&qq)[16]) -> uint32_t {
        const float* xr = xrow + off;
        uint32_t last = 0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          pf2 xv[KT];
          if constexpr (KT == 2) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + kk * XS);
            xv[0] = pk(x4.x, x4.y);
            xv[1] = pk(x4.z, x4.w);
          } else {
            const float2 x2 = *reinterpret_cast<const float2*>(xr + kk * XS);
            xv[0] = pk(x2.x, x2.y);
          }
          last = __float_as_uint(xv[0].x);
#pragma unroll
          for (int q = 0; q < KT; ++q)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[q][jj] = fma2(xv[q], dup(qq[4 * kk + jj]), acc[q][jj]);
        }
        xrow += 4 * XS;
        off = (off + RW * RG) & L.xmask;
        return last;
      };
      float qa[16], qb[16];
      // chunk c = rows 4c..4c+3 of the thread's 4 columns
      const float* qg = QSRC == QSRC_GMEM ? L.qs + j0 : nullptr;
      auto load_chunk = [&](int c, float (&dst)[16]) {
        if constexpr (QSRC == QSRC_TMEM) {
          tmem_ld16(tlane + 16 * c, dst);
        } else if constexpr (QSRC == QSRC_HYB) {
          if (c < HYB_TMEM_CHUNKS) {
            tmem_ld16(tlane + 16 * c, dst);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float4 v = *reinterpret_cast<const float4*>(
                  qtail + (size_t)(4 * (c - HYB_TMEM_CHUNKS) + kk) * HYB_LD + j0);
              dst[4 * kk] = v.x; dst[4 * kk + 1] = v.y; dst[4 * kk + 2] = v.z; dst[4 * kk + 3] = v.w;
            }
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(qg + (size_t)(4 * c + kk) * NP));
            dst[4 * kk] = v.x; dst[4 * kk + 1] = v.y; dst[4 * kk + 2] = v.z; dst[4 * kk + 3] = v.w;
          }
        }
      };
      auto wait_chunk = [&]() {
        if constexpr (HAS_TMEM) tmem_wait_ld();
      };
      load_chunk(0, qa);
      int kc = 0;
      if constexpr (PIPE) {
        // CG > 2*NQ (checked by the host): chunks 0 .. 2*NQ exist, no guards in the unrolled part
        constexpr int NQ = 2 * KT;
        const int tn = SOLVER == SOLVER_MF ? t + 1 : t;
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          wait_chunk();
          load_chunk(2 * u + 1, qb);
          // p.pin_mask is 0 at run time: the generator's counter formally depends on a state value
          // loaded in THIS pair of chunks, which keeps ptxas from hoisting all of the noise work
          // to the top of the iteration (it did) and spreads it over the contraction instead.
          const uint32_t pin = contract4(qa) & p.pin_mask;
          wait_chunk();
          load_chunk(2 * u + 2, qa);
          contract4(qb);
#if CCVM_SIMT_RNG
          rs.s0 ^= pin;
#endif
          if constexpr (SOLVER == SOLVER_MF) quantum(Wn, u >> 1, u & 1, tn ^ (int)pin);
          else quantum(W, u >> 1, u & 1, tn ^ (int)pin);
        }
        if constexpr (HOIST) precompute();
        kc = 2 * NQ;
      }
      for (; kc + 2 < CG; kc += 2) {  // steady state; the last one or two chunks are peeled off
        wait_chunk();
        load_chunk(kc + 1, qb);
        contract4(qa);
        wait_chunk();
        load_chunk(kc + 2, qa);
        contract4(qb);
      }
      if (kc < CG) {
        wait_chunk();
        if (kc + 2 == CG) {
          load_chunk(kc + 1, qb);
          contract4(qa);
          wait_chunk();
          contract4(qb);
        } else {
          contract4(qa);
        }
      }
    }

    // ---- elementwise SDE step
    if constexpr (SOLVER == SOLVER_DL) {
      if constexpr (!PIPE) draw(t);
      if constexpr (ADAM) {
        if constexpr (VSMEM) {
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 tv = vsm[(q * 2 + h) * 256];
              avv[q][2 * h] = pk(tv.x, tv.y);
              avv[q][2 * h + 1] = pk(tv.z, tv.w);
            }
        }
        adam_tile4(acc[0], am[0], avv[0], p, cb.y, cb.z);
        adam_tile4(acc[1], am[1], avv[1], p, cb.y, cb.z);
        if constexpr (VSMEM) {
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              vsm[(q * 2 + h) * 256] = make_float4(avv[q][2 * h].x, avv[q][2 * h].y, avv[q][2 * h + 1].x, avv[q][2 * h + 1].y);
        }
      }
      // every DL variant evaluates the step as  (drift-free part) + gain * drift  with the drift-free
      // part from `precompute` -- inside the contraction where it is hoisted, here otherwise -- so that
      // all of them (in-loop noise or not, single or batched launch) round identically
      if constexpr (!HOIST) precompute();
      const pf2 gain = dup(ca.x);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        st[0][jj] = fma2(gain, acc[0][jj], st[0][jj]);
        st[1][jj] = fma2(gain, acc[1][jj], st[1][jj]);
      }
      stage(buf ^ 1, st[0], st[1]);
    } else if constexpr (SOLVER == SOLVER_MF) {
      const pf2 fs = dup(p.fs);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[0][jj] = mul2(fs, acc[0][jj]);
      if constexpr (ADAM) adam_tile4(acc[0], am[0], avv[0], p, cb.y, cb.z);
      const pf2 pr = dup(ca.y), sj = dup(ca.w), opj = dup(cb.x), m2j = dup(-2.f * ca.z);
      const pf2 g2 = dup(p.g2), dtp = dup(p.dt), mhalf = dup(-0.5f);
      if constexpr (HOIST) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) st[0][jj] = fma2(dtp, acc[0][jj], st[0][jj]);
      } else {
        // mf_solver.py:158-233 with the constants folded (14 packed operations per element pair instead of 18:
        // the loop is issue-bound, every instruction counts):
        //   dmu = (p - g^2 mu^2) mu + fs grads + (sigma - 1/2) sqrt(j)/sqrt(dt) W
        //   dsg = 2 (p - 3 g^2 mu^2) sigma - 2 j (sigma - 1/2)^2 + (1 + j) + 2 g^2 mu^2
        const pf2 ng2 = dup(-p.g2), n3g2 = dup(-3.f * p.g2), p2g2 = dup(2.f * p.g2), two = dup(2.f);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 mu = st[0][jj], sg = st[1][jj];
          const pf2 m2 = mul2(mu, mu);
          const pf2 a1 = fma2(m2, ng2, pr);
          const pf2 sh = add2(sg, mhalf);
          const pf2 dmu = fma2(a1, mu, acc[0][jj]);
          st[0][jj] = fma2(dtp, fma2(mul2(sh, W[0][jj]), sj, dmu), mu);
          const pf2 a3 = fma2(m2, n3g2, pr);
          const pf2 t3 = fma2(m2, p2g2, opj);
          const pf2 inner = fma2(mul2(sh, sh), m2j, t3);
          st[1][jj] = fma2(dtp, fma2(mul2(a3, sg), two, inner), sg);
        }
      }
      if (t + 1 < T) {
        if constexpr (PIPE) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) W[0][jj] = Wn[0][jj];
        } else {
          draw(t + 1);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          meas[jj] = clamp2(fma2(dup(sa.x), W[0][jj], st[0][jj]), -sclamp[jj], sclamp[jj]);
        stage(buf ^ 1, meas, meas);
      }
    } else {
      if constexpr (!PIPE) draw(t);
      if constexpr (ADAM) adam_tile4(acc[0], am[0], avv[0], p, cb.y, cb.z);
      const pf2 dtfs = dup(p.dtfs), sig = dup(p.sig), mdt = dup(-p.dt), d1 = dup(ca.y);
      if constexpr (HOIST) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          st[0][jj] = clamp2(fma2(dtfs, acc[0][jj], st[0][jj]), -sclamp[jj], sclamp[jj]);
      } else
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const pf2 c = st[0][jj];
        pf2 inc = fma2(dtfs, acc[0][jj], mul2(sig, W[0][jj]));
        if constexpr (SOLVER == SOLVER_PLV) inc = fma2(c, fma2(mul2(c, c), mdt, d1), inc);
        st[0][jj] = clamp2(add2(c, inc), -sclamp[jj], sclamp[jj]);
      }
      stage(buf ^ 1, st[0], st[0]);
    }

    // ---- optional evolution snapshot (dl_solver.py:557-564)
    if (p.evolution_step > 0) {
      int sidx = -1;
      if (t % p.evolution_step == 0) sidx = t / p.evolution_step;
      else if (t + 1 >= T) sidx = (T - 1) / p.evolution_step + 1;
      if (sidx >= 0 && sidx < p.num_samples && active) {
        constexpr int NS = SolverTraits<SOLVER>::NSTATE;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const long long b = gb0 + i;
              if (b < p.batch && colok[jj] && (!SPLIT || a == 0))
                p.samples[(((size_t)(a + half) * p.num_samples + sidx) * p.batch + b) * N + j0 + jj] =
                    i ? st[a][jj].y : st[a][jj].x;
            }
      }
    }
    group_barrier(bar_id, bar_n);
  }

  // ------------------------------------------------------------------ results
  if (active) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long b = gb0 + i;
      if (b >= p.batch) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (!colok[jj]) continue;
        const size_t o = (size_t)b * N + j0 + jj;
        const float v0 = i ? st[0][jj].y : st[0][jj].x, v1 = i ? st[1][jj].y : st[1][jj].x;
        const float vm = i ? meas[jj].y : meas[jj].x;
        if constexpr (SOLVER == SOLVER_DL) {
          p.out0[o] = clampf(v0, -sclamp[jj], sclamp[jj]);
          p.out1[o] = v1;
        } else if constexpr (SOLVER == SOLVER_MF) {
          p.out0[o] = v0;
          p.out1[o] = vm;
          p.out2[o] = v1;
        } else {
          p.out0[o] = v0;
        }
      }
    }
  }
  }  // !idle
  // ------------------------------------------------------------------ fused tail of Solver.__call__
  // change of variables -> post-processor -> energy on the CTA's own trajectories (the stand-alone
  // epilogue's code, so the values are identical), then best / argmin / success counters merged across
  // CTAs with a handful of atomics: the whole call is this one launch (dl_solver.py:936-959,
  // grad_descent.py:58-64, adam.py:58-66, solution.py:125-136).
  if (f.epilogue) {
    __syncthreads();  // every group has written its outputs
    const unsigned long long t_loop = f.stats ? global_timer_ns() : 0ull;
    const long long per_cta = (long long)L.ng * 2 * L.rg;
    const long long b_begin = (long long)cta * per_cta;
    const long long b_end = b_begin + per_cta < p.batch ? b_begin + per_cta : p.batch;
    epilogue_run(f.epi, smem, b_begin, b_end, 0, 1);
    if (f.stats) {
      __syncthreads();  // the CTA's energies are in global memory
      StatsPartial sp;
      if (stats_block_reduce(f.epi.energy, b_begin, b_end, f.optimal, sp))
        stats_merge(sp, f.accum, f.out, f.total_ctas, t_loop - t_start, global_timer_ns() - t_loop);
    }
  }
  if constexpr (HAS_TMEM) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, L.tcols);
  }
}

template <int SOLVER, bool ADAM, int QSRC, bool PIPE, int CGC = 0, int KTAIL = 0>
__global__ void __launch_bounds__(QSRC == QSRC_GMEM ? 512 : 256, 1)
    sde_tmem_kernel(const SdeParams p, const TmemLaunch L, const FusedTail f) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint32_t tmem_slot;
  sde_tile_body<SOLVER, ADAM, QSRC, PIPE, CGC, KTAIL>(p, L, f, blockIdx.x, smem, &tmem_slot);
}

// PIPE needs Philox noise and more than 4*K chunks of four Q rows (DL: n >= 33, others: n >= 17).
template <int SOLVER>
__host__ __device__ __forceinline__ bool pipe_ok(int cg, bool philox) {
  return philox && cg > 4 * SolverTraits<SOLVER>::K;
}

// One launch over MANY problem instances (BatchItem, sde_launch.h): each CTA looks up (instance, CTA
// index inside the instance) and runs the same body with that instance's parameters.  CGC != 0: every
// instance of the bucket has that column-group count (compile-time variants of the PIPE kernels).
template <int SOLVER, bool ADAM, int QSRC, int CGC = 0, int KTAIL = 0>
__global__ void __launch_bounds__(256, 1)
    sde_tmem_batch_kernel(const BatchItem* __restrict__ items, const int2* __restrict__ cta_map) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ BatchItem s_item;
  const int2 m = cta_map[blockIdx.x];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(items + m.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_item);
    for (int i = threadIdx.x; i < (int)(sizeof(BatchItem) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const SdeParams p = s_item.p;
  const TmemLaunch L = s_item.L;
  if constexpr (CGC != 0) {
    sde_tile_body<SOLVER, ADAM, QSRC, true, CGC, KTAIL>(p, L, s_item.f, m.y, smem, &tmem_slot);
  } else {
    if (L.pipe)
      sde_tile_body<SOLVER, ADAM, QSRC, true>(p, L, s_item.f, m.y, smem, &tmem_slot);
    else
      sde_tile_body<SOLVER, ADAM, QSRC, false>(p, L, s_item.f, m.y, smem, &tmem_slot);
  }
}

}  // namespace ccvm
