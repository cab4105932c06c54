// Persistent Euler-Maruyama kernel, tensor-core variant (n >= 256, >= 1024 contraction rows):
// the drift  G = X . Qs  is a genuine dense GEMM there (BASELINE config 4: n = 1024, batch 8192),
// so it runs on the 5th-generation tensor cores as a 3xTF32 product
//
//        X . Qs  ~=  Xlo . Qhi  +  Xhi . Qlo  +  Xhi . Qhi        (hi = tf32(x), lo = x - hi)
//
// with FP32 accumulation in TMEM (tcgen05.mma kind::tf32, cta_group::1, M = 128, N = 256, K = 8).
// Reference loops: dl_solver.py:468-769, mf_solver.py:493-764, langevin_solver.py:368-561,
// pumped_langevin_solver.py:232-449 (the einsum "bi,ij->bj" + the elementwise SDE step).
//
// Decomposition.  The contraction rows are (trajectory, quadrature) pairs: DL interleaves them as
// row 2b = c_b, row 2b+1 = s_b (neighbouring lanes of one warp, so c^2 + s^2 is one shuffle); the
// other solvers have one row per trajectory.  A CTA owns 128 rows for ALL iterations -- rows are
// independent, there is no inter-CTA exchange.  Per iteration and 256-column output chunk:
//
//   warp 8 (one lane)  TMA producer: streams 16-wide k-blocks of the CTA's state rows (hi and lo,
//                      128 x 16) and of Qs^T (hi and lo, 256 x 16) into a 4-stage shared-memory ring
//                      (SWIZZLE_64B, K-major), 48 KB per stage;
//   warp 9 (one lane)  MMA issuer: 2 k-steps x 3 split terms per stage into one of two 128 x 256
//                      FP32 accumulators in TMEM; tcgen05.commit releases the stage / publishes
//                      the accumulator;
//   warps 0-7          epilogue: tcgen05.ld the accumulator (row = lane, 16 columns at a time),
//                      add the affine drift term, draw the Philox / replayed noise, apply the
//                      solver's SDE step in FP32 and write the new state as (hi, lo) into the
//                      OTHER ping-pong buffer -- overlapped with the MMAs of the next chunk.
//
// The state does not fit on chip at these sizes (DL, n = 1024: 8 KB per trajectory), so it lives
// in global memory (L2 / HBM) as two exact FP32 summands hi + lo, which are at the same time the
// two tensor-core operands; Qs^T hi / lo (8 MB at n = 1024) stays L2 resident.  A k-block of
// iteration t+1 only depends on the epilogue of the output chunk that covers its columns, so the
// producer waits per chunk (state_ready[c]) and the pipeline never drains between iterations.
#pragma once
#include <cuda.h>

#include "ccvm_common.cuh"
#include "sde_kernel_tmem.cuh"

namespace ccvm {

constexpr int TC_BM = 128;       // rows per CTA (= UMMA M)
constexpr int TC_BN = 256;       // output columns per accumulator (= UMMA N)
constexpr int TC_BK = 16;        // k-block: 16 floats = 64 B = one SWIZZLE_64B row
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_WARPS = 8;  // warps 0-7; warp 8 = TMA producer, warp 9 = MMA issuer
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 8 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   // 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // 48 KB
constexpr int TC_MAX_CHUNKS = 8;  // n <= 2048

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug must trap (a reported launch failure) rather than hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long start = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spins & 0xfff) == 0xfff) {
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_64B, rows of 64 B, 8-row atoms of 512 B
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;                    // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                             // layout type: SWIZZLE_64B
  return d;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---------------------------------------------------------------- one-off preparation
// Qs^T split:  qt_hi[j][k] + qt_lo[j][k] = Qs[k][j] = -alpha_k alpha_j Q[k][j]  (zero padded), and the
// per-column vectors h_j = -alpha_j ((u+l)/2 colsum_j(Q) + V_j), S_j.
static __global__ void tc_prepare_q_kernel(const float* __restrict__ q, const float* __restrict__ v,
                                    const float* __restrict__ drift_s_vec, float drift_s,
                                    const float* __restrict__ clamp_s_vec, float clamp_s, float a_half, float b_half,
                                    int n, int np, float* __restrict__ qt_hi, float* __restrict__ qt_lo,
                                    float* __restrict__ hvec, float* __restrict__ svec) {
  const int j = blockIdx.x;  // output column
  const float aj = j < n ? a_half / (drift_s_vec ? drift_s_vec[j] : drift_s) : 0.f;
  float cs = 0.f;
  for (int k = threadIdx.x; k < np; k += blockDim.x) {
    float val = 0.f;
    if (k < n && j < n) {
      const float qkj = q[(size_t)k * n + j];
      cs += qkj;
      const float ak = a_half / (drift_s_vec ? drift_s_vec[k] : drift_s);
      val = -ak * aj * qkj;
    }
    const float hi = tf32_rna(val);
    qt_hi[(size_t)j * np + k] = hi;
    qt_lo[(size_t)j * np + k] = val - hi;
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cs;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    hvec[j] = j < n ? -aj * (b_half * tot + v[j]) : 0.f;
    svec[j] = j < n ? (clamp_s_vec ? clamp_s_vec[j] : clamp_s) : 0.f;
  }
}

// Philox normals of (trajectory gb, iteration t, column group cg, quadrature qi): the SAME stream
// the SIMT kernels draw (sde_kernel_tmem.cuh `draw`), so both paths see identical noise.
__device__ __forceinline__ void tc_philox4(const SdeParams& p, unsigned long long gb, int t, int cg, uint32_t qi,
                                           float (&n)[4]) {
  noise_normals4(p.seed_lo, p.seed_hi ^ p.off_hi, p.off_lo, gb, (uint32_t)t, (uint32_t)cg, qi, n[0], n[1], n[2], n[3]);
}

// noise of 4 consecutive columns j0..j0+3 of one row at iteration t (0 beyond column n)
__device__ __forceinline__ void tc_noise4(const SdeParams& p, int K, long long b, uint32_t qi, int t, int j0,
                                          float (&w)[4]) {
  if (p.noise == nullptr) {
    tc_philox4(p, (unsigned long long)(p.traj_base + b), t, j0 >> 2, qi, w);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i;
      w[i] = (j < p.n && b < p.batch)
                 ? p.noise[(((size_t)t * K + qi) * p.n + j) * (size_t)p.noise_batch + (size_t)(p.traj_base + b)]
                 : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (j0 + i >= p.n) w[i] = 0.f;
}

// initial contraction input and in-place state (both ping-pong halves are cleared)
template <int SOLVER>
__global__ void tc_init_state_kernel(const SdeParams p, const TcParams tc, int n_aux) {
  constexpr int K = SolverTraits<SOLVER>::K;
  const size_t plane = (size_t)tc.rows_p * tc.np;
  const size_t idx4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx4 >= plane) return;
  const int row = (int)(idx4 / tc.np), j0 = (int)(idx4 - (size_t)row * tc.np);
  float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
  if constexpr (SOLVER == SOLVER_MF) {
    // measurement of iteration 0: clamp(mu + sqrt(1/(4 j_1)) W_0 / sqrt(dt)) with mu = 0 (mf_solver.py:551-554)
    float w[4];
    tc_noise4(p, K, row, 0u, 0, j0, w);
    const float sa = p.sched[SC_A];
    float m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float s = tc.svec[j0 + i];
      m[i] = clampf(sa * w[i], -s, s);
    }
    hi = make_float4(tf32_rna(m[0]), tf32_rna(m[1]), tf32_rna(m[2]), tf32_rna(m[3]));
    lo = make_float4(m[0] - hi.x, m[1] - hi.y, m[2] - hi.z, m[3] - hi.w);
  }
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(tc.xh + idx4) = hi;
  *reinterpret_cast<float4*>(tc.xl + idx4) = lo;
  *reinterpret_cast<float4*>(tc.xh + plane + idx4) = z;
  *reinterpret_cast<float4*>(tc.xl + plane + idx4) = z;
  for (int a = 0; a < n_aux; ++a) {
    const float init = (SOLVER == SOLVER_MF && a == 1) ? 0.5f : 0.f;  // sigma starts at 1/2
    *reinterpret_cast<float4*>(tc.aux + (size_t)a * plane + idx4) = make_float4(init, init, init, init);
  }
}

// 16 consecutive floats of a row-private array, as four 128-bit accesses
__device__ __forceinline__ void tc_ld16(const float* src, float (&r)[16]) {
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4) {
    const float4 t = *reinterpret_cast<const float4*>(src + 4 * v4);
    r[4 * v4 + 0] = t.x;
    r[4 * v4 + 1] = t.y;
    r[4 * v4 + 2] = t.z;
    r[4 * v4 + 3] = t.w;
  }
}
__device__ __forceinline__ void tc_st16(float* dst, const float (&r)[16]) {
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4)
    *reinterpret_cast<float4*>(dst + 4 * v4) = make_float4(r[4 * v4], r[4 * v4 + 1], r[4 * v4 + 2], r[4 * v4 + 3]);
}

// Adam transform of one gradient element (dl_solver.py:699-727 and siblings), scalar form of adam_tile4
__device__ __forceinline__ float tc_adam(float g, float& m, float& v, const SdeParams& p, float ib1, float ib2) {
  // branch-free (the per-element branch on beta2 == 1 fenced the 16 unrolled chains from each other):
  // beta2 == 1 -> update = alpha m_hat, the v / den chain is evaluated on harmless values and dropped
  const bool b2one = p.beta2_is_one != 0;
  m = fmaf(m, p.beta1, g * p.omb1);
  const float mh = m * ib1;
  v = b2one ? v : fmaf(v, p.beta2, (g * g) * p.omb2);
  const float den = fast_sqrt(v * ib2) + 1e-8f;
  const float upd = p.adam_alpha * (b2one ? mh : __fdividef(mh, den));
  return p.add_assign ? g + upd : upd;
}

// The solver's SDE step on 16 consecutive columns j0..j0+15 of one contraction row: G is the raw
// contraction x.Qs, x the old contraction input, W the noise of iteration t; xn receives the next
// contraction input.  In-place FP32 state (MF mu / sigma, Adam moments) lives in tc.aux and is only
// ever touched by the thread that owns the row.  Writes the solver outputs on the last iteration.
template <int SOLVER, bool ADAM>
__device__ __forceinline__ void tc_update16(const SdeParams& p, const TcParams& tc, size_t plane, int row, long long b,
                                            uint32_t qi, bool row_ok, int t, bool last, const float4 ca,
                                            const float4 cb, float next_a, int j0, const float (&G)[16],
                                            const float (&x)[16], const float (&W)[16], const float* hs,
                                            const float* ss, float (&xn)[16]) {
  constexpr int K = SolverTraits<SOLVER>::K;
  constexpr int AUX_ADAM = SOLVER == SOLVER_MF ? 2 : 0;  // first Adam array (after mu, sigma)
  const int NP = tc.np;
  if constexpr (SOLVER == SOLVER_DL) {
    float* am = tc.aux + (size_t)row * NP + j0;
    float* av = tc.aux + plane + (size_t)row * NP + j0;
    float m16[16], v16[16];
    if constexpr (ADAM) {
      tc_ld16(am, m16);
      tc_ld16(av, v16);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float g = G[i] + hs[j0 + i];
      if constexpr (ADAM) g = tc_adam(g, m16[i], v16[i], p, cb.y, cb.z);
      const float other = __shfl_xor_sync(0xffffffffu, x[i], 1);
      const float c = qi ? other : x[i], s = qi ? x[i] : other;
      const float r2 = fmaf(c, c, s * s);
      const float rt = fast_sqrt(r2 + 0.5f);
      const float u = fmaf(r2, -p.dt, qi ? ca.z : ca.y);
      const float nz = (rt * (qi ? cb.x : ca.w)) * W[i];
      xn[i] = x[i] + fmaf(x[i], u, fmaf(ca.x, g, nz));
    }
    if constexpr (ADAM) {
      tc_st16(am, m16);
      tc_st16(av, v16);
    }
    if (last && row_ok) {
      float* dst = (qi ? p.out1 : p.out0) + (size_t)b * p.n;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (j0 + i < p.n) dst[j0 + i] = qi ? xn[i] : clampf(xn[i], -ss[j0 + i], ss[j0 + i]);
    }
  } else if constexpr (SOLVER == SOLVER_MF) {
    float* mu_p = tc.aux + (size_t)row * NP + j0;
    float* sg_p = tc.aux + plane + (size_t)row * NP + j0;
    float* am = tc.aux + (size_t)AUX_ADAM * plane + (size_t)row * NP + j0;
    float* av = tc.aux + (size_t)(AUX_ADAM + 1) * plane + (size_t)row * NP + j0;
    float mun[16], sgn[16], m16[16], v16[16];
    tc_ld16(mu_p, mun);
    tc_ld16(sg_p, sgn);
    if constexpr (ADAM) {
      tc_ld16(am, m16);
      tc_ld16(av, v16);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float g = p.fs * (G[i] + hs[j0 + i]);
      if constexpr (ADAM) g = tc_adam(g, m16[i], v16[i], p, cb.y, cb.z);
      const float mu = mun[i], sg = sgn[i];
      const float g2m2 = (mu * mu) * p.g2;
      const float a1 = fmaf(g2m2, -1.f, ca.y);
      const float sh = sg + (-0.5f);
      const float dmu = fmaf(a1, mu, g);
      const float diff = (sh * ca.w) * W[i];
      mun[i] = fmaf(p.dt, dmu + diff, mu);
      const float a3 = fmaf(g2m2, -3.f, ca.y);
      const float t1 = (a3 * sg) * 2.f;
      const float t2 = (sh * sh) * (-2.f * ca.z);
      const float t3 = fmaf(g2m2, 2.f, cb.x);
      sgn[i] = fmaf(p.dt, (t1 + t2) + t3, sg);
    }
    tc_st16(mu_p, mun);
    tc_st16(sg_p, sgn);
    if constexpr (ADAM) {
      tc_st16(am, m16);
      tc_st16(av, v16);
    }
    if (!last) {
      // measurement of iteration t+1 (mf_solver.py:551-554) is the next contraction input
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        float w4[4];
        tc_noise4(p, K, b, 0u, t + 1, j0 + 4 * v4, w4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int jj = 4 * v4 + i;
          xn[jj] = clampf(fmaf(next_a, w4[i], mun[jj]), -ss[j0 + jj], ss[j0 + jj]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) xn[i] = x[i];
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (j0 + i < p.n) {
            const size_t o = (size_t)b * p.n + j0 + i;
            p.out0[o] = mun[i];
            p.out1[o] = x[i];  // the clamped measurement of the LAST iteration (mf_solver.py:591)
            p.out2[o] = sgn[i];
          }
      }
    }
  } else {
    float* am = tc.aux + (size_t)row * NP + j0;
    float* av = tc.aux + plane + (size_t)row * NP + j0;
    float m16[16], v16[16];
    if constexpr (ADAM) {
      tc_ld16(am, m16);
      tc_ld16(av, v16);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float g = G[i] + hs[j0 + i];
      if constexpr (ADAM) g = tc_adam(g, m16[i], v16[i], p, cb.y, cb.z);
      const float c = x[i];
      float inc = fmaf(p.dtfs, g, p.sig * W[i]);
      if constexpr (SOLVER == SOLVER_PLV) inc = fmaf(c, fmaf(c * c, -p.dt, ca.y), inc);
      xn[i] = clampf(c + inc, -ss[j0 + i], ss[j0 + i]);
    }
    if constexpr (ADAM) {
      tc_st16(am, m16);
      tc_st16(av, v16);
    }
    if (last && row_ok) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (j0 + i < p.n) p.out0[(size_t)b * p.n + j0 + i] = xn[i];
    }
  }
}

// ---------------------------------------------------------------- the kernel
template <int SOLVER, bool ADAM>
__global__ void __launch_bounds__(TC_THREADS, 1)
    sde_tc_kernel(const SdeParams p, const TcParams tc, const __grid_constant__ CUtensorMap map_xh,
                  const __grid_constant__ CUtensorMap map_xl, const __grid_constant__ CUtensorMap map_qh,
                  const __grid_constant__ CUtensorMap map_ql) {
  constexpr int K = SolverTraits<SOLVER>::K;

  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * TC_STAGES + 4 + TC_MAX_CHUNKS];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = tc.np, T = p.iterations;
  const int NC = NP / TC_BN;          // output chunks per iteration
  const int KB = NP / TC_BK;          // k-blocks per chunk
  const int KB_PER_CHUNK = TC_BN / TC_BK;
  const int row0 = blockIdx.x * TC_BM;
  const size_t plane = (size_t)tc.rows_p * NP;

  // 1024-byte aligned stage ring, then the per-column vectors
  const uint32_t smem_base = (smem_u32(tc_smem) + 1023u) & ~1023u;
  float* vecs = reinterpret_cast<float*>(tc_smem + (smem_base - smem_u32(tc_smem)) + TC_STAGES * TC_STAGE_BYTES);
  float* hs = vecs;        // [NP]
  float* ss = vecs + NP;   // [NP]

  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  auto accf_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + a); };
  auto acce_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + 2 + a); };
  auto ready_bar = [&](int c) { return bar0 + 8u * (2 * TC_STAGES + 4 + c); };

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf_bar(a), 1);
      mbar_init(acce_bar(a), TC_EPI_WARPS * 32);
    }
    for (int c = 0; c < TC_MAX_CHUNKS; ++c) mbar_init(ready_bar(c), TC_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 512);
  for (int j = tid; j < NP; j += TC_THREADS) {
    hs[j] = tc.hvec[j];
    ss[j] = tc.svec[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 8) {
    // ============================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        const int arow = (t & 1) * tc.rows_p + row0;  // iteration t contracts ping-pong half t&1
        for (int nc = 0; nc < NC; ++nc) {
          for (int kb = 0; kb < KB; ++kb) {
            if (t > 0 && nc == 0 && (kb % KB_PER_CHUNK) == 0)
              mbar_wait(ready_bar(kb / KB_PER_CHUNK), (uint32_t)((t - 1) & 1));  // columns written by epilogue(t-1)
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
            tma_load_2d(sa, &map_xh, kb * TC_BK, arow, full_bar(stage));
            tma_load_2d(sa + TC_A_BYTES, &map_xl, kb * TC_BK, arow, full_bar(stage));
            tma_load_2d(sa + 2 * TC_A_BYTES, &map_qh, kb * TC_BK, nc * TC_BN, full_bar(stage));
            tma_load_2d(sa + 2 * TC_A_BYTES + TC_B_BYTES, &map_ql, kb * TC_BK, nc * TC_BN, full_bar(stage));
            if (++stage == TC_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                 ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;  // global chunk counter
      for (int t = 0; t < T; ++t) {
        for (int nc = 0; nc < NC; ++nc, ++it) {
          const uint32_t ab = it & 1u;
          mbar_wait(acce_bar(ab), ((it >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ab * TC_BN;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
              const uint64_t a_hi = umma_desc_sw64(sa + ks * 32);
              const uint64_t a_lo = umma_desc_sw64(sa + TC_A_BYTES + ks * 32);
              const uint64_t b_hi = umma_desc_sw64(sa + 2 * TC_A_BYTES + ks * 32);
              const uint64_t b_lo = umma_desc_sw64(sa + 2 * TC_A_BYTES + TC_B_BYTES + ks * 32);
              umma_tf32(d_tmem, a_lo, b_hi, idesc, (kb | ks) != 0);  // small terms first
              umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
            }
            umma_commit(empty_bar(stage));  // frees the stage once these MMAs have read it
            if (++stage == TC_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(accf_bar(ab));  // accumulator complete
        }
      }
    }
  } else {
    // ============================================================ epilogue (warps 0-7)
    const int r = (warp & 3) * 32 + lane;        // accumulator row == TMEM lane
    const int colhalf = warp >> 2;               // which 128 columns of the 256-column chunk
    const int row = row0 + r;
    const long long b = K == 2 ? (row >> 1) : row;  // trajectory of this row
    const uint32_t qi = K == 2 ? (uint32_t)(row & 1) : 0u;
    const bool row_ok = row < tc.rows;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float4* sched4 = reinterpret_cast<const float4*>(p.sched);
    uint32_t it = 0;
    for (int t = 0; t < T; ++t) {
      const float4 ca = __ldg(sched4 + 2 * t), cb = __ldg(sched4 + 2 * t + 1);
      const float next_a = (t + 1 < T) ? __ldg(p.sched + (size_t)(t + 1) * SCHED_W + SC_A) : 0.f;
      const bool last = t + 1 == T;
      const float* xh_cur = tc.xh + (size_t)(t & 1) * plane + (size_t)row * NP;
      const float* xl_cur = tc.xl + (size_t)(t & 1) * plane + (size_t)row * NP;
      float* xh_nxt = tc.xh + (size_t)((t + 1) & 1) * plane + (size_t)row * NP;
      float* xl_nxt = tc.xl + (size_t)((t + 1) & 1) * plane + (size_t)row * NP;
      for (int nc = 0; nc < NC; ++nc, ++it) {
        const uint32_t ab = it & 1u;
        mbar_wait(accf_bar(ab), (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t tcol = tlane + ab * TC_BN + colhalf * (TC_BN / 2);
#pragma unroll 1
        for (int piece = 0; piece < TC_BN / 2 / 16; ++piece) {
          const int j0 = nc * TC_BN + colhalf * (TC_BN / 2) + piece * 16;
          float G[16];
          tmem_ld16(tcol + piece * 16, G);
          // old contraction input of this row (exact FP32 value = hi + lo)
          float x[16];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const float4 h4 = *reinterpret_cast<const float4*>(xh_cur + j0 + 4 * v4);
            const float4 l4 = *reinterpret_cast<const float4*>(xl_cur + j0 + 4 * v4);
            x[4 * v4 + 0] = h4.x + l4.x;
            x[4 * v4 + 1] = h4.y + l4.y;
            x[4 * v4 + 2] = h4.z + l4.z;
            x[4 * v4 + 3] = h4.w + l4.w;
          }
          float W[16];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            float w4[4];
            tc_noise4(p, K, b, qi, t, j0 + 4 * v4, w4);
#pragma unroll
            for (int i = 0; i < 4; ++i) W[4 * v4 + i] = w4[i];
          }
          tmem_wait_ld();
          if (piece == TC_BN / 2 / 16 - 1) {
            // last read of this accumulator: hand it back to the MMA issuer
            tc_fence_before();
            mbar_arrive(acce_bar(ab));
          }
          float xn[16];  // next contraction input
          tc_update16<SOLVER, ADAM>(p, tc, plane, row, b, qi, row_ok, t, last, ca, cb, next_a, j0, G, x, W, hs, ss, xn);
          // publish the next contraction input as exact (hi, lo) summands
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            float4 h4, l4;
            h4.x = tf32_rna(xn[4 * v4 + 0]);
            h4.y = tf32_rna(xn[4 * v4 + 1]);
            h4.z = tf32_rna(xn[4 * v4 + 2]);
            h4.w = tf32_rna(xn[4 * v4 + 3]);
            l4.x = xn[4 * v4 + 0] - h4.x;
            l4.y = xn[4 * v4 + 1] - h4.y;
            l4.z = xn[4 * v4 + 2] - h4.z;
            l4.w = xn[4 * v4 + 3] - h4.w;
            *reinterpret_cast<float4*>(xh_nxt + j0 + 4 * v4) = h4;
            *reinterpret_cast<float4*>(xl_nxt + j0 + 4 * v4) = l4;
          }
        }
        // the chunk's columns of the next state are written: make them visible to the TMA
        // engine (async proxy) and let the producer fetch them for iteration t+1
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_arrive(ready_bar(nc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_free(tmem_base, 512);
}

// =====================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster share every MMA.  Per pair and stage the
// accumulator tile is 256 rows x 256 columns: each CTA stages ITS 128 state rows (hi, lo) and ITS
// 128-row half of the Qs^T tile, the leader issues one tcgen05.mma.cta_group::2 (M = 256) that reads
// both halves, and each CTA's TMEM receives its own 128 accumulator rows.  Compared with the
// single-CTA kernel above (profiles/r1_ncu_sde_tc_v1.txt: L1->XBAR request port 76-89 % busy, tensor
// pipe 41 %) this
//   * halves the Qs^T bytes every SM pulls from L2 and the shared-memory bandwidth the MMAs need;
//   * moves 32-wide k-blocks (128 B rows, SWIZZLE_128B): half the TMA requests per byte;
//   * stores the new state through a swizzled shared-memory tile and TMA (full 64 B rows instead
//     of 16 B pieces of 32 B sectors) and reads the old state with 256-bit L1-bypassing loads.
constexpr int T2_BK = 32;                                // floats per k-block (128 B)
constexpr int T2_STAGES = 3;
constexpr int T2_TILE_BYTES = TC_BM * T2_BK * 4;          // 16 KB: 128 rows x 32 floats
constexpr int T2_STAGE_BYTES = 4 * T2_TILE_BYTES;         // A hi, A lo, B-half hi, B-half lo
constexpr int T2_OUT_TILE_BYTES = TC_BM * 16 * 4;         // 8 KB: 128 rows x 16 floats (SWIZZLE_64B)
constexpr int T2_SMEM_BYTES = 1024 + T2_STAGES * T2_STAGE_BYTES + 4 * T2_OUT_TILE_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier of the pair's LEADER CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                                 uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// K-major, SWIZZLE_128B, rows of 128 B, 8-row atoms of 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void ld_cg_256(const float* src, float* r) {
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(src)
               : "memory");
}

// Waits until a global chunk counter has reached `need` (acquire at GPU scope), then orders the TMA loads
// that follow behind it.  Bounded like mbar_wait: a protocol bug traps instead of hanging the device.
__device__ __forceinline__ void flag_wait(const unsigned int* flag, uint32_t need) {
  long long start = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= need) break;
    __nanosleep(64);
    if ((spins & 0xff) == 0xff) {
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > 60000000000ll) __trap();   // ~30 s: the partner may be held up by other streams' kernels
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

template <int SOLVER, bool ADAM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    sde_tc2_kernel(const SdeParams p, const TcParams tc, const __grid_constant__ CUtensorMap map_xh,
                   const __grid_constant__ CUtensorMap map_xl, const __grid_constant__ CUtensorMap map_qh,
                   const __grid_constant__ CUtensorMap map_ql, const __grid_constant__ CUtensorMap map_oh,
                   const __grid_constant__ CUtensorMap map_ol) {
  constexpr int K = SolverTraits<SOLVER>::K;
  constexpr int PIECES = TC_BN / 2 / 16;  // 16-column pieces per epilogue warpgroup and chunk

  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * T2_STAGES + 4 + TC_MAX_CHUNKS];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int NP = tc.np, T = p.iterations;
  const int NC = NP / TC_BN;
  const int KB = NP / T2_BK;
  constexpr int KB_PER_CHUNK = TC_BN / T2_BK;
  // column split (TcParams::col_split): CS pairs share one block of 256 rows, pair `cs` owns output chunks
  // [nc_lo, nc_hi); the chunks of the others reach it through global memory, published by chunk_flags
  const int CS = tc.col_split;
  const int pair_idx = (int)(blockIdx.x >> 1);
  const int cs = pair_idx % CS;
  const int row_cta = (pair_idx / CS) * 2 + (int)rank;   // index of this CTA's block of 128 rows
  const int row0 = row_cta * TC_BM;
  const int nc_lo = cs * (NC / CS), nc_hi = nc_lo + NC / CS;
  unsigned int* my_flags = tc.chunk_flags + (size_t)row_cta * NC;
  const size_t plane = (size_t)tc.rows_p * NP;

  const uint32_t smem_base = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const uint32_t out_base = smem_base + T2_STAGES * T2_STAGE_BYTES;  // [warpgroup][hi | lo] staging tiles

  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (T2_STAGES + s); };
  auto accf_bar = [&](int a) { return bar0 + 8u * (2 * T2_STAGES + a); };
  auto acce_bar = [&](int a) { return bar0 + 8u * (2 * T2_STAGES + 2 + a); };
  auto ready_bar = [&](int c) { return bar0 + 8u * (2 * T2_STAGES + 4 + c); };

  if (tid == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // used in the leader: its producer's expect_tx + both CTAs' bytes
      mbar_init(empty_bar(s), 1);   // multicast commit of the leader's MMA thread
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf_bar(a), 1);                      // multicast commit
      mbar_init(acce_bar(a), 2 * TC_EPI_WARPS * 32);  // used in the leader: both CTAs' epilogue threads
    }
    for (int c = 0; c < TC_MAX_CHUNKS; ++c) mbar_init(ready_bar(c), 2);  // one elected thread per warpgroup
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 8) {
    // ============================================================ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        const int arow = (t & 1) * tc.rows_p + row0;
        for (int nc = nc_lo; nc < nc_hi; ++nc) {
          const int brow = nc * TC_BN + (int)rank * TC_BM;  // this CTA's half of the Qs^T tile
          for (int kbi = 0; kbi < KB; ++kbi) {
            // k-blocks are consumed in the order their columns become ready: with a column split the h-th chunks
            // of all owners come first, own one leading (every pair finishes its h-th chunk at about the same
            // time), so the contraction of iteration t+1 starts under the last epilogues of iteration t
            const int ci = kbi / KB_PER_CHUNK;
            const int c = ((cs + ci % CS) % CS) * (NC / CS) + ci / CS;   // CS = 1: ci
            const int kb = c * KB_PER_CHUNK + kbi % KB_PER_CHUNK;
            if (t > 0 && nc == nc_lo && (kbi % KB_PER_CHUNK) == 0) {
              // the columns of this k-block were written by the epilogue of iteration t-1: this CTA's own
              // (shared-memory barrier) or the partner pair's (global counter, +2 per iteration)
              if (c >= nc_lo && c < nc_hi) mbar_wait(ready_bar(c), (uint32_t)((t - 1) & 1));
              else flag_wait(my_flags + c, 2u * (uint32_t)t);
            }
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * T2_STAGE_BYTES;
            const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * T2_STAGE_BYTES);
            tma_load_2d_pair(sa, &map_xh, kb * T2_BK, arow, lbar);
            tma_load_2d_pair(sa + T2_TILE_BYTES, &map_xl, kb * T2_BK, arow, lbar);
            tma_load_2d_pair(sa + 2 * T2_TILE_BYTES, &map_qh, kb * T2_BK, brow, lbar);
            tma_load_2d_pair(sa + 3 * T2_TILE_BYTES, &map_ql, kb * T2_BK, brow, lbar);
            if (++stage == T2_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================================================ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      // D = F32, A = B = TF32, K-major, N = 256, M = 256 (128 rows per CTA)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                 ((uint32_t)((2 * TC_BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = 0; t < T; ++t) {
        for (int nc = nc_lo; nc < nc_hi; ++nc, ++it) {
          const uint32_t ab = it & 1u;
          mbar_wait(acce_bar(ab), ((it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ab * TC_BN;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * T2_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < T2_BK / 8; ++ks) {
              const uint64_t a_hi = umma_desc_sw128(sa + ks * 32);
              const uint64_t a_lo = umma_desc_sw128(sa + T2_TILE_BYTES + ks * 32);
              const uint64_t b_hi = umma_desc_sw128(sa + 2 * T2_TILE_BYTES + ks * 32);
              const uint64_t b_lo = umma_desc_sw128(sa + 3 * T2_TILE_BYTES + ks * 32);
              umma_tf32_pair(d_tmem, a_lo, b_hi, idesc, (kb | ks) != 0);
              umma_tf32_pair(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_tf32_pair(d_tmem, a_hi, b_hi, idesc, 1u);
            }
            umma_commit_pair(empty_bar(stage));
            if (++stage == T2_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_pair(accf_bar(ab));
        }
      }
    }
  } else {
    // ============================================================ epilogue (warps 0-7, both CTAs)
    const int wg = warp >> 2;                    // warpgroup = 128-column half of the chunk
    const int r = (warp & 3) * 32 + lane;        // accumulator row == TMEM lane
    const int row = row0 + r;
    const long long b = K == 2 ? (row >> 1) : row;
    const uint32_t qi = K == 2 ? (uint32_t)(row & 1) : 0u;
    const bool row_ok = row < tc.rows;
    const bool elected = (tid & 127) == 0;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t out_hi = out_base + wg * 2 * T2_OUT_TILE_BYTES, out_lo = out_hi + T2_OUT_TILE_BYTES;
    const uint32_t my_row = (uint32_t)r * 64u;
    const uint32_t swz = ((uint32_t)r >> 1) & 3u;  // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
    const uint32_t acce_leader0 = mapa_cluster(acce_bar(0), 0), acce_leader1 = mapa_cluster(acce_bar(1), 0);
    const float4* sched4 = reinterpret_cast<const float4*>(p.sched);
    uint32_t it = 0;
    for (int t = 0; t < T; ++t) {
      const float4 ca = __ldg(sched4 + 2 * t), cb = __ldg(sched4 + 2 * t + 1);
      const float next_a = (t + 1 < T) ? __ldg(p.sched + (size_t)(t + 1) * SCHED_W + SC_A) : 0.f;
      const bool last = t + 1 == T;
      const float* xh_cur = tc.xh + (size_t)(t & 1) * plane + (size_t)row * NP;
      const float* xl_cur = tc.xl + (size_t)(t & 1) * plane + (size_t)row * NP;
      const int nrow0 = ((t + 1) & 1) * tc.rows_p + row0;  // first row of this CTA in the next state half
      for (int nc = nc_lo; nc < nc_hi; ++nc, ++it) {
        const uint32_t ab = it & 1u;
        mbar_wait(accf_bar(ab), (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t tcol = tlane + ab * TC_BN + wg * (TC_BN / 2);
#pragma unroll 1
        for (int piece = 0; piece < PIECES; ++piece) {
          const int j0 = nc * TC_BN + wg * (TC_BN / 2) + piece * 16;
          float G[16];
          tmem_ld16(tcol + piece * 16, G);
          // old contraction input: written by TMA (async proxy, L2), so it must not be served
          // from this SM's L1 -- 256-bit .cg loads fetch whole 32 B sectors straight from L2
          float x[16], xl[16];
          ld_cg_256(xh_cur + j0, x);
          ld_cg_256(xh_cur + j0 + 8, x + 8);
          ld_cg_256(xl_cur + j0, xl);
          ld_cg_256(xl_cur + j0 + 8, xl + 8);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] += xl[i];
          float W[16];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            float w4[4];
            tc_noise4(p, K, b, qi, t, j0 + 4 * v4, w4);
#pragma unroll
            for (int i = 0; i < 4; ++i) W[4 * v4 + i] = w4[i];
          }
          tmem_wait_ld();
          if (piece == PIECES - 1) {
            tc_fence_before();
            mbar_arrive_cluster(ab ? acce_leader1 : acce_leader0);  // accumulator drained (leader's barrier)
          }
          float xn[16];
          tc_update16<SOLVER, ADAM>(p, tc, plane, row, b, qi, row_ok, t, last, ca, cb, next_a, j0, G, x, W, tc.hvec,
                                    tc.svec, xn);
          // stage the new (hi, lo) rows in the swizzled tile and hand them to the TMA engine
          if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store left the tile
          group_barrier(1 + wg, 128);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const float h0 = tf32_rna(xn[4 * v4 + 0]), h1 = tf32_rna(xn[4 * v4 + 1]);
            const float h2 = tf32_rna(xn[4 * v4 + 2]), h3 = tf32_rna(xn[4 * v4 + 3]);
            const uint32_t off = my_row + (((uint32_t)v4 ^ swz) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_hi + off), "f"(h0), "f"(h1), "f"(h2),
                         "f"(h3)
                         : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_lo + off), "f"(xn[4 * v4 + 0] - h0),
                         "f"(xn[4 * v4 + 1] - h1), "f"(xn[4 * v4 + 2] - h2), "f"(xn[4 * v4 + 3] - h3)
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          group_barrier(1 + wg, 128);
          if (elected) {
            tma_store_2d(&map_oh, j0, nrow0, out_hi);
            tma_store_2d(&map_ol, j0, nrow0, out_lo);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (elected) {
          // this warpgroup's half of the chunk is in global memory: the producer may fetch it
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          mbar_arrive(ready_bar(nc));
          if (CS > 1) {
            // ... and so may the partner pair's: publish at GPU scope (the TMA stores above are complete)
            __threadfence();
            atomicAdd(my_flags + nc, 1u);
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 9)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace ccvm
