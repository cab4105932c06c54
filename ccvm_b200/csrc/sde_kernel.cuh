// The persistent Euler-Maruyama kernel: one launch runs ALL iterations of one
// Solver._solve / Solver._solve_adam call (reference: solvers/{dl,mf,langevin,pumped_langevin}_solver.py).
//
// Work decomposition (sm_100a, FP32 SIMT with packed FFMA2):
//   * a CTA owns Bt = RG*TB trajectories of the batch for the whole run;
//   * a thread owns a register tile of TB trajectories x 4 variables (x2 quadratures for DL)
//     of every state array (amplitudes, Adam moments) for all iterations;
//   * the drift contraction y.Q is a register-tiled outer-product loop over k:
//       - the box-scaled matrix Qs = -alpha_i alpha_j Q_ij lives in shared memory with every
//         element DUPLICATED into a float2 (q,q), so that one FFMA2 advances the same variable
//         of two neighbouring trajectories:  acc(b,b+1 ; j) += x(b,b+1 ; k) * (Qs_kj, Qs_kj);
//       - the contraction input (c,s / clamped mu_tilde / c) is staged k-major in a
//         double-buffered shared panel X[buf][k][row], rewritten once per iteration;
//       - the affine part of the drift, h_j = -alpha_j (b/2 colsum_j(Q) + V_j), initialises the
//         accumulators, so acc == -(gradient term) when the k loop ends;
//   * one __syncthreads per iteration; nothing touches HBM inside the loop except the
//     8-float schedule row (L1/L2 resident) and, in validation mode, the replayed noise.
#pragma once
#include "ccvm_common.cuh"

namespace ccvm {

template <int SOLVER>
struct SolverTraits {
  static constexpr int K = (SOLVER == SOLVER_DL) ? 2 : 1;        // contraction inputs per trajectory
  static constexpr int NSTATE = (SOLVER == SOLVER_LV || SOLVER == SOLVER_PLV) ? 1 : 2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Adam transform applied to a gradient tile (dl_solver.py:699-727 and siblings).
template <int P>
__device__ __forceinline__ void adam_tile(pf2 (&g)[P][4], pf2 (&m)[P][4], pf2 (&v)[P][4],
                                          const SdeParams& p, float ib1, float ib2) {
  const pf2 b1 = dup(p.beta1), b2 = dup(p.beta2), o1 = dup(p.omb1), o2 = dup(p.omb2);
  const pf2 i1 = dup(ib1), i2 = dup(ib2), al = dup(p.adam_alpha), eps = dup(1e-8f);
#pragma unroll
  for (int pp = 0; pp < P; ++pp)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const pf2 gr = g[pp][jj];
      m[pp][jj] = fma2(m[pp][jj], b1, mul2(gr, o1));
      const pf2 mh = mul2(m[pp][jj], i1);
      pf2 upd;
      if (!p.beta2_is_one) {
        v[pp][jj] = fma2(v[pp][jj], b2, mul2(mul2(gr, gr), o2));
        const pf2 den = add2(sqrt2(mul2(v[pp][jj], i2)), eps);
        upd = mul2(al, div2(mh, den));
      } else {
        upd = mul2(al, mh);
      }
      g[pp][jj] = p.add_assign ? add2(gr, upd) : upd;
    }
}

template <int SOLVER, bool ADAM, int TB>
__global__ void __launch_bounds__(256, 1) sde_kernel(const SdeParams p) {
  constexpr int K = SolverTraits<SOLVER>::K;
  constexpr int P = TB / 2;
  static_assert(TB == 2 || TB == 4 || TB == 8, "TB must be 2, 4 or 8");

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int N = p.n, CG = p.cg, NP = 4 * CG, RG = p.rg, XS = p.xs, Bt = RG * TB;
  const int T = p.iterations;

  float4* Qd = reinterpret_cast<float4*>(smem);           // [NP][2][CG] duplicated Qs
  float* X = smem + (size_t)NP * NP * 2;                  // [2][NP][XS] contraction input
  float* hv = X + (size_t)2 * NP * XS;                    // [NP] affine drift term
  float* av = hv + NP;                                    // [NP] alpha_j = (u-l)/(2 S_j)
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(av + NP);

  // ------------------------------------------------------------------ prologue
  for (int j = tid; j < NP; j += blockDim.x) {
    float a = 0.f;
    if (j < N) a = p.a_half / (p.drift_s_vec ? p.drift_s_vec[j] : p.drift_s);
    av[j] = a;
  }
  const float* rawq = p.q;
  if (p.use_tma) {
    // Stage the raw matrix with one bulk async copy (TMA engine) into the X region.
    const uint32_t bar = smem_u32(mbar), dst = smem_u32(X);
    const uint32_t bytes = (uint32_t)N * N * 4u;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
          "l"(p.q), "r"(bytes), "r"(bar)
          : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
    rawq = X;
  }
  __syncthreads();
  for (int idx = tid; idx < NP * 2 * CG; idx += blockDim.x) {
    const int k = idx / (2 * CG), r = idx - k * 2 * CG;
    const int plane = r / CG, cgi = r - plane * CG;
    const int j0 = 4 * cgi + 2 * plane;
    float q0 = 0.f, q1 = 0.f;
    if (k < N) {
      const float ak = -av[k];
      if (j0 < N) q0 = ak * av[j0] * rawq[k * N + j0];
      if (j0 + 1 < N) q1 = ak * av[j0 + 1] * rawq[k * N + j0 + 1];
    }
    Qd[idx] = make_float4(q0, q0, q1, q1);
  }
  for (int j = tid; j < NP; j += blockDim.x) {
    float h = 0.f;
    if (j < N) {
      float cs = 0.f;
      for (int i = 0; i < N; ++i) cs += rawq[i * N + j];
      h = -av[j] * (p.b_half * cs + p.v[j]);
    }
    hv[j] = h;
  }
  __syncthreads();
  for (int i = tid; i < 2 * NP * XS; i += blockDim.x) X[i] = 0.f;

  // ------------------------------------------------------------------ thread tile
  const int rg = tid % RG, cg = tid / RG;
  const bool active = cg < CG;  // idle lanes only keep the barriers company
  const int cgc = active ? cg : 0;
  const int row0 = rg * TB;                                  // first local trajectory
  const long long gb0 = (long long)blockIdx.x * Bt + row0;   // first trajectory (in this launch)
  const int j0 = 4 * cgc;

  float hreg[4], sclamp[4];
  bool colok[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    colok[jj] = active && (j0 + jj < N);
    sclamp[jj] = 0.f;
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    hreg[jj] = hv[j0 + jj];
    if (colok[jj]) sclamp[jj] = p.clamp_s_vec ? p.clamp_s_vec[j0 + jj] : p.clamp_s;
  }

  // state: st[0] = c | mu, st[1] = s | sigma ; Adam moments m,v per tracked array
  pf2 st[2][P][4];
  pf2 am[K][P][4], avv[K][P][4];
  pf2 W[K][P][4];    // noise of the current iteration
  pf2 meas[P][4];    // MF: clamped measurement mu_tilde of the current iteration
#pragma unroll
  for (int pp = 0; pp < P; ++pp)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      st[0][pp][jj] = dup(0.f);
      st[1][pp][jj] = dup(SOLVER == SOLVER_MF ? 0.5f : 0.f);
      meas[pp][jj] = dup(0.f);
#pragma unroll
      for (int q = 0; q < K; ++q) {
        am[q][pp][jj] = dup(0.f);
        avv[q][pp][jj] = dup(0.f);
        W[q][pp][jj] = dup(0.f);
      }
    }

  const uint2 key = make_uint2(p.seed_lo, p.seed_hi ^ p.off_hi);

  // draws the K*TB*4 normals of iteration `t` for this thread's tile
  auto draw = [&](int t) {
    if (p.noise == nullptr) {
#pragma unroll
      for (int q = 0; q < K; ++q)
#pragma unroll
        for (int i = 0; i < TB; ++i) {
          const unsigned long long gb = (unsigned long long)(p.traj_base + gb0 + i);
          const uint4 r = philox4x32_10(
              make_uint4((uint32_t)gb, (uint32_t)t, (uint32_t)cgc | ((uint32_t)q << 24) | ((uint32_t)(gb >> 32) << 25),
                         p.off_lo),
              key);
          float n0, n1, n2, n3;
          box_muller(r.x, r.y, n0, n1);
          box_muller(r.z, r.w, n2, n3);
          const float nn[4] = {n0, n1, n2, n3};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float w = colok[jj] ? nn[jj] : 0.f;
            if (i & 1) W[q][i / 2][jj].y = w; else W[q][i / 2][jj].x = w;
          }
        }
    } else {
#pragma unroll
      for (int q = 0; q < K; ++q)
#pragma unroll
        for (int i = 0; i < TB; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float w = 0.f;
            const long long b = gb0 + i;
            if (colok[jj] && b < p.batch)
              w = p.noise[(((size_t)t * K + q) * N + (j0 + jj)) * (size_t)p.noise_batch + (size_t)(p.traj_base + b)];
            if (i & 1) W[q][i / 2][jj].y = w; else W[q][i / 2][jj].x = w;
          }
    }
  };

  // stores this thread's tile of the next contraction input into panel `buf`
  auto stage = [&](int buf, const pf2 (&src)[P][4], int q) {
    if (!active) return;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float* dst = X + ((size_t)buf * NP + (j0 + jj)) * XS + ((cgc * Bt) & 31) + q * Bt + row0;
      if constexpr (TB == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(src[0][jj].x, src[0][jj].y);
      } else {
#pragma unroll
        for (int h = 0; h < TB / 4; ++h)
          *reinterpret_cast<float4*>(dst + 4 * h) =
              make_float4(src[2 * h][jj].x, src[2 * h][jj].y, src[2 * h + 1][jj].x, src[2 * h + 1][jj].y);
      }
    }
  };

  const float4* sched4 = reinterpret_cast<const float4*>(p.sched);
  float4 sa = __ldg(sched4), sb = __ldg(sched4 + 1);

  if constexpr (SOLVER == SOLVER_MF) {
    // measurement of iteration 0 (mf_solver.py:551-554): mu = 0
    draw(0);
#pragma unroll
    for (int pp = 0; pp < P; ++pp)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        meas[pp][jj] = clamp2(fma2(dup(sa.x), W[0][pp][jj], st[0][pp][jj]), -sclamp[jj], sclamp[jj]);
    __syncthreads();  // zero fill done
    stage(0, meas, 0);
  }
  __syncthreads();

  // ------------------------------------------------------------------ main loop
  for (int t = 0; t < T; ++t) {
    const int buf = t & 1;
    const float4 ca = sa, cb = sb;  // schedule row of this iteration
    if (t + 1 < T) {                // prefetch the next row
      sa = __ldg(sched4 + 2 * (t + 1));
      sb = __ldg(sched4 + 2 * (t + 1) + 1);
    }

    // ---- drift contraction: acc = h + X . Qs
    pf2 acc[K][P][4];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[q][pp][jj] = dup(hreg[jj]);

    {
      const float4* qp = Qd + cgc;
      const float* xp = X + (size_t)buf * NP * XS + row0;
      for (int kc = 0; kc < CG; ++kc) {
        const float* xrow = xp + ((kc * Bt) & 31);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 qa = qp[0], qb = qp[CG];
          const pf2 qd[4] = {pk(qa.x, qa.y), pk(qa.z, qa.w), pk(qb.x, qb.y), pk(qb.z, qb.w)};
          pf2 xv[K][P];
#pragma unroll
          for (int q = 0; q < K; ++q) {
            if constexpr (TB == 2) {
              const float2 x2 = *reinterpret_cast<const float2*>(xrow + q * Bt);
              xv[q][0] = pk(x2.x, x2.y);
            } else {
#pragma unroll
              for (int h = 0; h < TB / 4; ++h) {
                const float4 x4 = *reinterpret_cast<const float4*>(xrow + q * Bt + 4 * h);
                xv[q][2 * h] = pk(x4.x, x4.y);
                xv[q][2 * h + 1] = pk(x4.z, x4.w);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < K; ++q)
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) acc[q][pp][jj] = fma2(xv[q][pp], qd[jj], acc[q][pp][jj]);
          qp += 2 * CG;
          xrow += XS;
        }
        xp += 4 * XS;
      }
    }

    // ---- elementwise SDE step
    if constexpr (SOLVER == SOLVER_DL) {
      draw(t);
      // schedule: ca.x = dt*fs*(0.5+rate) | dt ; ca.y = dt(-1+p_t) ; ca.z = dt(-1-p_t) ;
      //           ca.w = 2g sqrt(dt) r_t ; cb.x = 2g sqrt(dt)/r_t ; cb.y, cb.z = Adam bias terms
      if constexpr (ADAM) {
        adam_tile<P>(acc[0], am[0], avv[0], p, cb.y, cb.z);
        adam_tile<P>(acc[1], am[1], avv[1], p, cb.y, cb.z);
      }
      const pf2 gain = dup(ca.x), d1 = dup(ca.y), d2 = dup(ca.z), n1 = dup(ca.w), n2 = dup(cb.x);
      const pf2 mdt = dup(-p.dt), half = dup(0.5f);
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 c = st[0][pp][jj], s = st[1][pp][jj];
          const pf2 r2 = fma2(c, c, mul2(s, s));
          const pf2 rt = sqrt2(add2(r2, half));
          const pf2 uc = fma2(r2, mdt, d1), us = fma2(r2, mdt, d2);
          const pf2 nc = mul2(mul2(rt, n1), W[0][pp][jj]);
          const pf2 ns = mul2(mul2(rt, n2), W[1][pp][jj]);
          st[0][pp][jj] = add2(c, fma2(c, uc, fma2(gain, acc[0][pp][jj], nc)));
          st[1][pp][jj] = add2(s, fma2(s, us, fma2(gain, acc[1][pp][jj], ns)));
        }
      stage(buf ^ 1, st[0], 0);
      stage(buf ^ 1, st[1], 1);
    } else if constexpr (SOLVER == SOLVER_MF) {
      // schedule: ca.x = sqrt(1/(4 j_t))/sqrt(dt) ; ca.y = pump*rate ; ca.z = j_t ;
      //           ca.w = sqrt(j_t)/sqrt(dt) ; cb.x = 1 + j_t
      const pf2 fs = dup(p.fs);
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[0][pp][jj] = mul2(fs, acc[0][pp][jj]);
      if constexpr (ADAM) adam_tile<P>(acc[0], am[0], avv[0], p, cb.y, cb.z);
      const pf2 pr = dup(ca.y), sj = dup(ca.w), opj = dup(cb.x), m2j = dup(-2.f * ca.z);
      const pf2 g2 = dup(p.g2), dtp = dup(p.dt), mhalf = dup(-0.5f);
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 mu = st[0][pp][jj], sg = st[1][pp][jj];
          const pf2 g2m2 = mul2(mul2(mu, mu), g2);
          const pf2 a1 = fma2(g2m2, dup(-1.f), pr);
          const pf2 sh = add2(sg, mhalf);
          const pf2 dmu = fma2(a1, mu, acc[0][pp][jj]);
          const pf2 diff = mul2(mul2(sh, sj), W[0][pp][jj]);
          st[0][pp][jj] = fma2(dtp, add2(dmu, diff), mu);
          const pf2 a3 = fma2(g2m2, dup(-3.f), pr);
          const pf2 t1 = mul2(mul2(a3, sg), dup(2.f));
          const pf2 t2 = mul2(mul2(sh, sh), m2j);
          const pf2 t3 = fma2(g2m2, dup(2.f), opj);
          st[1][pp][jj] = fma2(dtp, add2(add2(t1, t2), t3), sg);
        }
      if (t + 1 < T) {
        // measurement of the next iteration
        draw(t + 1);
#pragma unroll
        for (int pp = 0; pp < P; ++pp)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            meas[pp][jj] = clamp2(fma2(dup(sa.x), W[0][pp][jj], st[0][pp][jj]), -sclamp[jj], sclamp[jj]);
        stage(buf ^ 1, meas, 0);
      }
    } else {
      draw(t);
      // Langevin / pumped Langevin.  schedule: ca.y = dt(p_t - 1) (pumped only)
      if constexpr (ADAM) adam_tile<P>(acc[0], am[0], avv[0], p, cb.y, cb.z);
      const pf2 dtfs = dup(p.dtfs), sig = dup(p.sig), mdt = dup(-p.dt), d1 = dup(ca.y);
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const pf2 c = st[0][pp][jj];
          pf2 inc = fma2(dtfs, acc[0][pp][jj], mul2(sig, W[0][pp][jj]));
          if constexpr (SOLVER == SOLVER_PLV) inc = fma2(c, fma2(mul2(c, c), mdt, d1), inc);
          st[0][pp][jj] = clamp2(add2(c, inc), -sclamp[jj], sclamp[jj]);
        }
      stage(buf ^ 1, st[0], 0);
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
          for (int i = 0; i < TB; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const long long b = gb0 + i;
              if (b < p.batch && colok[jj]) {
                const pf2 val = st[a][i / 2][jj];
                p.samples[(((size_t)a * p.num_samples + sidx) * p.batch + b) * N + j0 + jj] = (i & 1) ? val.y : val.x;
              }
            }
      }
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ results
  if (active) {
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      const long long b = gb0 + i;
      if (b >= p.batch) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (!colok[jj]) continue;
        const size_t o = (size_t)b * N + j0 + jj;
        const pf2 a0 = st[0][i / 2][jj], a1 = st[1][i / 2][jj], mm = meas[i / 2][jj];
        const float v0 = (i & 1) ? a0.y : a0.x, v1 = (i & 1) ? a1.y : a1.x, vm = (i & 1) ? mm.y : mm.x;
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
}

}  // namespace ccvm
