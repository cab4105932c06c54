// Tail of Solver.__call__ as device code: change of variables, batched post-processors, BoxQP energy,
// solution statistics.  The SAME bodies serve the stand-alone kernels of the C ABI (ccvm_epilogue,
// ccvm_solution_stats: grid-stride over the whole batch) and the tail of the persistent SDE kernels
// (each CTA on its own trajectories, FusedTail) -- results are bit-identical by construction.
//
// Replaces: CCVMSolver.change_variables (dl_solver.py:219-235; Langevin's (c+S)/(2S),
// langevin_solver.py:717-722), PostProcessorGradDescent.postprocess (post_processor/grad_descent.py:58-64),
// PostProcessorAdam.postprocess (adam.py:58-66), ProblemInstance.compute_energy
// (problem_classes/boxqp/problem_instance.py:226-241), Solution.get_solution_stats (solution.py:65-146).
#pragma once
#include <math.h>

#include "../../include/ccvm_b200.h"
#include "ccvm_common.cuh"

namespace ccvm {

constexpr int EPI_WARPS = 8;  // warps per CTA of the stand-alone epilogue kernels
constexpr int EPT = 8;        // trajectories per tile of the tiled body
constexpr int EPW = 4;        // trajectories a warp of the per-warp body walks at once (one LDS.128 per row)

// shared-memory floats the two bodies need for `nwarps` warps (Q staged on chip or not)
__host__ __device__ inline size_t epi_tile_floats(int n, int nwarps, int wpt, bool q_in_smem) {
  const int tpc = nwarps / wpt;
  return (q_in_smem ? (((size_t)n * n + 3) & ~(size_t)3) : 0) + (size_t)tpc * 2 * n * EPT + (size_t)tpc * wpt * 4 * EPT * 2;
}
__host__ __device__ inline size_t epi_warp_floats(int n, int nwarps, bool q_in_smem) {
  return (q_in_smem ? (((size_t)n * (n | 1) + 3) & ~(size_t)3) : 0) + (size_t)nwarps * 2 * n * EPW;
}

// ------------------------------------------------------------------------- warp per trajectory
// One warp walks one trajectory at a time; lanes own variables j = lane, lane+32, ...
// The working vector x lives in a per-warp shared buffer; Q is staged in shared memory with an
// odd leading dimension (row and column sweeps are both conflict-free) when it fits, else read
// through L2.  The objective is reduced with warp shuffles.  Trajectories [b_begin, b_end) are
// dealt to the `n_slots * nwarps` warps of all participating CTAs; every thread of the CTA must call.
static __device__ __noinline__ void epilogue_warp_body(const EpiParams& p, float* esm, long long b_begin, long long b_end,
                                                       int cta_slot, int n_slots) {
  const int N = p.n, LD = p.ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* qs = esm;
  float* xbuf = esm + (p.q_in_smem ? (((size_t)N * LD + 3) & ~(size_t)3) : 0) + (size_t)warp * 2 * N * EPW;  // 16-byte aligned
  const float* Q = p.q;
  int ld = N;
  if (p.q_in_smem) {
    for (int i = warp; i < N; i += nwarps)
      for (int j = lane; j < N; j += 32) qs[i * LD + j] = p.q[(size_t)i * N + j];
    Q = qs;
    ld = LD;
  }
  __syncthreads();

  // a warp walks EPW trajectories AT ONCE (element [j][u] of the working vectors belongs to trajectory u): every
  // Q element it loads serves all of them and the EPW dot products are independent FMA chains -- the
  // one-trajectory-at-a-time version was latency-bound (36 us of tail behind a 3.2 ms loop at N = 70).  Each
  // trajectory still sees exactly the arithmetic below in exactly this order.
  const long long wslot = (long long)cta_slot * nwarps + warp, nslot = (long long)n_slots * nwarps;
  for (long long b0 = b_begin + wslot * EPW; b0 < b_end; b0 += nslot * EPW) {
    float* x = xbuf;
    float* y = xbuf + (size_t)N * EPW;
    bool ok[EPW];
#pragma unroll
    for (int u = 0; u < EPW; ++u) ok[u] = b0 + u < b_end;
    for (int j = lane; j < N; j += 32) {
#pragma unroll
      for (int u = 0; u < EPW; ++u) {
        float val = ok[u] ? p.state[(size_t)(b0 + u) * N + j] : 0.f;
        if (p.map1) val = val * (p.m1vec ? p.m1vec[j] : p.m1s) + p.m1o;
        x[j * EPW + u] = val;
      }
    }
    __syncwarp();
    if (p.pp == CCVM_PP_GRAD_DESCENT) {
      // x <- clamp(x - step (xQ + V), lo, hi), all variables from the OLD x (grad_descent.py:61-64)
      for (int it = 0; it < p.pp_iters; ++it) {
        for (int j = lane; j < N; j += 32) {
          float acc[EPW];
#pragma unroll
          for (int u = 0; u < EPW; ++u) acc[u] = 0.f;
          for (int i = 0; i < N; ++i) {
            const float q = Q[i * ld + j];
            const float4 x4 = *reinterpret_cast<const float4*>(x + i * EPW);
            const float xi[EPW] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int u = 0; u < EPW; ++u) acc[u] = fmaf(xi[u], q, acc[u]);
          }
#pragma unroll
          for (int u = 0; u < EPW; ++u) {
            const float g = acc[u] + p.v[j];
            y[j * EPW + u] = clampf(x[j * EPW + u] + (-p.step) * g, p.lo, p.hi);
          }
        }
        __syncwarp();
        float* tmp = x;
        x = y;
        y = tmp;
      }
    } else if (p.pp == CCVM_PP_ADAM) {
      // one torch.optim.Adam step on 1/2 xQx + Vx then clamp (adam.py:58-66):
      // g = 1/2 (xQ + x Q^T) + V ; x <- clamp(x - lr * (m/(1-b1)) / (sqrt(v/(1-b2)) + eps))
      for (int j = lane; j < N; j += 32) {
        float a1[EPW], a2[EPW];
#pragma unroll
        for (int u = 0; u < EPW; ++u) a1[u] = a2[u] = 0.f;
        for (int i = 0; i < N; ++i) {
          const float qc = Q[i * ld + j], qr = Q[j * ld + i];
          const float4 x4 = *reinterpret_cast<const float4*>(x + i * EPW);
          const float xi[EPW] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int u = 0; u < EPW; ++u) {
            a1[u] = fmaf(xi[u], qc, a1[u]);
            a2[u] = fmaf(xi[u], qr, a2[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < EPW; ++u) {
          const float g = 0.5f * (a1[u] + a2[u]) + p.v[j];
          const float m = (1.f - 0.9f) * g, vv = (1.f - 0.99f) * g * g;
          const float den = sqrtf(vv) / sqrtf(1.f - 0.99f) + 1e-8f;
          y[j * EPW + u] = clampf(x[j * EPW + u] - (p.step / (1.f - 0.9f)) * (m / den), p.lo, p.hi);
        }
      }
      __syncwarp();
      float* tmp = x;
      x = y;
      y = tmp;
    }
    if (p.pv)
      for (int j = lane; j < N; j += 32) {
#pragma unroll
        for (int u = 0; u < EPW; ++u)
          if (ok[u]) p.pv[(size_t)(b0 + u) * N + j] = x[j * EPW + u];
      }
    if (p.energy) {
      if (p.map2) {
        for (int j = lane; j < N; j += 32) {
#pragma unroll
          for (int u = 0; u < EPW; ++u) y[j * EPW + u] = x[j * EPW + u] * (p.m2vec ? p.m2vec[j] : p.m2s) + p.m2o;
        }
        __syncwarp();
        x = y;
      }
      // E = (1/2 x Q x + V x) * scaled_by   (problem_instance.py:226-241)
      float e1[EPW], e2[EPW];
#pragma unroll
      for (int u = 0; u < EPW; ++u) e1[u] = e2[u] = 0.f;
      for (int j = lane; j < N; j += 32) {
        float acc[EPW];
#pragma unroll
        for (int u = 0; u < EPW; ++u) acc[u] = 0.f;
        for (int i = 0; i < N; ++i) {
          const float q = Q[i * ld + j];
          const float4 x4 = *reinterpret_cast<const float4*>(x + i * EPW);
          const float xi[EPW] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int u = 0; u < EPW; ++u) acc[u] = fmaf(xi[u], q, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < EPW; ++u) {
          e1[u] = fmaf(acc[u], x[j * EPW + u], e1[u]);
          e2[u] = fmaf(p.v[j], x[j * EPW + u], e2[u]);
        }
      }
      const float sb = p.scaled_by_ptr ? *p.scaled_by_ptr : p.scaled_by;
#pragma unroll
      for (int u = 0; u < EPW; ++u) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          e1[u] += __shfl_xor_sync(0xffffffffu, e1[u], o);
          e2[u] += __shfl_xor_sync(0xffffffffu, e2[u], o);
        }
        if (lane == 0 && ok[u]) p.energy[b0 + u] = 0.5f * (e1[u] * sb) + e2[u] * sb;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------- tiles of 8 trajectories
// Change of variables + gradient-descent post-processor + energy (everything but the Adam
// post-processor, whose gradient also walks Q by rows).  The body above reads the whole matrix once per
// TRAJECTORY and iteration (2.5 GB through L2 for B = 1000, n = 250, 10 iterations: 660 us, a third of
// the solve it follows); here a tile of 8 trajectories shares every Q element: `wpt` warps split the
// columns of the tile (one column per lane and 32 * wpt columns per pass, CPL passes), the 8 x-values
// of a row come from one broadcast LDS.128 pair, and the dot products keep the single-accumulator,
// ascending-i order of the reference einsum (same values as the body above; only the final energy
// sum is associated differently).  Named barriers 1 .. 8 synchronise the warps of a tile.
template <int CPL>
static __device__ __noinline__ void epilogue_tile_body(const EpiParams& p, float* esm, long long b_begin, long long b_end,
                                                       int cta_slot, int n_slots) {
  const int N = p.n, wpt = p.wpt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int tpc = nwarps / wpt;                      // tiles per CTA
  const int tile = warp / wpt, wt = warp - tile * wpt;
  float* qs = esm;
  float* xall = esm + (p.q_in_smem ? (((size_t)N * N + 3) & ~(size_t)3) : 0);  // 16-byte aligned (LDS.128)
  float* red = xall + (size_t)tpc * 2 * N * EPT;  // [tpc][wpt * CPL column blocks][EPT][2]
  if (p.q_in_smem) {
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) qs[idx] = p.q[idx];
  }
  __syncthreads();
  if (tile >= tpc) return;                      // warps that do not fill a tile (nwarps % wpt != 0)
  const float* Q = p.q_in_smem ? qs : p.q;
  float* x = xall + (size_t)tile * 2 * N * EPT;
  float* y = x + (size_t)N * EPT;
  const int nthr = wpt * 32, tl = wt * 32 + lane;
  auto tile_sync = [&]() {
    if (wpt == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "r"(nthr) : "memory");
  };
  int jc[CPL];
  bool jok[CPL];
  float vj[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = (wt * CPL + c) * 32 + lane;
    jok[c] = j < N;
    jc[c] = jok[c] ? j : 0;
    vj[c] = p.v[jc[c]];
  }
  // acc[c][tb] = sum_i x[i][tb] * Q[i][j_c]   (single accumulator per element, ascending i)
  auto contract = [&](const float* xs, float (&acc)[CPL][EPT]) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int tb = 0; tb < EPT; ++tb) acc[c][tb] = 0.f;
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      const float4 x0 = *reinterpret_cast<const float4*>(xs + (size_t)i * EPT);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + (size_t)i * EPT + 4);
      const float xv[EPT] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const float q = Q[(size_t)i * N + jc[c]];
#pragma unroll
        for (int tb = 0; tb < EPT; ++tb) acc[c][tb] = fmaf(xv[tb], q, acc[c][tb]);
      }
    }
  };

  for (long long b0 = b_begin + ((long long)cta_slot * tpc + tile) * EPT; b0 < b_end; b0 += (long long)n_slots * tpc * EPT) {
    for (int idx = tl; idx < N * EPT; idx += nthr) {
      const int tb = idx / N, j = idx - tb * N;
      float val = 0.f;
      if (b0 + tb < b_end) {
        val = p.state[(size_t)(b0 + tb) * N + j];
        if (p.map1) val = val * (p.m1vec ? p.m1vec[j] : p.m1s) + p.m1o;
      }
      x[(size_t)j * EPT + tb] = val;
    }
    tile_sync();
    float acc[CPL][EPT];
    if (p.pp == CCVM_PP_GRAD_DESCENT) {
      // x <- clamp(x - step (xQ + V), lo, hi), all variables from the OLD x (grad_descent.py:61-64)
      for (int it = 0; it < p.pp_iters; ++it) {
        contract(x, acc);
#pragma unroll
        for (int c = 0; c < CPL; ++c)
          if (jok[c]) {
#pragma unroll
            for (int tb = 0; tb < EPT; ++tb) {
              const float g = acc[c][tb] + vj[c];
              y[(size_t)jc[c] * EPT + tb] = clampf(x[(size_t)jc[c] * EPT + tb] + (-p.step) * g, p.lo, p.hi);
            }
          }
        tile_sync();
        float* tmp = x;
        x = y;
        y = tmp;
      }
    }
    if (p.pv)
      for (int idx = tl; idx < N * EPT; idx += nthr) {
        const int tb = idx / N, j = idx - tb * N;
        if (b0 + tb < b_end) p.pv[(size_t)(b0 + tb) * N + j] = x[(size_t)j * EPT + tb];
      }
    if (p.energy) {
      if (p.map2) {
        for (int idx = tl; idx < N * EPT; idx += nthr) {
          const int j = idx / EPT;
          y[idx] = x[idx] * (p.m2vec ? p.m2vec[j] : p.m2s) + p.m2o;
        }
        tile_sync();
        float* tmp = x;
        x = y;
        y = tmp;
      }
      // E = (1/2 x Q x + V x) * scaled_by   (problem_instance.py:226-241)
      // The sum over columns is associated the same way for EVERY tile geometry (so that a fused tail
      // in CTAs of any size and the stand-alone kernel agree bit for bit): each block of 32 consecutive
      // columns is reduced by the warp's xor tree, the block sums are added in ascending column order.
      contract(x, acc);
      float* rt = red + (size_t)tile * wpt * CPL * EPT * 2;   // [block = wt * CPL + c][EPT][2]
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
#pragma unroll
        for (int tb = 0; tb < EPT; ++tb) {
          float e1 = 0.f, e2 = 0.f;
          if (jok[c]) {
            const float xj = x[(size_t)jc[c] * EPT + tb];
            e1 = acc[c][tb] * xj;
            e2 = vj[c] * xj;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            e1 += __shfl_xor_sync(0xffffffffu, e1, o);
            e2 += __shfl_xor_sync(0xffffffffu, e2, o);
          }
          if (lane == 0) {
            rt[((wt * CPL + c) * EPT + tb) * 2] = e1;
            rt[((wt * CPL + c) * EPT + tb) * 2 + 1] = e2;
          }
        }
      }
      tile_sync();
      if (wt == 0 && lane < EPT && b0 + lane < b_end) {
        const int nblocks = (N + 31) / 32;
        float s1 = 0.f, s2 = 0.f;
        for (int m = 0; m < nblocks; ++m) {
          s1 += rt[(m * EPT + lane) * 2];
          s2 += rt[(m * EPT + lane) * 2 + 1];
        }
        const float sb = p.scaled_by_ptr ? *p.scaled_by_ptr : p.scaled_by;
        p.energy[b0 + lane] = 0.5f * (s1 * sb) + s2 * sb;
      }
    }
    tile_sync();
  }
}

// which body serves `p`, and its geometry for CTAs of `nwarps` warps: fills p.wpt / p.cpl / p.ld and
// returns the shared-memory floats needed with (q_in_smem = true) or without the matrix on chip
struct EpiGeometry {
  bool tiled;
  size_t floats_q, floats_noq;
};
__host__ __device__ inline EpiGeometry epi_geometry(EpiParams& p, int nwarps, bool allow_tiled = true) {
  EpiGeometry g;
  const int N = p.n;
  g.tiled = allow_tiled && p.pp != CCVM_PP_ADAM && N <= 1024;
  if (g.tiled) {
    int wpt = (N + 31) / 32;
    if (wpt > nwarps) wpt = nwarps;
    int cpl = (N + 32 * wpt - 1) / (32 * wpt);
    cpl = cpl <= 1 ? 1 : cpl == 2 ? 2 : 4;
    if ((size_t)cpl * 32 * wpt < (size_t)N) g.tiled = false;  // more than 4 columns per lane: warp body
    p.wpt = wpt;
    p.cpl = cpl;
  }
  if (g.tiled) {
    g.floats_q = epi_tile_floats(N, nwarps, p.wpt, true);
    g.floats_noq = epi_tile_floats(N, nwarps, p.wpt, false);
  } else {
    p.ld = N | 1;
    g.floats_q = epi_warp_floats(N, nwarps, true);
    g.floats_noq = epi_warp_floats(N, nwarps, false);
  }
  return g;
}

// runs the body chosen by epi_geometry (p.wpt != 0: tiled)
static __device__ __forceinline__ void epilogue_run(const EpiParams& p, float* esm, long long b_begin, long long b_end,
                                                    int cta_slot, int n_slots) {
  if (p.wpt > 0) {
    if (p.cpl == 1) epilogue_tile_body<1>(p, esm, b_begin, b_end, cta_slot, n_slots);
    else if (p.cpl == 2) epilogue_tile_body<2>(p, esm, b_begin, b_end, cta_slot, n_slots);
    else epilogue_tile_body<4>(p, esm, b_begin, b_end, cta_slot, n_slots);
  } else {
    epilogue_warp_body(p, esm, b_begin, b_end, cta_slot, n_slots);
  }
}

// ------------------------------------------------------------------------- solution statistics
// best = max_b(-E_b) (NaN propagates like torch.max), arg_best = its lowest index, counts[k] =
// #{b : gap_b <= thr_k}, gap_b = (optimal - (-E_b)) * 100 / |E_b| in fp32 like the reference
// (solution.py:113-136).  Block-level reduction over energies [b_begin, b_end); thread 0 returns true
// and holds the result in `out` (arg_best is the GLOBAL index; 0x7fffffff if every value was NaN).
struct StatsPartial {
  float best;
  int arg;
  int counts[7];
  int saw_nan;
};

static __device__ __noinline__ bool stats_block_reduce(const float* __restrict__ energy, long long b_begin, long long b_end,
                                                       float optimal, StatsPartial& out) {
  __shared__ float s_best[32];
  __shared__ int s_arg[32];
  __shared__ int s_cnt[7];
  __shared__ int s_nan;
  const float thr[7] = {0.1f, 1.f, 2.f, 3.f, 4.f, 5.f, 10.f};
  if (threadIdx.x < 7) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  float best = -INFINITY;
  int arg = 0x7fffffff, cnt[7] = {0, 0, 0, 0, 0, 0, 0};
  bool saw_nan = false;
  for (long long b = b_begin + threadIdx.x; b < b_end; b += blockDim.x) {
    const float val = -energy[b];
    if (val != val) saw_nan = true;
    if (val > best) {
      best = val;
      arg = (int)b;
    }
    const float gap = __fdiv_rn(__fmul_rn(__fsub_rn(optimal, val), 100.f), fabsf(val));
#pragma unroll
    for (int k = 0; k < 7; ++k) cnt[k] += (gap <= thr[k]) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
  }
  if (saw_nan) atomicOr(&s_nan, 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_best[warp] = best;
    s_arg[warp] = arg;
#pragma unroll
    for (int k = 0; k < 7; ++k) atomicAdd(&s_cnt[k], cnt[k]);
  }
  __syncthreads();
  if (threadIdx.x != 0) return false;
  float bb = s_best[0];
  int ba = s_arg[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (s_best[w] > bb || (s_best[w] == bb && s_arg[w] < ba)) {
      bb = s_best[w];
      ba = s_arg[w];
    }
  out.best = bb;
  out.arg = ba;
  out.saw_nan = s_nan;
  for (int k = 0; k < 7; ++k) out.counts[k] = s_cnt[k];
  return true;
}

__device__ __forceinline__ void stats_finalize(const StatsPartial& s, StatsOut* out) {
  out->best = s.saw_nan ? NAN : s.best;  // torch.max propagates NaN
  out->arg_best = s.arg == 0x7fffffff ? 0 : s.arg;
  for (int k = 0; k < 7; ++k) out->counts[k] = s.counts[k];
}

// order-preserving map float -> uint32 (larger float <-> larger key; no NaN expected)
__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Fused tail, statistics part: the CTA's partial (thread 0) is merged into the cross-CTA accumulators
// with one atomicMax on a packed (value, index) key and seven atomicAdds; the CTA that arrives last
// writes the result block.  `index_base` makes the trajectory index local to the instance.
__device__ __forceinline__ void stats_merge(const StatsPartial& s, StatsAccum* acc, FusedOut* out, unsigned int total_ctas,
                                            unsigned long long loop_ns, unsigned long long tail_ns) {
  if (s.arg != 0x7fffffff) {
    const unsigned long long key = ((unsigned long long)orderable(s.best) << 32) | (0xffffffffu - (uint32_t)s.arg);
    atomicMax(&acc->key, key);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (s.counts[k]) atomicAdd(&acc->counts[k], s.counts[k]);
  if (s.saw_nan) atomicOr(&acc->saw_nan, 1);
  atomicMax(&acc->loop_ns, loop_ns);
  atomicMax(&acc->tail_ns, tail_ns);
  __threadfence();
  const unsigned int prev = atomicAdd(&acc->done, 1u);
  if (prev + 1 == total_ctas) {
    __threadfence();
    const volatile StatsAccum* a = acc;
    const unsigned long long key = a->key;
    StatsPartial tot;
    tot.saw_nan = a->saw_nan;
    tot.best = key ? from_orderable((uint32_t)(key >> 32)) : -INFINITY;
    tot.arg = key ? (int)(0xffffffffu - (uint32_t)key) : 0x7fffffff;
    for (int k = 0; k < 7; ++k) tot.counts[k] = a->counts[k];
    stats_finalize(tot, &out->stats);
    out->ctas = total_ctas;
    out->loop_ns = a->loop_ns;
    out->tail_ns = a->tail_ns;
  }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace ccvm
