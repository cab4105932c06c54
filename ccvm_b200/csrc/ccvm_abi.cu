// C ABI of libccvm_b200.so (see include/ccvm_b200.h) and the small kernels around the
// persistent SDE kernel: schedule builder, fused epilogue (change of variables, projected
// GD / Adam post-processing, BoxQP energy), solution statistics, scaling factor, FP32 probe.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ccvm_b200.h"
#include "epilogue.cuh"
#include "sde_kernel_mma.cuh"
#include "sde_kernel_tc.cuh"
#include "sde_kernel_tmem.cuh"
#include "sde_launch.h"

using namespace ccvm;

// ------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

int ccvm::set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define fail(...) ::ccvm::set_error(__VA_ARGS__)

// NVTX range around an entry point (tracing hooks); header-only NVTX3, a no-op without a tool attached
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

extern "C" const char* ccvm_last_error(void) { return g_err; }
extern "C" int ccvm_abi_version(void) { return CCVM_ABI_VERSION; }

struct DeviceInfo {
  int device = -1, sms = 0, max_smem = 0;
  cudaMemPool_t pool = nullptr;
};
// Scratch comes from a PRIVATE stream-ordered pool per device (not the device's default pool, which
// other libraries in the process share): freed blocks stay cached up to CCVM_POOL_KEEP_MB (default
// 1024 MiB) across synchronisations -- the default pool hands memory back to the driver at every sync,
// which cost ~50 ms per host-buffer solve -- and anything above that is returned to the driver.
static int device_info(DeviceInfo& di) {
  static thread_local DeviceInfo cache[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(CCVM_E_CUDA, "unsupported device ordinal %d", dev);
  DeviceInfo& c = cache[dev];
  if (c.device != dev) {
    int sms = 0, smem = 0, major = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
      return fail(CCVM_E_CUDA, "ccvm_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
    static cudaMemPool_t pools[64];  // one per device for the whole process
    static std::mutex pools_mutex;
    std::lock_guard<std::mutex> lock(pools_mutex);
    if (!pools[dev]) {
      cudaMemPoolProps props;
      memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      CUDA_TRY(cudaMemPoolCreate(&pools[dev], &props));
      unsigned long long keep = 1024ull << 20;
      if (const char* e = getenv("CCVM_POOL_KEEP_MB")) keep = (unsigned long long)atoll(e) << 20;
      CUDA_TRY(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c.device = dev;
    c.sms = sms;
    c.max_smem = smem;
    c.pool = pools[dev];
  }
  di = c;
  return CCVM_OK;
}

// Stream-ordered scratch that is returned on EVERY exit path of an entry point (an early CUDA_TRY
// return must not leak it into the pool for the life of the process).
struct StreamBuf {
  void* p = nullptr;
  cudaStream_t st;
  explicit StreamBuf(cudaStream_t s) : st(s) {}
  StreamBuf(const StreamBuf&) = delete;
  StreamBuf& operator=(const StreamBuf&) = delete;
  ~StreamBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  cudaError_t alloc(size_t bytes) {
    DeviceInfo di;
    if (device_info(di) != CCVM_OK) return cudaErrorInvalidDevice;
    return cudaMallocFromPoolAsync(&p, bytes, di.pool, st);
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ------------------------------------------------------------------- schedule builder
// (SchedArgs / schedule_row: ccvm_common.cuh.)  Stand-alone kernels for the paths that do not
// evaluate the table inside the persistent kernel: the tcgen05 path and batched launches.
__global__ void build_schedule_kernel(SchedArgs a, float* __restrict__ out) {
  schedule_row(a, blockIdx.x * blockDim.x + threadIdx.x, out);
}

// ---- tiled paths (sde_kernel_tmem.cuh): Q slice from TMEM (n <= 128) or streamed from L2 (any n)
enum { PATH_TMEM = 0, PATH_GMEM = 1, PATH_TC = 3, PATH_HYB = 4, PATH_MMA = 5 };

// tcgen05 3xTF32 drift (sde_kernel_tc.cuh) where batch x N x N is a genuine dense GEMM:
// n > 256 and at least 1024 contraction rows (8 CTAs of 128 rows).  n = 256 itself is also served by
// the hybrid SIMT kernel, which keeps all SMs busy at small batches: measured crossover at ~10k rows
// (one tensor-core iteration of a 128-row CTA takes ~34 us at n = 256, latency-bound with a single
// output chunk; one hybrid wave of 1184 trajectories 4-7 us).  CCVM_TC=0 / 1 / 2 overrides the size
// rule (experiments and tests), evolution sampling stays on the SIMT path.
static bool tc_size_rule(const ccvm_solve_desc& d) {
  const long long rows = (long long)d.batch * (d.solver == CCVM_SOLVER_DL ? 2 : 1);
  if (d.n > 256) return rows >= 1024;
  return d.n == 256 && rows >= 10240;
}

static bool tc_eligible(const ccvm_solve_desc& d) {
  if (d.n > TC_MAX_CHUNKS * TC_BN || d.evolution_step > 0) return false;
  if (const char* e = getenv("CCVM_TC")) {
    if (e[0] == 'a') return tc_size_rule(d);  // "auto"
    return atoi(e) != 0;  // 0: never, 1 / 2: always (1 = single-CTA kernel, 2 = CTA-pair kernel)
  }
  return tc_size_rule(d);
}

// Small-n tensor-core kernel (sde_kernel_mma.cuh): single-instance launches in production (Philox) mode whose batch
// gives every SM a few dozen trajectories.  There the register-tile kernels are issue-bound on the contraction's
// FFMA2 stream; the tensor core takes it off the SIMT pipes.  Its iteration is latency-bound (~0.6-0.7 us: mbarrier ->
// MMA -> commit -> tcgen05.ld per warpgroup) and hardly depends on the batch up to one wave of CTAs, while the tiled
// kernel's time grows with the trajectories per SM and with n^2: measured crossovers (profiles/r2y_*, r2z_batch_crossover.txt;
// tensor-core / tiled ms at T = 1500):
//   n = 40:  B = 3000  Langevin 0.98 / 0.66, DL-adam 1.41 / 1.53;   B = 4096  0.97 / 1.11, 1.53 / 2.36
//   n = 70:  B = 2048  Langevin 1.03 / 0.90, DL-adam 1.67 / 2.14;   B = 3000  1.06 / 1.58, 2.01 / 3.19;  B = 4096  1.07 / 1.59, 2.00 / 3.22
//   n = 128: B = 2048  Langevin 1.25 / 2.59, DL-adam 2.23 / 4.94;   B = 4096  1.42 / 5.10, 3.61 / 9.88
// Below ~40 variables the handshake is longer than the whole tiled iteration.  Noise replay, evolution sampling and
// batched (many-instance) launches stay on the tiled kernels.  CCVM_MMA=0 / 1 overrides the rule ("1": whenever the
// kernel can run).
static bool mma_eligible(const ccvm_solve_desc& d) {
  if (d.n > MMA_N_MAX || d.rng_mode != CCVM_RNG_PHILOX || d.evolution_step > 0) return false;
  if (const char* e = getenv("CCVM_MMA")) {
    if (e[0] != 'a') return atoi(e) != 0;
  }
  if (d.n < 40) return false;
  // two M tiles (128 < n <= 192): twice the MMAs per iteration, but the hybrid FP32 kernel they replace needs a wave of
  // CTAs per ~590 trajectories (3-6 ms each at T = 1500) -- measured (profiles/r2z_two_m_tiles.txt): DL wins from the smallest
  // batches on, the K = 1 loops from the hybrid kernel's second wave
  if (d.n > 128) return d.batch >= (d.solver == CCVM_SOLVER_DL ? 256 : 640);
  const int min_batch = (d.solver == CCVM_SOLVER_DL || d.n >= 96) ? 2048 : d.n >= 56 ? 2560 : 3584;
  return d.batch >= min_batch;
}

static int choose_path(const ccvm_solve_desc& d, bool single_launch = false) {
  if (tc_eligible(d)) return PATH_TC;
  if (single_launch && mma_eligible(d)) return PATH_MMA;
  if (d.n <= 128 && getenv("CCVM_NO_TMEM") == nullptr) return PATH_TMEM;
  // 128 < n <= 256: first 128 rows of every Q slice in TMEM, the rest in shared memory
  if (d.n <= 4 * 64 && getenv("CCVM_NO_TMEM") == nullptr && getenv("CCVM_NO_HYB") == nullptr) return PATH_HYB;
  return PATH_GMEM;
}

// `share_hint` > 0 overrides the trajectories-per-SM estimate (batched launches plan every
// instance against the load of the whole batch, not its own).
static int plan_tmem(const ccvm_solve_desc& d, const DeviceInfo& di, int path, TmemPlan& P, int share_hint = 0,
                     bool batched = false) {
  const int K = d.solver == CCVM_SOLVER_DL ? 2 : 1, RW = 2 * K;
  const int cg = (d.n + 3) / 4, np = 4 * cg;
  const bool tm = path == PATH_TMEM || path == PATH_HYB;    // Q slices (partly) TMEM resident
  const int max_threads = tm ? 256 : 512;
  const int lanes = tm ? 128 : max_threads;   // threads one group may span
  if (cg > lanes) return fail(CCVM_E_TOO_LARGE, "n=%d exceeds the tiled SIMT path (n <= %d)", d.n, 4 * lanes);
  const int rg_max = lanes / cg;
  const int share = share_hint > 0 ? share_hint : (d.batch + di.sms - 1) / di.sms;
  const int pairs = (share + 1) / 2;
  auto round32 = [](int x) { return ((x + 31) / 32) * 32; };
  int ng = 1, rg = pairs < rg_max ? pairs : rg_max;
  if (pairs > rg_max) {
    // a second, independently synchronised group if the CTA has room for it
    const int rg2 = (pairs + 1) / 2 < rg_max ? (pairs + 1) / 2 : rg_max;
    const int gt2 = tm ? 128 : round32(rg2 * cg);
    if (2 * gt2 <= max_threads) {
      ng = 2;
      rg = rg2;
    }
  }
  if (const char* e = getenv("CCVM_NG")) {
    const int v = atoi(e);
    if (v == 1 || v == 2) ng = v;
  }
  if (const char* e = getenv("CCVM_RG")) {
    const int v = atoi(e);
    if (v > 0) rg = v;
  }
  if (rg > rg_max) rg = rg_max;
  if (rg < 1) rg = 1;
  if (ng == 2 && !tm && 2 * round32(rg * cg) > max_threads) ng = 1;
  // in-loop noise generation (sde_kernel_tmem.cuh, PIPE): Philox mode with enough Q chunks to hide it in
  // small compile-time column-group variants (CG = 5, 8: all noise quanta unpinned at the top of the
  // iteration; every tile has them), single and batched launches (bucketed by column-group count)
  const bool cgc_off = getenv("CCVM_NO_CGC") != nullptr || (batched && getenv("CCVM_NO_BATCH_CGC") != nullptr);
  const bool small_cgc = path == PATH_TMEM && d.rng_mode == CCVM_RNG_PHILOX && (cg == 5 || cg == 8) && !cgc_off;
  const bool pipe = getenv("CCVM_NO_PIPE") == nullptr &&
                    (small_cgc ||
                     (d.solver == CCVM_SOLVER_DL ? pipe_ok<SOLVER_DL>(cg, d.rng_mode == CCVM_RNG_PHILOX)
                                                 : pipe_ok<SOLVER_LV>(cg, d.rng_mode == CCVM_RNG_PHILOX)));
  int xs = 0, xmask = 31;
  const size_t tail = path == PATH_HYB ? (size_t)(np - 4 * HYB_TMEM_CHUNKS) * HYB_LD : 0;  // floats
  const bool fixed_xs = pipe && tm;  // compile-time panel stride, no row rotation
  // fixed-stride panels of the K = 1 solvers in the hybrid kernel hold two k rows per panel row
  // (sde_kernel_tmem.cuh, KP)
  const bool cgc_variant = fixed_xs && path == PATH_TMEM && !cgc_off && (cg == 10 || cg == 13 || cg == 15 || cg == 18);
  const int kp = (fixed_xs && K == 1 && path == PATH_HYB) ? 2 : 1;
  // DL + Adam parks its second moments in shared memory (sde_kernel_tmem.cuh, VSMEM): 64 B per thread
  const size_t vsm = (fixed_xs && path == PATH_TMEM && d.solver == CCVM_SOLVER_DL && d.algorithm == CCVM_ALG_ADAM)
                         ? (size_t)256 * 16 : 0;  // floats
  auto smem_of = [&](int xs_) {
    return ((size_t)2 * np + (size_t)ng * 2 * (np / kp) * xs_ + tail + vsm) * sizeof(float);
  };
  const int pipe_xs = path == PATH_HYB ? HYB_PIPE_XS : TMEM_PIPE_XS;
  if (fixed_xs && RW * kp * rg > pipe_xs) rg = pipe_xs / (RW * kp);  // (DL at CG = 5: at most 17 pairs per row)
  if (fixed_xs && (RW * kp * rg > pipe_xs || smem_of(pipe_xs) > (size_t)di.max_smem))
    return fail(CCVM_E_INVALID, "internal: the fixed state-panel stride does not fit (n=%d rg=%d)", d.n, rg);
  for (; !fixed_xs;) {
    xmask = 31;
    xs = ((RW * rg + 32 + 3) / 4) * 4;
    if (smem_of(xs) <= (size_t)di.max_smem) break;
    xmask = 0;  // drop the bank-spreading slack before giving up trajectories
    xs = ((RW * rg + 3) / 4) * 4;
    if (smem_of(xs) <= (size_t)di.max_smem) break;
    if (ng == 2) { ng = 1; continue; }
    if (rg == 1) return fail(CCVM_E_TOO_LARGE, "n=%d: the state panel does not fit shared memory", d.n);
    --rg;
  }
  int tcols = 32;
  while (tcols < 4 * np) tcols *= 2;
  memset(&P.L, 0, sizeof(P.L));
  P.L.rg = rg;
  P.L.ng = ng;
  P.L.gt = tm ? (ng > 1 ? 128 : round32(rg * cg)) : round32(rg * cg);
  if (path == PATH_HYB) tcols = 512;
  P.L.tcols = tcols;
  if (fixed_xs) {
    xs = pipe_xs;
    xmask = 0;
  }
  P.L.xs = xs;
  P.L.xmask = xmask;
  P.L.pipe = pipe ? 1 : 0;
  P.L.tcols = tcols;
  P.L.phase_ns = 0;
  if (const char* e = getenv("CCVM_PHASE_NS")) P.L.phase_ns = atoi(e);
  P.qsrc = path == PATH_TMEM ? QSRC_TMEM : path == PATH_HYB ? QSRC_HYB : QSRC_GMEM;
  P.cg = cg;
  P.cgc = (cgc_variant || (small_cgc && pipe)) ? cg : 0;
  // N = 30, 50, 70: the variant that leaves out the two padding rows of the last chunk of four
  P.ktail = (P.cgc == 8 || P.cgc == 13 || P.cgc == 18) && d.n == 4 * cg - 2 && getenv("CCVM_NO_KTAIL") == nullptr ? 2 : 0;
  P.threads = ng * P.L.gt;
  P.ctas = (d.batch + ng * 2 * rg - 1) / (ng * 2 * rg);
  P.smem = smem_of(xs);
  return CCVM_OK;
}

// Launch geometry of the small-n tensor-core kernel: a CTA advances 4 nbp trajectories (two warpgroups of nbp pairs,
// nbp <= 8): the smallest nbp that needs the fewest waves of CTAs (B = 4096 on 148 SMs: nbp = 7, 147 CTAs).
static void plan_mma(const ccvm_solve_desc& d, const DeviceInfo& di, MmaPlan& P) {
  auto waves = [&](int nbp) {
    const long long ctas = ((long long)d.batch + 4 * nbp - 1) / (4 * nbp);
    return (ctas + di.sms - 1) / di.sms;
  };
  // a lane owns at most `ipl_max` (variable, pair) items: the per-item state has to stay in registers (compiled variants:
  // 2 ... 8 for every tile, 9 ... 11 for the tiles with the smallest state)
  const bool light = d.algorithm != CCVM_ALG_ADAM || d.solver == CCVM_SOLVER_LANGEVIN || d.solver == CCVM_SOLVER_PUMPED_LANGEVIN;
  int ipl_max = light ? MMA_IPL_MAX_LIGHT : MMA_IPL_MAX;
  if (const char* e = getenv("CCVM_MMA_IPL_MAX")) {   // tuning aid; never above what the tile has compiled in
    const int v = atoi(e);
    if (v >= 2 && v < ipl_max) ipl_max = v;
  }
  const int per_quadrant = d.n <= 128 ? (d.n + 3) / 4 : 32 + (d.n - 128 + 3) / 4;
  int nbp_max = 32 * ipl_max / per_quadrant;
  if (nbp_max > 8) nbp_max = 8;
  if (nbp_max < 1) nbp_max = 1;
  P.nbp = nbp_max;
  for (int nbp = nbp_max - 1; nbp >= 1; --nbp)
    if (waves(nbp) <= waves(P.nbp)) P.nbp = nbp;
  if (const char* e = getenv("CCVM_MMA_NBP")) {
    const int v = atoi(e);
    if (v >= 1 && v <= nbp_max) P.nbp = v;
  }
  P.kd = ((d.n + 15) / 16) * 16;
  P.mt = d.n > 128 ? 2 : 1;
  P.tcols = P.mt == 1 ? 256 : 512;   // mt (64 accumulator columns + kd columns of A)
  P.ipl = (per_quadrant * P.nbp + 31) / 32;
  if (P.ipl < 2) P.ipl = 2;
  P.threads = MMA_THREADS;
  // the two update warpgroups start half an iteration apart (sde_kernel_mma.cuh; profiles/r2z_issuer_protocol.txt);
  // CCVM_MMA_STAGGER=0 releases them together
  P.stagger = 1;
  if (const char* e = getenv("CCVM_MMA_STAGGER")) P.stagger = atoi(e) != 0;
  P.ctas = (int)(((long long)d.batch + 4 * P.nbp - 1) / (4 * P.nbp));
  P.smem = mma_loop_smem_bytes(P.mt);
}

// Qs[k][j] = -alpha_k alpha_j Q_kj, zero padded to NP x NP (QSRC_GMEM operand)
__global__ void scale_q_kernel(const float* __restrict__ q, const float* __restrict__ svec, float s, float a_half,
                               int n, int np, float* __restrict__ qs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= np * np) return;
  const int k = idx / np, j = idx - k * np;
  float val = 0.f;
  if (k < n && j < n) {
    const float ak = a_half / (svec ? svec[k] : s), aj = a_half / (svec ? svec[j] : s);
    val = -ak * aj * q[k * n + j];
  }
  qs[idx] = val;
}

// ------------------------------------------------------------------ tensor-core path
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// fp32 [rows][cols] row-major; box = box_rows x box_cols with box_cols * 4 bytes == the swizzle span
static int make_map_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(CCVM_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CCVM_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CCVM_OK;
}

static int tc_version() {
  if (const char* e = getenv("CCVM_TC")) {
    const int v = atoi(e);
    if (v == 1 || v == 2) return v;
  }
  return 2;
}

static void plan_tc(const ccvm_solve_desc& d, TcPlan& P) {
  const int K = d.solver == CCVM_SOLVER_DL ? 2 : 1;
  P.version = tc_version();
  const int row_block = P.version == 2 ? 2 * TC_BM : TC_BM;
  P.np = ((d.n + TC_BN - 1) / TC_BN) * TC_BN;
  P.rows = K * d.batch;
  P.rows_p = ((P.rows + row_block - 1) / row_block) * row_block;
  const bool adam = d.algorithm == CCVM_ALG_ADAM;
  P.n_aux = (d.solver == CCVM_SOLVER_MF ? 2 : 0) + (adam ? 2 : 0);
  P.ctas = P.rows_p / TC_BM;
  // Few row blocks (Langevin-type loops have one row per trajectory: 64 CTAs at B = 8192): two or four CTA pairs
  // share a block of 256 rows and split its output chunks, exchanging the new state through L2 with per-chunk
  // flags.  They wait for each other, so every CTA must be resident: col_split x ctas <= the SM count.
  P.col_split = 1;
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int nc = P.np / TC_BN;
    if (P.version == 2 && getenv("CCVM_TC_NO_SPLIT") == nullptr)
      for (int cs = 4; cs >= 2; cs >>= 1)
        // (2-CTA clusters cannot use the odd SM of a GPC: leave a margin so that ALL clusters are resident)
        if (nc % cs == 0 && cs * P.ctas <= sms - 12) {
          P.col_split = cs;
          break;
        }
  }
  P.ctas *= P.col_split;
  P.smem = P.version == 2 ? (size_t)T2_SMEM_BYTES
                          : 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)2 * P.np * sizeof(float);
}

// `p` carries everything but the launch geometry; sched is already being built on `st`
static int solve_tc(const ccvm_solve_desc* d, SdeParams& p, cudaStream_t st) {
  TcPlan P;
  plan_tc(*d, P);
  const size_t plane = (size_t)P.rows_p * P.np;
  const size_t n_flags = (size_t)(P.rows_p / TC_BM) * (P.np / TC_BN);
  const size_t floats = 4 * plane + (size_t)P.n_aux * plane + 2 * (size_t)P.np * P.np + 2 * (size_t)P.np + n_flags;
  StreamBuf scratch_buf(st);
  CUDA_TRY(scratch_buf.alloc(floats * sizeof(float)));
  float* scratch = scratch_buf.as<float>();
  TcParams tc;
  tc.col_split = P.col_split;
  tc.chunk_flags = reinterpret_cast<unsigned int*>(scratch + floats - n_flags);
  CUDA_TRY(cudaMemsetAsync(tc.chunk_flags, 0, n_flags * sizeof(unsigned int), st));
  tc.xh = scratch;
  tc.xl = tc.xh + 2 * plane;
  tc.aux = tc.xl + 2 * plane;
  float* qt_hi = tc.aux + (size_t)P.n_aux * plane;
  float* qt_lo = qt_hi + (size_t)P.np * P.np;
  float* hvec = qt_lo + (size_t)P.np * P.np;
  float* svec = hvec + P.np;
  tc.hvec = hvec;
  tc.svec = svec;
  tc.np = P.np;
  tc.rows = P.rows;
  tc.rows_p = P.rows_p;
  tc_prepare_q_kernel<<<P.np, 256, 0, st>>>(p.q, p.v, p.drift_s_vec, p.drift_s, p.clamp_s_vec, p.clamp_s, p.a_half,
                                            p.b_half, p.n, P.np, qt_hi, qt_lo, hvec, svec);
  int rc = CCVM_OK;
  if (cudaGetLastError() != cudaSuccess) rc = fail(CCVM_E_CUDA, "tc_prepare_q_kernel launch failed");
  TcMaps M;
  if (P.version == 2) {
    if (!rc) rc = make_map_2d(&M.xh, tc.xh, 2ull * P.rows_p, P.np, TC_BM, T2_BK);
    if (!rc) rc = make_map_2d(&M.xl, tc.xl, 2ull * P.rows_p, P.np, TC_BM, T2_BK);
    if (!rc) rc = make_map_2d(&M.qh, qt_hi, P.np, P.np, TC_BM, T2_BK);
    if (!rc) rc = make_map_2d(&M.ql, qt_lo, P.np, P.np, TC_BM, T2_BK);
    if (!rc) rc = make_map_2d(&M.oh, tc.xh, 2ull * P.rows_p, P.np, TC_BM, 16);
    if (!rc) rc = make_map_2d(&M.ol, tc.xl, 2ull * P.rows_p, P.np, TC_BM, 16);
  } else {
    if (!rc) rc = make_map_2d(&M.xh, tc.xh, 2ull * P.rows_p, P.np, TC_BM, TC_BK);
    if (!rc) rc = make_map_2d(&M.xl, tc.xl, 2ull * P.rows_p, P.np, TC_BM, TC_BK);
    if (!rc) rc = make_map_2d(&M.qh, qt_hi, P.np, P.np, TC_BN, TC_BK);
    if (!rc) rc = make_map_2d(&M.ql, qt_lo, P.np, P.np, TC_BN, TC_BK);
  }
  if (!rc) {
    const bool adam = d->algorithm == CCVM_ALG_ADAM;
#define LAUNCH_TC(S, A) rc = launch_tc<S, A>(p, tc, P, M, st)
    CCVM_DISPATCH_TILE(d->solver, adam, LAUNCH_TC)
#undef LAUNCH_TC
  }
  return rc;
}

static int validate_solve(const ccvm_solve_desc* d) {
  if (!d) return fail(CCVM_E_INVALID, "null descriptor");
  if (d->solver < 0 || d->solver > 3) return fail(CCVM_E_INVALID, "unknown solver id %d", d->solver);
  if (d->algorithm != CCVM_ALG_ORIGINAL && d->algorithm != CCVM_ALG_ADAM)
    return fail(CCVM_E_INVALID, "unknown algorithm id %d", d->algorithm);
  if (d->n < 1 || d->batch < 1 || d->iterations < 1)
    return fail(CCVM_E_INVALID, "n, batch and iterations must be >= 1 (got %d, %d, %d)", d->n, d->batch, d->iterations);
  if (!d->q || !d->v || !d->out0) return fail(CCVM_E_INVALID, "q, v and out0 are required");
  if (d->solver == CCVM_SOLVER_DL && !d->out1) return fail(CCVM_E_INVALID, "DL needs out1 (s)");
  if (d->solver == CCVM_SOLVER_MF && (!d->out1 || !d->out2)) return fail(CCVM_E_INVALID, "MF needs out1 and out2");
  if (!(d->upper > d->lower)) return fail(CCVM_E_INVALID, "solution bounds must satisfy lower < upper");
  if (d->rng_mode == CCVM_RNG_REPLAY) {
    if (!d->noise) return fail(CCVM_E_INVALID, "replay mode needs a noise tensor");
    if (d->noise_batch < d->traj_base + d->batch) return fail(CCVM_E_INVALID, "noise tensor is too small for the batch");
  } else if (d->rng_mode != CCVM_RNG_PHILOX) {
    return fail(CCVM_E_INVALID, "unknown rng mode %d", d->rng_mode);
  } else if (d->traj_base & 1) {
    // the SIMT kernels key one noise stream per PAIR of trajectories (2p, 2p+1)
    return fail(CCVM_E_INVALID, "traj_base must be even in Philox mode (got %lld): split batches at even indices",
                (long long)d->traj_base);
  }
  if (d->evolution_step > 0 && (!d->samples || d->num_samples < 1))
    return fail(CCVM_E_INVALID, "evolution sampling needs a samples buffer");
  return CCVM_OK;
}

// Resolve which S the drift and the clamp see:
//   DL  _solve      drift S = sqrt(pump-1) if pump>1 else 1 ; final clamp = the S handed in
//   DL  _solve_adam S := sqrt(pump-1) if pump>1 (drift AND clamp) else the S handed in
//   MF / Langevin / PumpedLangevin: the S handed in everywhere
static void resolve_saturation(const ccvm_solve_desc& d, SdeParams& p) {
  p.drift_s_vec = d.s_vec;
  p.clamp_s_vec = d.s_vec;
  p.drift_s = (float)d.s;
  p.clamp_s = (float)d.s;
  if (d.solver == CCVM_SOLVER_DL) {
    const bool pumped = d.pump > 1.0;
    const double sp = pumped ? sqrt(d.pump - 1.0) : 1.0;
    if (d.algorithm == CCVM_ALG_ORIGINAL) {
      p.drift_s_vec = nullptr;
      p.drift_s = (float)sp;
    } else if (pumped) {
      p.drift_s_vec = p.clamp_s_vec = nullptr;
      p.drift_s = p.clamp_s = (float)sp;
    }
  }
}

static void fill_params(const ccvm_solve_desc* d, const float* sched, int cg, SdeParams& p) {
  memset(&p, 0, sizeof(p));
  p.q = d->q;
  p.v = d->v;
  resolve_saturation(*d, p);
  p.sched = sched;
  p.noise = d->rng_mode == CCVM_RNG_REPLAY ? d->noise : nullptr;
  p.noise_batch = d->noise_batch;
  p.traj_base = d->traj_base;
  p.out0 = d->out0;
  p.out1 = d->out1;
  p.out2 = d->out2;
  p.samples = d->evolution_step > 0 ? d->samples : nullptr;
  p.evolution_step = d->evolution_step > 0 ? d->evolution_step : 0;
  p.num_samples = d->num_samples;
  p.n = d->n;
  p.batch = d->batch;
  p.iterations = d->iterations;
  p.cg = cg;
  p.a_half = (float)((d->upper - d->lower) * 0.5);
  p.b_half = (float)((d->upper + d->lower) * 0.5);
  p.dt = (float)d->dt;
  p.fs = (float)d->feedback_scale;
  p.g2 = (float)(d->g * d->g);
  p.sig = (float)(d->sigma * sqrt(d->dt));
  p.dtfs = (float)(d->dt * d->feedback_scale);
  p.beta1 = (float)d->beta1;
  p.beta2 = (float)d->beta2;
  p.omb1 = (float)(1.0 - d->beta1);
  p.omb2 = (float)(1.0 - d->beta2);
  p.adam_alpha = (float)d->alpha;
  p.add_assign = d->add_assign != 0;
  p.beta2_is_one = d->beta2 == 1.0;
  p.seed_lo = (uint32_t)d->seed;
  p.seed_hi = (uint32_t)(d->seed >> 32);
  p.off_lo = (uint32_t)d->offset;
  p.off_hi = (uint32_t)(d->offset >> 32);
}

static SchedArgs sched_args(const ccvm_solve_desc* d) {
  SchedArgs sa;
  sa.solver = d->solver;
  sa.adam = d->algorithm == CCVM_ALG_ADAM;
  sa.iterations = d->iterations;
  sa.flag = d->solver == CCVM_SOLVER_LANGEVIN ? 0 : (d->pump_rate_flag != 0);
  sa.pump = d->pump;
  sa.dt = d->dt;
  sa.noise_ratio = d->noise_ratio;
  sa.j = d->j;
  sa.fs = d->feedback_scale;
  sa.g = d->g;
  sa.beta1 = d->beta1;
  sa.beta2 = d->beta2;
  return sa;
}

extern "C" int ccvm_query_launch(const ccvm_solve_desc* d, int32_t* info5) {
  int rc = validate_solve(d);
  if (rc) return rc;
  DeviceInfo di;
  if ((rc = device_info(di))) return rc;
  const int path = choose_path(*d, true);
  const bool adam = d->algorithm == CCVM_ALG_ADAM;
  int regs = -1;
  if (path == PATH_MMA) {
    MmaPlan P;
    plan_mma(*d, di, P);
    info5[0] = P.threads;
    info5[1] = P.ctas;
    info5[2] = 4 * P.nbp;
    info5[3] = (int)P.smem;
#define REGS_MMA(S, A) regs = regs_mma<S, A>(P.ipl, P.mt)
    CCVM_DISPATCH_TILE(d->solver, adam, REGS_MMA)
#undef REGS_MMA
    info5[4] = regs;
    return CCVM_OK;
  }
  if (path == PATH_TC) {
    TcPlan P;
    plan_tc(*d, P);
    info5[0] = TC_THREADS;
    info5[1] = P.ctas;
    info5[2] = TC_BM / (d->solver == CCVM_SOLVER_DL ? 2 : 1) / P.col_split;
    info5[3] = (int)P.smem;
#define REGS_TC(S, A) regs = regs_tc<S, A>(P.version)
    CCVM_DISPATCH_TILE(d->solver, adam, REGS_TC)
#undef REGS_TC
    info5[4] = regs;
    return CCVM_OK;
  }
  TmemPlan P;
  if ((rc = plan_tmem(*d, di, path, P))) return rc;
  info5[0] = P.threads;
  info5[1] = P.ctas;
  info5[2] = P.L.ng * 2 * P.L.rg;
  info5[3] = (int)P.smem;
#define REGS_TMEM(S, A) regs = regs_tmem<S, A>(P.qsrc)
  CCVM_DISPATCH_TILE(d->solver, adam, REGS_TMEM)
#undef REGS_TMEM
  info5[4] = regs;
  return CCVM_OK;
}

// ------------------------------------------------------------------------- epilogue (stand-alone)
// The bodies live in epilogue.cuh (shared with the tail of the persistent kernels).
__global__ void __launch_bounds__(EPI_WARPS * 32) epilogue_kernel(const EpiParams p) {
  extern __shared__ __align__(16) float esm[];
  epilogue_run(p, esm, 0, p.batch, blockIdx.x, gridDim.x);
}

static int epi_params_from_desc(const ccvm_epilogue_desc* d, EpiParams& p) {
  if (!d) return fail(CCVM_E_INVALID, "null descriptor");
  if (d->post_processor < CCVM_PP_NONE || d->post_processor > CCVM_PP_ADAM)
    return fail(CCVM_E_INVALID, "unknown post-processor id %d", d->post_processor);
  memset(&p, 0, sizeof(p));
  p.q = d->q;
  p.v = d->v;
  p.state = d->state;
  p.m1vec = d->map1_scale_vec;
  p.m2vec = d->map2_scale_vec;
  p.pv = d->problem_variables;
  p.energy = d->energy;
  p.n = d->n;
  p.batch = d->batch;
  p.map1 = d->apply_map1 != 0;
  p.map2 = d->apply_map2 != 0;
  p.pp = d->post_processor;
  p.pp_iters = d->pp_iterations;
  p.m1s = (float)d->map1_scale;
  p.m1o = (float)d->map1_shift;
  p.m2s = (float)d->map2_scale;
  p.m2o = (float)d->map2_shift;
  p.step = (float)d->pp_step;
  p.lo = (float)d->pp_lower;
  p.hi = (float)d->pp_upper;
  p.scaled_by = (float)d->scaled_by;
  p.scaled_by_ptr = d->scaled_by_dev;
  return CCVM_OK;
}

// Picks the body (tiles of 8 trajectories, or a warp per trajectory for the Adam post-processor),
// decides whether Q is staged in shared memory and returns the dynamic shared memory `p` needs in
// CTAs of `nwarps` warps.
static int plan_epilogue(EpiParams& p, int nwarps, int max_smem, size_t& smem) {
  bool allow_tiled = getenv("CCVM_EPILOGUE_LEGACY") == nullptr;
  for (;;) {
    p.wpt = p.cpl = 0;
    const EpiGeometry g = epi_geometry(p, nwarps, allow_tiled);
    p.q_in_smem = g.floats_q * sizeof(float) <= (size_t)max_smem;
    smem = (p.q_in_smem ? g.floats_q : g.floats_noq) * sizeof(float);
    if (smem <= (size_t)max_smem) return CCVM_OK;
    if (!g.tiled) return fail(CCVM_E_TOO_LARGE, "n=%d too large for the epilogue", p.n);
    allow_tiled = false;  // the tile buffers alone do not fit: warp per trajectory
  }
}

static int run_epilogue(const EpiParams& p0, cudaStream_t st) {
  NvtxRange range("ccvm_epilogue");
  EpiParams p = p0;
  DeviceInfo di;
  int rc = device_info(di);
  if (rc) return rc;
  size_t smem = 0;
  if ((rc = plan_epilogue(p, EPI_WARPS, di.max_smem, smem))) return rc;
  long long grid;
  if (p.wpt > 0) {
    const int per_cta = (EPI_WARPS / p.wpt) * EPT;
    grid = ((long long)p.batch + per_cta - 1) / per_cta;
    if (grid > 4LL * di.sms) grid = 4LL * di.sms;
  } else {
    grid = (p.batch + EPI_WARPS - 1) / EPI_WARPS;
    if (grid > 2LL * di.sms) grid = 2LL * di.sms;
  }
  CUDA_TRY(cudaFuncSetAttribute(epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  epilogue_kernel<<<(unsigned)grid, EPI_WARPS * 32, smem, st>>>(p);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

extern "C" int ccvm_epilogue(const ccvm_epilogue_desc* d, void* stream) {
  EpiParams p;
  int rc = epi_params_from_desc(d, p);
  if (rc) return rc;
  if (d->n < 1 || d->batch < 1) return fail(CCVM_E_INVALID, "n and batch must be >= 1");
  if (!d->q || !d->v || !d->state) return fail(CCVM_E_INVALID, "q, v and state are required");
  return run_epilogue(p, (cudaStream_t)stream);
}

extern "C" int ccvm_compute_energy(const float* x, const float* q, const float* v, double scaled_by,
                                   int32_t batch, int32_t n, float* energy, void* stream) {
  if (!x || !q || !v || !energy || batch < 1 || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_compute_energy");
  EpiParams p;
  memset(&p, 0, sizeof(p));
  p.q = q;
  p.v = v;
  p.state = x;
  p.energy = energy;
  p.n = n;
  p.batch = batch;
  p.scaled_by = (float)scaled_by;
  return run_epilogue(p, (cudaStream_t)stream);
}

extern "C" int ccvm_postprocess_grad_descent(float* x, const float* q, const float* v, int32_t batch,
                                             int32_t n, int32_t iterations, double step_size, double lower,
                                             double upper, void* stream) {
  if (!x || !q || !v || batch < 1 || n < 1 || iterations < 0)
    return fail(CCVM_E_INVALID, "bad argument to ccvm_postprocess_grad_descent");
  EpiParams p;
  memset(&p, 0, sizeof(p));
  p.q = q;
  p.v = v;
  p.state = x;
  p.pv = x;
  p.n = n;
  p.batch = batch;
  p.pp = CCVM_PP_GRAD_DESCENT;
  p.pp_iters = iterations;
  p.step = (float)step_size;
  p.lo = (float)lower;
  p.hi = (float)upper;
  return run_epilogue(p, (cudaStream_t)stream);
}

extern "C" int ccvm_postprocess_adam(float* x, const float* q, const float* v, int32_t batch, int32_t n,
                                     double lr, double lower, double upper, void* stream) {
  if (!x || !q || !v || batch < 1 || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_postprocess_adam");
  EpiParams p;
  memset(&p, 0, sizeof(p));
  p.q = q;
  p.v = v;
  p.state = x;
  p.pv = x;
  p.n = n;
  p.batch = batch;
  p.pp = CCVM_PP_ADAM;
  p.pp_iters = 1;
  p.step = (float)lr;
  p.lo = (float)lower;
  p.hi = (float)upper;
  return run_epilogue(p, (cudaStream_t)stream);
}

// -------------------------------------------------------------------- solution stats
__global__ void __launch_bounds__(1024) stats_kernel(const float* __restrict__ energy, int batch, float optimal,
                                                     StatsOut* out) {
  StatsPartial sp;
  if (stats_block_reduce(energy, 0, batch, optimal, sp)) stats_finalize(sp, out);
}

// one block per instance: energies concatenated, instance i = energy[offsets[i] .. offsets[i+1])
__global__ void __launch_bounds__(1024) stats_batch_kernel(const float* __restrict__ energy,
                                                           const long long* __restrict__ offsets,
                                                           const float* __restrict__ optimal, StatsOut* out) {
  const int i = blockIdx.x;
  const long long lo = offsets[i];
  StatsPartial sp;
  if (stats_block_reduce(energy + lo, 0, offsets[i + 1] - lo, optimal[i], sp)) stats_finalize(sp, out + i);
}

extern "C" int ccvm_solution_stats_batch(const float* energy, const int64_t* offsets, const float* optimal_values,
                                         int32_t count, void* result, void* stream) {
  if (!energy || !offsets || !optimal_values || !result || count < 1)
    return fail(CCVM_E_INVALID, "bad argument to ccvm_solution_stats_batch");
  stats_batch_kernel<<<count, 1024, 0, (cudaStream_t)stream>>>(energy, (const long long*)offsets, optimal_values,
                                                              (StatsOut*)result);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

extern "C" int ccvm_solution_stats(const float* energy, int32_t batch, double optimal_value, void* result,
                                   void* stream) {
  if (!energy || !result || batch < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_solution_stats");
  stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(energy, batch, (float)optimal_value, (StatsOut*)result);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// --------------------------------------------------------------------------- one Solver.__call__
// Fills the tail of a persistent kernel from an epilogue descriptor: the epilogue runs on the solve's
// own q / v / final state (`epi`'s q, v, state, n and batch are ignored) in CTAs of `threads` threads
// whose loop needs `loop_smem` bytes; `smem` returns the dynamic shared memory of the launch.
static int plan_fused_tail(const ccvm_solve_desc& d, const ccvm_epilogue_desc* epi, double optimal, int threads,
                           size_t loop_smem, int max_smem, unsigned ctas, StatsAccum* accum, FusedOut* out,
                           FusedTail& f, size_t& smem) {
  smem = loop_smem;
  f.epilogue = f.stats = 0;
  if (!epi) return CCVM_OK;
  int rc = epi_params_from_desc(epi, f.epi);
  if (rc) return rc;
  EpiParams& e = f.epi;
  e.q = d.q;
  e.v = d.v;
  e.n = d.n;
  e.batch = d.batch;
  e.state = d.solver == CCVM_SOLVER_MF ? d.out1 : d.out0;
  if (out && !e.energy) return fail(CCVM_E_INVALID, "solution statistics need the energy buffer");
  size_t need = 0;
  if ((rc = plan_epilogue(e, threads / 32, max_smem, need))) return rc;
  if (need > smem) smem = need;
  f.epilogue = 1;
  f.stats = out != nullptr;
  f.total_ctas = ctas;
  f.optimal = (float)optimal;
  f.accum = accum;
  f.out = out;
  return CCVM_OK;
}

static int solve_impl(const ccvm_solve_desc* d, const ccvm_epilogue_desc* epi, double optimal, FusedOut* result,
                      cudaStream_t st) {
  NvtxRange range(epi ? "ccvm_solve_fused" : "ccvm_solve");
  int rc = validate_solve(d);
  if (rc) return rc;
  DeviceInfo di;
  if ((rc = device_info(di))) return rc;
  const int path = choose_path(*d, true);
  const bool adam = d->algorithm == CCVM_ALG_ADAM;
  const size_t sched_row_bytes = (size_t)d->iterations * SCHED_W * sizeof(float);
  StreamBuf sched_buf(st), qs_buf(st), accum_buf(st);
  SdeParams p;
  bool fused = false;

  if (path == PATH_MMA) {
    MmaPlan MP;
    plan_mma(*d, di, MP);
    FusedTail f;
    memset(&f, 0, sizeof(f));
    const bool inline_sched = sched_row_bytes * (size_t)MP.ctas <= ((size_t)64 << 20) && getenv("CCVM_NO_SCHED_INLINE") == nullptr;
    CUDA_TRY(sched_buf.alloc(inline_sched ? sched_row_bytes * MP.ctas : sched_row_bytes));
    f.sched_inline = inline_sched ? 1 : 0;
    f.sa = sched_args(d);
    f.sched_scratch = sched_buf.as<float>();
    if (!inline_sched) {
      build_schedule_kernel<<<(d->iterations + 127) / 128, 128, 0, st>>>(f.sa, sched_buf.as<float>());
      CUDA_TRY(cudaGetLastError());
    }
    fill_params(d, inline_sched ? nullptr : sched_buf.as<float>(), (d->n + 3) / 4, p);
    if (epi && getenv("CCVM_NO_FUSE") == nullptr) {
      StatsAccum* accum = nullptr;
      if (result) {
        CUDA_TRY(accum_buf.alloc(sizeof(StatsAccum)));
        accum = accum_buf.as<StatsAccum>();
        CUDA_TRY(cudaMemsetAsync(accum, 0, sizeof(StatsAccum), st));
      }
      if ((rc = plan_fused_tail(*d, epi, optimal, MP.threads, MP.smem, di.max_smem, (unsigned)MP.ctas, accum, result, f,
                                MP.smem)))
        return rc;
      fused = true;
    }
#define LAUNCH_MMA(S, A) rc = launch_mma<S, A>(p, MP, f, st)
    CCVM_DISPATCH_TILE(d->solver, adam, LAUNCH_MMA)
#undef LAUNCH_MMA
    if (rc) return rc;
  } else if (path == PATH_TC) {
    CUDA_TRY(sched_buf.alloc(sched_row_bytes));
    build_schedule_kernel<<<(d->iterations + 127) / 128, 128, 0, st>>>(sched_args(d), sched_buf.as<float>());
    CUDA_TRY(cudaGetLastError());
    fill_params(d, sched_buf.as<float>(), 0, p);
    if ((rc = solve_tc(d, p, st))) return rc;
  } else {
    TmemPlan TP;
    if ((rc = plan_tmem(*d, di, path, TP))) return rc;
    FusedTail f;
    memset(&f, 0, sizeof(f));
    // the schedule table is evaluated by every CTA in its prologue (one launch per solve) -- unless the per-CTA
    // copies would add up (very large batches x very long runs): then ONE table from a kernel of its own
    const bool inline_sched = sched_row_bytes * (size_t)TP.ctas <= ((size_t)64 << 20) && getenv("CCVM_NO_SCHED_INLINE") == nullptr;
    CUDA_TRY(sched_buf.alloc(inline_sched ? sched_row_bytes * TP.ctas : sched_row_bytes));
    f.sched_inline = inline_sched ? 1 : 0;
    f.sa = sched_args(d);
    f.sched_scratch = sched_buf.as<float>();
    if (!inline_sched) {
      build_schedule_kernel<<<(d->iterations + 127) / 128, 128, 0, st>>>(f.sa, sched_buf.as<float>());
      CUDA_TRY(cudaGetLastError());
    }
    fill_params(d, inline_sched ? nullptr : sched_buf.as<float>(), TP.cg, p);
    if (TP.qsrc == QSRC_GMEM) {
      const int np = 4 * TP.cg;
      CUDA_TRY(qs_buf.alloc((size_t)np * np * sizeof(float)));
      scale_q_kernel<<<(np * np + 255) / 256, 256, 0, st>>>(p.q, p.drift_s_vec, p.drift_s, p.a_half, p.n, np,
                                                            qs_buf.as<float>());
      CUDA_TRY(cudaGetLastError());
      TP.L.qs = qs_buf.as<float>();
    }
    if (epi && getenv("CCVM_NO_FUSE") == nullptr) {
      StatsAccum* accum = nullptr;
      if (result) {
        CUDA_TRY(accum_buf.alloc(sizeof(StatsAccum)));
        accum = accum_buf.as<StatsAccum>();
        CUDA_TRY(cudaMemsetAsync(accum, 0, sizeof(StatsAccum), st));
      }
      if ((rc = plan_fused_tail(*d, epi, optimal, TP.threads, TP.smem, di.max_smem, (unsigned)TP.ctas, accum, result, f,
                                TP.smem)))
        return rc;
      fused = true;
    }
#define LAUNCH_TMEM(S, A) rc = launch_tmem<S, A>(p, TP, f, st)
    CCVM_DISPATCH_TILE(d->solver, adam, LAUNCH_TMEM)
#undef LAUNCH_TMEM
    if (rc) return rc;
  }
  if (epi && !fused) {
    ccvm_epilogue_desc ed = *epi;
    ed.n = d->n;
    ed.batch = d->batch;
    ed.q = d->q;
    ed.v = d->v;
    ed.state = d->solver == CCVM_SOLVER_MF ? d->out1 : d->out0;
    if ((rc = ccvm_epilogue(&ed, st))) return rc;
    if (result) {
      if (!ed.energy) return fail(CCVM_E_INVALID, "solution statistics need the energy buffer");
      CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(FusedOut), st));
      if ((rc = ccvm_solution_stats(ed.energy, d->batch, optimal, &result->stats, st))) return rc;
    }
  }
  return CCVM_OK;
}

extern "C" int ccvm_solve(const ccvm_solve_desc* d, void* stream) {
  return solve_impl(d, nullptr, 0.0, nullptr, (cudaStream_t)stream);
}

extern "C" int ccvm_solve_fused(const ccvm_solve_desc* d, const ccvm_epilogue_desc* epi, double optimal_value,
                                void* result, void* stream) {
  if (!epi) return fail(CCVM_E_INVALID, "ccvm_solve_fused needs an epilogue descriptor");
  return solve_impl(d, epi, optimal_value, (FusedOut*)result, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ batched instances
// side streams of the library (per device, created once): the buckets of a batched launch run concurrently
struct SideStreams {
  static constexpr int N = 8;
  cudaStream_t s[N];
  bool ok = false;
};
static int side_streams(int device, SideStreams*& out) {
  static SideStreams pools[64];
  static std::mutex pools_mutex;
  std::lock_guard<std::mutex> lock(pools_mutex);
  SideStreams& p = pools[device];
  if (!p.ok) {
    for (int i = 0; i < SideStreams::N; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking));
    p.ok = true;
  }
  out = &p;
  return CCVM_OK;
}

struct SchedJob {
  SchedArgs a;
  long long offset;  // first row of this problem in the shared schedule table
};

__global__ void build_schedule_batch_kernel(const SchedJob* __restrict__ jobs, float* __restrict__ out) {
  const SchedJob job = jobs[blockIdx.y];
  if (blockIdx.x * blockDim.x >= job.a.iterations) return;
  schedule_row(job.a, blockIdx.x * blockDim.x + threadIdx.x, out + job.offset * SCHED_W);
}

static int solve_batch_impl(const ccvm_solve_desc* descs, const ccvm_epilogue_desc* epis, const double* optimal,
                            int32_t count, FusedOut* results, cudaStream_t st) {
  NvtxRange range(epis ? "ccvm_solve_batch_fused" : "ccvm_solve_batch");
  if (!descs || count < 1) return fail(CCVM_E_INVALID, "ccvm_solve_batch needs at least one descriptor");
  if (results && (!epis || !optimal)) return fail(CCVM_E_INVALID, "statistics need epilogue descriptors and optimal values");
  DeviceInfo di;
  int rc = device_info(di);
  if (rc) return rc;
  const int solver = descs[0].solver, alg = descs[0].algorithm;
  std::vector<int> batched, single;
  for (int i = 0; i < count; ++i) {
    if ((rc = validate_solve(&descs[i]))) return rc;
    if (descs[i].solver != solver || descs[i].algorithm != alg)
      return fail(CCVM_E_INVALID, "all descriptors of a batch must share solver and algorithm");
    if (descs[i].rng_mode != CCVM_RNG_PHILOX || descs[i].evolution_step > 0)
      return fail(CCVM_E_INVALID, "batched solves use Philox noise and no evolution sampling");
    const int path = choose_path(descs[i]);
    (path == PATH_TMEM || path == PATH_HYB ? batched : single).push_back(i);
  }
  for (int i : single)
    if ((rc = solve_impl(&descs[i], epis ? &epis[i] : nullptr, optimal ? optimal[i] : 0.0, results ? results + i : nullptr, st)))
      return rc;
  if (batched.empty()) return CCVM_OK;

  // plans, schedule table offsets
  std::vector<TmemPlan> plans(batched.size());
  std::vector<SchedJob> jobs(batched.size());
  long long rows = 0, total_traj = 0;
  int max_t = 0;
  for (int i : batched) total_traj += descs[i].batch;
  const int share = (int)((total_traj + di.sms - 1) / di.sms);
  for (size_t b = 0; b < batched.size(); ++b) {
    const ccvm_solve_desc& d = descs[batched[b]];
    if ((rc = plan_tmem(d, di, choose_path(d), plans[b], share, true))) return rc;
    jobs[b].a = sched_args(&d);
    jobs[b].offset = rows;
    rows += d.iterations;
    if (d.iterations > max_t) max_t = d.iterations;
  }
  StreamBuf sched_buf(st), jobs_buf(st), items_buf(st), map_buf(st), accum_buf(st);
  CUDA_TRY(sched_buf.alloc((size_t)rows * SCHED_W * sizeof(float)));
  CUDA_TRY(jobs_buf.alloc(jobs.size() * sizeof(SchedJob)));
  float* sched = sched_buf.as<float>();
  SchedJob* d_jobs = jobs_buf.as<SchedJob>();
  CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(SchedJob), cudaMemcpyHostToDevice, st));
  build_schedule_batch_kernel<<<dim3((max_t + 127) / 128, (unsigned)jobs.size()), 128, 0, st>>>(d_jobs, sched);
  CUDA_TRY(cudaGetLastError());
  const bool fuse = epis != nullptr && getenv("CCVM_NO_FUSE") == nullptr;
  StatsAccum* accum = nullptr;
  if (fuse && results) {
    CUDA_TRY(accum_buf.alloc(batched.size() * sizeof(StatsAccum)));
    accum = accum_buf.as<StatsAccum>();
    CUDA_TRY(cudaMemsetAsync(accum, 0, batched.size() * sizeof(StatsAccum), st));
  }

  // items + CTA maps, bucketed by (Q source, block size, compiled-in column-group count) so that small
  // instances do not pay for big blocks and the benchmarking sizes get the fully unrolled kernels
  struct Bucket {
    int threads, qsrc, cgc, ktail;
    size_t smem;
    std::vector<int2> map;
  };
  std::vector<Bucket> buckets;
  std::vector<BatchItem> items(batched.size());
  const int bucket_threads[4] = {32, 64, 128, 256};
  for (size_t b = 0; b < batched.size(); ++b) {
    const ccvm_solve_desc& d = descs[batched[b]];
    fill_params(&d, sched + jobs[b].offset * SCHED_W, plans[b].cg, items[b].p);
    items[b].L = plans[b].L;
    memset(&items[b].f, 0, sizeof(FusedTail));
    int k = 0;
    while (bucket_threads[k] < plans[b].threads) ++k;
    size_t smem = plans[b].smem;
    if (fuse) {
      if ((rc = plan_fused_tail(d, &epis[batched[b]], optimal ? optimal[batched[b]] : 0.0, bucket_threads[k], plans[b].smem,
                                di.max_smem, (unsigned)plans[b].ctas, accum ? accum + b : nullptr,
                                results ? results + batched[b] : nullptr, items[b].f, smem)))
        return rc;
    }
    size_t which = buckets.size();
    for (size_t u = 0; u < buckets.size(); ++u)
      if (buckets[u].threads == bucket_threads[k] && buckets[u].qsrc == plans[b].qsrc && buckets[u].cgc == plans[b].cgc &&
          buckets[u].ktail == plans[b].ktail)
        which = u;
    if (which == buckets.size())
      buckets.push_back(Bucket{bucket_threads[k], plans[b].qsrc, plans[b].cgc, plans[b].ktail, 0, {}});
    Bucket& B = buckets[which];
    for (int c = 0; c < plans[b].ctas; ++c) B.map.push_back(make_int2((int)b, c));
    if (smem > B.smem) B.smem = smem;
  }
  // largest buckets first: the long kernels start early, the small ones fill the gaps
  std::sort(buckets.begin(), buckets.end(), [](const Bucket& a, const Bucket& b) {
    return (size_t)a.threads * a.map.size() > (size_t)b.threads * b.map.size();
  });
  size_t total_ctas = 0;
  for (const Bucket& B : buckets) total_ctas += B.map.size();
  CUDA_TRY(items_buf.alloc(items.size() * sizeof(BatchItem)));
  CUDA_TRY(map_buf.alloc(total_ctas * sizeof(int2)));
  BatchItem* d_items = items_buf.as<BatchItem>();
  int2* d_map = map_buf.as<int2>();
  CUDA_TRY(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(BatchItem), cudaMemcpyHostToDevice, st));
  {
    std::vector<int2> all;
    all.reserve(total_ctas);
    for (const Bucket& B : buckets) all.insert(all.end(), B.map.begin(), B.map.end());
    CUDA_TRY(cudaMemcpyAsync(d_map, all.data(), total_ctas * sizeof(int2), cudaMemcpyHostToDevice, st));
    // `all` is pageable host memory: the copy is staged before cudaMemcpyAsync returns
  }
  // The buckets of a chunk are independent kernels of very different sizes (a bucket may hold a single
  // small instance: a dozen CTAs): on ONE stream they would run one after the other and leave most SMs idle.
  // They are forked onto side streams of the library (after an event that orders them behind the copies and
  // the schedule kernel above) and joined back into the caller's stream.
  const bool adam = alg == CCVM_ALG_ADAM;
  const bool fork = buckets.size() > 1 && getenv("CCVM_NO_FORK") == nullptr;
  SideStreams* side = nullptr;
  cudaEvent_t ready = nullptr;
  if (fork) {
    if ((rc = side_streams(di.device, side))) return rc;
    CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ready, st));
  }
  size_t off = 0;
  int launched = 0;
  for (const Bucket& B : buckets) {
    BatchBucket bb;
    bb.items = d_items;
    bb.map = d_map + off;
    bb.ctas = (unsigned)B.map.size();
    bb.threads = B.threads;
    bb.qsrc = B.qsrc;
    bb.cgc = B.cgc;
    bb.ktail = B.ktail;
    bb.smem = B.smem;
    cudaStream_t ls = st;
    if (fork) {
      ls = side->s[launched % SideStreams::N];
      if (launched < SideStreams::N) cudaStreamWaitEvent(ls, ready, 0);  // later buckets follow on the same side stream
    }
#define LAUNCH_BATCH(S, A) rc = launch_tmem_batch<S, A>(bb, ls)
    CCVM_DISPATCH_TILE(solver, adam, LAUNCH_BATCH)
#undef LAUNCH_BATCH
    if (rc) break;
    off += B.map.size();
    ++launched;
  }
  if (fork) {
    // join: the caller's stream continues (and the scratch is freed) only after every side stream is done
    const int used = launched < SideStreams::N ? launched : SideStreams::N;
    for (int i = 0; i < used; ++i) {
      cudaEvent_t done;
      if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) continue;
      cudaEventRecord(done, side->s[i]);
      cudaStreamWaitEvent(st, done, 0);
      cudaEventDestroy(done);   // released by the runtime once the wait has been satisfied
    }
    cudaEventDestroy(ready);
  }
  if (rc) return rc;
  if (epis && !fuse) {
    for (int i : batched) {
      ccvm_epilogue_desc ed = epis[i];
      ed.n = descs[i].n;
      ed.batch = descs[i].batch;
      ed.q = descs[i].q;
      ed.v = descs[i].v;
      ed.state = descs[i].solver == CCVM_SOLVER_MF ? descs[i].out1 : descs[i].out0;
      if ((rc = ccvm_epilogue(&ed, st))) return rc;
      if (results) {
        CUDA_TRY(cudaMemsetAsync(results + i, 0, sizeof(FusedOut), st));
        if ((rc = ccvm_solution_stats(ed.energy, descs[i].batch, optimal[i], &results[i].stats, st))) return rc;
      }
    }
  }
  return CCVM_OK;
}

extern "C" int ccvm_solve_batch(const ccvm_solve_desc* descs, int32_t count, void* stream) {
  return solve_batch_impl(descs, nullptr, nullptr, count, nullptr, (cudaStream_t)stream);
}

extern "C" int ccvm_solve_batch_fused(const ccvm_solve_desc* descs, const ccvm_epilogue_desc* epis,
                                      const double* optimal_values, int32_t count, void* results, void* stream) {
  if (!epis) return fail(CCVM_E_INVALID, "ccvm_solve_batch_fused needs epilogue descriptors");
  return solve_batch_impl(descs, epis, optimal_values, count, (FusedOut*)results, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- noise dump (validation)
// The normals a production solve draws, written in the replay layout noise[T][K][n][batch]: feeding this
// tensor to the CPU oracle (or back to the engine in replay mode) reproduces a production run exactly,
// which pins the production kernel variants (in-loop noise, compile-time column groups) to the oracle.
// Counter mode (the generator of the tcgen05 path): one thread per (iteration, quadrature, column group,
// trajectory).
__global__ void dump_noise_counter_kernel(uint32_t k0, uint32_t k1, uint32_t off_lo, long long traj_base, int n,
                                          int batch, int iterations, int K, float* __restrict__ noise) {
  const int cg_count = (n + 3) / 4;
  const size_t total = (size_t)iterations * K * cg_count * batch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx % batch);
    size_t r = idx / batch;
    const int cg = (int)(r % cg_count);
    r /= cg_count;
    const int q = (int)(r % K);
    const int t = (int)(r / K);
    float w[4];
    noise_normals4(k0, k1, off_lo, (unsigned long long)(traj_base + b), (uint32_t)t, (uint32_t)cg, (uint32_t)q, w[0], w[1],
                   w[2], w[3]);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * cg + jj;
      if (j < n) noise[(((size_t)t * K + q) * n + j) * (size_t)batch + b] = w[jj];
    }
  }
}

// Stream mode (the tiled SIMT kernels): one thread per (trajectory pair, column group) walks its
// xoshiro128+ stream through all iterations in the kernels' draw order (quadrature, trajectory of the pair).
__global__ void dump_noise_stream_kernel(uint32_t k0, uint32_t k1, uint32_t off_lo, long long traj_base, int n,
                                         int batch, int iterations, int K, float* __restrict__ noise) {
  const int cg_count = (n + 3) / 4;
  const int pairs = (batch + 1) / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pairs * cg_count) return;
  const int pr = idx % pairs, cg = idx / pairs;   // neighbouring threads: neighbouring trajectories (coalesced stores)
  NoiseStream rs = stream_init(k0, k1, off_lo, (unsigned long long)(traj_base + 2 * (long long)pr) >> 1, (uint32_t)cg);
  for (int t = 0; t < iterations; ++t)
    for (int q = 0; q < K; ++q)
      for (int i = 0; i < 2; ++i) {
        float w[4];
        stream_normals4(rs, w[0], w[1], w[2], w[3]);
        const int b = 2 * pr + i;
        if (b >= batch) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * cg + jj;
          if (j < n) noise[(((size_t)t * K + q) * n + j) * (size_t)batch + b] = w[jj];
        }
      }
}

// Streams of the small-n tensor-core kernel (sde_kernel_mma.cuh): one thread per (trajectory pair, variable);
// a Box-Muller pair per iteration and quadrature serves the two trajectories of the pair.
__global__ void dump_noise_mma_kernel(uint32_t k0, uint32_t k1, uint32_t off_lo, long long traj_base, int n, int batch,
                                      int iterations, int K, float* __restrict__ noise) {
  const int pairs = (batch + 1) / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pairs * n) return;
  const int pr = idx % pairs, v = idx / pairs;
  NoiseStream rs = stream_init(k0, k1, off_lo, (unsigned long long)(traj_base + 2 * (long long)pr) >> 1,
                               (uint32_t)v | MMA_STREAM_TAG);
  for (int t = 0; t < iterations; ++t)
    for (int q = 0; q < K; ++q) {
      const pf2 w = stream_normal_pair(rs);
      float* dst = noise + (((size_t)t * K + q) * n + v) * (size_t)batch + 2 * pr;
      dst[0] = w.x;
      if (2 * pr + 1 < batch) dst[1] = w.y;
    }
}

extern "C" int ccvm_dump_noise(const ccvm_solve_desc* d, float* noise, void* stream) {
  if (!d || !noise) return fail(CCVM_E_INVALID, "bad argument to ccvm_dump_noise");
  if (d->n < 1 || d->batch < 1 || d->iterations < 1) return fail(CCVM_E_INVALID, "n, batch and iterations must be >= 1");
  if (d->solver < 0 || d->solver > 3) return fail(CCVM_E_INVALID, "unknown solver id %d", d->solver);
  const int K = d->solver == CCVM_SOLVER_DL ? 2 : 1;
  const int cg_count = (d->n + 3) / 4;
  const uint32_t k0 = (uint32_t)d->seed, k1 = (uint32_t)(d->seed >> 32) ^ (uint32_t)(d->offset >> 32);
  cudaStream_t st = (cudaStream_t)stream;
  // the generator ccvm_solve(desc) would use: counter mode on the tcgen05 path, streams on the tiled SIMT path
  ccvm_solve_desc launch = *d;   // the launch whose draws are wanted (noise_batch: its batch, when this is a slice)
  launch.rng_mode = CCVM_RNG_PHILOX;
  if (d->noise_batch > 0 && d->noise_batch <= 0x7fffffff) launch.batch = (int32_t)d->noise_batch;
  const int path = choose_path(launch, true);
  const bool stream_mode = CCVM_SIMT_RNG && path != PATH_TC;
  if (path == PATH_MMA) {
    if (d->traj_base & 1) return fail(CCVM_E_INVALID, "traj_base must be even (noise streams belong to trajectory pairs)");
    const int total = ((d->batch + 1) / 2) * d->n;
    dump_noise_mma_kernel<<<(total + 127) / 128, 128, 0, st>>>(k0, k1, (uint32_t)d->offset, d->traj_base, d->n, d->batch,
                                                              d->iterations, K, noise);
  } else if (stream_mode) {
    if (d->traj_base & 1) return fail(CCVM_E_INVALID, "traj_base must be even (noise streams belong to trajectory pairs)");
    const int total = ((d->batch + 1) / 2) * cg_count;
    dump_noise_stream_kernel<<<(total + 127) / 128, 128, 0, st>>>(k0, k1, (uint32_t)d->offset, d->traj_base, d->n, d->batch,
                                                                 d->iterations, K, noise);
  } else {
    const size_t total = (size_t)d->iterations * K * cg_count * d->batch;
    size_t grid = (total + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    dump_noise_counter_kernel<<<(unsigned)grid, 256, 0, st>>>(k0, k1, (uint32_t)d->offset, d->traj_base, d->n, d->batch,
                                                             d->iterations, K, noise);
  }
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// ------------------------------------------------------- multi-GPU result records
// One record per rank: [min energy, global trajectory index of the winner (two 32-bit halves), 7
// success counters, winner's solution vector (n)] -- CCVM_RECORD_HEADER + n 32-bit words; the integer
// fields are stored as INTEGERS (bit patterns inside the float buffer), so indices and counts beyond
// 2^24 stay exact.  Written by ONE kernel from the statistics block of ccvm_solution_stats,
// all-gathered by the host layer, and reduced by ONE kernel on every rank (ties go to the lowest rank,
// a NaN objective wins like torch.max(-E) propagates it).
constexpr int REC_H = CCVM_RECORD_HEADER;

__global__ void pack_record_kernel(const StatsOut* __restrict__ stats, const float* __restrict__ pv, int n,
                                   long long traj_base, float* __restrict__ rec) {
  const int arg = stats->arg_best;
  int* reci = reinterpret_cast<int*>(rec);
  if (threadIdx.x == 0) {
    const unsigned long long g = (unsigned long long)(traj_base + arg);
    rec[0] = -stats->best;
    reci[1] = (int)(uint32_t)g;
    reci[2] = (int)(uint32_t)(g >> 32);
  }
  if (threadIdx.x < 7) reci[3 + threadIdx.x] = stats->counts[threadIdx.x];
  for (int j = threadIdx.x; j < n; j += blockDim.x) rec[REC_H + j] = pv[(size_t)arg * n + j];
}

__global__ void merge_records_kernel(const float* __restrict__ gathered, int world, int n, float* __restrict__ out) {
  __shared__ int s_owner;
  const int len = REC_H + n;
  const int* gi = reinterpret_cast<const int*>(gathered);
  int* outi = reinterpret_cast<int*>(out);
  if (threadIdx.x == 0) {
    int owner = 0;
    float best = gathered[0];
    for (int r = 1; r < world && !(best != best); ++r) {
      const float e = gathered[(size_t)r * len];
      if (e != e || e < best) {
        best = e;
        owner = r;
      }
    }
    s_owner = owner;
    out[0] = -best;
    outi[1] = gi[(size_t)owner * len + 1];
    outi[2] = gi[(size_t)owner * len + 2];
  }
  if (threadIdx.x < 7) {
    int tot = 0;
    for (int r = 0; r < world; ++r) tot += gi[(size_t)r * len + 3 + threadIdx.x];
    outi[3 + threadIdx.x] = tot;
  }
  __syncthreads();
  const float* src = gathered + (size_t)s_owner * len + REC_H;
  for (int j = threadIdx.x; j < n; j += blockDim.x) out[REC_H + j] = src[j];
}

extern "C" int ccvm_pack_record(const void* stats, const float* problem_variables, int32_t n, int64_t traj_base,
                                float* record, void* stream) {
  if (!stats || !problem_variables || !record || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_pack_record");
  pack_record_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const StatsOut*)stats, problem_variables, n,
                                                          (long long)traj_base, record);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

extern "C" int ccvm_merge_records(const float* gathered, int32_t world_size, int32_t n, float* merged, void* stream) {
  if (!gathered || !merged || world_size < 1 || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_merge_records");
  merge_records_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(gathered, world_size, n, merged);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// -------------------------------------------------------------------- scaling factor
__global__ void __launch_bounds__(1024) scaling_factor_kernel(const float* __restrict__ q, int nn, float mult,
                                                              float* out) {
  __shared__ double part[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) acc += (double)fabsf(q[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
    *out = __fmul_rn(sqrtf((float)tot), mult);  // sqrt(sum|Q|) * multiplier, fp32 (ccvm_solver.py:146-149)
  }
}

extern "C" int ccvm_scaling_factor(const float* q, int32_t n, double multiplier, float* out, void* stream) {
  if (!q || !out || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_scaling_factor");
  scaling_factor_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(q, n * n, (float)multiplier, out);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// ------------------------------------------------------------ synthetic instance generator
// Dense symmetric BoxQP coefficients with the statistics of the reference's bundled instances
// (off-diagonal std 28.5/sqrt(N), diagonal std sqrt(2) x that -- the law of
// (A + A^T)/sqrt(2) for i.i.d. normal A -- and V std 20), drawn on the device with Philox keyed by
// (seed, unordered index pair): element (i, j) and (j, i) evaluate the same counter, so no transpose
// pass is needed.  Stored in the reference's in-memory convention (negated, problem_instance.py:183-188).
__global__ void generate_boxqp_kernel(float* __restrict__ q, float* __restrict__ v, int n, uint32_t seed_lo,
                                      uint32_t seed_hi, float q_std, float v_std) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n + n) return;
  int i, j;
  uint32_t plane;
  if (idx < n * n) {
    i = idx / n;
    j = idx - i * n;
    if (i > j) {
      const int t = i;
      i = j;
      j = t;
    }
    plane = 0u;
  } else {
    i = idx - n * n;
    j = 0;
    plane = 1u;
  }
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)j, plane, 0x51424f58u), make_uint2(seed_lo, seed_hi));
  float z0, z1;
  box_muller(r.x, r.y, z0, z1);
  if (plane == 0u) q[idx] = -(i == j ? 1.41421356237f : 1.f) * q_std * z0;
  else v[i] = -v_std * z0;
}

extern "C" int ccvm_generate_boxqp(float* q, float* v, int32_t n, uint64_t seed, double q_offdiag_std, double v_std,
                                   void* stream) {
  if (!q || !v || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_generate_boxqp");
  const int total = n * n + n;
  generate_boxqp_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      q, v, n, (uint32_t)seed, (uint32_t)(seed >> 32), (float)q_offdiag_std, (float)v_std);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// ------------------------------------------------------------------ host-buffer entry
extern "C" int ccvm_solve_host(const ccvm_solve_desc* solve, const ccvm_epilogue_desc* epi, const float* h_q,
                               const float* h_v, double optimal_value, float* h_energy, void* h_stats,
                               void* stream) {
  if (!solve || !epi || !h_q || !h_v || !h_energy || !h_stats) return fail(CCVM_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = solve->n, b = solve->batch;
  if (solve->n < 1 || solve->batch < 1) return fail(CCVM_E_INVALID, "n and batch must be >= 1");
  const size_t words = n * n + n + 3 * b * n + b * n + b + 16 + sizeof(FusedOut) / 4;
  StreamBuf host_buf(st);
  CUDA_TRY(host_buf.alloc(words * sizeof(float)));
  float* buf = host_buf.as<float>();
  float* d_q = buf;
  float* d_v = d_q + n * n;
  float* d_o0 = d_v + n;
  float* d_o1 = d_o0 + b * n;
  float* d_o2 = d_o1 + b * n;
  float* d_pv = d_o2 + b * n;
  float* d_e = d_pv + b * n;
  float* d_stats = d_e + b;
  d_stats = (float*)(((uintptr_t)d_stats + 15) & ~(uintptr_t)15);
  int rc = CCVM_OK;
  cudaError_t ce;
  if ((ce = cudaMemcpyAsync(d_q, h_q, n * n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (ce = cudaMemcpyAsync(d_v, h_v, n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
    rc = fail(CCVM_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(ce));
  }
  ccvm_solve_desc sd = *solve;
  sd.q = d_q;
  sd.v = d_v;
  sd.out0 = d_o0;
  sd.out1 = d_o1;
  sd.out2 = d_o2;
  sd.evolution_step = 0;
  sd.samples = nullptr;
  ccvm_epilogue_desc ed = *epi;
  ed.problem_variables = d_pv;
  ed.energy = d_e;
  // ONE launch: schedule, all iterations, change of variables, post-processor, energy, statistics
  if (!rc) rc = solve_impl(&sd, &ed, optimal_value, (FusedOut*)d_stats, st);
  if (!rc) {
    if ((ce = cudaMemcpyAsync(h_energy, d_e, b * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (ce = cudaMemcpyAsync(h_stats, d_stats, sizeof(StatsOut), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      rc = fail(CCVM_E_CUDA, "D2H copy failed: %s", cudaGetErrorString(ce));
  }
  ce = cudaStreamSynchronize(st);
  if (!rc && ce != cudaSuccess) rc = fail(CCVM_E_CUDA, "stream sync failed: %s", cudaGetErrorString(ce));
  return rc;
}

// ---------------------------------------------------------------------- FP32 probe
template <int MODE>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, int iters, float seed) {
  if constexpr (MODE == 0) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    const float x = 0.999f + seed * 1e-3f, y = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;
  } else {
    pf2 a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = pk(seed + i * 0.001f, seed + threadIdx.x * 1e-6f);
    const pf2 x = dup(0.999f + seed * 1e-3f), y = dup(1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
  }
}

extern "C" int ccvm_microbench_fp32(int32_t mode, double* tflops, void* stream) {
  if (!tflops || (mode != 0 && mode != 1)) return fail(CCVM_E_INVALID, "bad argument to ccvm_microbench_fp32");
  DeviceInfo di;
  int rc = device_info(di);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  StreamBuf probe_buf(st);
  CUDA_TRY(probe_buf.alloc(64));
  float* d = probe_buf.as<float>();
  const int iters = 4096, grid = di.sms * 8, block = 256;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    if (mode == 0) fp32_probe_kernel<0><<<grid, block, 0, st>>>(d, iters, 0.5f);
    else fp32_probe_kernel<1><<<grid, block, 0, st>>>(d, iters, 0.5f);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * (mode == 0 ? 1.0 : 2.0) * 16.0 * 8.0 * iters * (double)grid * block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CUDA_TRY(cudaStreamSynchronize(st));
  *tflops = best;
  return CCVM_OK;
}

// ------------------------------------------------------------------ tensor-pipe probe
// Roofline denominator of the tcgen05 path ("measure dense TF32 peak on the box"):
// every CTA (or CTA pair) issues back-to-back tcgen05.mma kind::tf32 of the shape the SDE kernel
// uses (M = 128 per CTA, N = 256, K = 8) on operands that stay in shared memory, alternating two
// TMEM accumulators -- no loads, no epilogue.  What it reports is the rate at which the tensor pipe
// retires this instruction, i.e. the ceiling of sde_tc_kernel (cta_group::1) and sde_tc2_kernel
// (cta_group::2); three MMAs make one logical 3xTF32 product.
template <int PAIR>
__device__ __forceinline__ void tf32_probe_body(int iters) {
  extern __shared__ __align__(1024) uint8_t probe_smem[];
  __shared__ __align__(8) unsigned long long done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(probe_smem) + 1023u) & ~1023u;
  // A: 128 rows x 32 floats (16 KB), B: 256 (single CTA) or 128 (pair: this CTA's half) rows x 32 floats
  constexpr int B_ROWS = PAIR ? 128 : 256;
  float* tiles = reinterpret_cast<float*>(probe_smem + (base - smem_u32(probe_smem)));
  for (int i = tid; i < (128 + B_ROWS) * 32; i += blockDim.x) tiles[i] = 1.0f / (float)(1 + (i & 7));
  uint32_t rank = 0;
  if (PAIR) rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_slot, 512);
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy tile writes -> tensor-core reads
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 32 && rank == 0) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) |
                               ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint32_t a0 = base, b0 = base + 128 * 32 * 4;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (uint32_t)(it & 1) * 256u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = umma_desc_sw128(a0 + ks * 32), bd = umma_desc_sw128(b0 + ks * 32);
        if (PAIR) umma_tf32_pair(d, ad, bd, idesc, (it > 1 || ks > 0) ? 1u : 0u);
        else umma_tf32(d, ad, bd, idesc, (it > 1 || ks > 0) ? 1u : 0u);
      }
    }
    if (PAIR) umma_commit_pair(smem_u32(&done_bar)); else umma_commit(smem_u32(&done_bar));
  }
  mbar_wait(smem_u32(&done_bar), 0u);
  tc_fence_after();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else tmem_free(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(int iters) { tf32_probe_body<0>(iters); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tf32_probe_pair_kernel(int iters) {
  tf32_probe_body<1>(iters);
}

extern "C" int ccvm_microbench_tf32(int32_t mode, double* tflops, void* stream) {
  if (!tflops || (mode != 1 && mode != 2)) return fail(CCVM_E_INVALID, "bad argument to ccvm_microbench_tf32");
  DeviceInfo di;
  int rc = device_info(di);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int iters = 2048;
  const int grid = mode == 2 ? (di.sms / 2) * 2 : di.sms;
  const size_t smem = 1024 + (size_t)(128 + (mode == 2 ? 128 : 256)) * 32 * 4;
  if (mode == 1) CUDA_TRY(cudaFuncSetAttribute(tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else CUDA_TRY(cudaFuncSetAttribute(tf32_probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    if (mode == 1) tf32_probe_kernel<<<grid, 128, smem, st>>>(iters);
    else tf32_probe_pair_kernel<<<grid, 128, smem, st>>>(iters);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    // MMAs per launch x flops per MMA (pair: M = 256 per instruction, one instruction per two CTAs)
    const double mmas = (double)(mode == 2 ? grid / 2 : grid) * iters * 4.0;
    const double flops = mmas * 2.0 * (mode == 2 ? 256.0 : 128.0) * 256.0 * 8.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return CCVM_OK;
}

// ------------------------------------------------------------------ operator hooks
// One thread per output element, arithmetic in the reference's fp32 operation order
// (host scalars are folded in fp64 first, exactly as Python does before they meet a tensor).
struct HookParams {
  const float* q;
  const float* v;
  const float* in0;
  const float* in1;
  const float* in2;
  const float* svec;
  float* out0;
  float* out1;
  int solver, kind, n, batch;
  float s, a, b, half_b, two_s;
  float c_lin, c_lin2;   // DL: -1 + pump*rate and -1 - pump*rate ; MF: -(1+j) + pump ; PLV: -1 + p
  float fsd;             // DL: -(fs*(0.5+rate)) ; MF / PLV: fs
  float g2, g2x2, g2x3, m2j, opj;
};

__device__ __forceinline__ float hook_contract(const HookParams& p, const float* in, int b, int j, bool half) {
  // sum_i (in_i * a / S_i + b) Q_ij     (half: in_i * a / (2 S_i) + b/2)
  const int N = p.n;
  float acc = 0.f;
  for (int i = 0; i < N; ++i) {
    const float si = p.svec ? p.svec[i] : p.s;
    const float den = half ? __fmul_rn(2.f, si) : si;
    const float x = __fadd_rn(__fdiv_rn(__fmul_rn(in[(size_t)b * N + i], p.a), den), half ? p.half_b : p.b);
    acc = fmaf(x, p.q[(size_t)i * N + j], acc);
  }
  return acc;
}

__global__ void hook_kernel(const HookParams p) {
  const int N = p.n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)p.batch * N) return;
  const int b = (int)(idx / N), j = (int)(idx - (size_t)b * N);
  const float sj = p.svec ? p.svec[j] : p.s;
  const float two_sj = __fmul_rn(2.f, sj);
  if (p.solver == SOLVER_DL) {
    const float c = p.in0[idx], s = p.in1[idx];
    const float ec = hook_contract(p, p.in0, b, j, false), es = hook_contract(p, p.in1, b, j, false);
    const float g1c = __fdiv_rn(__fmul_rn(__fmul_rn(0.25f, ec), p.a), sj);
    const float g1s = __fdiv_rn(__fmul_rn(__fmul_rn(0.25f, es), p.a), sj);
    const float g3 = __fdiv_rn(__fmul_rn(p.v[j], p.a), two_sj);
    if (p.kind == 0) {
      p.out0[idx] = __fsub_rn(-g1c, g3);
      p.out1[idx] = __fsub_rn(-g1s, g3);
    } else {
      const float c2 = __fmul_rn(c, c), s2 = __fmul_rn(s, s);
      const float g2c = __fmul_rn(__fsub_rn(__fsub_rn(p.c_lin, c2), s2), c);
      const float g2s = __fmul_rn(__fsub_rn(__fsub_rn(p.c_lin2, c2), s2), s);
      p.out0[idx] = __fadd_rn(__fmul_rn(p.fsd, __fadd_rn(g1c, g3)), g2c);
      p.out1[idx] = __fadd_rn(__fmul_rn(p.fsd, __fadd_rn(g1s, g3)), g2s);
    }
  } else if (p.solver == SOLVER_MF) {
    const float* mt = p.kind == 0 ? p.in0 : p.in1;
    const float e = hook_contract(p, mt, b, j, false);
    const float t21 = __fdiv_rn(__fmul_rn(__fmul_rn(-0.25f, e), p.a), sj);
    const float t22 = __fdiv_rn(__fmul_rn(-p.v[j], p.a), two_sj);
    const float fb = __fmul_rn(p.fsd, __fadd_rn(t21, t22));
    if (p.kind == 0) {
      p.out0[idx] = fb;
    } else {
      const float mu = p.in0[idx], sg = p.in2[idx];
      const float mu2 = __fmul_rn(mu, mu);
      const float term1 = __fmul_rn(__fsub_rn(p.c_lin, __fmul_rn(p.g2, mu2)), mu);
      p.out0[idx] = __fadd_rn(term1, fb);
      const float s1 = __fmul_rn(__fmul_rn(2.f, __fsub_rn(p.c_lin, __fmul_rn(p.g2x3, mu2))), sg);
      const float sh = __fsub_rn(sg, 0.5f);
      const float s2 = __fmul_rn(p.m2j, __fmul_rn(sh, sh));
      const float s3 = __fadd_rn(p.opj, __fmul_rn(p.g2x2, mu2));
      p.out1[idx] = __fadd_rn(__fadd_rn(s1, s2), s3);
    }
  } else if (p.solver == SOLVER_LV) {
    const float e = hook_contract(p, p.in0, b, j, true);
    p.out0[idx] = __fdiv_rn(__fmul_rn(-__fadd_rn(e, p.v[j]), p.a), two_sj);
  } else {
    const float e = hook_contract(p, p.in0, b, j, true);
    const float g1 = __fdiv_rn(__fmul_rn(e, p.a), two_sj);
    const float g2 = __fdiv_rn(__fmul_rn(p.v[j], p.a), two_sj);
    const float grads = __fsub_rn(-g1, g2);
    if (p.kind == 0) {
      p.out0[idx] = grads;
    } else {
      const float c = p.in0[idx];
      const float d0 = __fmul_rn(__fsub_rn(p.c_lin, __fmul_rn(c, c)), c);
      p.out0[idx] = __fadd_rn(d0, __fmul_rn(p.fsd, grads));
    }
  }
}

extern "C" int ccvm_eval_hook(const ccvm_hook_desc* d, void* stream) {
  if (!d) return fail(CCVM_E_INVALID, "null descriptor");
  if (d->solver < 0 || d->solver > 3 || (d->kind != 0 && d->kind != 1)) return fail(CCVM_E_INVALID, "bad solver/kind");
  if (d->n < 1 || d->batch < 1 || !d->q || !d->v || !d->in0 || !d->out0) return fail(CCVM_E_INVALID, "bad hook arguments");
  if (d->solver == CCVM_SOLVER_DL && (!d->in1 || !d->out1)) return fail(CCVM_E_INVALID, "DL hooks need c, s and two outputs");
  if (d->solver == CCVM_SOLVER_MF && d->kind == 1 && (!d->in1 || !d->in2 || !d->out1))
    return fail(CCVM_E_INVALID, "MF drift needs mu, mu_tilde, sigma and two outputs");
  HookParams p;
  memset(&p, 0, sizeof(p));
  p.q = d->q; p.v = d->v; p.in0 = d->in0; p.in1 = d->in1; p.in2 = d->in2;
  p.out0 = d->out0; p.out1 = d->out1;
  p.solver = d->solver; p.kind = d->kind; p.n = d->n; p.batch = d->batch;
  double s = d->s;
  p.svec = d->s_vec;
  if (d->solver == CCVM_SOLVER_DL && d->kind == 1 && d->pump > 1.0) {  // dl_solver.py:140-141
    s = sqrt(d->pump - 1.0);
    p.svec = nullptr;
  }
  p.s = (float)s;
  p.two_s = (float)(2.0 * s);
  p.a = (float)(d->upper - d->lower);
  p.b = (float)(d->upper + d->lower);
  p.half_b = (float)((d->upper + d->lower) / 2.0);
  if (d->solver == CCVM_SOLVER_DL) {
    p.c_lin = (float)(-1.0 + d->pump * d->rate);
    p.c_lin2 = (float)(-1.0 - d->pump * d->rate);
    p.fsd = (float)(-(d->feedback_scale * (0.5 + d->rate)));
  } else if (d->solver == CCVM_SOLVER_MF) {
    p.c_lin = (float)(-(1.0 + d->j) + d->pump);
    p.fsd = (float)d->feedback_scale;
    p.g2 = (float)(d->g * d->g);
    p.g2x2 = (float)(2.0 * d->g * d->g);
    p.g2x3 = (float)(3.0 * d->g * d->g);
    p.m2j = (float)(-2.0 * d->j);
    p.opj = (float)(1.0 + d->j);
  } else {
    p.c_lin = (float)(-1.0 + d->pump);
    p.fsd = (float)d->feedback_scale;
  }
  const size_t total = (size_t)d->batch * d->n;
  hook_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

__global__ void change_variables_kernel(const float* __restrict__ x, float* __restrict__ out, size_t total, int n,
                                        float a, float half_b, float s, const float* __restrict__ svec) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float sj = svec ? svec[idx % n] : s;
  // 0.5 * y / S * (u - l) + 0.5 * (u + l), left to right
  out[idx] = __fadd_rn(__fmul_rn(__fdiv_rn(__fmul_rn(0.5f, x[idx]), sj), a), half_b);
}

extern "C" int ccvm_change_variables(const float* x, float* out, int32_t batch, int32_t n, double lower,
                                     double upper, double s, const float* s_vec, void* stream) {
  if (!x || !out || batch < 1 || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_change_variables");
  const size_t total = (size_t)batch * n;
  change_variables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, out, total, n, (float)(upper - lower), (float)(0.5 * (upper + lower)), (float)s, s_vec);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

__global__ void clamp_kernel(const float* __restrict__ x, float* __restrict__ out, size_t total, int n, float lo,
                             float hi, const float* __restrict__ lo_t, const float* __restrict__ hi_t,
                             long long bound_len) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t bi = bound_len == (long long)total ? idx : idx % n;
  const float l = lo_t ? lo_t[bi] : lo, h = hi_t ? hi_t[bi] : hi;
  out[idx] = clampf(x[idx], l, h);
}

extern "C" int ccvm_fit_to_constraints(const float* x, float* out, int32_t batch, int32_t n, double lo, double hi,
                                       const float* lo_t, const float* hi_t, int64_t bound_len, void* stream) {
  if (!x || !out || batch < 1 || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_fit_to_constraints");
  const size_t total = (size_t)batch * n;
  if ((lo_t || hi_t) && bound_len != n && bound_len != (int64_t)total)
    return fail(CCVM_E_INVALID, "tensor bounds must have n or batch*n elements");
  clamp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, total, n, (float)lo,
                                                                                    (float)hi, lo_t, hi_t, bound_len);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

__global__ void scale_coefs_kernel(const float* __restrict__ q, const float* __restrict__ v, int n,
                                   const float* __restrict__ f, long long flen, float* __restrict__ qo,
                                   float* __restrict__ vo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nn = n * n;
  if (idx >= nn) return;
  const float fi = flen == 1 ? f[0] : f[idx];
  qo[idx] = __fdiv_rn(q[idx], fi);
  if (flen == 1) {
    if (idx < n) vo[idx] = __fdiv_rn(v[idx], fi);
  } else {
    vo[idx] = __fdiv_rn(v[idx % n], fi);
  }
}

extern "C" int ccvm_scale_coefs(const float* q, const float* v, int32_t n, const float* factor, int64_t factor_len,
                                float* q_out, float* v_out, void* stream) {
  if (!q || !v || !factor || !q_out || !v_out || n < 1) return fail(CCVM_E_INVALID, "bad argument to ccvm_scale_coefs");
  if (factor_len != 1 && factor_len != (int64_t)n * n)
    return fail(CCVM_E_INVALID, "scaling factor must have 1 or n*n elements");
  scale_coefs_kernel<<<(n * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(q, v, n, factor, factor_len, q_out, v_out);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}
