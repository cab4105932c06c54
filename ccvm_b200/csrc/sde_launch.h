// Seam between the C ABI (ccvm_abi.cu: validation, launch plans, small kernels) and the translation
// units that instantiate the persistent SDE kernels.  The kernels are heavy templates (eight loops x
// operand source x noise placement x compile-time column-group counts), so every (solver, algorithm)
// pair is compiled as its own object -- sde_tmem_inst.cu / sde_tc_inst.cu built once per pair with
// -DCCVM_INST_SOLVER / -DCCVM_INST_ADAM -- and the objects are linked into libccvm_b200.so
// (__graft_entry__.build compiles them in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ccvm_b200.h"
#include "ccvm_common.cuh"

namespace ccvm {

// thread-local error message of the ABI (ccvm_last_error); returns `code`
int set_error(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                             \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess)                                                                                         \
      return ::ccvm::set_error(CCVM_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- tiled SIMT kernels (sde_kernel_tmem.cuh)
struct TmemLaunch {
  int rg;      // trajectory pairs per group
  int ng;      // groups per CTA
  int gt;      // threads per group (TMEM source: 128 when ng > 1)
  int xs;      // floats per k-row of a group's X panel
  int tcols;   // TMEM columns to allocate (power of two >= 4*NP, >= 32); unused for QSRC_GMEM
  int phase_ns;  // start delay of odd groups (experiment knob; 0 in production)
  int xmask;        // 31: per-column-group bank offsets inside an X row (needs 32 floats of slack); 0: none
  int pipe;         // 1: in-loop noise generation (PIPE kernels); decided once by the host plan
  const float* qs;  // QSRC_GMEM: the scaled matrix Qs[NP][NP] (zero padded) in global memory
};

// Where the thread's Q slice comes from.
//   QSRC_TMEM (n <= 128): the thread's own TMEM lane (tcgen05.ld), see sde_kernel_tmem.cuh.
//   QSRC_GMEM (any n):    streamed from global memory / L2 with read-only 128-bit loads; the matrix
//                         is too large for on-chip replication, so it stays L2-resident (4 MB at
//                         n = 1024) and every CTA re-reads it once per iteration.
//   QSRC_HYB (128 < n <= 256): rows k < 128 of the slice in the thread's TMEM lane (all 512 columns),
//                         rows k >= 128 in shared memory, zero padded to HYB_LD columns per row so
//                         that every address of the tail is base + immediate (4 LDS.128 per chunk;
//                         lanes of one column group broadcast, neighbouring groups are contiguous).
enum : int { QSRC_TMEM = 0, QSRC_GMEM = 1, QSRC_HYB = 2 };

struct TmemPlan {
  TmemLaunch L;
  int cg, threads, ctas, qsrc;
  int cgc;  // column-group count compiled into the kernel variant to launch (0: run-time loop)
  int ktail;  // 2: the variant that skips the two padding rows of the last chunk (n = 4 cgc - 2: N = 30, 50, 70); 0: none
  size_t smem;
};

// One launch over MANY problem instances (grid = sum of the instances' CTAs): the reference's user
// loop over instance files (examples/ccvm_boxqp_*.py) folded into the grid.  Each CTA looks up
// (instance, CTA index inside the instance) and runs the same body with that instance's parameters.
struct BatchItem {
  SdeParams p;
  TmemLaunch L;
  FusedTail f;
};

// one bucket of a batched launch: CTAs [0, ctas) of `map` share block size, Q source and (when
// cgc != 0) the compiled-in column-group count
struct BatchBucket {
  const BatchItem* items;
  const int2* map;
  unsigned ctas;
  int threads, qsrc, cgc, ktail;
  size_t smem;
};

template <int SOLVER, bool ADAM>
int launch_tmem(const SdeParams& p, const TmemPlan& P, const FusedTail& f, cudaStream_t st);
template <int SOLVER, bool ADAM>
int launch_tmem_batch(const BatchBucket& b, cudaStream_t st);
template <int SOLVER, bool ADAM>
int regs_tmem(int qsrc);

// ---- tcgen05 3xTF32 kernels (sde_kernel_tc.cuh)
struct TcParams {
  float* xh;        // [2][rows_p][np]  hi part of the contraction input (ping-pong)
  float* xl;        // [2][rows_p][np]  lo part
  float* aux;       // [n_aux][rows_p][np] FP32 in-place state: MF mu, sigma; Adam m, v
  const float* hvec;    // [np] affine drift term h_j (0 in the padding)
  const float* svec;    // [np] clamp bound S_j (0 in the padding)
  int np;           // n rounded up to a multiple of TC_BN
  int rows;         // valid rows = K * batch
  int rows_p;       // rows rounded up to a multiple of TC_BM
  int col_split;    // CTA pairs (1, 2 or 4) that share one block of 256 rows, each owning np / 256 / col_split output chunks
  unsigned int* chunk_flags;  // col_split > 1: [rows_p / TC_BM][np / 256] counters, +2 per iteration and chunk written
};

struct TcPlan {
  int version;  // 1: single-CTA kernel (sde_tc_kernel), 2: CTA-pair kernel (sde_tc2_kernel)
  int np, rows, rows_p, n_aux, ctas;
  int col_split;  // see TcParams
  size_t smem;
};

struct TcMaps {
  CUtensorMap xh, xl, qh, ql, oh, ol;
};

template <int SOLVER, bool ADAM>
int launch_tc(const SdeParams& p, const TcParams& tc, const TcPlan& P, const TcMaps& M, cudaStream_t st);
template <int SOLVER, bool ADAM>
int regs_tc(int version);

// ---- small-n tensor-core kernel (sde_kernel_mma.cuh): contraction on tcgen05 with Qs^T resident in TMEM
constexpr int MMA_IPL_MAX = 8;          // items per lane compiled in for every tile
constexpr int MMA_IPL_MAX_LIGHT = 11;   // ... for every tile but DL-adam and MF-adam
struct MmaPlan {
  int nbp;     // trajectory pairs per warpgroup (<= 8): a CTA advances 4 nbp trajectories
  int ipl;     // (variable, pair) items per lane compiled into the kernel variant: ceil(ceil(n / 4) nbp / 32), >= 2
  int kd, tcols, mt;   // MmaLaunch
  int stagger;         // MmaLaunch: warpgroup 1 starts half an iteration after warpgroup 0
  int ctas, threads;
  size_t smem;
};
template <int SOLVER, bool ADAM>
int launch_mma(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st);
template <int SOLVER, bool ADAM>
int regs_mma(int ipl, int mt);

// dispatch on run-time (solver, algorithm): `CALL` is a macro taking (SOLVER, ADAM)
#define CCVM_DISPATCH_TILE(solver, adam, CALL)                 \
  switch ((solver) * 2 + ((adam) ? 1 : 0)) {                   \
    case 0: CALL(::ccvm::SOLVER_DL, false); break;             \
    case 1: CALL(::ccvm::SOLVER_DL, true); break;              \
    case 2: CALL(::ccvm::SOLVER_MF, false); break;             \
    case 3: CALL(::ccvm::SOLVER_MF, true); break;              \
    case 4: CALL(::ccvm::SOLVER_LV, false); break;             \
    case 5: CALL(::ccvm::SOLVER_LV, true); break;              \
    case 6: CALL(::ccvm::SOLVER_PLV, false); break;            \
    default: CALL(::ccvm::SOLVER_PLV, true); break;            \
  }

}  // namespace ccvm
