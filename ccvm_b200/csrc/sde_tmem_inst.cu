// Instantiations of the tiled SIMT kernels (sde_kernel_tmem.cuh) for ONE (solver, algorithm) pair:
// compiled eight times, with -DCCVM_INST_SOLVER=0..3 -DCCVM_INST_ADAM=0/1 (see sde_launch.h).
#include "sde_kernel_tmem.cuh"

#ifndef CCVM_INST_SOLVER
#error "compile with -DCCVM_INST_SOLVER=<0..3> -DCCVM_INST_ADAM=<0|1>"
#endif

namespace ccvm {

template <int SOLVER, bool ADAM, int QSRC, bool PIPE, int CGC = 0, int KTAIL = 0>
static int launch_tmem_variant(const SdeParams& p, const TmemPlan& P, const FusedTail& f, cudaStream_t st) {
  auto kern = sde_tmem_kernel<SOLVER, ADAM, QSRC, PIPE, CGC, KTAIL>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  kern<<<P.ctas, P.threads, P.smem, st>>>(p, P.L, f);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// Column-group counts of the reference's benchmarking sizes (N = 20 ... 70; examples/
// benchmarking_instances/Size*) compiled in: fully unrolled contraction with immediate addresses
// and (mostly) unpinned noise, see sde_kernel_tmem.cuh.  Measured at N = 70: DL + Adam 3.57 -> 3.39 ms,
// MF 2.05 -> 1.86, Langevin + Adam 2.03 -> 1.78, Langevin 1.73 -> 1.67; more at N = 20 ... 60.
// does this tile have the KTAIL = 2 variants (CCVM_KTAIL_MASK)?
template <int SOLVER, bool ADAM>
constexpr int ktail_of(int requested) {
  return ((CCVM_KTAIL_MASK >> (SOLVER * 2 + (ADAM ? 1 : 0))) & 1) ? requested : 0;
}

template <int SOLVER, bool ADAM>
int launch_tmem(const SdeParams& p, const TmemPlan& P, const FusedTail& f, cudaStream_t st) {
  const bool pipe = P.L.pipe != 0;
  constexpr int KT2 = ktail_of<SOLVER, ADAM>(2);   // 2, or 0 when the tile keeps all four rows
  if (P.qsrc == QSRC_TMEM && pipe) {
    switch (P.cgc) {
      case 5: return launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 5>(p, P, f, st);
      case 8: return P.ktail == 2 ? launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 8, KT2>(p, P, f, st)
                                  : launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 8>(p, P, f, st);
      case 10: return launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 10>(p, P, f, st);
      case 13: return P.ktail == 2 ? launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 13, KT2>(p, P, f, st)
                                   : launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 13>(p, P, f, st);
      case 15: return launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 15>(p, P, f, st);
      case 18: return P.ktail == 2 ? launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 18, KT2>(p, P, f, st)
                                   : launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true, 18>(p, P, f, st);
      default: break;
    }
  }
  if (P.qsrc == QSRC_TMEM)
    return pipe ? launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, true>(p, P, f, st)
                : launch_tmem_variant<SOLVER, ADAM, QSRC_TMEM, false>(p, P, f, st);
  if (P.qsrc == QSRC_HYB)
    return pipe ? launch_tmem_variant<SOLVER, ADAM, QSRC_HYB, true>(p, P, f, st)
                : launch_tmem_variant<SOLVER, ADAM, QSRC_HYB, false>(p, P, f, st);
  return pipe ? launch_tmem_variant<SOLVER, ADAM, QSRC_GMEM, true>(p, P, f, st)
              : launch_tmem_variant<SOLVER, ADAM, QSRC_GMEM, false>(p, P, f, st);
}

template <int SOLVER, bool ADAM, int QSRC, int CGC, int KTAIL = 0>
static int launch_batch_variant(const BatchBucket& b, cudaStream_t st) {
  auto kern = sde_tmem_batch_kernel<SOLVER, ADAM, QSRC, CGC, KTAIL>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem));
  kern<<<b.ctas, b.threads, b.smem, st>>>(b.items, b.map);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

template <int SOLVER, bool ADAM>
int launch_tmem_batch(const BatchBucket& b, cudaStream_t st) {
  constexpr int KT2 = ktail_of<SOLVER, ADAM>(2);
  if (b.qsrc == QSRC_HYB) return launch_batch_variant<SOLVER, ADAM, QSRC_HYB, 0>(b, st);
  switch (b.cgc) {
    case 5: return launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 5>(b, st);
    case 8: return b.ktail == 2 ? launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 8, KT2>(b, st)
                                : launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 8>(b, st);
    case 10: return launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 10>(b, st);
    case 13: return b.ktail == 2 ? launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 13, KT2>(b, st)
                                 : launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 13>(b, st);
    case 15: return launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 15>(b, st);
    case 18: return b.ktail == 2 ? launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 18, KT2>(b, st)
                                 : launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 18>(b, st);
    default: return launch_batch_variant<SOLVER, ADAM, QSRC_TMEM, 0>(b, st);
  }
}

template <int SOLVER, bool ADAM>
int regs_tmem(int qsrc) {
  cudaFuncAttributes fa;
  cudaError_t e = qsrc == QSRC_TMEM  ? cudaFuncGetAttributes(&fa, sde_tmem_kernel<SOLVER, ADAM, QSRC_TMEM, true>)
                  : qsrc == QSRC_HYB ? cudaFuncGetAttributes(&fa, sde_tmem_kernel<SOLVER, ADAM, QSRC_HYB, true>)
                                     : cudaFuncGetAttributes(&fa, sde_tmem_kernel<SOLVER, ADAM, QSRC_GMEM, true>);
  return e == cudaSuccess ? fa.numRegs : -1;
}

template int launch_tmem<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(const SdeParams&, const TmemPlan&, const FusedTail&,
                                                                   cudaStream_t);
template int launch_tmem_batch<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(const BatchBucket&, cudaStream_t);
template int regs_tmem<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(int);

}  // namespace ccvm
