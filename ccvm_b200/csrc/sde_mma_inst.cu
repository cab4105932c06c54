// Instantiations of the small-n tensor-core kernel (sde_kernel_mma.cuh) for ONE (solver, algorithm) pair:
// compiled eight times, with -DCCVM_INST_SOLVER=0..3 -DCCVM_INST_ADAM=0/1 (see sde_launch.h).
#include "sde_kernel_mma.cuh"

#ifndef CCVM_INST_SOLVER
#error "compile with -DCCVM_INST_SOLVER=<0..3> -DCCVM_INST_ADAM=<0|1>"
#endif

namespace ccvm {

template <int SOLVER, bool ADAM, int NBP>
static int launch_mma_variant(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  auto kern = sde_mma_kernel<SOLVER, ADAM, NBP>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  MmaLaunch L;
  L.kd = P.kd;
  L.tcols = P.tcols;
  kern<<<P.ctas, MMA_THREADS, P.smem, st>>>(p, L, f);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

template <int SOLVER, bool ADAM>
int launch_mma(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  if (P.nbp == 7) return launch_mma_variant<SOLVER, ADAM, 7>(p, P, f, st);
  return launch_mma_variant<SOLVER, ADAM, 8>(p, P, f, st);
}

template <int SOLVER, bool ADAM>
int regs_mma(int nbp) {
  cudaFuncAttributes fa;
  cudaError_t e = nbp == 7 ? cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 7>)
                           : cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 8>);
  return e == cudaSuccess ? fa.numRegs : -1;
}

#ifdef CCVM_MMA_TRACE
}  // namespace ccvm
extern "C" int ccvm_debug_mma_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, ccvm::g_mma_trace, sizeof(long long) * 32 * 8);
}
namespace ccvm {
#endif

template int launch_mma<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(const SdeParams&, const MmaPlan&, const FusedTail&,
                                                                  cudaStream_t);
template int regs_mma<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(int);

}  // namespace ccvm
