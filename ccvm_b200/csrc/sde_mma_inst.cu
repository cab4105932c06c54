// Instantiations of the small-n tensor-core kernel (sde_kernel_mma.cuh) for ONE (solver, algorithm) pair:
// compiled eight times, with -DCCVM_INST_SOLVER=0..3 -DCCVM_INST_ADAM=0/1 (see sde_launch.h).
#include "sde_kernel_mma.cuh"

#ifndef CCVM_INST_SOLVER
#error "compile with -DCCVM_INST_SOLVER=<0..3> -DCCVM_INST_ADAM=<0|1>"
#endif

namespace ccvm {

template <int SOLVER, bool ADAM, int IPL>
static int launch_mma_variant(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  auto kern = sde_mma_kernel<SOLVER, ADAM, IPL>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  MmaLaunch L;
  L.kd = P.kd;
  L.tcols = P.tcols;
  L.nbp = P.nbp;
  L.stagger = P.stagger;
  kern<<<P.ctas, MMA_THREADS, P.smem, st>>>(p, L, f);
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

// items per lane compiled in: 2 ... 8 (n = 70 with 7 pairs per warpgroup: 4)
template <int SOLVER, bool ADAM>
int launch_mma(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  switch (P.ipl) {
    case 2: return launch_mma_variant<SOLVER, ADAM, 2>(p, P, f, st);
    case 3: return launch_mma_variant<SOLVER, ADAM, 3>(p, P, f, st);
    case 4: return launch_mma_variant<SOLVER, ADAM, 4>(p, P, f, st);
    case 5: return launch_mma_variant<SOLVER, ADAM, 5>(p, P, f, st);
    case 6: return launch_mma_variant<SOLVER, ADAM, 6>(p, P, f, st);
    case 7: return launch_mma_variant<SOLVER, ADAM, 7>(p, P, f, st);
    default: return launch_mma_variant<SOLVER, ADAM, 8>(p, P, f, st);
  }
}

template <int SOLVER, bool ADAM>
int regs_mma(int ipl) {
  cudaFuncAttributes fa;
  cudaError_t e;
  switch (ipl) {
    case 2: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 2>); break;
    case 3: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 3>); break;
    case 4: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 4>); break;
    case 5: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 5>); break;
    case 6: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 6>); break;
    case 7: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 7>); break;
    default: e = cudaFuncGetAttributes(&fa, sde_mma_kernel<SOLVER, ADAM, 8>); break;
  }
  return e == cudaSuccess ? fa.numRegs : -1;
}

#ifdef CCVM_MMA_TRACE
}  // namespace ccvm
extern "C" int ccvm_debug_mma_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, ccvm::g_mma_trace, sizeof(long long) * 32 * 16);
}
namespace ccvm {
#endif

template int launch_mma<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(const SdeParams&, const MmaPlan&, const FusedTail&,
                                                                  cudaStream_t);
template int regs_mma<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(int);

}  // namespace ccvm
