// Instantiations of the tcgen05 3xTF32 kernels (sde_kernel_tc.cuh) for ONE solver, both algorithms:
// compiled four times, with -DCCVM_INST_SOLVER=0..3 (see sde_launch.h).
#include "sde_kernel_tc.cuh"

#ifndef CCVM_INST_SOLVER
#error "compile with -DCCVM_INST_SOLVER=<0..3>"
#endif

namespace ccvm {

template <int SOLVER, bool ADAM>
int launch_tc(const SdeParams& p, const TcParams& tc, const TcPlan& P, const TcMaps& M, cudaStream_t st) {
  const size_t plane = (size_t)tc.rows_p * tc.np;
  tc_init_state_kernel<SOLVER><<<(unsigned)((plane / 4 + 255) / 256), 256, 0, st>>>(p, tc, P.n_aux);
  CUDA_TRY(cudaGetLastError());
  if (P.version == 2) {
    auto kern = sde_tc2_kernel<SOLVER, ADAM>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    kern<<<P.ctas, TC_THREADS, P.smem, st>>>(p, tc, M.xh, M.xl, M.qh, M.ql, M.oh, M.ol);
  } else {
    auto kern = sde_tc_kernel<SOLVER, ADAM>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    kern<<<P.ctas, TC_THREADS, P.smem, st>>>(p, tc, M.xh, M.xl, M.qh, M.ql);
  }
  CUDA_TRY(cudaGetLastError());
  return CCVM_OK;
}

template <int SOLVER, bool ADAM>
int regs_tc(int version) {
  cudaFuncAttributes fa;
  const cudaError_t e = version == 2 ? cudaFuncGetAttributes(&fa, sde_tc2_kernel<SOLVER, ADAM>)
                                     : cudaFuncGetAttributes(&fa, sde_tc_kernel<SOLVER, ADAM>);
  return e == cudaSuccess ? fa.numRegs : -1;
}

template int launch_tc<CCVM_INST_SOLVER, false>(const SdeParams&, const TcParams&, const TcPlan&, const TcMaps&, cudaStream_t);
template int launch_tc<CCVM_INST_SOLVER, true>(const SdeParams&, const TcParams&, const TcPlan&, const TcMaps&, cudaStream_t);
template int regs_tc<CCVM_INST_SOLVER, false>(int);
template int regs_tc<CCVM_INST_SOLVER, true>(int);

}  // namespace ccvm
