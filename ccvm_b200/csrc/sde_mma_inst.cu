// Instantiations of the small-n tensor-core kernel (sde_kernel_mma.cuh) for ONE (solver, algorithm) pair:
// compiled eight times, with -DCCVM_INST_SOLVER=0..3 -DCCVM_INST_ADAM=0/1 (see sde_launch.h).
#include "sde_kernel_mma.cuh"

#ifndef CCVM_INST_SOLVER
#error "compile with -DCCVM_INST_SOLVER=<0..3> -DCCVM_INST_ADAM=<0|1>"
#endif

namespace ccvm {

template <int SOLVER, bool ADAM, int IPL, int MT>
static int launch_mma_tiles(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  auto kern = sde_mma_kernel<SOLVER, ADAM, IPL, MT>;
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

// one or two M tiles of 128 variables
template <int SOLVER, bool ADAM, int IPL>
static int launch_mma_variant(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  return P.mt == 2 ? launch_mma_tiles<SOLVER, ADAM, IPL, 2>(p, P, f, st) : launch_mma_tiles<SOLVER, ADAM, IPL, 1>(p, P, f, st);
}
template <int SOLVER, bool ADAM, int IPL>
static cudaError_t attrs_mma_variant(cudaFuncAttributes* fa, int mt) {
  return mt == 2 ? cudaFuncGetAttributes(fa, sde_mma_kernel<SOLVER, ADAM, IPL, 2>)
                 : cudaFuncGetAttributes(fa, sde_mma_kernel<SOLVER, ADAM, IPL, 1>);
}

// items per lane compiled in: 2 ... 8 (n = 70 with 7 pairs per warpgroup: 4), 9 ... 11 for every tile but DL-adam and MF-adam
// (n = 129 ... 192 with 7 pairs per warpgroup: one wave of CTAs at B = 4096; with the Adam moments of two quadratures, or of
// MF's mean next to its variance, the spills cost more than the second wave: profiles/r2z_two_m_tiles.txt)
template <int SOLVER, bool ADAM>
constexpr bool mma_light_tile() { return !ADAM || SOLVER == SOLVER_LV || SOLVER == SOLVER_PLV; }

template <int SOLVER, bool ADAM>
int launch_mma(const SdeParams& p, const MmaPlan& P, const FusedTail& f, cudaStream_t st) {
  if constexpr (mma_light_tile<SOLVER, ADAM>()) {
    switch (P.ipl) {
      case 9: return launch_mma_variant<SOLVER, ADAM, 9>(p, P, f, st);
      case 10: return launch_mma_variant<SOLVER, ADAM, 10>(p, P, f, st);
      case 11: return launch_mma_variant<SOLVER, ADAM, 11>(p, P, f, st);
      default: break;
    }
  }
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
int regs_mma(int ipl, int mt) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaErrorInvalidValue;
  if constexpr (mma_light_tile<SOLVER, ADAM>()) {
    switch (ipl) {
      case 9: e = attrs_mma_variant<SOLVER, ADAM, 9>(&fa, mt); break;
      case 10: e = attrs_mma_variant<SOLVER, ADAM, 10>(&fa, mt); break;
      case 11: e = attrs_mma_variant<SOLVER, ADAM, 11>(&fa, mt); break;
      default: break;
    }
    if (ipl >= 9 && ipl <= 11) return e == cudaSuccess ? fa.numRegs : -1;
  }
  switch (ipl) {
    case 2: e = attrs_mma_variant<SOLVER, ADAM, 2>(&fa, mt); break;
    case 3: e = attrs_mma_variant<SOLVER, ADAM, 3>(&fa, mt); break;
    case 4: e = attrs_mma_variant<SOLVER, ADAM, 4>(&fa, mt); break;
    case 5: e = attrs_mma_variant<SOLVER, ADAM, 5>(&fa, mt); break;
    case 6: e = attrs_mma_variant<SOLVER, ADAM, 6>(&fa, mt); break;
    case 7: e = attrs_mma_variant<SOLVER, ADAM, 7>(&fa, mt); break;
    default: e = attrs_mma_variant<SOLVER, ADAM, 8>(&fa, mt); break;
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
template int regs_mma<CCVM_INST_SOLVER, (CCVM_INST_ADAM != 0)>(int, int);

}  // namespace ccvm
