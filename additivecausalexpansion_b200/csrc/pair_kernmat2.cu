// pair_kernmat2.cu -- instantiations of kernmat2_kernel (kernmat2_kernel.cuh) for one kernel kind, B = 1..16 terms.
// Compiled once per -DKB2_KIND={0,1} (see the Makefile); each object exports one launcher.  The caller holds the
// per-device serialisation of the constant table (ace_b200.cu: ConstChain).
#include "kernmat2_kernel.cuh"

#if !defined(KB2_KIND)
#error "compile with -DKB2_KIND=<0|1>"
#endif

namespace ace {

template <int BX, bool CUBE, bool SYM>
static int kb2_launch(const KernArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  auto kern = kernmat2_kernel<KB2_KIND, BX, CUBE, SYM>;
  ACE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ACE_CUDA(cudaMemcpyToSymbolAsync(cKB, a.tab + TAB_G3, sizeof(double) * G3_SIZE, 0, cudaMemcpyDeviceToDevice, st));
  kern<<<grid, 256, smem, st>>>(a);
  ACE_CUDA(cudaGetLastError());
  return 0;
}

template <int BX>
static int kb2_go(const KernArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  if (a.sym) return a.cube != nullptr ? kb2_launch<BX, true, true>(a, grid, smem, st) : kb2_launch<BX, false, true>(a, grid, smem, st);
  return a.cube != nullptr ? kb2_launch<BX, true, false>(a, grid, smem, st) : kb2_launch<BX, false, false>(a, grid, smem, st);
}

#define KB2_CAT2(a, b) a##b
#define KB2_CAT(a, b) KB2_CAT2(a, b)
int KB2_CAT(launch_kernmat2_k, KB2_KIND)(const KernArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  switch (a.B) {
    case 1: return kb2_go<1>(a, grid, smem, st);
    case 2: return kb2_go<2>(a, grid, smem, st);
    case 3: return kb2_go<3>(a, grid, smem, st);
    case 4: return kb2_go<4>(a, grid, smem, st);
    case 5: return kb2_go<5>(a, grid, smem, st);
    case 6: return kb2_go<6>(a, grid, smem, st);
    case 7: return kb2_go<7>(a, grid, smem, st);
    case 8: return kb2_go<8>(a, grid, smem, st);
    case 9: return kb2_go<9>(a, grid, smem, st);
    case 10: return kb2_go<10>(a, grid, smem, st);
    case 11: return kb2_go<11>(a, grid, smem, st);
    case 12: return kb2_go<12>(a, grid, smem, st);
    case 13: return kb2_go<13>(a, grid, smem, st);
    case 14: return kb2_go<14>(a, grid, smem, st);
    case 15: return kb2_go<15>(a, grid, smem, st);
    case 16: return kb2_go<16>(a, grid, smem, st);
    default: set_error("kernmat2: B out of range"); return -1;
  }
}

}  // namespace ace
