// common.cuh -- shared device helpers for the ACE B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace ace {

constexpr int TB = 128;  // block size of every blocked dense algorithm (diagonal tile edge)

// ---------------------------------------------------------------------------------------------
// error handling: CUDA failures become negative status codes + a message for ace_last_error()
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);

#define ACE_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::ace::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                       ":" + std::to_string(__LINE__) + ")");                                  \
      return -(int)_e - 1000;                                                                  \
    }                                                                                          \
  } while (0)

#define ACE_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy (cp.async.bulk -> SASS UBLKCP), FP64 tensor-core MMA
// (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind, so this is the FP64
// tensor path on Blackwell).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// order generic-proxy smem accesses before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core.
//   a : A[lane/4][lane%4]      b : B[lane%4][lane/4]      c0,c1 : C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ace
