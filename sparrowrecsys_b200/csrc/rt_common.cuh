// rt_common.cuh - pieces shared by the row-tile DIN kernels (din_rt.cu: E <= 32, din_rt64.cu:
// E <= 64): operand descriptors beyond umma.cuh (MN-major, SWIZZLE_64B), cp.async, mbarrier
// arrive, named barriers.
#pragma once

#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int a_mn) {
  return idesc_bf16(M, N) | ((uint32_t)a_mn << 15);
}
// K-major, 64-byte rows (32 bf16), SWIZZLE_64B, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t chunk) {
  return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}
// MN-major SWIZZLE_128B: 64 MN elements (128 B) contiguous per K row, 8 K rows per 1024-byte
// atom, atoms along K `sbo` bytes apart (the history tile read "transposed")
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// non-blocking phase test (mbarrier.try_wait may suspend the thread for a system-dependent time before
// it answers "not yet": a loop that polls SEVERAL barriers must not sit in one of them)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__device__ __forceinline__ int rt_f32_roundtrip_id(int id) {   // DIN.py:95,125: ids pass through float32
  return __float2int_rz(__int2float_rn(id));
}

}  // namespace srs
