// umma_bench.cu - micro-benchmark of small tcgen05.mma shapes (one CTA, one issuing thread):
// cycles for a chain of `n_mma` MMAs, M = 128, K = 16 (bf16), for N in {32, 64, 128},
// A from shared memory or tensor memory, all into one accumulator or alternating two.
// Used to size the DIN activation-unit MMAs; srs_debug_umma_bench.
#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

__global__ void __launch_bounds__(128) umma_bench_kernel(unsigned long long* out, int N, int n_mma,
                                                         int a_in_tmem, int two_acc, int uniform) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = base;             // 16 KB
  uint8_t* sB = base + 16384;     // up to 128 rows x 128 B = 16 KB
  for (int i = tid; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  uint32_t zero[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) zero[i] = 0;
  tmem_st16(tb + 256 + ((uint32_t)(warp * 32) << 16), zero);
  tmem_st16(tb + 272 + ((uint32_t)(warp * 32) << 16), zero);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (uniform) {
    // warp-uniform issue: the whole warp runs the loop, one elected lane issues
    if (warp == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_bf16(128, N);
      const uint64_t ad = smem_desc_sw128(smem_u32(sA)), bd = smem_desc_sw128(smem_u32(sB));
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tb + ((two_acc && (i & 1)) ? 128 : 0);
        const int ks = i & 3;
        if (elect_one()) {
          if (a_in_tmem) mma_ts(d, tb + 256 + 8 * ks, bd + 2 * ks, idesc, i >= 2);
          else mma_ss(d, ad + 2 * ks, bd + 2 * ks, idesc, i >= 2);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 0) {
        out[0] = (unsigned long long)(t1 - t0);
        out[1] = (unsigned long long)(t2 - t0);
      }
    }
  } else if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, N);
    const uint64_t ad = smem_desc_sw128(smem_u32(sA)), bd = smem_desc_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tb + ((two_acc && (i & 1)) ? 128 : 0);
      const int ks = i & 3;
      if (a_in_tmem) mma_ts(d, tb + 256 + 8 * ks, bd + 2 * ks, idesc, i >= 2);
      else mma_ss(d, ad + 2 * ks, bd + 2 * ks, idesc, i >= 2);
    }
    const long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = (unsigned long long)(t1 - t0);
    out[1] = (unsigned long long)(t2 - t0);
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

cudaError_t launch_umma_bench(unsigned long long* out, int N, int n_mma, int a_in_tmem, int two_acc,
                              int uniform, cudaStream_t s) {
  const size_t smem = 1024 + 32768;
  cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  umma_bench_kernel<<<1, 128, smem, s>>>(out, N, n_mma, a_in_tmem, two_acc, uniform);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace srs
