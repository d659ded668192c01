// umma_selftest.cu - one-CTA known-answer kernel for the tcgen05 plumbing in umma.cuh:
// operand layouts (SW128 K-major smem tiles, packed-bf16 A operand in TMEM), descriptors,
// accumulator read-back.  D[128][N] = bf16(A)[128][K] * bf16(B)[N][K]^T with fp32 accumulate.
// Exposed as srs_selftest_umma (include/srs_ctr.h) and checked by tests/test_gpu_umma.py.
#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A,
                                                            const float* __restrict__ Bm,
                                                            float* __restrict__ D, int N, int KB,
                                                            int a_in_tmem) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = KB * 64;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = base;                          // KB tiles of 128 rows x 128 B
  uint8_t* sB = base + KB * 16384;             // KB tiles of N rows x 128 B (1024-aligned)

  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t d_tmem = tbase;               // columns [0, N)
  const uint32_t a_tmem = tbase + 64;          // columns [64, 64 + K/2)

  // A operand: row = tid
  for (int kb = 0; kb < KB; ++kb) {
    const float* arow = A + (size_t)tid * K + kb * 64;
    if (a_in_tmem) {
      uint32_t v[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pack_hi(arow[h * 32 + 2 * i], arow[h * 32 + 2 * i + 1]);
        tmem_st16(tmem_addr(a_tmem, (warp & 3) * 32, kb * 32 + h * 16), v);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 q;
        q.x = pack_hi(arow[8 * c + 0], arow[8 * c + 1]);
        q.y = pack_hi(arow[8 * c + 2], arow[8 * c + 3]);
        q.z = pack_hi(arow[8 * c + 4], arow[8 * c + 5]);
        q.w = pack_hi(arow[8 * c + 6], arow[8 * c + 7]);
        *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_offset(tid, c)) = q;
      }
    }
    if (tid < N) {
      const float* brow = Bm + (size_t)tid * K + kb * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 q;
        q.x = pack_hi(brow[8 * c + 0], brow[8 * c + 1]);
        q.y = pack_hi(brow[8 * c + 2], brow[8 * c + 3]);
        q.z = pack_hi(brow[8 * c + 4], brow[8 * c + 5]);
        q.w = pack_hi(brow[8 * c + 6], brow[8 * c + 7]);
        *reinterpret_cast<uint4*>(sB + kb * (N * 128) + sw128_offset(tid, c)) = q;
      }
    }
  }
  if (a_in_tmem) tmem_st_wait();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, N);
    uint32_t acc = 0;
    for (int kb = 0; kb < KB; ++kb) {
      const uint64_t ad = smem_desc_sw128(smem_u32(sA + kb * 16384));
      const uint64_t bd = smem_desc_sw128(smem_u32(sB + kb * (N * 128)));
#pragma unroll
      for (int k = 0; k < 4; ++k) {            // 4 K steps of 16 elements = 32 bytes each
        if (a_in_tmem) mma_ts(d_tmem, a_tmem + kb * 32 + k * 8, bd + 2 * k, idesc, acc);
        else mma_ss(d_tmem, ad + 2 * k, bd + 2 * k, idesc, acc);
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  if (N == 32) {
    tmem_ld32(tmem_addr(d_tmem, (warp & 3) * 32, 0), r);
  } else {
    uint32_t r16[16];
    tmem_ld16(tmem_addr(d_tmem, (warp & 3) * 32, 0), r16);
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = r16[i];
  }
  tmem_ld_wait();
  for (int j = 0; j < N; ++j) D[(size_t)tid * N + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

cudaError_t launch_umma_selftest(const float* A, const float* B, float* D, int N, int KB,
                                 int a_in_tmem, cudaStream_t s) {
  if ((N != 16 && N != 32) || KB < 1 || KB > 3) return cudaErrorInvalidValue;
  const size_t smem = 1024 + (size_t)KB * 16384 + (size_t)KB * 32 * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  umma_selftest_kernel<<<1, 128, smem, s>>>(A, B, D, N, KB, a_in_tmem);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace srs
