// util.cu - small device utilities around the forward path.
#include "kernels.h"

namespace srs {

int64_t g_launch_count = 0;

// Counter-based uniform fill (splitmix64 of seed + (i+1)*golden): the synthetic
// 10^8-row movie table of BASELINE cfg 5 is generated in place in HBM; the oracle
// regenerates any row it needs from the same formula.
__global__ void fill_uniform_kernel(float* __restrict__ x, int64_t n, uint64_t seed, float lo,
                                    float hi) {
  const float span = hi - lo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    const float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
    x[i] = __fadd_rn(lo, __fmul_rn(span, u));
  }
}

cudaError_t launch_fill_uniform(float* x, int64_t n, uint64_t seed, float lo, float hi,
                                cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  fill_uniform_kernel<<<(int)blocks, threads, 0, s>>>(x, n, seed, lo, hi);
  ++g_launch_count;
  return cudaGetLastError();
}

// uint16 -> int32 history ids (srs_batch::hist16): the narrow form crosses PCIe, the kernels
// read int32.  `src` is 4-byte aligned (packed layout); two ids per thread.
__global__ void widen_u16_kernel(const uint16_t* __restrict__ src, int32_t* __restrict__ dst,
                                 int64_t n) {
  const int64_t pairs = n >> 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
    *reinterpret_cast<int2*>(dst + 2 * i) = make_int2((int)(v & 0xFFFFu), (int)(v >> 16));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = src[n - 1];
}

cudaError_t launch_widen_u16(const uint16_t* src, int32_t* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  int64_t blocks = ((n >> 1) + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  widen_u16_kernel<<<(int)blocks, threads, 0, s>>>(src, dst, n);
  ++g_launch_count;
  return cudaGetLastError();
}

// Cosine similarity of one query against n candidates, one warp per candidate.
// Reference: online/model/Embedding.java:33-47 - float products accumulated in double,
// dot / (sqrt(n1) * sqrt(n2)).
__global__ void cosine_kernel(const float* __restrict__ q, const float* __restrict__ c, int n,
                              int dim, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* v = c + (size_t)warp * dim;
  double dot = 0.0, n1 = 0.0, n2 = 0.0;
  for (int k = lane; k < dim; k += 32) {
    const float a = __ldg(q + k), bb = __ldg(v + k);
    dot += (double)__fmul_rn(a, bb);
    n1 += (double)__fmul_rn(a, a);
    n2 += (double)__fmul_rn(bb, bb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if (lane == 0) out[warp] = (float)(dot / (sqrt(n1) * sqrt(n2)));
}

cudaError_t launch_cosine(const float* q, const float* c, int n, int dim, float* out,
                          cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  const int blocks = (n * 32 + threads - 1) / threads;
  cosine_kernel<<<blocks, threads, 0, s>>>(q, c, n, dim, out);
  ++g_launch_count;
  return cudaGetLastError();
}

}  // namespace srs
