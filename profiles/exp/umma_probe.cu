// umma_probe.cu - standalone probe (not part of the library) for the operand forms the
// row-tile DIN kernel relies on:
//   1. a [positions][64 bf16] SW128 tile used K-major (M = positions) AND MN-major (M = the 64
//      columns, K = positions) by two different tcgen05.mma instructions,
//   2. the accumulator lane map of M = 64,
//   3. the issue cost of small MMAs when 1, 2 or 4 warps issue concurrently.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I sparrowrecsys_b200/csrc -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace srs::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__host__ __device__ constexpr uint32_t idesc_full(int M, int N, int a_mn, int b_mn) {
  return idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
// MN-major SW128: 64 MN elements (128 B) contiguous per K row, 8 K rows per 1024-B atom,
// K groups `sbo` bytes apart, MN groups `lbo` bytes apart
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------- correctness: tile[128 pos][64] bf16; D1 = tile * B1^T (M=128,N=32,K=64);
// D2 = tile[0:64]^T-as-A (M=64 cols, K=64 pos) * B2[8][64 pos]^T
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {   // K-major, 64-byte rows, 8-row groups 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ inline uint32_t sw64_offset(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

__global__ void __launch_bounds__(128) probe_kernel(const float* tile_f, const float* b1_f, const float* b2_f,
                                                    float* d1, float* d2raw, const float* b3_f, float* d3) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sT = base;             // 128 x 128 B = 16 KB
  uint8_t* sB1 = base + 16384;    // 32 x 128 B = 4 KB
  uint8_t* sB2 = base + 20480;    // 8 x 128 B = 1 KB
  uint8_t* sB3 = base + 22528;    // 128 x 64 B = 8 KB, SW64 (1024-aligned)
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  auto put_row = [&](uint8_t* dst, int row, const float* src) {
    for (int c = 0; c < 8; ++c) {
      uint4 q;
      q.x = pack_hi(src[8 * c + 0], src[8 * c + 1]);
      q.y = pack_hi(src[8 * c + 2], src[8 * c + 3]);
      q.z = pack_hi(src[8 * c + 4], src[8 * c + 5]);
      q.w = pack_hi(src[8 * c + 6], src[8 * c + 7]);
      *reinterpret_cast<uint4*>(dst + sw128_offset(row, c)) = q;
    }
  };
  put_row(sT, tid, tile_f + tid * 64);
  if (tid < 32) put_row(sB1, tid, b1_f + tid * 64);
  if (tid < 8) put_row(sB2, tid, b2_f + tid * 64);
  for (int c = 0; c < 4; ++c) {
    const float* src = b3_f + tid * 32 + 8 * c;
    uint4 q;
    q.x = pack_hi(src[0], src[1]); q.y = pack_hi(src[2], src[3]); q.z = pack_hi(src[4], src[5]); q.w = pack_hi(src[6], src[7]);
    *reinterpret_cast<uint4*>(sB3 + sw64_offset(tid, c)) = q;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    const uint32_t i1 = idesc_full(128, 32, 0, 0);
    const uint64_t ad = smem_desc_sw128(smem_u32(sT)), bd = smem_desc_sw128(smem_u32(sB1));
    for (int k = 0; k < 4; ++k) mma_ss(tb, ad + 2 * k, bd + 2 * k, i1, k > 0);
    const uint32_t i2 = idesc_full(64, 8, 1, 0);
    const uint64_t b2d = smem_desc_sw128(smem_u32(sB2));
    for (int k = 0; k < 4; ++k) {
      const uint64_t a2 = smem_desc_mn_sw128(smem_u32(sT) + k * 2048, 1024, 1024);
      mma_ss(tb + 64, a2, b2d + 2 * k, i2, k > 0);
    }
    // D3[128][128] = tile[:, 0:32] * B3^T, then lo-style: tile[:, 32:64] * B3[0:64]^T accumulated into cols 0..63
    const uint32_t i3 = idesc_full(128, 128, 0, 0), i4 = idesc_full(128, 64, 0, 0);
    const uint64_t b3d = smem_desc_sw64(smem_u32(sB3));
    for (int k = 0; k < 2; ++k) mma_ss(tb + 128, ad + 2 * k, b3d + 2 * k, i3, k > 0);
    for (int k = 0; k < 2; ++k) mma_ss(tb + 128, ad + 2 * (k + 2), b3d + 2 * k, i4, 1);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int cb = 0; cb < 4; ++cb) {
    tmem_ld32(tmem_addr(tb, warp * 32, 128 + 32 * cb), r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d3[tid * 128 + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tmem_ld32(tmem_addr(tb, warp * 32, 0), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) d1[tid * 32 + j] = __uint_as_float(r[j]);
  uint32_t r8[8];
  tmem_ld8(tmem_addr(tb, warp * 32, 64), r8);
  tmem_ld_wait();
  for (int j = 0; j < 8; ++j) d2raw[tid * 8 + j] = __uint_as_float(r8[j]);   // by TMEM lane
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

// ---------------- timing: W warps issue n MMAs each, own accumulators
// shape 0: M=128 N=Nn SS K-major;  shape 1: M=64 N=8, A MN-major;  shape 2: alternate both
__global__ void __launch_bounds__(128) issue_kernel(unsigned long long* out, int W, int n, int shape, int Nn, int nacc) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < W) {
    const uint32_t i_big = idesc_full(128, Nn, 0, 0), i_pool = idesc_full(64, 8, 1, 0);
    const uint32_t sA = smem_u32(base), sB = smem_u32(base) + 16384 + warp * 0;
    const uint64_t ad = smem_desc_sw128(sA), bd = smem_desc_sw128(sB + 16384);
    const uint32_t d = tb + warp * 128;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const int ks = i & 3;
      const bool pool = shape == 1 || (shape == 2 && (i & 1));
      if (elect_one()) {
        if (pool) mma_ss(d + 120, smem_desc_mn_sw128(sA + ks * 2048, 1024, 1024), bd + 2 * ks, i_pool, i >= 4);
        else mma_ss(d + (uint32_t)((i / 4) % nacc) * Nn, ad + 2 * ks, bd + 2 * ks, i_big, i >= 4 * nacc);
      }
      __syncwarp();
    }
    t1 = clock64();
    if (elect_one()) mma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    t2 = clock64();
    if ((tid & 31) == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static float bf16t(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFF0000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
  std::vector<float> tile(128 * 64), b1(32 * 64), b2(8 * 64), b3(128 * 32);
  srand(1);
  auto rnd = [] { return bf16t((float)(rand() % 2001 - 1000) / 512.f); };
  for (auto& v : tile) v = rnd();
  for (auto& v : b1) v = rnd();
  for (auto& v : b2) v = rnd();
  for (auto& v : b3) v = rnd();
  float *dt, *db1, *db2, *dd1, *dd2;
  CK(cudaMalloc(&dt, tile.size() * 4)); CK(cudaMalloc(&db1, b1.size() * 4)); CK(cudaMalloc(&db2, b2.size() * 4));
  CK(cudaMalloc(&dd1, 128 * 32 * 4)); CK(cudaMalloc(&dd2, 128 * 8 * 4));
  CK(cudaMemcpy(dt, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd2, 0xFF, 128 * 8 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  float *db3, *dd3;
  CK(cudaMalloc(&db3, b3.size() * 4)); CK(cudaMalloc(&dd3, 128 * 128 * 4));
  CK(cudaMemcpy(db3, b3.data(), b3.size() * 4, cudaMemcpyHostToDevice));
  probe_kernel<<<1, 128, 32768>>>(dt, db1, db2, dd1, dd2, db3, dd3);
  CK(cudaDeviceSynchronize());
  std::vector<float> d1(128 * 32), d2(128 * 8);
  CK(cudaMemcpy(d1.data(), dd1, d1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(d2.data(), dd2, d2.size() * 4, cudaMemcpyDeviceToHost));
  double e1 = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
    double s = 0; for (int k = 0; k < 64; ++k) s += (double)tile[m * 64 + k] * b1[n * 64 + k];
    e1 = fmax(e1, fabs(s - d1[m * 32 + n]));
  }
  printf("K-major M=128 N=32: max err %g\n", e1);
  // pool reference: D2[m][n] = sum_t tile[t][m] * b2[n][t], t < 64
  double e2 = 0; int bad = 0;
  for (int m = 0; m < 64; ++m) for (int n = 0; n < 8; ++n) {
    double s = 0; for (int t = 0; t < 64; ++t) s += (double)tile[t * 64 + m] * b2[n * 64 + t];
    const int lane = (m % 16) + 32 * (m / 16);
    const double err = fabs(s - d2[lane * 8 + n]);
    if (err > 1e-3) ++bad;
    e2 = fmax(e2, err);
  }
  printf("MN-major A M=64 N=8 (lane = m%%16 + 32*(m/16)): max err %g, bad %d\n", e2, bad);
  if (bad) {
    printf("  lane dump col0: "); for (int l = 0; l < 128; ++l) printf("%g ", d2[l * 8]); printf("\n  ref col0: ");
    for (int m = 0; m < 64; ++m) { double s = 0; for (int t = 0; t < 64; ++t) s += (double)tile[t * 64 + m] * b2[t]; printf("%g ", s); }
    printf("\n");
  }
  {
    std::vector<float> d3(128 * 128);
    CK(cudaMemcpy(d3.data(), dd3, d3.size() * 4, cudaMemcpyDeviceToHost));
    double e3 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
      double sum = 0;
      for (int k = 0; k < 32; ++k) sum += (double)tile[m * 64 + k] * b3[n * 32 + k];
      if (n < 64) for (int k = 0; k < 32; ++k) sum += (double)tile[m * 64 + 32 + k] * b3[n * 32 + k];
      e3 = fmax(e3, fabs(sum - d3[m * 128 + n]));
    }
    printf("B in SWIZZLE_64B [128][32], N=128 then N=64 accumulate: max err %g\n", e3);
  }
  unsigned long long* dout; CK(cudaMalloc(&dout, 64));
  CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 65536));
  if (argc < 2) return 0;
  printf("%-6s %-3s %-4s %-5s %10s %10s %10s\n", "shape", "W", "N", "n", "issue_cyc", "total_cyc", "cyc/mma/SM");
  for (int shape = 0; shape < 3; ++shape)
    for (int Nn : {32, 64, 128, 256}) {
      if (shape == 1 && Nn != 32) continue;
      for (int W : {1, 2, 4}) {
        if (Nn == 256 && W > 2) continue;   // 2 x 256 columns
        for (int n : {16, 128}) {
          unsigned long long best_i = 0, best_t = ~0ull;
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(dout, 0, 64));
            issue_kernel<<<1, 128, 1024 + 65536>>>(dout, W, n, shape, Nn, 1);
            CK(cudaDeviceSynchronize());
            unsigned long long o[8]; CK(cudaMemcpy(o, dout, 64, cudaMemcpyDeviceToHost));
            unsigned long long mi = 0, mt = 0;
            for (int w = 0; w < W; ++w) { mi = o[2 * w] > mi ? o[2 * w] : mi; mt = o[2 * w + 1] > mt ? o[2 * w + 1] : mt; }
            if (mt < best_t) { best_t = mt; best_i = mi; }
          }
          printf("%-6d %-3d %-4d %-5d %10llu %10llu %10.1f\n", shape, W, Nn, n, best_i, best_t, (double)best_t / (n * W));
        }
      }
    }
  printf("single warp, chains of 4 accumulating MMAs round-robin over nacc accumulators (M=128)\n");
  for (int Nn : {32, 64, 128})
    for (int nacc : {1, 2, 4}) {
      if (Nn * nacc > 512) continue;
      unsigned long long best_t = ~0ull, best_i = 0;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(dout, 0, 64));
        issue_kernel<<<1, 128, 1024 + 65536>>>(dout, 1, 128, 0, Nn, nacc);
        CK(cudaDeviceSynchronize());
        unsigned long long o[8]; CK(cudaMemcpy(o, dout, 64, cudaMemcpyDeviceToHost));
        if (o[1] < best_t) { best_t = o[1]; best_i = o[0]; }
      }
      printf("N=%-4d nacc=%d: issue %llu total %llu -> %.1f cyc/mma\n", Nn, nacc, best_i, best_t, best_t / 128.0);
    }
  return 0;
}
