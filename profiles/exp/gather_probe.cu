// gather_probe.cu - standalone probe (not part of the library): how fast can one SM / the chip
// gather 128-byte table rows into a 128B-swizzled shared-memory tile?
//   method 0: TMA tile::gather4 (tensor map, 4 rows per instruction, hardware swizzle)
//   method 1: cp.async.bulk 1D, one 128-byte row per instruction (no swizzle; throughput only)
//   method 2: cp.async (LDGSTS) 16 bytes per thread, software swizzle
// and checks that method 0 lands the rows in the SW128 K-major layout tcgen05.mma expects.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "umma.cuh"
using namespace srs::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int DEPTH = 8;          // tiles in flight per CTA
constexpr int TILE_POS = 64;      // positions (rows of 128 B) per tile

__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int x, int y0, int y1, int y2, int y3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y0), "r"(y1), "r"(y2), "r"(y3)
      : "memory");
}

__global__ void __launch_bounds__(128) gather_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* table,
                                                     const int* ids, int tiles_per_cta, int method, int vmax,
                                                     uint8_t* dump, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t full[DEPTH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  if (tid == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  const int* my_ids = ids + (size_t)(blockIdx.x % 148) * tiles_per_cta * TILE_POS;
  const long long t0 = clock64();
  if (method == 0 || method == 1) {
    if (warp == 0) {
      for (int i = 0; i < tiles_per_cta; ++i) {
        const int slot = i % DEPTH;
        if (i >= DEPTH) mbar_wait(&full[slot], ((i / DEPTH) - 1) & 1);
        uint8_t* tile = base + slot * (TILE_POS * 128);
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], TILE_POS * 128);
        __syncwarp();
        if (method == 0) {
          if (lane < TILE_POS / 4) {
            const int4 id4 = __ldg(reinterpret_cast<const int4*>(my_ids + i * TILE_POS) + lane);
            tma_gather4(tile + lane * 512, &map, 0, id4.x, id4.y, id4.z, id4.w, &full[slot]);
          }
        } else {
          int2 id2 = __ldg(reinterpret_cast<const int2*>(my_ids + i * TILE_POS) + lane);
          id2.x = min(max(id2.x, 0), vmax); id2.y = min(max(id2.y, 0), vmax);
          bulk_g2s(tile + (2 * lane) * 128, table + (size_t)id2.x * 128, 128, &full[slot]);
          bulk_g2s(tile + (2 * lane + 1) * 128, table + (size_t)id2.y * 128, 128, &full[slot]);
        }
      }
      for (int i = max(0, tiles_per_cta - DEPTH); i < tiles_per_cta; ++i) mbar_wait(&full[i % DEPTH], (i / DEPTH) & 1);
    }
  } else if (method == 4) {
    for (int i = 0; i < tiles_per_cta; ++i) {
      const int slot = i % DEPTH;
      uint8_t* tile = base + slot * (TILE_POS * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = 16 * j + (tid >> 3), c = tid & 7;
        const int id = min(max(__ldg(my_ids + i * TILE_POS + pos), 0), vmax);
        const uint8_t* src = table + (size_t)id * 128 + c * 16;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + sw128_offset(pos, c))), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (method == 3) {
    for (int i = 0; i < tiles_per_cta; ++i) {
      const int slot = i % DEPTH;
      uint8_t* tile = base + slot * (TILE_POS * 128);
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = 16 * j + (tid >> 3), c = tid & 7;
        const int id = min(max(__ldg(my_ids + i * TILE_POS + pos), 0), vmax);
        v[j] = __ldg(reinterpret_cast<const uint4*>(table + (size_t)id * 128 + c * 16));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = 16 * j + (tid >> 3), c = tid & 7;
        *reinterpret_cast<uint4*>(tile + sw128_offset(pos, c)) = v[j];
      }
    }
  } else {
    // 128 threads x 4 chunks of 16 B = 64 positions x 128 B
    for (int i = 0; i < tiles_per_cta; ++i) {
      const int slot = i % DEPTH;
      uint8_t* tile = base + slot * (TILE_POS * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = 16 * j + (tid >> 3), c = tid & 7;
        const int id = min(max(__ldg(my_ids + i * TILE_POS + pos), 0), vmax);
        const uint8_t* src = table + (size_t)id * 128 + c * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + sw128_offset(pos, c))), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (dump && blockIdx.x == 0) {
    const int slot = (tiles_per_cta - 1) % DEPTH;
    for (int i = tid; i < TILE_POS * 128; i += 128) dump[i] = base[slot * TILE_POS * 128 + i];
  }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int zipf = argc > 2 ? atoi(argv[2]) : 0;
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int n_sms = 148;
  for (int big = 0; big < (zipf ? 1 : 2); ++big) {
    const size_t V = big ? (size_t)1 << 25 : 27279;        // rows of 128 B: 4 GB or 3.5 MB
    uint8_t* table; CK(cudaMalloc(&table, V * 128));
    std::vector<uint16_t> host;
    if (!big) {
      host.resize(V * 64);
      for (size_t i = 0; i < host.size(); ++i) host[i] = (uint16_t)(i * 2654435761u >> 7);
      CK(cudaMemcpy(table, host.data(), V * 128, cudaMemcpyHostToDevice));
    } else CK(cudaMemset(table, 1, V * 128));
    CUtensorMap map;
    cuuint64_t gdim[2] = {64, (cuuint64_t)V};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, table, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("table %zu rows: cuTensorMapEncodeTiled(box rows %d) -> %d\n", V, box_rows, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int tiles_per_cta = 256;
    const size_t n_ids = (size_t)n_sms * tiles_per_cta * TILE_POS;
    std::vector<int> ids(n_ids);
    srand(7);
    for (auto& v : ids) {
      if (zipf) {   // ~Zipf(1): id = floor(V^u) - 1, half of the cells are padding id 0
        const double u = (double)rand() / RAND_MAX;
        v = (rand() & 1) ? 0 : (int)(pow((double)V, u)) - 1;
        if (v < 0) v = 0; if (v >= (int)V) v = (int)V - 1;
      } else v = (int)(((size_t)rand() * 32768u + rand()) % V);
    }
    ids[n_ids - 1] = -1; ids[(size_t)(tiles_per_cta - 1) * TILE_POS + 5] = -1; ids[(size_t)(tiles_per_cta - 1) * TILE_POS + 6] = (int)V;  // out of bounds -> zeros?
    int* dids; CK(cudaMalloc(&dids, n_ids * 4));
    CK(cudaMemcpy(dids, ids.data(), n_ids * 4, cudaMemcpyHostToDevice));
    uint8_t* ddump; CK(cudaMalloc(&ddump, TILE_POS * 128));
    unsigned long long* dcyc; CK(cudaMalloc(&dcyc, n_sms * 8));
    const int smem = 1024 + DEPTH * TILE_POS * 128;
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int method = 2; method < 5; method += 2) {
      if (method != 0 && false) continue;
      for (int grid : {1, n_sms, 2 * n_sms, 3 * n_sms}) {
        
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          gather_kernel<<<grid, 128, smem>>>(map, table, dids, tiles_per_cta, method, (int)V - 1, ddump, dcyc);
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          best = ms < best ? ms : best;
        }
        unsigned long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        const double bytes = (double)grid * tiles_per_cta * TILE_POS * 128;
        printf("  method %d grid %3d: %8.3f us  %8.1f GB/s  CTA0 %llu cycles = %.1f cyc/tile (%.1f cyc per 4 rows)\n", method, grid,
               best * 1e3, bytes / best * 1e-6, cyc, (double)cyc / tiles_per_cta, (double)cyc / tiles_per_cta / 16);
      }
      if (method == 0 && !big) {
        std::vector<uint8_t> dump(TILE_POS * 128);
        CK(cudaMemcpy(dump.data(), ddump, dump.size(), cudaMemcpyDeviceToHost));
        int bad = 0, zero_ok = 0;
        for (int p = 0; p < TILE_POS; ++p) {
          const int id = ids[(size_t)(tiles_per_cta - 1) * TILE_POS + p];
          for (int c = 0; c < 8; ++c) {
            const uint8_t* got = dump.data() + sw128_offset(p, c);
            if (id < 0 || id >= (int)V) {
              bool z = true; for (int b = 0; b < 16; ++b) z &= got[b] == 0;
              zero_ok += z;
            } else if (memcmp(got, (const uint8_t*)host.data() + (size_t)id * 128 + c * 16, 16) != 0) ++bad;
          }
        }
        printf("  gather4 layout check: %d bad chunks of %d; out-of-bounds rows zero-filled chunks: %d of 16\n", bad, TILE_POS * 8, zero_ok);
      }
    }
    CK(cudaFree(table)); CK(cudaFree(dids));
  }
  return 0;
}
