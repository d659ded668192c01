// embmlp_tc.cu - EmbeddingMLP / Wide&Deep forward on the tensor cores (tcgen05 + TMEM) for the
// reference shape (E <= 12: ten 12-float embedding slots = 120 of 128 K columns).
//
// Reference: EmbeddingMLP.py:72-77 and WideNDeep.py:101-107
// (TFRecModel/src/com/sparrowrecsys/offline/tensorflow/).  Same structure as phase 2 of
// din_tc.cu: both Dense(128) layers are computed transposed - D[128 units x rows] =
// W^T[128 x 128] . X^T - so the 64 rows of a super-group are the MMA's N and the weights,
// resident in shared memory as bf16 hi/lo images, its M.  Operands are split x = hi + lo
// (bf16x3, see din_tc.cu); the activations' hi and lo halves are stacked along N
// ([64 rows hi | 64 rows lo]), so a layer is 8 K steps x 2 MMAs of N = 128.  The 7 raw-scale
// numerics never enter an MMA: their contribution is added in fp32 in the layer-1 epilogue.
//
// One persistent CTA per SM, 256 threads; per super-group of 64 rows: 30 row gathers per row
// straight into the X operand tile, 16 MMAs, epilogue (thread = unit, 32 rows each) -> H1
// operand tile over the X tile, 16 MMAs, epilogue, Dense(1) (+ the W&D wide weight), sigmoid.
#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

constexpr int kEtRows = 64;                              // rows per super-group = half of the MMA N
// shared-memory image: four 32 KB operands, each 2 K blocks x [128 units][64 k] bf16 SW128
constexpr uint32_t EIMG_W1_HI = 0, EIMG_W1_LO = 32768, EIMG_W2_HI = 65536, EIMG_W2_LO = 98304;
constexpr uint32_t EIMG_BYTES = 131072;
// scratch
constexpr uint32_t ES_X = 0;                             // 2 K blocks x [64 hi | 64 lo rows][64 k] = 32 KB
constexpr uint32_t ES_RED = 32768;                       // f32 [128 units][64 rows] = 32 KB
constexpr uint32_t ES_NUMS = 65536;                      // f32 [64][8]
constexpr uint32_t ES_ZP = 67584;                        // f32 [4][64]
constexpr uint32_t ES_BYTES = 68608;

__global__ void __launch_bounds__(256, 1) embmlp_tc_kernel(const __grid_constant__ EmbMlpTcParams p,
                                                            BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t wbar, mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, wg = tid >> 7, tw = tid & 127, warp_w = tw >> 5;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* img = base;
  uint8_t* sc = base + EIMG_BYTES;
  float* nums = reinterpret_cast<float*>(sc + ES_NUMS);
  float* red = reinterpret_cast<float*>(sc + ES_RED);
  float* zp = reinterpret_cast<float*>(sc + ES_ZP);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid < 32) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&wbar, EIMG_BYTES);
    for (uint32_t off = 0; off < EIMG_BYTES; off += 32768u) bulk_g2s(img + off, p.image + off, 32768u, &wbar);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tD = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t idesc = idesc_bf16(128, 2 * kEtRows);
  const uint32_t s_img = smem_u32(img), s_x = smem_u32(sc + ES_X);
  uint32_t phase = 0;
  bool weights_ready = false;
  // this thread is unit `tw` of both layers
  const float b1 = __ldg(p.b1 + tw), b2 = __ldg(p.b2 + tw), w3 = __ldg(p.w3 + tw);
  float w1n[kNumNumerics];
#pragma unroll
  for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);

  const int n_sg = (b.B + kEtRows - 1) / kEtRows;
  for (int sg = blockIdx.x; sg < n_sg; sg += gridDim.x) {
    const int row0 = sg * kEtRows;
    // ---- gathers -> X operand tile (K = slot * 12 + e; columns 120..127 are zero) ---------
    for (int i = tid; i < kEtRows * 32; i += 256) {      // 64 rows x 32 float4 (30 gathered + 2 zero)
      const int r = i >> 5, f = i & 31;
      const int row = row0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < 30 && row < b.B) {
        const int slot = f / 3, q = f - slot * 3;
        int id;
        const float* table;
        if (slot < 3) {
          id = __ldg(b.movie_genre + row * 3 + slot);
          table = p.genre[slot];
        } else if (slot == 3) {
          id = checked_id(__ldg(b.movie_id + row), p.n_movies, b.err_flag);
          table = p.movie;
        } else if (slot < 9) {
          id = __ldg(b.user_genre + row * 5 + (slot - 4));
          table = p.genre[slot - 1];
        } else {
          id = checked_id(__ldg(b.user_id + row), p.n_users, b.err_flag);
          table = p.user;
        }
        if (slot != 3 && slot != 9) {
          if (id >= p.n_genres) { atomicExch(b.err_flag, 1); id = -1; }
        }
        if (id >= 0) v = ldg4(table + (size_t)id * 12 + 4 * q);
      }
      const int k = 4 * f;                               // K index of v.x (slot*12 + 4q == 4f)
      const uint32_t off = (uint32_t)(k >> 6) * 16384u + sw128_offset(r, (k & 63) >> 3) + ((k & 4) ? 8u : 0u);
      const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
      *reinterpret_cast<uint2*>(sc + ES_X + off) = make_uint2(s0.hi, s1.hi);
      *reinterpret_cast<uint2*>(sc + ES_X + off + 8192u) = make_uint2(s0.lo, s1.lo);   // row + 64
    }
    for (int i = tid; i < kEtRows * 8; i += 256) {
      const int r = i >> 3, j = i & 7;
      const int row = row0 + r;
      nums[i] = (j < kNumNumerics && row < b.B) ? __ldg(b.numerics + row * kNumNumerics + j) : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (!weights_ready) { mbar_wait(&wbar, 0); weights_ready = true; }

    // ---- two Dense(128, relu) layers ----------------------------------------------------
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a_hi = s_img + (layer == 0 ? EIMG_W1_HI : EIMG_W2_HI);
        const uint32_t a_lo = s_img + (layer == 0 ? EIMG_W1_LO : EIMG_W2_LO);
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ah = smem_desc_sw128(a_hi + kb * 16384), al = smem_desc_sw128(a_lo + kb * 16384);
          const uint64_t xs = smem_desc_sw128(s_x + kb * 16384);            // [X hi | X lo], N = 128
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ss(tD, ah + 2 * ks, xs + 2 * ks, idesc, acc);
            acc = 1;
            mma_ss(tD, al + 2 * ks, xs + 2 * ks, idesc, 1);
          }
        }
        mma_commit(&mbar);
      }
      __syncwarp();
      mbar_wait(&mbar, phase);
      phase ^= 1;
      __syncwarp();
      tc_fence_after();
      // epilogue: unit tw, rows 32*wg .. 32*wg+31, in two halves of 16 rows
      const uint32_t koff = (uint32_t)(tw >> 6) * 16384u;
      const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t dh[16], dl[16];
        const int rbase = 32 * wg + 16 * half;
        tmem_ld16(tD + rbase + lane_base, dh);               // W . X hi
        tmem_ld16(tD + kEtRows + rbase + lane_base, dl);     // W . X lo
        tmem_ld_wait();
        if (layer == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = rbase + i;
            const float4 n0 = *reinterpret_cast<const float4*>(nums + r * 8);
            const float4 n1 = *reinterpret_cast<const float4*>(nums + r * 8 + 4);
            float v = (__uint_as_float(dh[i]) + __uint_as_float(dl[i])) + b1;
            v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
            v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
            v = fmaf(n1.z, w1n[6], v);
            v = fmaxf(v, 0.f);
            dh[i] = __float_as_uint(v);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            dh[i] = __float_as_uint(fmaxf((__uint_as_float(dh[i]) + __uint_as_float(dl[i])) + b2, 0.f) * w3);
        }
        if (layer == 0) {
          // the X tile's MMAs have completed (every thread waited on mbar), but other threads may
          // still be reading D; H1 only overwrites shared memory, which is safe
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = rbase + i;
            const float v = __uint_as_float(dh[i]);
            const uint32_t off = koff + sw128_offset(r, chunk) + within;
            const __nv_bfloat16 vh = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16*>(sc + ES_X + off) = vh;
            *reinterpret_cast<__nv_bfloat16*>(sc + ES_X + off + 8192u) =
                __float2bfloat16_rn(v - __bfloat162float(vh));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(red + tw * kEtRows + rbase + i) =
                make_float4(__uint_as_float(dh[i]), __uint_as_float(dh[i + 1]), __uint_as_float(dh[i + 2]),
                            __uint_as_float(dh[i + 3]));
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    // ---- Dense(1): sum over the 128 units, + wide weight (W&D), sigmoid ------------------------
    {
      const int r = tid & 63, part = tid >> 6;             // 4 parts x 32 units
      float s = 0.f;
#pragma unroll 8
      for (int u = 0; u < 32; ++u) s += red[(part * 32 + u) * kEtRows + r];
      zp[part * kEtRows + r] = s;
    }
    __syncthreads();
    if (tid < kEtRows) {
      const int row = row0 + tid;
      if (row < b.B) {
        float z = p.b3 + ((zp[tid] + zp[kEtRows + tid]) + (zp[2 * kEtRows + tid] + zp[3 * kEtRows + tid]));
        if (p.wide) {
          const int mid = checked_id(__ldg(b.movie_id + row), p.n_movies, b.err_flag);
          const int rated = checked_id(__ldg(b.hist + (size_t)row * b.hist_stride), p.n_movies, b.err_flag);
          z += __ldg(p.wide + crossed_bucket(mid, rated, (uint32_t)p.cross_buckets));
        }
        store_score(b, row, sigmoidf_acc(z));
        if (b.logits) b.logits[row] = z;
      }
    }
    __syncthreads();
  }
  if (!weights_ready) mbar_wait(&wbar, 0);
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, 128);
}

static size_t embmlp_tc_smem() { return 1024 + EIMG_BYTES + ES_BYTES; }

cudaError_t launch_embmlp_tc(const EmbMlpTcParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  const int n_sg = (b.B + kEtRows - 1) / kEtRows;
  const int grid = n_sg < p.num_sms ? n_sg : p.num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = embmlp_tc_smem();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, embmlp_tc_kernel, p, b);
}

cudaError_t setup_embmlp_tc_attributes() {
  return cudaFuncSetAttribute(embmlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)embmlp_tc_smem());
}

}  // namespace srs
