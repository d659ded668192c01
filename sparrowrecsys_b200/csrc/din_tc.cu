// din_tc.cu - DIN forward on the 5th-generation tensor cores (tcgen05 + TMEM), one kernel.
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167.
// Same algebra as din.cu (activation unit folded to h.(Wsub+Wh) + (h*c).Wp + cst_b), but
// the three matrix products run as tcgen05.mma with fp32 accumulators in tensor memory:
//
//   activation unit  D[128 (row,position) pairs x 32 units] = A[128 x 64] * Bau[32 x 64]^T
//                    A = [h | h*c] built per pair in registers and written to TMEM
//                    (tcgen05.st) - the history rows go global -> registers -> TMEM and
//                    never touch shared memory; Bau = [Wsub+Wh ; Wp]^T resident in smem.
//   top layer 1      D1[128 units x 16 rows] = W1^T[128 x 192] * X[16 x 192]^T  (transposed
//                    so that the rows of a group are the MMA's N and the weights its M)
//   top layer 2      D2[(64 units hi | 64 units lo) x 16 rows] = W2stack[128 x 128] * H1^T
//
// Precision: every operand is split x = hi + lo into two bf16 halves and each product is
// evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation ("bf16x3", relative error
// ~2^-17 per product); numerics (raw releaseYear, rating counts) never enter an MMA - their
// rank-7 contribution is added in fp32 in the layer-1 epilogue.  Measured against the
// oracle in tests/test_gpu_parity.py.
//
// Work decomposition: one persistent CTA per SM holds the weight images in shared memory
// (bulk-copied once per launch) and runs four independent warpgroups ("workers"), each
// with 128 tensor-memory columns and 13 KB of scratch.  A worker owns groups of G (8 or 16)
// consecutive rows; the four workers interleave on the SM, which is what hides the
// latencies of the serial per-group chain:
//   phase 0  candidate/user/genre row gathers (kept in registers), cst_b
//   phase 1  per tile of 4 chunks x 32 positions, thread = (row, position) pair: its
//            history row arrives by 8 x 128-bit loads prefetched one tile ahead, A operand
//            -> TMEM, 12 MMAs, read back 32 accumulators, PReLU / sigmoid gate,
//            butterfly-pool w_t * h_t over the chunk
//   phase 2  X operand tile (bf16 hi/lo, SW128) <- pooled + side rows, layer-1 MMAs (36),
//            epilogue (bias, numerics, PReLU) -> H1 operand tile, layer-2 MMAs (16),
//            epilogue, sigmoid, scores out.
#include <climits>

#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

constexpr int kTcWG = 4;                 // warpgroups (workers) per CTA
constexpr int kTcMaxCPR = 4;             // chunks (of 32 positions) per row: T <= 128
constexpr int kNoPair = INT_MIN;         // sentinel: this thread has no (row, position) pair

// tensor-memory map of one worker (128 columns)
constexpr uint32_t TM_A_HI = 0;          // 32 cols: bf16 pairs of [h | h*c], hi halves
constexpr uint32_t TM_A_LO = 32;         // 32 cols: lo halves
constexpr uint32_t TM_D = 64;            // 32 cols: activation-unit accumulators; top layer 1 (16 cols)
constexpr uint32_t TM_HS = 96;           // 32 cols: fp32 stash of h for pooling; top layer 2 (16 cols)

// shared-memory image (bulk-copied from global; built by build_din_tc in model.cu)
constexpr uint32_t IMG_AUB_HI = 0;                       // [32 units][64 k] bf16, SW128
constexpr uint32_t IMG_AUB_LO = 4096;
constexpr uint32_t IMG_W1_HI = 8192;                     // 3 K blocks x [128 units][64 k]
constexpr uint32_t IMG_W1_LO = IMG_W1_HI + 3 * 16384;
constexpr uint32_t IMG_W2 = IMG_W1_LO + 3 * 16384;       // 2 K blocks x [64 hi | 64 lo units][64 k]
constexpr uint32_t IMG_ALPHAW = IMG_W2 + 2 * 16384;      // f32 [32 units][TP] alpha*wout, TP = CPR*32
// per-worker scratch: phases 0/1 and phase 2 overlay the same 12 KB
constexpr uint32_t WS_CAND = 0;                          // f32 [16][32]              (phase 0/1)
constexpr uint32_t WS_CST = 2048;                        // f32 [16][32]              (phase 0/1)
constexpr uint32_t WS_PART = 4096;                       // f32 [16][4][32] partials  (phase 1)
constexpr uint32_t WS_XB_HI = 0;                         // 3 K blocks x [16 rows][64 k] bf16 (phase 2)
constexpr uint32_t WS_XB_LO = 6144;
constexpr uint32_t WS_H1_HI = 0;                         // 2 K blocks x [16 rows][64 k]; after layer 1
constexpr uint32_t WS_H1_LO = 4096;
constexpr uint32_t WS_RED = 8192;                        // f32 [64][16]; after layer 1
constexpr uint32_t WS_NUMS = 12288;                      // f32 [16][8]
constexpr uint32_t WS_ZP = 12800;                        // f32 [8][16] final partial sums
constexpr uint32_t WS_BYTES = 13312;

// Phase timestamps (SM clock) of worker 0 of CTA 0, written when DinTcParams::trace != 0:
// [0] kernel entry, [1] prologue done, [2] phase 0 done, [3] weight image landed,
// [4 + k] tile k read back and pooled, [30] top layer 1 done, [31] group done, [32] exit.
__device__ unsigned long long g_din_tc_trace[40];
#define TC_TRACE(slot)                                                             \
  do {                                                                             \
    if (p.trace && blockIdx.x == 0 && tid == 0) g_din_tc_trace[slot] = clock64();  \
  } while (0)

__host__ __device__ inline uint32_t din_tc_image_bytes(int cpr) { return IMG_ALPHAW + 32u * cpr * 32u * 4u; }

__device__ __forceinline__ void wg_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

// write 4 consecutive K elements (col % 4 == 0) of row rs into a bf16 hi/lo SW128 operand
__device__ __forceinline__ void store_x4(uint8_t* hi, uint8_t* lo, int block, int rs, int col,
                                         float4 v) {
  const uint32_t off = block * 2048u + sw128_offset(rs, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(lo + off) = make_uint2(s0.lo, s1.lo);
}

__device__ __forceinline__ int f32_roundtrip_id(int id) {   // DIN.py:95,125: ids pass through float32
  return __float2int_rz(__int2float_rn(id));
}

__global__ void __launch_bounds__(kTcWG * 128, 1) din_tc_kernel(const __grid_constant__ DinTcParams p,
                                                                 BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t wbar;                 // weight image landed
  __shared__ uint64_t mbar[kTcWG];          // per-worker "MMAs complete"
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  TC_TRACE(0);
  const int wg = tid >> 7, tw = tid & 127, warp_w = tw >> 5, lane = tw & 31;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* img = base;
  const uint32_t img_bytes = din_tc_image_bytes(p.CPR);
  uint8_t* ws = base + ((img_bytes + 1023u) & ~1023u) + wg * WS_BYTES;
  const float* alphaw = reinterpret_cast<const float*>(img + IMG_ALPHAW);
  const int TP = p.CPR * 32;
  float* cand = reinterpret_cast<float*>(ws + WS_CAND);
  float* cst = reinterpret_cast<float*>(ws + WS_CST);
  float* part = reinterpret_cast<float*>(ws + WS_PART);
  float* nums = reinterpret_cast<float*>(ws + WS_NUMS);

  // ---- prologue ---------------------------------------------------------------------
  if (tid < 32) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&wbar, 1);
    for (int i = 0; i < kTcWG; ++i) mbar_init(&mbar[i], 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&wbar, img_bytes);
    // activation-unit operand + alpha table first (needed first), then the top-MLP images
    bulk_g2s(img, p.image, 8192, &wbar);
    bulk_g2s(img + IMG_ALPHAW, p.image + IMG_ALPHAW, img_bytes - IMG_ALPHAW, &wbar);
    for (uint32_t off = 8192; off < IMG_ALPHAW; off += 32768u)
      bulk_g2s(img + off, p.image + off, min(32768u, IMG_ALPHAW - off), &wbar);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TC_TRACE(1);
  const uint32_t tbase = tmem_slot + wg * 128;          // this worker's 128 TMEM columns
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  uint64_t* my_bar = &mbar[wg];
  uint32_t phase = 0;
  bool weights_ready = false;

  const uint32_t idesc_au = idesc_bf16(128, 32), idesc_top = idesc_bf16(128, 16);
  const uint32_t s_img = smem_u32(img), s_ws = smem_u32(ws);

  const int G = p.G;                                    // rows per group: 8 or 16
  const int n_groups = (b.B + G - 1) / G;
  const int n_workers = gridDim.x * kTcWG;
  const int CPR = p.CPR, T = p.T;
  const int n_tiles = (G * CPR) >> 2;                   // G rows * CPR chunks / 4 chunks per tile

  for (int g = blockIdx.x * kTcWG + wg; g < n_groups; g += n_workers) {
    const int row0 = g * G;
    // thread <-> (chunk = 4*tile + warp_w, position = lane): row slot rs = chunk / CPR
    auto pair_of = [&](int tile, int& rs, int& cq, int& t) {
      const int ch = 4 * tile + warp_w;
      rs = ch / CPR;
      cq = ch - rs * CPR;
      t = cq * 32 + lane;
    };
    // raw history id of this thread's pair in `tile` (kNoPair if none); the value is not
    // touched here so the load stays in flight until fix_id() a tile later
    auto raw_id = [&](int tile) -> int {
      if (tile >= n_tiles) return kNoPair;
      int rs, cq, t;
      pair_of(tile, rs, cq, t);
      const int row = row0 + rs;
      if (row >= b.B || t >= T) return kNoPair;
      return __ldg(b.hist + (size_t)row * b.hist_stride + t);
    };
    auto fix_id = [&](int raw) -> int {
      if (raw == kNoPair) return -1;
      return checked_id(f32_roundtrip_id(raw), p.n_movies, b.err_flag);
    };
    auto load_row = [&](int id, float4 (&h)[8]) {
      if (id >= 0) {
        const float* src = p.movie + (size_t)id * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = ldg4(src + 4 * q);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    int raw0 = raw_id(0), raw1 = raw_id(1);             // ids come from HBM: start them now

    // ================= phase 0: side gathers (registers), candidate rows, cst ============
    // thread (rs = tw / 8, q = tw % 8) owns floats [4q, 4q+4) of the side rows of row rs
    const int srs = tw >> 3, sq = tw & 7;
    float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), ug4 = u4, mg4 = u4;
    {
      const int row = row0 + srs;
      const bool vr = srs < G && row < b.B;
      float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float nv = 0.f;
      if (vr) {
        const int cid_raw = __ldg(b.movie_id + row), uid_raw = __ldg(b.user_id + row);
        int ug = __ldg(b.user_genre + row * 5), mg = __ldg(b.movie_genre + row * 3);
        if (sq < kNumNumerics) nv = __ldg(b.numerics + row * kNumNumerics + sq);
        const int cid = checked_id(f32_roundtrip_id(cid_raw), p.n_movies, b.err_flag);
        const int uid = checked_id(uid_raw, p.n_users, b.err_flag);
        c4 = ldg4(p.movie + (size_t)cid * 32 + 4 * sq);
        u4 = ldg4(p.user + (size_t)uid * 32 + 4 * sq);
        if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
        if (ug >= 0) ug4 = ldg4(p.ugenre + ug * 32 + 4 * sq);
        if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
        if (mg >= 0) mg4 = ldg4(p.mgenre + mg * 32 + 4 * sq);
      }
      *reinterpret_cast<float4*>(cand + srs * 32 + 4 * sq) = c4;
      nums[srs * 8 + sq] = nv;
    }
    float4 hn[8];                                        // history row of the next tile to build
    bool valid_nxt;
    {
      const int id0 = fix_id(raw0);
      valid_nxt = id0 >= 0;
      load_row(id0, hn);
      raw0 = raw_id(2);                                  // raw1 = tile 1, raw0 = tile 2
    }
    wg_sync(wg);
    {  // cst[rs][j] = au_b[j] + sum_e cand[rs][e] * (Wc - Wsub)[e][j]
      float acc[4];
      const float ab = __ldg(p.au_b + lane);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = ab;
#pragma unroll 8
      for (int e = 0; e < 32; ++e) {
        const float wc = __ldg(p.au_wc + e * 32 + lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(cand[(warp_w + 4 * i) * 32 + e], wc, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) cst[(warp_w + 4 * i) * 32 + lane] = acc[i];
    }
    wg_sync(wg);
    TC_TRACE(2);
    if (!weights_ready) { mbar_wait(&wbar, 0); weights_ready = true; }
    TC_TRACE(3);

    // ================= phase 1: activation unit + pooling ================================
    for (int tile = 0; tile < n_tiles; ++tile) {
      int rs, cq, t;
      pair_of(tile, rs, cq, t);
      const bool valid = valid_nxt;
      // ---- A operand [h | h*c] as bf16 hi / lo (two K elements per column) + fp32 stash of h
      {
        const float* c = cand + rs * 32;
        uint32_t ahi[16], alo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ahi[4 * q] = __float_as_uint(hn[q].x); ahi[4 * q + 1] = __float_as_uint(hn[q].y);
          ahi[4 * q + 2] = __float_as_uint(hn[q].z); ahi[4 * q + 3] = __float_as_uint(hn[q].w);
          alo[4 * q] = __float_as_uint(hn[q + 4].x); alo[4 * q + 1] = __float_as_uint(hn[q + 4].y);
          alo[4 * q + 2] = __float_as_uint(hn[q + 4].z); alo[4 * q + 3] = __float_as_uint(hn[q + 4].w);
        }
        tmem_st16(tbase + TM_HS + lane_base, ahi);
        tmem_st16(tbase + TM_HS + 16 + lane_base, alo);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const Split2 s0 = split_pack(hn[q].x, hn[q].y), s1 = split_pack(hn[q].z, hn[q].w);
          ahi[2 * q] = s0.hi; ahi[2 * q + 1] = s1.hi;
          alo[2 * q] = s0.lo; alo[2 * q + 1] = s1.lo;
        }
        tmem_st16(tbase + TM_A_HI + lane_base, ahi);
        tmem_st16(tbase + TM_A_LO + lane_base, alo);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 c4 = *reinterpret_cast<const float4*>(c + 4 * q);
          const Split2 s0 = split_pack(hn[q].x * c4.x, hn[q].y * c4.y);
          const Split2 s1 = split_pack(hn[q].z * c4.z, hn[q].w * c4.w);
          ahi[2 * q] = s0.hi; ahi[2 * q + 1] = s1.hi;
          alo[2 * q] = s0.lo; alo[2 * q + 1] = s1.lo;
        }
        tmem_st16(tbase + TM_A_HI + 16 + lane_base, ahi);
        tmem_st16(tbase + TM_A_LO + 16 + lane_base, alo);
      }
      tmem_st_wait();
      tc_fence_before();
      wg_sync(wg);
      if (tw == 0) {
        tc_fence_after();
        const uint64_t bh = smem_desc_sw128(s_img + IMG_AUB_HI), bl = smem_desc_sw128(s_img + IMG_AUB_LO);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(tbase + TM_D, tbase + TM_A_HI + 8 * ks, bh + 2 * ks, idesc_au, ks > 0);
          mma_ts(tbase + TM_D, tbase + TM_A_LO + 8 * ks, bh + 2 * ks, idesc_au, 1);
          mma_ts(tbase + TM_D, tbase + TM_A_HI + 8 * ks, bl + 2 * ks, idesc_au, 1);
        }
        mma_commit(my_bar);
      }
      __syncwarp();
      // ---- prefetch the next tile's history rows while the MMAs run
      {
        const int idn = fix_id(raw1);
        valid_nxt = idn >= 0;
        load_row(tile + 1 < n_tiles ? idn : -1, hn);
        raw1 = raw0;
        raw0 = raw_id(tile + 3);
      }
      mbar_wait(my_bar, phase);
      phase ^= 1;
      __syncwarp();
      tc_fence_after();
      uint32_t d[32];
      tmem_ld32(tbase + TM_D + lane_base, d);
      tmem_ld_wait();
      // ---- epilogue: + cst, PReLU (alpha per position), Dense(1), sigmoid gate
      float s0 = p.au_bout, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      {
        const float* cs = cst + rs * 32;
        const float* aw = alphaw + t;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cs + j);
          const float v0 = __uint_as_float(d[j]) + c4.x, v1 = __uint_as_float(d[j + 1]) + c4.y;
          const float v2 = __uint_as_float(d[j + 2]) + c4.z, v3 = __uint_as_float(d[j + 3]) + c4.w;
          s0 = fmaf(fmaxf(v0, 0.f), p.au_wout[j], s0);
          s1 = fmaf(fmaxf(v1, 0.f), p.au_wout[j + 1], s1);
          s2 = fmaf(fmaxf(v2, 0.f), p.au_wout[j + 2], s2);
          s3 = fmaf(fmaxf(v3, 0.f), p.au_wout[j + 3], s3);
          s0 = fmaf(fminf(v0, 0.f), aw[j * TP], s0);
          s1 = fmaf(fminf(v1, 0.f), aw[(j + 1) * TP], s1);
          s2 = fmaf(fminf(v2, 0.f), aw[(j + 2) * TP], s2);
          s3 = fmaf(fminf(v3, 0.f), aw[(j + 3) * TP], s3);
        }
      }
      const float s = (s0 + s1) + (s2 + s3);
      const float w = valid ? 1.f / (1.f + __expf(-s)) : 0.f;
      // ---- pooling: out[e = lane] = sum over the chunk's 32 positions of w_t * h_t[e]
      tmem_ld32(tbase + TM_HS + lane_base, d);
      tmem_ld_wait();
      float h[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) h[e] = __uint_as_float(d[e]) * w;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
          const float send = up ? h[i] : h[i + o];
          const float keep = up ? h[i + o] : h[i];
          h[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      part[(rs * kTcMaxCPR + cq) * 32 + lane] = h[0];
      if (tile < 24) TC_TRACE(4 + tile);
    }
    wg_sync(wg);

    // ================= phase 2: top MLP on the group's rows ================================
    {  // pooled and candidate leave the phase-1 scratch before the X operand overlays it
      float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < (srs < G ? CPR : 0); ++c) {
        const float4 v = *reinterpret_cast<const float4*>(part + (srs * kTcMaxCPR + c) * 32 + 4 * sq);
        pl.x += v.x; pl.y += v.y; pl.z += v.z; pl.w += v.w;
      }
      const float4 c4 = *reinterpret_cast<const float4*>(cand + srs * 32 + 4 * sq);
      wg_sync(wg);
      uint8_t* xh = ws + WS_XB_HI;
      uint8_t* xl = ws + WS_XB_LO;
      store_x4(xh, xl, 0, srs, 4 * sq, ug4);           // K block 0: [userGenre1 | userId]
      store_x4(xh, xl, 0, srs, 32 + 4 * sq, u4);
      store_x4(xh, xl, 1, srs, 4 * sq, pl);            // K block 1: [pooled | candidate]
      store_x4(xh, xl, 1, srs, 32 + 4 * sq, c4);
      store_x4(xh, xl, 2, srs, 4 * sq, mg4);           // K block 2: [movieGenre1 | 0]
      const uint32_t zoff = 2 * 2048u + sw128_offset(srs, 4 + (sq >> 1)) + ((sq & 1) ? 8u : 0u);
      *reinterpret_cast<uint2*>(xh + zoff) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(xl + zoff) = make_uint2(0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (tw == 0) {
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const uint64_t ah = smem_desc_sw128(s_img + IMG_W1_HI + kb * 16384);
        const uint64_t al = smem_desc_sw128(s_img + IMG_W1_LO + kb * 16384);
        const uint64_t xh = smem_desc_sw128(s_ws + WS_XB_HI + kb * 2048);
        const uint64_t xl = smem_desc_sw128(s_ws + WS_XB_LO + kb * 2048);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ss(tbase + TM_D, ah + 2 * ks, xh + 2 * ks, idesc_top, acc);
          acc = 1;
          mma_ss(tbase + TM_D, al + 2 * ks, xh + 2 * ks, idesc_top, 1);
          mma_ss(tbase + TM_D, ah + 2 * ks, xl + 2 * ks, idesc_top, 1);
        }
      }
      mma_commit(my_bar);
    }
    __syncwarp();
    // this thread is unit `tw` of layer 1: its constants arrive while the MMAs run
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    mbar_wait(my_bar, phase);
    phase ^= 1;
    __syncwarp();
    tc_fence_after();
    TC_TRACE(30);
    {
      uint32_t d[16];
      tmem_ld16(tbase + TM_D + lane_base, d);
      tmem_ld_wait();
      // layer-1 epilogue for unit tw: bias + numerics (fp32) + PReLU -> H1 operand (bf16 hi/lo),
      // which overlays the X operand (its MMAs have completed)
      const uint32_t koff = (uint32_t)(tw >> 6) * 2048u;
      const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float4 n0 = *reinterpret_cast<const float4*>(nums + r * 8);
        const float4 n1 = *reinterpret_cast<const float4*>(nums + r * 8 + 4);
        float v = __uint_as_float(d[r]) + b1;
        v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
        v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
        v = fmaf(n1.z, w1n[6], v);
        v = v > 0.f ? v : a1 * v;
        const uint32_t off = koff + sw128_offset(r, chunk) + within;
        const __nv_bfloat16 vh = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(ws + WS_H1_HI + off) = vh;
        *reinterpret_cast<__nv_bfloat16*>(ws + WS_H1_LO + off) = __float2bfloat16_rn(v - __bfloat162float(vh));
      }
    }
    fence_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (tw == 0) {
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t a = smem_desc_sw128(s_img + IMG_W2 + kb * 16384);
        const uint64_t hh = smem_desc_sw128(s_ws + WS_H1_HI + kb * 2048);
        const uint64_t hl = smem_desc_sw128(s_ws + WS_H1_LO + kb * 2048);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ss(tbase + TM_HS, a + 2 * ks, hh + 2 * ks, idesc_top, acc);
          acc = 1;
          mma_ss(tbase + TM_HS, a + 2 * ks, hl + 2 * ks, idesc_top, 1);
        }
      }
      mma_commit(my_bar);
    }
    __syncwarp();
    const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
    mbar_wait(my_bar, phase);
    phase ^= 1;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[16];
      tmem_ld16(tbase + TM_HS + lane_base, d);
      tmem_ld_wait();
      float* red = reinterpret_cast<float*>(ws + WS_RED);     // [64 units][16 rows]
      float* zp = reinterpret_cast<float*>(ws + WS_ZP);       // [8][16]
      if (tw >= 64) {
#pragma unroll
        for (int r = 0; r < 16; r += 4)
          *reinterpret_cast<float4*>(red + (tw - 64) * 16 + r) =
              make_float4(__uint_as_float(d[r]), __uint_as_float(d[r + 1]), __uint_as_float(d[r + 2]),
                          __uint_as_float(d[r + 3]));
      }
      wg_sync(wg);
      if (tw < 64) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float v = __uint_as_float(d[r]) + red[tw * 16 + r] + b2;   // (W2hi + W2lo) . (H1hi + H1lo)
          v = v > 0.f ? v : a2 * v;
          red[tw * 16 + r] = v * w3;
        }
      }
      wg_sync(wg);
      {  // 16 rows x 8 partial sums of 8 units
        const int r = tw & 15, pt = tw >> 4;
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += red[(pt * 8 + u) * 16 + r];
        zp[pt * 16 + r] = s;
      }
      wg_sync(wg);
      if (tw < G) {
        float z = p.b3;
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) z += zp[pt * 16 + tw];
        const int row = row0 + tw;
        if (row < b.B) {
          b.probs[row] = sigmoidf_acc(z);
          if (b.logits) b.logits[row] = z;
        }
      }
    }
    tc_fence_before();
    wg_sync(wg);                                         // scratch is reused by the next group
    TC_TRACE(31);
  }
  if (!weights_ready) mbar_wait(&wbar, 0);               // never exit with the bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, 512);
  TC_TRACE(32);
}

cudaError_t read_din_tc_trace(unsigned long long* out40) {
  return cudaMemcpyFromSymbol(out40, g_din_tc_trace, sizeof(unsigned long long) * 40);
}

size_t din_tc_smem_bytes(int cpr) {
  return 1024 + ((din_tc_image_bytes(cpr) + 1023u) & ~1023u) + (size_t)kTcWG * WS_BYTES;
}

cudaError_t launch_din_tc(const DinTcParams& p0, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  DinTcParams p = p0;
  // rows per group: 16 when there is enough work to keep every worker busy, else 8
  const int workers = p.num_sms * kTcWG;
  p.G = (b.B >= 16 * workers) ? 16 : 8;
  const int n_groups = (b.B + p.G - 1) / p.G;
  int grid = (n_groups + kTcWG - 1) / kTcWG;
  if (grid > p.num_sms) grid = p.num_sms;
  din_tc_kernel<<<grid, kTcWG * 128, din_tc_smem_bytes(p.CPR), s>>>(p, b);
  ++g_launch_count;
  return cudaGetLastError();
}

cudaError_t setup_din_tc_attributes() {
  return cudaFuncSetAttribute(din_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)din_tc_smem_bytes(kTcMaxCPR));
}

}  // namespace srs
