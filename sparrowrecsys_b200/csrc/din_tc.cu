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
//                    so that the 16 rows of a group are the MMA's N and the weights its M)
//   top layer 2      D2[(64 units hi | 64 units lo) x 16 rows] = W2stack[128 x 128] * H1^T
//
// Precision: every operand is split x = hi + lo into two bf16 halves and each product is
// evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation ("bf16x3", relative error
// ~2^-16 per product); numerics (raw releaseYear, rating counts) never enter an MMA - their
// rank-7 contribution is added in fp32 in the layer-1 epilogue.  Measured against the
// oracle in tests/test_gpu_parity.py.
//
// Work decomposition: a CTA is two independent warpgroups ("workers") sharing the weight
// images in shared memory; a worker owns groups of 16 consecutive rows.  Per group:
//   phase 0  gather candidate/user/genre rows -> X operand tile (bf16 hi/lo, SW128) + cst
//   phase 1  for each tile of 4 chunks x 32 positions: thread = (row, position) pair:
//            gather its history row (8 x 128-bit loads, prefetched one tile ahead), build A,
//            MMA (12 instr), read back 32 accumulators, PReLU / gate, butterfly-pool
//   phase 2  pooled -> X tile, layer-1 MMA (36), epilogue (bias, numerics, PReLU) -> H1
//            operand tile, layer-2 MMA (16), epilogue, sigmoid, store 16 scores.
#include <climits>

#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

constexpr int kTcRows = 16;              // rows per group (N of the top-MLP MMAs)
constexpr int kTcWG = 2;                 // warpgroups (workers) per CTA
constexpr int kTcMaxCPR = 4;             // chunks (of 32 positions) per row: T <= 128
constexpr int kNoPair = INT_MIN;         // sentinel: this thread has no (row, position) pair

// tensor-memory map of one worker: 2 buffers x 128 columns
constexpr uint32_t TM_A_HI = 0;          // 32 cols: bf16 pairs of [h | h*c], hi halves
constexpr uint32_t TM_A_LO = 32;         // 32 cols: lo halves
constexpr uint32_t TM_D = 64;            // 32 cols: activation-unit accumulators (fp32)
constexpr uint32_t TM_HS = 96;           // 32 cols: fp32 stash of h for pooling

// shared-memory image (bulk-copied from global; built by build_din_tc_image in model.cu)
constexpr uint32_t IMG_AUB_HI = 0;                       // [32 units][64 k] bf16, SW128
constexpr uint32_t IMG_AUB_LO = 4096;
constexpr uint32_t IMG_W1_HI = 8192;                     // 3 K blocks x [128 units][64 k]
constexpr uint32_t IMG_W1_LO = IMG_W1_HI + 3 * 16384;
constexpr uint32_t IMG_W2 = IMG_W1_LO + 3 * 16384;       // 2 K blocks x [64 hi | 64 lo units][64 k]
constexpr uint32_t IMG_ALPHAW = IMG_W2 + 2 * 16384;      // f32 [32 units][TP] alpha*wout, TP = CPR*32
// per-worker scratch
constexpr uint32_t WS_XB_HI = 0;                         // 3 K blocks x [16 rows][64 k] bf16
constexpr uint32_t WS_XB_LO = 6144;
constexpr uint32_t WS_H1_HI = 12288;                     // 2 K blocks x [16 rows][64 k]
constexpr uint32_t WS_H1_LO = 16384;
constexpr uint32_t WS_CAND = 20480;                      // f32 [16][32]
constexpr uint32_t WS_CST = 22528;                       // f32 [16][32]
constexpr uint32_t WS_PART = 24576;                      // f32 [16][4][32] pooled partials / reduce scratch
constexpr uint32_t WS_NUMS = 32768;                      // f32 [16][8]
constexpr uint32_t WS_BYTES = 33792;

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
  __shared__ uint64_t mbar[kTcWG][2];       // per-worker "MMAs complete", one per TMEM buffer
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
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
    for (int i = 0; i < kTcWG; ++i) { mbar_init(&mbar[i][0], 1); mbar_init(&mbar[i][1], 1); }
    fence_mbar_init();
    mbar_arrive_expect_tx(&wbar, img_bytes);
    for (uint32_t off = 0; off < img_bytes; off += 32768u) {
      const uint32_t n = min(32768u, img_bytes - off);
      bulk_g2s(img + off, p.image + off, n, &wbar);
    }
  }
  // zero the K padding of the X operand (K block 2, columns 32..63): never rewritten
  for (int i = tw; i < 16 * 4; i += 128) {
    const int rs = i >> 2, ch = 4 + (i & 3);
    const uint32_t off = 2 * 2048u + sw128_offset(rs, ch);
    *reinterpret_cast<uint4*>(ws + WS_XB_HI + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(ws + WS_XB_LO + off) = make_uint4(0, 0, 0, 0);
  }
  // per-thread constants: this thread is unit `tw` of layer 1 and unit `tw & 63` of layer 2
  const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
  float w1n[kNumNumerics];
#pragma unroll
  for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
  const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot + wg * 256;          // this worker's 256 TMEM columns
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t tD1 = tbase + TM_D, tD2 = tbase + 128 + TM_D;   // top MLP reuses the AU accumulators
  uint64_t* my_bar = mbar[wg];
  uint32_t phase = 0;                                   // bit i = parity to wait for on my_bar[i]
  bool weights_ready = false;

  const uint32_t idesc_au = idesc_bf16(128, 32), idesc_top = idesc_bf16(128, 16);
  const uint32_t s_img = smem_u32(img), s_ws = smem_u32(ws);

  const int n_groups = (b.B + kTcRows - 1) / kTcRows;
  const int n_workers = gridDim.x * kTcWG;
  const int CPR = p.CPR, T = p.T;
  const int n_tiles = 4 * CPR;                          // 16 rows * CPR chunks / 4 chunks per tile

  for (int g = blockIdx.x * kTcWG + wg; g < n_groups; g += n_workers) {
    const int row0 = g * kTcRows;
    // thread <-> (chunk = 4*tile + warp_w, position = lane): row slot rs = chunk / CPR
    auto pair_of = [&](int tile, int& rs, int& cq, int& t) {
      const int ch = 4 * tile + warp_w;
      rs = ch / CPR;
      cq = ch - rs * CPR;
      t = cq * 32 + lane;
    };
    // raw history id of this thread's pair in `tile` (kNoPair if none); the value is not
    // touched here so the load stays in flight until fix_id() one or two tiles later
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
    int raw0 = raw_id(0), raw1 = raw_id(1), raw2 = raw_id(2);   // ids come from HBM: start now

    // ================= phase 0: side gathers -> X operand, candidate rows, cst ==========
    {
      const int rs = tw >> 3, q = tw & 7;
      const int row = row0 + rs;
      const bool vr = row < b.B;
      float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), u4 = c4, ug4 = c4, mg4 = c4;
      float nv = 0.f;
      if (vr) {
        const int cid_raw = __ldg(b.movie_id + row), uid_raw = __ldg(b.user_id + row);
        int ug = __ldg(b.user_genre + row * 5), mg = __ldg(b.movie_genre + row * 3);
        if (q < kNumNumerics) nv = __ldg(b.numerics + row * kNumNumerics + q);
        const int cid = checked_id(f32_roundtrip_id(cid_raw), p.n_movies, b.err_flag);
        const int uid = checked_id(uid_raw, p.n_users, b.err_flag);
        c4 = ldg4(p.movie + (size_t)cid * 32 + 4 * q);
        u4 = ldg4(p.user + (size_t)uid * 32 + 4 * q);
        if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
        if (ug >= 0) ug4 = ldg4(p.ugenre + ug * 32 + 4 * q);
        if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
        if (mg >= 0) mg4 = ldg4(p.mgenre + mg * 32 + 4 * q);
      }
      *reinterpret_cast<float4*>(cand + rs * 32 + 4 * q) = c4;
      uint8_t* xh = ws + WS_XB_HI;
      uint8_t* xl = ws + WS_XB_LO;
      store_x4(xh, xl, 0, rs, 4 * q, ug4);            // K block 0: [userGenre1 | userId]
      store_x4(xh, xl, 0, rs, 32 + 4 * q, u4);
      store_x4(xh, xl, 1, rs, 32 + 4 * q, c4);        // K block 1: [pooled | candidate]
      store_x4(xh, xl, 2, rs, 4 * q, mg4);            // K block 2: [movieGenre1 | 0]
      nums[rs * 8 + q] = nv;
    }
    float4 hn[8];                                        // history row of the next tile to build
    bool valid_nxt;
    {
      const int id0 = fix_id(raw0);
      valid_nxt = id0 >= 0;
      load_row(id0, hn);
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
    if (!weights_ready) { mbar_wait(&wbar, 0); weights_ready = true; }

    // ================= phase 1: activation unit + pooling ================================
    // Two TMEM buffers: the MMAs of tile k+1 run while tile k is read back and pooled.
    // build(k): hn (history row of this thread's pair) -> A operand [h | h*c] hi/lo + fp32
    // stash of h in TMEM buffer k&1, then one thread issues the 12 MMAs.
    auto build_and_issue = [&](int tile) {
      int rs, cq, t;
      pair_of(tile, rs, cq, t);
      const uint32_t buf = tbase + (tile & 1) * 128;
      const float* c = cand + rs * 32;
      uint32_t ahi[16], alo[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ahi[4 * q] = __float_as_uint(hn[q].x); ahi[4 * q + 1] = __float_as_uint(hn[q].y);
        ahi[4 * q + 2] = __float_as_uint(hn[q].z); ahi[4 * q + 3] = __float_as_uint(hn[q].w);
        alo[4 * q] = __float_as_uint(hn[q + 4].x); alo[4 * q + 1] = __float_as_uint(hn[q + 4].y);
        alo[4 * q + 2] = __float_as_uint(hn[q + 4].z); alo[4 * q + 3] = __float_as_uint(hn[q + 4].w);
      }
      tmem_st16(buf + TM_HS + lane_base, ahi);           // fp32 stash of h for the pooling step
      tmem_st16(buf + TM_HS + 16 + lane_base, alo);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const Split2 s0 = split_pack(hn[q].x, hn[q].y), s1 = split_pack(hn[q].z, hn[q].w);
        ahi[2 * q] = s0.hi; ahi[2 * q + 1] = s1.hi;
        alo[2 * q] = s0.lo; alo[2 * q + 1] = s1.lo;
      }
      tmem_st16(buf + TM_A_HI + lane_base, ahi);
      tmem_st16(buf + TM_A_LO + lane_base, alo);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 c4 = *reinterpret_cast<const float4*>(c + 4 * q);
        const Split2 s0 = split_pack(hn[q].x * c4.x, hn[q].y * c4.y);
        const Split2 s1 = split_pack(hn[q].z * c4.z, hn[q].w * c4.w);
        ahi[2 * q] = s0.hi; ahi[2 * q + 1] = s1.hi;
        alo[2 * q] = s0.lo; alo[2 * q + 1] = s1.lo;
      }
      tmem_st16(buf + TM_A_HI + 16 + lane_base, ahi);
      tmem_st16(buf + TM_A_LO + 16 + lane_base, alo);
      tmem_st_wait();
      tc_fence_before();
      wg_sync(wg);
      if (tw == 0) {
        tc_fence_after();
        const uint64_t bh = smem_desc_sw128(s_img + IMG_AUB_HI), bl = smem_desc_sw128(s_img + IMG_AUB_LO);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(buf + TM_D, buf + TM_A_HI + 8 * ks, bh + 2 * ks, idesc_au, ks > 0);
          mma_ts(buf + TM_D, buf + TM_A_LO + 8 * ks, bh + 2 * ks, idesc_au, 1);
          mma_ts(buf + TM_D, buf + TM_A_HI + 8 * ks, bl + 2 * ks, idesc_au, 1);
        }
        mma_commit(&my_bar[tile & 1]);
      }
      __syncwarp();
    };

    bool valid_cur = valid_nxt;
    build_and_issue(0);
    {
      const int id1 = fix_id(raw1);
      valid_nxt = id1 >= 0;
      load_row(n_tiles > 1 ? id1 : -1, hn);
      raw1 = raw2;                                       // raw1 := id of tile k+2, raw2 := tile k+3
      raw2 = raw_id(3);
    }
    for (int tile = 0; tile < n_tiles; ++tile) {
      const bool valid_next_tile = valid_nxt;
      if (tile + 1 < n_tiles) {
        build_and_issue(tile + 1);
        const int idn = fix_id(raw1);
        valid_nxt = idn >= 0;
        load_row(tile + 2 < n_tiles ? idn : -1, hn);
        raw1 = raw2;
        raw2 = raw_id(tile + 4);
      }
      int rs, cq, t;
      pair_of(tile, rs, cq, t);
      const uint32_t buf = tbase + (tile & 1) * 128;
      mbar_wait(&my_bar[tile & 1], (phase >> (tile & 1)) & 1);
      phase ^= 1u << (tile & 1);
      __syncwarp();
      tc_fence_after();
      uint32_t d[32];
      tmem_ld32(buf + TM_D + lane_base, d);
      tmem_ld_wait();
      // ---- epilogue: + cst, PReLU (alpha per position), Dense(1), sigmoid gate
      float s = p.au_bout;
      {
        const float* cs = cst + rs * 32;
        const float* aw = alphaw + t;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(d[j]) + cs[j];
          s = fmaf(fmaxf(v, 0.f), p.au_wout[j], s);
          s = fmaf(fminf(v, 0.f), aw[j * TP], s);
        }
      }
      const float w = valid_cur ? 1.f / (1.f + __expf(-s)) : 0.f;
      valid_cur = valid_next_tile;
      // ---- pooling: out[e = lane] = sum over the chunk's 32 positions of w_t * h_t[e]
      tmem_ld32(buf + TM_HS + lane_base, d);
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
    }
    tc_fence_before();
    wg_sync(wg);

    // ================= phase 2: top MLP on the group's 16 rows ============================
    {
      const int rs = tw >> 3, q = tw & 7;
      float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < CPR; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(part + (rs * kTcMaxCPR + c) * 32 + 4 * q);
        pl.x += v.x; pl.y += v.y; pl.z += v.z; pl.w += v.w;
      }
      store_x4(ws + WS_XB_HI, ws + WS_XB_LO, 1, rs, 4 * q, pl);
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
          mma_ss(tD1, ah + 2 * ks, xh + 2 * ks, idesc_top, acc);
          acc = 1;
          mma_ss(tD1, al + 2 * ks, xh + 2 * ks, idesc_top, 1);
          mma_ss(tD1, ah + 2 * ks, xl + 2 * ks, idesc_top, 1);
        }
      }
      mma_commit(&my_bar[0]);
    }
    mbar_wait(&my_bar[0], phase & 1);
    phase ^= 1;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[16];
      tmem_ld16(tD1 + lane_base, d);
      tmem_ld_wait();
      // layer-1 epilogue for unit tw: bias + numerics (fp32) + PReLU -> H1 operand (bf16 hi/lo)
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
          mma_ss(tD2, a + 2 * ks, hh + 2 * ks, idesc_top, acc);
          acc = 1;
          mma_ss(tD2, a + 2 * ks, hl + 2 * ks, idesc_top, 1);
        }
      }
      mma_commit(&my_bar[1]);
    }
    mbar_wait(&my_bar[1], (phase >> 1) & 1);
    phase ^= 2;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[16];
      tmem_ld16(tD2 + lane_base, d);
      tmem_ld_wait();
      float* red = part;                                  // [64 units][16 rows]
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
        cst[pt * 16 + r] = s;
      }
      wg_sync(wg);
      if (tw < 16) {
        float z = p.b3;
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) z += cst[pt * 16 + tw];
        const int row = row0 + tw;
        if (row < b.B) {
          b.probs[row] = sigmoidf_acc(z);
          if (b.logits) b.logits[row] = z;
        }
      }
    }
  }
  if (!weights_ready) mbar_wait(&wbar, 0);               // never exit with the bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, 512);
}

size_t din_tc_smem_bytes(int cpr) {
  return 1024 + ((din_tc_image_bytes(cpr) + 1023u) & ~1023u) + (size_t)kTcWG * WS_BYTES;
}

cudaError_t launch_din_tc(const DinTcParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  const int n_groups = (b.B + kTcRows - 1) / kTcRows;
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
