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
// (bulk-copied once per launch) and owns super-groups of 32 consecutive rows.  Its four
// warpgroups ("workers", 128 tensor-memory columns each) take 8 rows apiece through
//   phase 0  candidate/user/genre row gathers (kept in registers), cst_b
//   phase 1  per tile of 4 chunks x 32 positions, thread = (row, position) pair: its
//            history row arrives by 8 x 128-bit loads prefetched one tile ahead, A operand
//            -> TMEM, 12 MMAs, read back 32 accumulators, PReLU / sigmoid gate,
//            butterfly-pool w_t * h_t over the chunk
// and then run the top MLP together, once per super-group, with the 32 rows as the MMA N:
//   phase 2  X operand tile (bf16 hi/lo, SW128) <- pooled + side rows, layer-1 MMAs (36),
//            epilogue (bias, numerics, PReLU) -> H1 operand tile, layer-2 MMAs (16),
//            epilogue, sigmoid, scores out.
// The ids of the next super-group are requested from HBM before phase 2 of the current one.
#include <climits>

#include "kernels.h"
#include "umma.cuh"

namespace srs {
using namespace umma;

constexpr int kTcWG = 4;                 // warpgroups (workers) per CTA
constexpr int kTcG = 8;                  // rows per worker per super-group
constexpr int kTcSG = kTcWG * kTcG;      // rows per super-group = N of the top-MLP MMAs
constexpr int kTcMaxCPR = 4;             // chunks (of 32 positions) per row: T <= 128
constexpr int kNoPair = INT_MIN;         // sentinel: this thread has no (row, position) pair

// tensor-memory map of one worker (128 columns)
constexpr uint32_t TM_A_HI = 0;          // 32 cols: bf16 pairs of [h | h*c], hi halves
constexpr uint32_t TM_A_LO = 32;         // 32 cols: lo halves
constexpr uint32_t TM_D = 64;            // 64 cols: accumulators, N-stacked: [A.Bhi (32) | A.Blo (32)];
                                         //          worker 0: top layer 1, worker 1: top layer 2

// shared-memory image (bulk-copied from global; built by build_din_tc in model.cu)
constexpr uint32_t IMG_AUB_HI = 0;                       // [32 units][64 k] bf16, SW128
// (lo halves follow at +4096: rows 32..63 of the N-stacked operand)
constexpr uint32_t IMG_W1_HI = 8192;                     // 3 K blocks x [128 units][64 k]
constexpr uint32_t IMG_W1_LO = IMG_W1_HI + 3 * 16384;
constexpr uint32_t IMG_W2 = IMG_W1_LO + 3 * 16384;       // 2 K blocks x [64 hi | 64 lo units][64 k]
constexpr uint32_t IMG_PQ = IMG_W2 + 2 * 16384;          // f32 [TP positions][68]: per position 16 x (P_2m, P_2m+1,
                                                         // Q_2m, Q_2m+1) + 4 pad floats (conflict-free LDS.128)
constexpr uint32_t kPqStride = 68;                       // floats per position row
// CTA scratch (28 KB).  Phase 0/1 view: worker w owns 6 KB at w*6144; phase 2 overlays all 24 KB.
constexpr uint32_t WS_W_STRIDE = 6144;
constexpr uint32_t WS_CAND = 0;                          // f32 [8][32]        (worker, phase 0/1)
constexpr uint32_t WS_CST = 1024;                        // f32 [8][32]        (worker, phase 0/1)
constexpr uint32_t WS_PART = 2048;                       // f32 [8][4][32]     (worker, phase 1)
// operand tiles of phase 2 are N-stacked per K block: [32 rows hi | 32 rows lo] x 128 B = 8 KB,
// so one MMA with N = 64 multiplies a weight tile by both halves of the activations
constexpr uint32_t WS_XB = 0;                            // 3 K blocks x 8 KB (phase 2)
constexpr uint32_t WS_H1 = 0;                            // 2 K blocks x 8 KB; after layer 1
constexpr uint32_t WS_RED = 16384;                       // f32 [64][32]; after layer 1
constexpr uint32_t WS_NUMS = 24576;                      // f32 [32][8]
constexpr uint32_t WS_ZP = 25600;                        // f32 [16][32] final partial sums
constexpr uint32_t WS_BYTES = 28672;

// Phase timestamps (SM clock) of thread 0 of CTA 0, written when DinTcParams::trace != 0:
// [0] kernel entry, [1] prologue done, [2] phase 0 done, [3] weight image landed,
// [4 + k] tile k read back and pooled, [30] top layer 1 done, [31] super-group done, [32] exit.
__device__ unsigned long long g_din_tc_trace[40];
#define TC_TRACE(slot)                                                             \
  do {                                                                             \
    if (p.trace && blockIdx.x == 0 && tid == 0) g_din_tc_trace[slot] = clock64();  \
  } while (0)

__host__ __device__ inline uint32_t din_tc_image_bytes(int cpr) { return IMG_PQ + cpr * 32u * kPqStride * 4u; }

__device__ __forceinline__ void wg_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

// write 4 consecutive K elements (col % 4 == 0) of row `row` (0..31) into an N-stacked bf16
// operand: K block `block` is 8 KB, hi halves in rows 0..31, lo halves in rows 32..63
__device__ __forceinline__ void store_x4(uint8_t* tile, int block, int row, int col, float4 v) {
  const uint32_t off = block * 8192u + sw128_offset(row, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(tile + off + 4096u) = make_uint2(s0.lo, s1.lo);   // row + 32: same swizzle phase
}

// 256-bit read-only global load (sm_100 LDG.256): a 128-byte embedding row in 4 instructions,
// half the L1 tag passes of 8 x LDG.128 for a gather in which every lane reads a different row
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

__device__ __forceinline__ int f32_roundtrip_id(int id) {   // DIN.py:95,125: ids pass through float32
  return __float2int_rz(__int2float_rn(id));
}

__device__ __forceinline__ int div_cpr(int ch, int cpr) {   // ch < 32, cpr in 1..4
  return cpr == 1 ? ch : cpr == 2 ? ch >> 1 : cpr == 4 ? ch >> 2 : (ch * 43) >> 7;
}

// values a thread requests from HBM for a super-group before it needs them
struct GroupLoads {
  int cid_raw, uid_raw, ug, mg;     // side ids of row (wg*8 + tw/8); kNoPair when no such row
  float nv;                         // numeric tw%8 of that row
  int raw0, raw1;                   // history ids of this thread's pair in tiles 0 and 1
};

template <int CPR>
__global__ void __launch_bounds__(kTcWG * 128, 1) din_tc_kernel(const __grid_constant__ DinTcParams p,
                                                                 BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t wbar;                 // weight image landed
  __shared__ uint64_t mbar[kTcWG];          // per-worker "MMAs complete"
  __shared__ uint64_t cbar;                 // CTA-wide "top-MLP MMAs complete"
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  TC_TRACE(0);
  const int wg = tid >> 7, tw = tid & 127, warp_w = tw >> 5, lane = tw & 31;
  const int srs = tw >> 3, sq = tw & 7;                 // side-feature role: row slot / float4 index
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* img = base;
  const uint32_t img_bytes = din_tc_image_bytes(CPR);
  uint8_t* cs_base = base + ((img_bytes + 1023u) & ~1023u);     // CTA scratch
  uint8_t* ws = cs_base + wg * WS_W_STRIDE;                      // this worker's phase-0/1 view
  const int T = p.T;
  const float4* PQtab = reinterpret_cast<const float4*>(img + IMG_PQ);
  float* cand = reinterpret_cast<float*>(ws + WS_CAND);
  float* cst = reinterpret_cast<float*>(ws + WS_CST);
  float* part = reinterpret_cast<float*>(ws + WS_PART);
  float* nums = reinterpret_cast<float*>(cs_base + WS_NUMS);
  constexpr int n_tiles = 2 * CPR;                      // 8 rows * CPR chunks / 4 chunks per tile
  const int n_sg = (b.B + kTcSG - 1) / kTcSG;

  // thread <-> (chunk = 4*tile + warp_w, position = lane): row slot rs = chunk / CPR
  auto pair_of = [&](int tile, int& rs, int& cq, int& t) {
    const int ch = 4 * tile + warp_w;
    rs = div_cpr(ch, CPR);
    cq = ch - rs * CPR;
    t = cq * 32 + lane;
  };
  // raw history id of this thread's pair (kNoPair if none); the value is not touched here so
  // the load stays in flight until fix_id() a tile later
  auto raw_id = [&](int row0, int tile) -> int {
    if (tile >= n_tiles) return kNoPair;
    int rs, cq, t;
    pair_of(tile, rs, cq, t);
    const int row = row0 + rs;
    if (row >= b.B || t >= T) return kNoPair;
    return __ldg(b.hist + (size_t)row * b.hist_stride + t);
  };
  auto issue_loads = [&](int sg) -> GroupLoads {
    GroupLoads L;
    L.cid_raw = L.uid_raw = kNoPair; L.ug = L.mg = -1; L.nv = 0.f; L.raw0 = L.raw1 = kNoPair;
    if (sg >= n_sg) return L;
    const int row0 = sg * kTcSG + wg * kTcG;
    const int row = row0 + srs;
    if (srs < kTcG && row < b.B) {
      L.cid_raw = __ldg(b.movie_id + row);
      L.uid_raw = __ldg(b.user_id + row);
      L.ug = __ldg(b.user_genre + row * 5);
      L.mg = __ldg(b.movie_genre + row * 3);
      if (sq < kNumNumerics) L.nv = __ldg(b.numerics + row * kNumNumerics + sq);
    }
    L.raw0 = raw_id(row0, 0);
    L.raw1 = raw_id(row0, 1);
    return L;
  };
  // ---- prologue ---------------------------------------------------------------------
  // Programmatic dependent launch: the next launch in the stream may start its own prologue
  // (TMEM allocation, barrier init, weight image copy - nothing that depends on this launch)
  // as soon as an SM frees up; it reads no input before its griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid < 32) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&cbar, 1);
    for (int i = 0; i < kTcWG; ++i) mbar_init(&mbar[i], 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&wbar, img_bytes);
    // activation-unit operand + P/Q tables first (needed first), then the top-MLP images
    bulk_g2s(img, p.image, 8192, &wbar);
    bulk_g2s(img + IMG_PQ, p.image + IMG_PQ, img_bytes - IMG_PQ, &wbar);
    for (uint32_t off = 8192; off < IMG_PQ; off += 32768u)
      bulk_g2s(img + off, p.image + off, min(32768u, IMG_PQ - off), &wbar);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");    // inputs may come from the previous kernel
  GroupLoads pre = issue_loads(blockIdx.x);             // ids come from HBM: in flight during the sync
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TC_TRACE(1);
  const uint32_t tbase = tmem_slot + wg * 128;          // this worker's 128 TMEM columns
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  uint64_t* my_bar = &mbar[wg];
  uint32_t phase = 0, cphase = 0;
  bool weights_ready = false;

  const uint32_t idesc_au = idesc_bf16(128, 64), idesc_top = idesc_bf16(128, 2 * kTcSG);
  const uint32_t s_img = smem_u32(img), s_cs = smem_u32(cs_base);

  auto fix_id = [&](int raw_v) -> int {
    if (raw_v == kNoPair) return -1;
    return checked_id(f32_roundtrip_id(raw_v), p.n_movies, b.err_flag);
  };
  auto load_row = [&](int id, float4 (&h)[8]) {
    if (id >= 0) {
      const float* src = p.movie + (size_t)id * 32;
#pragma unroll
      for (int q = 0; q < 8; q += 2) ldg8(src + 4 * q, h[q], h[q + 1]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) h[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  for (int sg = blockIdx.x; sg < n_sg; sg += gridDim.x) {
    const int row0 = sg * kTcSG + wg * kTcG;            // this worker's first row
    int raw1 = pre.raw1;
    int raw0 = raw_id(row0, 2);                          // raw1 = tile 1, raw0 = tile 2

    // ================= phase 0: side gathers (registers), candidate rows, cst ============
    float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), ug4 = u4, mg4 = u4;
    if (srs < kTcG) {
      float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pre.cid_raw != kNoPair) {
        const int cid = checked_id(f32_roundtrip_id(pre.cid_raw), p.n_movies, b.err_flag);
        const int uid = checked_id(pre.uid_raw, p.n_users, b.err_flag);
        int ug = pre.ug, mg = pre.mg;
        c4 = ldg4(p.movie + (size_t)cid * 32 + 4 * sq);
        u4 = ldg4(p.user + (size_t)uid * 32 + 4 * sq);
        if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
        if (ug >= 0) ug4 = ldg4(p.ugenre + ug * 32 + 4 * sq);
        if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
        if (mg >= 0) mg4 = ldg4(p.mgenre + mg * 32 + 4 * sq);
      }
      *reinterpret_cast<float4*>(cand + srs * 32 + 4 * sq) = c4;
      nums[(wg * kTcG + srs) * 8 + sq] = pre.nv;
    }
    float4 hn[8];                                        // history row of the next tile to build
    bool valid_nxt;
    {
      const int id0 = fix_id(pre.raw0);
      valid_nxt = id0 >= 0;
      load_row(id0, hn);
    }
    wg_sync(wg);
    // cst[rs][j] = au_b[j] + sum_e cand[rs][e] * (Wc - Wsub)[e][j]; warp w: rows w and w+4.
    // First needed by the first epilogue: computed while the first tile's MMAs run.
    auto compute_cst = [&]() {
      float acc0 = __ldg(p.au_b + lane), acc1 = acc0;
#pragma unroll 8
      for (int e = 0; e < 32; ++e) {
        const float wc = __ldg(p.au_wc + e * 32 + lane);
        acc0 = fmaf(cand[warp_w * 32 + e], wc, acc0);
        acc1 = fmaf(cand[(warp_w + 4) * 32 + e], wc, acc1);
      }
      cst[warp_w * 32 + lane] = acc0;
      cst[(warp_w + 4) * 32 + lane] = acc1;
    };
    const bool cst_early = false;
    TC_TRACE(2);
    if (!weights_ready) { mbar_wait(&wbar, 0); weights_ready = true; }
    TC_TRACE(3);

    // ================= phase 1: activation unit + pooling ================================
    for (int tile = 0; tile < n_tiles; ++tile) {
      int rs, cq, t;
      pair_of(tile, rs, cq, t);
      const bool valid = valid_nxt;
      // ---- A operand [h | h*c] as bf16 hi / lo (two K elements per column)
      {
        const float* c = cand + rs * 32;
        uint32_t ahi[16], alo[16];
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
          const float2 g01 = mul2(make_float2(hn[q].x, hn[q].y), make_float2(c4.x, c4.y));
          const float2 g23 = mul2(make_float2(hn[q].z, hn[q].w), make_float2(c4.z, c4.w));
          const Split2 s0 = split_pack(g01.x, g01.y);
          const Split2 s1 = split_pack(g23.x, g23.y);
          ahi[2 * q] = s0.hi; ahi[2 * q + 1] = s1.hi;
          alo[2 * q] = s0.lo; alo[2 * q + 1] = s1.lo;
        }
        tmem_st16(tbase + TM_A_HI + 16 + lane_base, ahi);
        tmem_st16(tbase + TM_A_LO + 16 + lane_base, alo);
      }
      if (tile == 1) TC_TRACE(20);
      tmem_st_wait();
      tc_fence_before();
      if (tile == 1) TC_TRACE(21);
      wg_sync(wg);
      if (tile == 1) TC_TRACE(22);
      if (tw == 0) {
        // B operand N-stacked: rows 0..31 = Bhi, rows 32..63 = Blo  ->  D[:, :32] = A.Bhi, D[:, 32:] = A.Blo;
        // two MMAs per K step (A hi, A lo) instead of three
        tc_fence_after();
        const uint64_t bst = smem_desc_sw128(s_img + IMG_AUB_HI);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(tbase + TM_D, tbase + TM_A_HI + 8 * ks, bst + 2 * ks, idesc_au, ks > 0);
          mma_ts(tbase + TM_D, tbase + TM_A_LO + 8 * ks, bst + 2 * ks, idesc_au, 1);
        }
        mma_commit(my_bar);
      }
      __syncwarp();
      if (tile == 1) TC_TRACE(23);
      // ---- the next tile's history rows are requested now: they land while the MMAs run and
      //      the accumulators are read back
      float4 hnext[8];
      {
        const int idn = fix_id(raw1);
        valid_nxt = idn >= 0;
        load_row(tile + 1 < n_tiles ? idn : -1, hnext);
        raw1 = raw0;
        raw0 = raw_id(row0, tile + 3);
      }
      if (tile == 0 && !cst_early) compute_cst();
      if (tile == 1) TC_TRACE(24);
      mbar_wait(my_bar, phase);
      phase ^= 1;
      if (tile == 1) TC_TRACE(25);
      if (tile == 0) wg_sync(wg);                       // cst of all 8 rows is in place
      __syncwarp();
      tc_fence_after();
      // ---- epilogue: v = A.Bhi + A.Blo + cst; PReLU (alpha per position) and Dense(1) folded into
      //      two tables: sum_j wout_j max(v,0) + alpha_tj wout_j min(v,0) = sum_j v P_tj + |v| Q_tj
      float2 sa = make_float2(p.au_bout, 0.f), sb = make_float2(0.f, 0.f);
      {
        const float* cs = cst + rs * 32;
        const float4* pq = PQtab + t * (kPqStride / 4);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t dh[16], dl[16];
          tmem_ld16(tbase + TM_D + 16 * half + lane_base, dh);
          tmem_ld16(tbase + TM_D + 32 + 16 * half + lane_base, dl);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; jj += 4) {
            const int j = 16 * half + jj;
            const float4 c4 = *reinterpret_cast<const float4*>(cs + j);
            const float4 pq0 = pq[j >> 1], pq1 = pq[(j >> 1) + 1];   // (P_j, P_j+1, Q_j, Q_j+1), next pair
            float2 v01 = add2(make_float2(__uint_as_float(dh[jj]), __uint_as_float(dh[jj + 1])),
                              make_float2(__uint_as_float(dl[jj]), __uint_as_float(dl[jj + 1])));
            float2 v23 = add2(make_float2(__uint_as_float(dh[jj + 2]), __uint_as_float(dh[jj + 3])),
                              make_float2(__uint_as_float(dl[jj + 2]), __uint_as_float(dl[jj + 3])));
            v01 = add2(v01, make_float2(c4.x, c4.y));
            v23 = add2(v23, make_float2(c4.z, c4.w));
            sa = fma2(v01, make_float2(pq0.x, pq0.y), sa);
            sb = fma2(v23, make_float2(pq1.x, pq1.y), sb);
            sa = fma2(make_float2(fabsf(v01.x), fabsf(v01.y)), make_float2(pq0.z, pq0.w), sa);
            sb = fma2(make_float2(fabsf(v23.x), fabsf(v23.y)), make_float2(pq1.z, pq1.w), sb);
          }
        }
      }
      const float s = (sa.x + sa.y) + (sb.x + sb.y);
      if (tile == 1) TC_TRACE(26);
      const float w = valid ? 1.f / (1.f + __expf(-s)) : 0.f;
      // ---- pooling: out[e = lane] = sum over the chunk's 32 positions of w_t * h_t[e]
      float h[32];
      {
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 a = mul2(make_float2(hn[q].x, hn[q].y), w2);
          const float2 bq = mul2(make_float2(hn[q].z, hn[q].w), w2);
          h[4 * q] = a.x; h[4 * q + 1] = a.y; h[4 * q + 2] = bq.x; h[4 * q + 3] = bq.y;
        }
      }
#pragma unroll
      for (int o = 16; o >= 2; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; i += 2) {
          const float send0 = up ? h[i] : h[i + o], send1 = up ? h[i + 1] : h[i + 1 + o];
          const float keep0 = up ? h[i + o] : h[i], keep1 = up ? h[i + 1 + o] : h[i + 1];
          const float2 r = add2(make_float2(keep0, keep1),
                                make_float2(__shfl_xor_sync(0xffffffffu, send0, o),
                                            __shfl_xor_sync(0xffffffffu, send1, o)));
          h[i] = r.x; h[i + 1] = r.y;
        }
      }
      {
        const bool up = (lane & 1) != 0;
        const float send = up ? h[0] : h[1];
        const float keep = up ? h[1] : h[0];
        h[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) hn[q] = hnext[q];
      part[(rs * kTcMaxCPR + cq) * 32 + lane] = h[0];
      if (tile < 24) TC_TRACE(4 + tile);
    }
    wg_sync(wg);

    // ================= phase 2: top MLP on the super-group's 32 rows, whole CTA ===========
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f), c4s = pl;
    if (srs < kTcG) {   // pooled and candidate leave the phase-1 scratch before the X operand overlays it
      for (int c = 0; c < CPR; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(part + (srs * kTcMaxCPR + c) * 32 + 4 * sq);
        pl.x += v.x; pl.y += v.y; pl.z += v.z; pl.w += v.w;
      }
      c4s = *reinterpret_cast<const float4*>(cand + srs * 32 + 4 * sq);
    }
    pre = issue_loads(sg + gridDim.x);                   // next super-group's ids: request from HBM now
    tc_fence_before();
    __syncthreads();
    if (srs < kTcG) {
      uint8_t* xb = cs_base + WS_XB;
      const int xr = wg * kTcG + srs;                    // row slot in the super-group
      store_x4(xb, 0, xr, 4 * sq, ug4);                  // K block 0: [userGenre1 | userId]
      store_x4(xb, 0, xr, 32 + 4 * sq, u4);
      store_x4(xb, 1, xr, 4 * sq, pl);                   // K block 1: [pooled | candidate]
      store_x4(xb, 1, xr, 32 + 4 * sq, c4s);
      store_x4(xb, 2, xr, 4 * sq, mg4);                  // K block 2: [movieGenre1 | 0]
      const uint32_t zoff = 2 * 8192u + sw128_offset(xr, 4 + (sq >> 1)) + ((sq & 1) ? 8u : 0u);
      *reinterpret_cast<uint2*>(xb + zoff) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(xb + zoff + 4096u) = make_uint2(0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tD1 = tmem_slot + TM_D, tD2 = tmem_slot + 128 + TM_D;   // workers 0 / 1 accumulators
    if (tid == 0) {
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const uint64_t ah = smem_desc_sw128(s_img + IMG_W1_HI + kb * 16384);
        const uint64_t al = smem_desc_sw128(s_img + IMG_W1_LO + kb * 16384);
        const uint64_t xs = smem_desc_sw128(s_cs + WS_XB + kb * 8192);     // [X hi | X lo], N = 64
#pragma unroll
        for (int ks = 0; ks < (kb == 2 ? 2 : 4); ++ks) {                   // K block 2: columns 32..63 are zero
          mma_ss(tD1, ah + 2 * ks, xs + 2 * ks, idesc_top, acc);           // W1hi.(Xhi | Xlo)
          acc = 1;
          mma_ss(tD1, al + 2 * ks, xs + 2 * ks, idesc_top, 1);             // W1lo.(Xhi | Xlo)
        }
      }
      mma_commit(&cbar);
    }
    __syncwarp();
    // this thread is unit `tw` of layer 1 for rows 8*wg .. 8*wg+7: constants arrive while the MMAs run
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    TC_TRACE(30);
    {
      uint32_t d[8], d2[8];
      tmem_ld8(tD1 + 8 * wg + lane_base, d);             // W1 . X hi
      tmem_ld8(tD1 + 32 + 8 * wg + lane_base, d2);       // W1 . X lo
      tmem_ld_wait();
      // layer-1 epilogue: bias + numerics (fp32) + PReLU -> H1 operand (bf16 hi/lo, N-stacked),
      // which overlays the X operand (its MMAs have completed)
      const uint32_t koff = (uint32_t)(tw >> 6) * 8192u;
      const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int xr = wg * kTcG + r;
        const float4 n0 = *reinterpret_cast<const float4*>(nums + xr * 8);
        const float4 n1 = *reinterpret_cast<const float4*>(nums + xr * 8 + 4);
        float v = (__uint_as_float(d[r]) + __uint_as_float(d2[r])) + b1;
        v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
        v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
        v = fmaf(n1.z, w1n[6], v);
        v = v > 0.f ? v : a1 * v;
        const uint32_t off = koff + sw128_offset(xr, chunk) + within;
        const __nv_bfloat16 vh = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(cs_base + WS_H1 + off) = vh;
        *reinterpret_cast<__nv_bfloat16*>(cs_base + WS_H1 + off + 4096u) = __float2bfloat16_rn(v - __bfloat162float(vh));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t a = smem_desc_sw128(s_img + IMG_W2 + kb * 16384);
        const uint64_t hs = smem_desc_sw128(s_cs + WS_H1 + kb * 8192);      // [H1 hi | H1 lo], N = 64
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ss(tD2, a + 2 * ks, hs + 2 * ks, idesc_top, acc);            // (W2hi ; W2lo).(H1hi | H1lo)
          acc = 1;
        }
      }
      mma_commit(&cbar);
    }
    __syncwarp();
    const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[8], d2[8];
      tmem_ld8(tD2 + 8 * wg + lane_base, d);
      tmem_ld8(tD2 + 32 + 8 * wg + lane_base, d2);
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 8; ++r) d[r] = __float_as_uint(__uint_as_float(d[r]) + __uint_as_float(d2[r]));
      float* red = reinterpret_cast<float*>(cs_base + WS_RED);    // [64 units][32 rows]
      float* zp = reinterpret_cast<float*>(cs_base + WS_ZP);      // [16][32]
      if (tw >= 64) {                                              // lo halves of W2 -> smem
        *reinterpret_cast<float4*>(red + (tw - 64) * 32 + 8 * wg) =
            make_float4(__uint_as_float(d[0]), __uint_as_float(d[1]), __uint_as_float(d[2]), __uint_as_float(d[3]));
        *reinterpret_cast<float4*>(red + (tw - 64) * 32 + 8 * wg + 4) =
            make_float4(__uint_as_float(d[4]), __uint_as_float(d[5]), __uint_as_float(d[6]), __uint_as_float(d[7]));
      }
      __syncthreads();
      if (tw < 64) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float v = __uint_as_float(d[r]) + red[tw * 32 + 8 * wg + r] + b2;   // (W2hi + W2lo) . (H1hi + H1lo)
          v = v > 0.f ? v : a2 * v;
          red[tw * 32 + 8 * wg + r] = v * w3;
        }
      }
      __syncthreads();
      {  // 32 rows x 16 partial sums of 4 units
        const int r = tid & 31, pt = tid >> 5;
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) s += red[(pt * 4 + u) * 32 + r];
        zp[pt * 32 + r] = s;
      }
      __syncthreads();
      if (tid < kTcSG) {
        float z = p.b3;
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) z += zp[pt * 32 + tid];
        const int row = sg * kTcSG + tid;
        if (row < b.B) {
          store_score(b, row, sigmoidf_acc(z));
          if (b.logits) b.logits[row] = z;
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                     // scratch is reused by the next super-group
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
  return 1024 + ((din_tc_image_bytes(cpr) + 1023u) & ~1023u) + WS_BYTES;
}

template <int CPR>
static cudaError_t launch_din_tc_t(const DinTcParams& p, const BatchView& b, cudaStream_t s) {
  const int n_sg = (b.B + kTcSG - 1) / kTcSG;
  const int grid = n_sg < p.num_sms ? n_sg : p.num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcWG * 128);
  cfg.dynamicSmemBytes = din_tc_smem_bytes(CPR);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see the kernel prologue
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, din_tc_kernel<CPR>, p, b);
}

cudaError_t launch_din_tc(const DinTcParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  switch (p.CPR) {
    case 1: return launch_din_tc_t<1>(p, b, s);
    case 2: return launch_din_tc_t<2>(p, b, s);
    case 3: return launch_din_tc_t<3>(p, b, s);
    case 4: return launch_din_tc_t<4>(p, b, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t setup_din_tc_attributes() {
  cudaError_t e;
#define SRS_ATTR(C_)                                                                     \
  e = cudaFuncSetAttribute(din_tc_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)din_tc_smem_bytes(C_));                                  \
  if (e != cudaSuccess) return e;
  SRS_ATTR(1) SRS_ATTR(2) SRS_ATTR(3) SRS_ATTR(4)
#undef SRS_ATTR
  return cudaSuccess;
}

}  // namespace srs
