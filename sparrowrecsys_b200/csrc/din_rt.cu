// din_rt.cu - DIN forward, "row tile" kernel: history rows go HBM/L2 -> shared memory by
// cp.async straight into tcgen05 operand tiles and are never touched by a CUDA core.
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167.
// Same math as din_tc.cu / din.cu.  The change is where the per-(row, position) work runs:
//
//   * the movie table is stored pre-split, one 128-byte row [32 x bf16 hi | 32 x bf16 lo]
//     per movie (x = hi + lo, both round-to-nearest), so a gathered row IS an MMA operand row;
//   * activation unit: (h*c).Wp = h.(diag(c) Wp), so per batch row r the weight
//         W_r = (Wsub + Wh) + diag(c_r) Wp            [32 x 32]
//     is built once (fp32, then split to bf16 hi/lo) and the 64 positions of two batch rows
//     form one M = 128 tile:  D[128 x 128] = [H_hi | H_lo] * B, B N-stacked per row and half;
//   * gate: v = D_hi + D_lo + cst_r, s = sum_j v_j P_tj + |v_j| Q_tj, w = sigmoid(s)  (CUDA cores,
//     thread = position, P/Q of its position in registers);
//   * pooling: sum_t w_t h_t is a second MMA on the SAME tile read MN-major:
//         D2[64 = hi e | lo e][8] = tile^T[64 x 64 positions] * [w_hi | w_lo | ...]
//     so h is never loaded into registers at all;
//   * top MLP: as in din_tc.cu (transposed, rows of the group are the MMA N); its weight images
//     are copied over the ring when the group's tiles are done.
//
// CTA = 512 threads, one per SM, persistent over groups of up to 32 batch rows; tile K (a
// CTA-global counter) lives in ring slot K % 8.  Warp roles during the tiles:
//   warps 0-3,7  gatherers: cp.async the 100 history rows of tile K into the slot, 3 tiles ahead
//   warp  4      issuer: every tcgen05.mma / commit of the tile phase
//   warps 5-6    builders: the B operand (two W_r) of tile K; afterwards warp 5 streams the top-MLP
//                weight images into ring slots as their last tiles retire
//   warps 8-15   consumers: tile K belongs to consumer K & 1: gate epilogue -> pooling weights,
//                pooled accumulators of its previous tile -> shared memory
// Precision: bf16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate) everywhere, as in din_tc.cu.
#include <climits>

#include "rt_common.cuh"

namespace srs {

constexpr int kRtThreads = 512;
constexpr int kRtRows = 32;                 // row slots per group = N/2 of the top-MLP MMAs
constexpr int kRtSlots = 8;                 // ring slots (history tiles in flight or being consumed)
constexpr int kRtAhead = 3;                 // tiles the gatherers keep in flight ahead of the one they deliver
constexpr int kRtGatherThreads = 160;       // warps 0-3 and 7
constexpr int kRtCopiesPerThread = 7;       // ceil(2 * 64 positions * 8 chunks / 160)
constexpr int kRtIdsLd = 64;                // ints per row of the staged history ids
constexpr int kRtHistPerThread = kRtRows * kRtIdsLd / kRtThreads;

// weight image in global memory (built by build_din_rt in model.cu): W2 | W1 hi | W1 lo.  It is
// copied over the ring when a group's tiles are done - during the tiles that space is ring slots.
constexpr uint32_t RI_BYTES = 131072;
constexpr uint32_t RW2 = 0;                          // ring offsets in phase 2: 2 K blocks x [64 hi | 64 lo units][64 k]
constexpr uint32_t RW1_HI = 32768;                   // 3 K blocks x [128 units][64 k]
constexpr uint32_t RW1_LO = RW1_HI + 49152;
// ring slot: A = [2 rows x 64 positions][hi 32 | lo 32] bf16, SW128 K-major (16 KB)
//            B = [hi: row0 32 units, row1 32 units | lo: same][32 k] bf16, SWIZZLE_64B K-major (8 KB)
constexpr uint32_t RS_A = 16384, RS_B = 8192, RS_SLOT = RS_A + RS_B;
constexpr uint32_t RING_BYTES = kRtSlots * RS_SLOT;  // 196608
// phase-2 scratch behind the weight images
constexpr uint32_t P2_XB = RI_BYTES;                 // 3 K blocks x [32 rows hi | 32 rows lo][64 k]  (24 KB)
constexpr uint32_t P2_H1 = P2_XB;                    // 2 K blocks, after layer 1
constexpr uint32_t P2_RED = P2_XB + 24576;           // f32 [64][32]
constexpr uint32_t P2_ZP = P2_XB + 32768;            // f32 [16][32]
static_assert(P2_ZP + 2048 <= RING_BYTES, "phase-2 scratch must fit in the ring");
// scratch behind the ring
constexpr uint32_t RX_IDS = 0;                       // int [32][64]
constexpr uint32_t RX_CAND = 8192;                   // f32 [32][32]
constexpr uint32_t RX_CST = 12288;                   // f32 [32][32]
constexpr uint32_t RX_POOL = 16384;                  // f32 [32][hi | lo][32]
constexpr uint32_t RX_B2 = 24576;                    // [consumer][buffer] x 2 K blocks x [8 n][64 positions] bf16, SW128
constexpr uint32_t RX_NUMS = 32768;                  // f32 [32][8]
constexpr uint32_t RX_BYTES = 33792;
// tensor memory columns
constexpr uint32_t TMC_D1 = 0;                       // tile K: [128 (K % 3), + 128)
constexpr uint32_t TMC_D2 = 384;                     // consumer q, buffer u, tile row r: 384 + 32 q + 16 u + 8 r
constexpr uint32_t TMC_TOP1 = 0, TMC_TOP2 = 64;      // top-MLP accumulators (phase 2), 64 columns each

__device__ unsigned long long g_din_rt_trace[40];
#define RT_TRACE(slot, cond)                                                     \
  do {                                                                           \
    if (p.trace && blockIdx.x == 0 && (cond)) g_din_rt_trace[slot] = clock64();  \
  } while (0)

__device__ __forceinline__ void rt_store_x4(uint8_t* tile, int block, int row, int col, float4 v) {
  const uint32_t off = block * 8192u + sw128_offset(row, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(tile + off + 4096u) = make_uint2(s0.lo, s1.lo);   // row + 32: same swizzle phase
}

// inputs of a group a thread requests from HBM before it needs them
struct RtGroupLoads {
  int id_a, id_b;                  // which 0: (movieId, userId) raw; which 1: (userGenre1, movieGenre1)
  float nv;                        // which 1: numeric sq of the row
  int hraw[kRtHistPerThread];      // raw history ids of cells tid + u * 512
};

__global__ void __launch_bounds__(kRtThreads, 1) din_rt_kernel(const __grid_constant__ DinRtParams p,
                                                               BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t wbar;                 // top-MLP weight image landed (once per group)
  __shared__ uint64_t cbar;                 // top-MLP MMAs complete
  __shared__ uint64_t full[kRtSlots];       // tile operands in place (160 gatherer + 64 builder arrivals)
  __shared__ uint64_t empty[kRtSlots];      // both MMAs of the tile in the slot have completed
  __shared__ uint64_t d1_full[3];           // tile K: activation-unit accumulators in buffer K % 3 ready
  __shared__ uint64_t w_ready[2];           // consumer q: pooling weights written (128 arrivals)
#ifdef SRS_WREADY_SPLIT
  // one barrier per (consumer, pooled buffer): see din_rt64.cu and profiles/exp/rt_protocol_sim.py
  __shared__ uint64_t w_ready_u1[2];
#define W_READY(q, u) ((u) ? &w_ready_u1[q] : &w_ready[q])
#define W_READY_PAR(K) (((K) >> 2) & 1)
#else
#define W_READY(q, u) (&w_ready[q])
#define W_READY_PAR(K) (((K) >> 1) & 1)
#endif
  __shared__ uint64_t d2_full[2][2];        // consumer q, buffer u: pooled accumulators ready
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  RT_TRACE(0, tid == 0);
  const int lane = tid & 31;
  const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);          // warp-uniform by construction
  const int warp_w = __shfl_sync(0xffffffffu, (tid >> 5) & 3, 0);
  const int tw = tid & 127;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  uint8_t* xs = ring + RING_BYTES;
  int* ids_s = reinterpret_cast<int*>(xs + RX_IDS);
  float* cand = reinterpret_cast<float*>(xs + RX_CAND);
  float* cst = reinterpret_cast<float*>(xs + RX_CST);
  float* pooled = reinterpret_cast<float*>(xs + RX_POOL);
  uint8_t* b2s = xs + RX_B2;
  float* nums = reinterpret_cast<float*>(xs + RX_NUMS);
  const int T = p.T;
  const int RPG = p.rows_per_group;
  const int n_groups = (b.B + RPG - 1) / RPG;
  const bool is_gather = wg == 0 || (wg == 1 && warp_w == 3), is_issuer = wg == 1 && warp_w == 0;
  const bool is_builder = wg == 1 && (warp_w == 1 || warp_w == 2), is_consumer = wg >= 2;
  // phase-0 / phase-2 role: row slot, feature pair, float4 index
  const int xr = tid >> 4, which = (tid >> 3) & 1, sq = tid & 7;

  auto issue_group_loads = [&](int g) -> RtGroupLoads {
    RtGroupLoads L;
    L.id_a = L.id_b = -1; L.nv = 0.f;
#pragma unroll
    for (int u = 0; u < kRtHistPerThread; ++u) L.hraw[u] = 0;
    if (g >= n_groups) return L;
    const int row0 = g * RPG;
    const int nrows = min(RPG, b.B - row0);
    if (xr < nrows) {
      const int row = row0 + xr;
      if (which == 0) {
        L.id_a = __ldg(b.movie_id + row);
        L.id_b = __ldg(b.user_id + row);
      } else {
        L.id_a = __ldg(b.user_genre + row * 5);
        L.id_b = __ldg(b.movie_genre + row * 3);
        if (sq < kNumNumerics) L.nv = __ldg(b.numerics + row * kNumNumerics + sq);
      }
    }
#pragma unroll
    for (int u = 0; u < kRtHistPerThread; ++u) {
      const int i = tid + u * kRtThreads;
      const int r = i >> 6, t = i & 63;
      if (r < nrows && t < T) L.hraw[u] = __ldg(b.hist + (size_t)(row0 + r) * b.hist_stride + t);
    }
    return L;
  };

  // ---- prologue ---------------------------------------------------------------------------
  // One CTA per SM fills the machine, so the next launch cannot start before this one ends anyway:
  // the dependency wait comes first and the ids (HBM) are requested before anything else.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");    // inputs may come from the previous kernel
  RtGroupLoads pre = issue_group_loads(blockIdx.x);
  if (tid < 32) tmem_alloc(&tmem_slot, 512);
  if (tid >= 32 && tid < 64) {                           // warp 1: one mbarrier per lane
    const int l = tid - 32;
    if (l < kRtSlots) mbar_init(&full[l], kRtGatherThreads + 64);
    else if (l < 2 * kRtSlots) mbar_init(&empty[l - kRtSlots], 1);
    else if (l < 2 * kRtSlots + 3) mbar_init(&d1_full[l - 2 * kRtSlots], 1);
    else if (l < 2 * kRtSlots + 5) mbar_init(&w_ready[l - 2 * kRtSlots - 3], 128);
    else if (l < 2 * kRtSlots + 9) mbar_init(&d2_full[(l - 2 * kRtSlots - 5) >> 1][(l - 2 * kRtSlots - 5) & 1], 1);
    else if (l == 2 * kRtSlots + 9) mbar_init(&wbar, 1);
    else if (l == 2 * kRtSlots + 10) mbar_init(&cbar, 1);
#ifdef SRS_WREADY_SPLIT
    else if (l < 2 * kRtSlots + 13) mbar_init(&w_ready_u1[l - 2 * kRtSlots - 11], 128);
#endif
    fence_mbar_init();
  }
  // per-thread constants of the roles
  //   builders : rc[16 c + 0..7] = (Wsub+Wh)[8 cq .. 8 cq + 7][j], rc[16 c + 8..15] = Wp[..][j],
  //              j = bt >> 1, cq = 2 (bt & 1) + c, c = 0, 1
  //   consumers: rc[0..31] = P_t[j], rc[32..63] = Q_t[j] of position t = tw & 63
  float rc[64];
  if (is_consumer) {
    const float* src = p.pq + (size_t)min(tw & 63, T - 1) * 64;      // positions >= T: any finite values do (w is forced to 0)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = ldg4(src + 4 * i);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
  } else if (is_builder) {
    const int bt = tid & 63, j = bt >> 1, cq0 = 2 * (bt & 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {                                     // index 4 i = 16 c + 8 part + 4 h
      const int c = i >> 2, part = (i >> 1) & 1, h = i & 1;
      const float4 v = ldg4((part ? p.wpT : p.waT) + j * 32 + 8 * (cq0 + c) + 4 * h);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 32; i < 64; ++i) rc[i] = 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) rc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RT_TRACE(1, tid == 0);
  const uint32_t tbase = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t s_ring = smem_u32(ring);
  const uint32_t idesc_top = idesc_bf16(128, 2 * kRtRows);
  uint32_t cphase = 0, wphase = 0;
  int kbase = 0;                                        // tiles of earlier groups of this CTA

  // gatherer constants: copy n of this thread moves chunk (i & 7) of history cell (i >> 3), i = gt + 160 n,
  // cells counted row 0 positions 0..T-1, then row 1
  const int gt = tid < 128 ? tid : tid - 96;            // warp 7 (tid 224..255) -> 128..159
  uint32_t g_dst[kRtCopiesPerThread];                   // byte offset in the A tile
  int g_ids[kRtCopiesPerThread];                        // index into the tile's two id rows, -1: no copy
#pragma unroll
  for (int n = 0; n < kRtCopiesPerThread; ++n) {
    const int i = gt + kRtGatherThreads * n, cell = i >> 3, c = i & 7;
    const int r = cell >= T ? 1 : 0, pos = cell - r * T;
    g_ids[n] = cell < 2 * T ? r * kRtIdsLd + pos : -1;
    g_dst[n] = (uint32_t)(r * 64 + pos) * 128u + (uint32_t)((c ^ (pos & 7)) << 4) ;
  }
  const uint32_t g_src = (uint32_t)(gt & 7) * 16u;
  auto gather = [&](int k) {                            // local tile k -> slot of global tile kbase + k
    const int K = kbase + k, slot = K % kRtSlots;
    if (K >= kRtSlots) mbar_wait(&empty[slot], ((K / kRtSlots) + 1) & 1);
    uint8_t* A = ring + slot * RS_SLOT;
    const int* idrow = ids_s + 2 * k * kRtIdsLd;
#pragma unroll
    for (int n = 0; n < kRtCopiesPerThread; ++n)
      if (g_ids[n] >= 0) cp_async16(A + g_dst[n], p.movie_split + (size_t)idrow[g_ids[n]] * 128 + g_src);
  };

  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int row0 = g * RPG;
    const int nrows = min(RPG, b.B - row0);
    const int n_tiles = (nrows + 1) >> 1;

    // ================= phase 0: ids, side rows, candidate rows =============================
    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;     // which 0: (candidate, user); 1: (userGenre1, movieGenre1)
    {
      const bool live = xr < nrows;
      if (which == 0) {
        if (live) {
          const int cid = checked_id(rt_f32_roundtrip_id(pre.id_a), p.n_movies, b.err_flag);
          const int uid = checked_id(pre.id_b, p.n_users, b.err_flag);
          fa = ldg4(p.movie + (size_t)cid * 32 + 4 * sq);
          fb = ldg4(p.user + (size_t)uid * 32 + 4 * sq);
        }
      } else if (live) {
        int ug = pre.id_a, mg = pre.id_b;
        if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
        if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
        if (ug >= 0) fa = ldg4(p.ugenre + ug * 32 + 4 * sq);
        if (mg >= 0) fb = ldg4(p.mgenre + mg * 32 + 4 * sq);
      }
      RT_TRACE(30, tid == 0);
      // history ids (float32 round trip, range check) of the whole group
#pragma unroll
      for (int u = 0; u < kRtHistPerThread; ++u)
        ids_s[tid + u * kRtThreads] = checked_id(rt_f32_roundtrip_id(pre.hraw[u]), p.n_movies, b.err_flag);
      RT_TRACE(31, tid == 0);
      // tile rows of positions >= T are read by both MMAs: keep them zero (phase 2 of the previous
      // group used the ring for the weight images and as scratch)
      const int pad = 64 - T;                                     // thread -> (pad row tid >> 3, chunk tid & 7)
      if (tid < pad * 8) {
#pragma unroll
        for (int sr = 0; sr < kRtSlots * 2; ++sr)                 // sr = slot * 2 + row
          *reinterpret_cast<uint4*>(ring + (sr >> 1) * RS_SLOT + ((sr & 1) * 64 + T + (tid >> 3)) * 128 +
                                    ((tid & 7) << 4)) = make_uint4(0, 0, 0, 0);
      }
    }
    RT_TRACE(32, tid == 0);
    if (which == 0) *reinterpret_cast<float4*>(cand + xr * 32 + 4 * sq) = fa;
    else nums[xr * 8 + sq] = pre.nv;
    RT_TRACE(33, tid == 0);
    __syncthreads();                                        // history ids, candidate rows staged
    RT_TRACE(2, tid == 0);

    // ================= phase 1: tiles ====================================================
    if (is_gather) {
#pragma unroll
      for (int a = 0; a < kRtAhead; ++a) {
        if (a < n_tiles) gather(a);
        cp_async_commit();
      }
      RT_TRACE(34, tid == 0);
      for (int k = 0; k < n_tiles; ++k) {
        const int slot = (kbase + k) % kRtSlots;
        cp_async_wait<kRtAhead - 1>();                      // this tile's rows have landed (later tiles' may be in flight)
        fence_async_smem();
        mbar_arrive(&full[slot]);
        if (k == 0) RT_TRACE(20, tid == 0);
        if (k == 2) RT_TRACE(21, tid == 0);
        if (k == 4) RT_TRACE(22, tid == 0);
        if (k + kRtAhead < n_tiles) gather(k + kRtAhead);   // its slot frees when tile K + kRtAhead - kRtSlots retires
        cp_async_commit();
      }
      cp_async_wait<0>();
    } else if (is_builder) {
      // ---- B operand of every tile: W_r = (Wsub+Wh) + diag(c_r) Wp, bf16 hi / lo
      const int bt = tid & 63, pj = bt >> 1, cq0 = 2 * (bt & 1);
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, slot = K % kRtSlots;
        if (K >= kRtSlots) mbar_wait(&empty[slot], ((K / kRtSlots) + 1) & 1);
        uint8_t* Bt = ring + slot * RS_SLOT + RS_A;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int cq = cq0 + c;
            const float* cv = cand + (2 * k + r) * 32 + 8 * cq;
            const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
            const float* wa = rc + 16 * c;
            const float* wp = rc + 16 * c + 8;
            const float2 v0 = fma2(make_float2(c0.x, c0.y), make_float2(wp[0], wp[1]), make_float2(wa[0], wa[1]));
            const float2 v1 = fma2(make_float2(c0.z, c0.w), make_float2(wp[2], wp[3]), make_float2(wa[2], wa[3]));
            const float2 v2 = fma2(make_float2(c1.x, c1.y), make_float2(wp[4], wp[5]), make_float2(wa[4], wa[5]));
            const float2 v3 = fma2(make_float2(c1.z, c1.w), make_float2(wp[6], wp[7]), make_float2(wa[6], wa[7]));
            const Split2 s0 = split_pack(v0.x, v0.y), s1 = split_pack(v1.x, v1.y);
            const Split2 s2 = split_pack(v2.x, v2.y), s3 = split_pack(v3.x, v3.y);
            const uint32_t n = r * 32 + pj;
            *reinterpret_cast<uint4*>(Bt + sw64_offset(n, cq)) = make_uint4(s0.hi, s1.hi, s2.hi, s3.hi);
            *reinterpret_cast<uint4*>(Bt + sw64_offset(64 + n, cq)) = make_uint4(s0.lo, s1.lo, s2.lo, s3.lo);
          }
        fence_async_smem();
        mbar_arrive(&full[slot]);
      }
      // ---- top-MLP weight images: each ring slot receives its part as soon as its last tile retires
      if (warp_w == 1) {
    if (lane == 0) {
          mbar_arrive_expect_tx(&wbar, RI_BYTES);
          const int tail = n_tiles < kRtSlots ? n_tiles : kRtSlots;      // the last `tail` tiles hold distinct slots
          auto load_slot = [&](int slot) {
            const uint32_t off = slot * RS_SLOT;
            if (off < RI_BYTES) bulk_g2s(ring + off, p.image + off, min(RS_SLOT, RI_BYTES - off), &wbar);
          };
          for (int sl = 0; sl < kRtSlots; ++sl) {                         // slots no tile of this group uses
            bool used = false;
            for (int j = 0; j < tail; ++j) used |= ((kbase + n_tiles - tail + j) % kRtSlots) == sl;
            if (!used) load_slot(sl);
          }
          for (int j = 0; j < tail; ++j) {
            const int K = kbase + n_tiles - tail + j, slot = K % kRtSlots;
            mbar_wait(&empty[slot], (K / kRtSlots) & 1);
            load_slot(slot);
          }
        }
        __syncwarp();
      }
    } else if (is_issuer) {
      // ---- every MMA of the tile phase, in the order the operands become ready
      auto mma1 = [&](int k) {
        const int K = kbase + k, slot = K % kRtSlots, db = K % 3;
        mbar_wait(&full[slot], (K / kRtSlots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD1 = tbase + TMC_D1 + 128u * db;
          const uint64_t ad = smem_desc_sw128(s_ring + slot * RS_SLOT);
          const uint64_t bd = smem_desc_sw64(s_ring + slot * RS_SLOT + RS_A);
          mma_ss(tD1, ad + 0, bd + 0, idesc_bf16(128, 128), 0);     // H_hi . [W_hi | W_lo]
          mma_ss(tD1, ad + 2, bd + 2, idesc_bf16(128, 128), 1);
          mma_ss(tD1, ad + 4, bd + 0, idesc_bf16(128, 64), 1);      // H_lo . W_hi
          mma_ss(tD1, ad + 6, bd + 2, idesc_bf16(128, 64), 1);
          mma_commit(&d1_full[db]);
        }
        __syncwarp();
      };
      if (0 < n_tiles) mma1(0);
      if (1 < n_tiles) mma1(1);
      if (2 < n_tiles) mma1(2);
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, slot = K % kRtSlots, q = K & 1, u = (K >> 1) & 1;
        mbar_wait(W_READY(q, u), W_READY_PAR(K));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD2 = tbase + TMC_D2 + 32u * q + 16u * u;
          const uint32_t s_b2 = smem_u32(b2s) + (q * 2 + u) * 2048;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ss(tD2 + 8 * r, smem_desc_mn_sw128(s_ring + slot * RS_SLOT + r * 8192 + ks * 2048),
                     smem_desc_sw128(s_b2 + r * 1024) + 2 * ks, idesc_mn(64, 8, 1), ks > 0);
          mma_commit(&d2_full[q][u]);
          mma_commit(&empty[slot]);
        }
        __syncwarp();
        if (k + 3 < n_tiles) mma1(k + 3);                   // accumulator buffer K % 3 was read before w_ready
      }
    } else if (is_consumer) {
      const int q = wg - 2;
      const int r_t = warp_w >> 1, t = tw & 63;             // this thread's tile row and position
      // cst[xr][j] = au_b[j] + sum_e cand[xr][e] (Wc - Wsub)[e][j]: 256 threads x 4 outputs
      {
        const int ct = tid - 256, cr = ct >> 3, j0 = (ct & 7) * 4;
        float4 acc = ldg4(p.au_b + j0);
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
          const float cv = cand[cr * 32 + e];
          const float4 w4 = ldg4(p.au_wc + e * 32 + j0);
          acc.x = fmaf(cv, w4.x, acc.x); acc.y = fmaf(cv, w4.y, acc.y);
          acc.z = fmaf(cv, w4.z, acc.z); acc.w = fmaf(cv, w4.w, acc.w);
        }
        *reinterpret_cast<float4*>(cst + cr * 32 + j0) = acc;
      }
      named_sync(5, 256);
      // pooled accumulators of local tile k -> shared memory
      auto pool_out = [&](int k) {
        const int K = kbase + k, u = (K >> 1) & 1;
        mbar_wait(&d2_full[q][u], (K >> 2) & 1);
        tc_fence_after();
        // D2 row m = 16 warp_w + lane (lane < 16): m < 32 -> hi e = m, else lo e = m - 32;
        // columns 8 r + {0: . w_hi, 1: . w_lo}
        uint32_t d[16];
        tmem_ld16(tbase + TMC_D2 + 32u * q + 16u * u + lane_base, d);
        tmem_ld_wait();
        if (lane < 16) {
          const int m = 16 * warp_w + lane;                  // pooled[row][m]: hi part (m < 32) or lo part
          const bool hi = warp_w < 2;
          pooled[(2 * k) * 64 + m] = hi ? __uint_as_float(d[0]) + __uint_as_float(d[1]) : __uint_as_float(d[0]);
          pooled[(2 * k + 1) * 64 + m] = hi ? __uint_as_float(d[8]) + __uint_as_float(d[9]) : __uint_as_float(d[8]);
        }
        tc_fence_before();
      };
      const int first = (q - kbase) & 1;                    // local tiles k with (kbase + k) & 1 == q
      for (int k = first; k < n_tiles; k += 2) {
        const int K = kbase + k, u = (K >> 1) & 1;
        const uint32_t tD1 = tbase + TMC_D1 + 128u * (K % 3);
        mbar_wait(&d1_full[K % 3], (K / 3) & 1);
        tc_fence_after();
        if (k == first) RT_TRACE(3, tid == 256);
        if (k == first + 2) RT_TRACE(10, tid == 256);
        // ---- gate: v = D_hi + D_lo + cst; s = sum_j v_j P_tj + |v_j| Q_tj
        float2 sa = make_float2(p.au_bout, 0.f), sb = make_float2(0.f, 0.f);
        {
          const float* cs = cst + (2 * k + r_t) * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t dh[16], dl[16];
            tmem_ld16(tD1 + r_t * 32 + 16 * half + lane_base, dh);
            tmem_ld16(tD1 + 64 + r_t * 32 + 16 * half + lane_base, dl);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const int j = 16 * half + jj;
              const float4 c4 = *reinterpret_cast<const float4*>(cs + j);
              float2 v01 = add2(make_float2(__uint_as_float(dh[jj]), __uint_as_float(dh[jj + 1])),
                                make_float2(__uint_as_float(dl[jj]), __uint_as_float(dl[jj + 1])));
              float2 v23 = add2(make_float2(__uint_as_float(dh[jj + 2]), __uint_as_float(dh[jj + 3])),
                                make_float2(__uint_as_float(dl[jj + 2]), __uint_as_float(dl[jj + 3])));
              v01 = add2(v01, make_float2(c4.x, c4.y));
              v23 = add2(v23, make_float2(c4.z, c4.w));
              sa = fma2(v01, make_float2(rc[j], rc[j + 1]), sa);
              sb = fma2(v23, make_float2(rc[j + 2], rc[j + 3]), sb);
              sa = fma2(make_float2(fabsf(v01.x), fabsf(v01.y)), make_float2(rc[32 + j], rc[32 + j + 1]), sa);
              sb = fma2(make_float2(fabsf(v23.x), fabsf(v23.y)), make_float2(rc[32 + j + 2], rc[32 + j + 3]), sb);
            }
          }
        }
        const float s = (sa.x + sa.y) + (sb.x + sb.y);
        const float w = (t < T) ? 1.f / (1.f + __expf(-s)) : 0.f;
        {
          // pooling weights operand: K block r_t, row 0 = w hi, row 1 = w lo, column = position
          // (rows 2..7 feed accumulator columns nobody reads)
          const __nv_bfloat16 wh = __float2bfloat16_rn(w);
          const __nv_bfloat16 wl = __float2bfloat16_rn(w - __bfloat162float(wh));
          uint8_t* dstw = b2s + (q * 2 + u) * 2048 + r_t * 1024 + (t & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(0, t >> 3)) = wh;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(1, t >> 3)) = wl;
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(W_READY(q, u));
        if (k == first + 2) RT_TRACE(11, tid == 256);
        if (k - 2 >= 0) pool_out(k - 2);                    // the previous tile's pooling MMAs finished long ago
        if (k == first + 2) RT_TRACE(16, tid == 256);
      }
      {
        const int last = first + ((n_tiles - 1 - first) & ~1);
        if (first < n_tiles) pool_out(last);
      }
      RT_TRACE(4, tid == 256);
    }
    kbase += n_tiles;
    tc_fence_before();
    __syncthreads();
    RT_TRACE(5, tid == 0);
    pre = issue_group_loads(g + gridDim.x);               // next group's ids: request from HBM now

    // ================= phase 2: top MLP on the group's 32 row slots, whole CTA ===============
    {
      uint8_t* xb = ring + P2_XB;
      if (which == 0) {
        rt_store_x4(xb, 0, xr, 32 + 4 * sq, fb);             // K block 0: [userGenre1 | userId]
        rt_store_x4(xb, 1, xr, 32 + 4 * sq, fa);             // K block 1: [pooled | candidate]
        const uint32_t zoff = 2 * 8192u + sw128_offset(xr, 4 + (sq >> 1)) + ((sq & 1) ? 8u : 0u);
        *reinterpret_cast<uint2*>(xb + zoff) = make_uint2(0u, 0u);          // K block 2: [movieGenre1 | 0]
        *reinterpret_cast<uint2*>(xb + zoff + 4096u) = make_uint2(0u, 0u);
      } else {
        float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xr < nrows) {
          const float4 ph = *reinterpret_cast<const float4*>(pooled + xr * 64 + 4 * sq);
          const float4 pw = *reinterpret_cast<const float4*>(pooled + xr * 64 + 32 + 4 * sq);
          pl = make_float4(ph.x + pw.x, ph.y + pw.y, ph.z + pw.z, ph.w + pw.w);
        }
        rt_store_x4(xb, 0, xr, 4 * sq, fa);
        rt_store_x4(xb, 1, xr, 4 * sq, pl);
        rt_store_x4(xb, 2, xr, 4 * sq, fb);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tT1 = tbase + TMC_TOP1, tT2 = tbase + TMC_TOP2;
    RT_TRACE(24, tid == 0);
    if (wg == 0 && warp_w == 0) {
      mbar_wait(&wbar, wphase);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t ah = smem_desc_sw128(s_ring + RW1_HI + kb * 16384);
          const uint64_t al = smem_desc_sw128(s_ring + RW1_LO + kb * 16384);
          const uint64_t xd = smem_desc_sw128(s_ring + P2_XB + kb * 8192);      // [X hi | X lo], N = 64
#pragma unroll
          for (int ks = 0; ks < (kb == 2 ? 2 : 4); ++ks) {                     // K block 2: columns 32..63 are zero
            mma_ss(tT1, ah + 2 * ks, xd + 2 * ks, idesc_top, acc);             // W1hi.(Xhi | Xlo)
            acc = 1;
            mma_ss(tT1, al + 2 * ks, xd + 2 * ks, idesc_top, 1);               // W1lo.(Xhi | Xlo)
          }
        }
        mma_commit(&cbar);
      }
      __syncwarp();
    }
    RT_TRACE(25, tid == 0);
    wphase ^= 1;
    // this thread is unit `tw` of layer 1 for row slots 8*wg .. 8*wg+7
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    RT_TRACE(6, tid == 0);
    {
      uint32_t d[8], d2[8];
      tmem_ld8(tT1 + 8 * wg + lane_base, d);               // W1 . X hi
      tmem_ld8(tT1 + 32 + 8 * wg + lane_base, d2);         // W1 . X lo
      tmem_ld_wait();
      const uint32_t koff = (uint32_t)(tw >> 6) * 8192u;
      const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int sr = wg * 8 + r;
        const float4 n0 = *reinterpret_cast<const float4*>(nums + sr * 8);
        const float4 n1 = *reinterpret_cast<const float4*>(nums + sr * 8 + 4);
        float v = (__uint_as_float(d[r]) + __uint_as_float(d2[r])) + b1;
        v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
        v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
        v = fmaf(n1.z, w1n[6], v);
        v = v > 0.f ? v : a1 * v;
        const uint32_t off = koff + sw128_offset(sr, chunk) + within;
        const __nv_bfloat16 vh = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(ring + P2_H1 + off) = vh;
        *reinterpret_cast<__nv_bfloat16*>(ring + P2_H1 + off + 4096u) = __float2bfloat16_rn(v - __bfloat162float(vh));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (wg == 0 && warp_w == 0) {
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a = smem_desc_sw128(s_ring + RW2 + kb * 16384);
          const uint64_t hs = smem_desc_sw128(s_ring + P2_H1 + kb * 8192);      // [H1 hi | H1 lo], N = 64
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ss(tT2, a + 2 * ks, hs + 2 * ks, idesc_top, acc);              // (W2hi ; W2lo).(H1hi | H1lo)
            acc = 1;
          }
        }
        mma_commit(&cbar);
      }
      __syncwarp();
    }
    const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[8], d2[8];
      tmem_ld8(tT2 + 8 * wg + lane_base, d);
      tmem_ld8(tT2 + 32 + 8 * wg + lane_base, d2);
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 8; ++r) d[r] = __float_as_uint(__uint_as_float(d[r]) + __uint_as_float(d2[r]));
      float* red = reinterpret_cast<float*>(ring + P2_RED);    // [64 units][32 rows]
      float* zp = reinterpret_cast<float*>(ring + P2_ZP);      // [16][32]
      if (tw >= 64) {                                          // lo halves of W2 -> smem
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
        float sum = 0.f;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) sum += red[(pt * 4 + uu) * 32 + r];
        zp[pt * 32 + r] = sum;
      }
      __syncthreads();
      if (tid < kRtRows) {
        float z = p.b3;
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) z += zp[pt * 32 + tid];
        if (tid < nrows) {
          store_score(b, row0 + tid, sigmoidf_acc(z));
          if (b.logits) b.logits[row0 + tid] = z;
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                     // ring and scratch are reused by the next group
    RT_TRACE(7, tid == 0);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, 512);
  RT_TRACE(8, tid == 0);
  gather_signal_tail(b);                                  // spanning ranking call: publish "slice complete"
}

// fp32 table [rows][32] -> [rows][32 x bf16 hi | 32 x bf16 lo]
__global__ void split_table_kernel(const float* __restrict__ src, uint32_t* __restrict__ dst, int64_t n_pairs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // pair index: row * 16 + pair
  if (i >= n_pairs) return;
  const int64_t row = i >> 4;
  const int pr = (int)(i & 15);
  const float2 v = *reinterpret_cast<const float2*>(src + row * 32 + 2 * pr);
  const Split2 s = split_pack(v.x, v.y);
  dst[row * 32 + pr] = s.hi;
  dst[row * 32 + 16 + pr] = s.lo;
}

cudaError_t launch_split_table(const float* src, void* dst, int64_t rows, cudaStream_t s) {
  const int64_t n_pairs = rows * 16;
  const int threads = 256;
  const int64_t blocks = (n_pairs + threads - 1) / threads;
  split_table_kernel<<<(unsigned)blocks, threads, 0, s>>>(src, reinterpret_cast<uint32_t*>(dst), n_pairs);
  ++g_launch_count;
  return cudaGetLastError();
}

cudaError_t read_din_rt_trace(unsigned long long* out40) {
  return cudaMemcpyFromSymbol(out40, g_din_rt_trace, sizeof(unsigned long long) * 40);
}

size_t din_rt_smem_bytes() { return 1024 + RING_BYTES + RX_BYTES; }

cudaError_t launch_din_rt(const DinRtParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  DinRtParams q = p;
  // rows per group: as even as possible over the SMs, at most 32, even
  const int waves = (b.B + kRtRows * p.num_sms - 1) / (kRtRows * p.num_sms);
  int rpg = (b.B + waves * p.num_sms - 1) / (waves * p.num_sms);
  rpg = (rpg + 1) & ~1;
  if (rpg > kRtRows) rpg = kRtRows;
  if (rpg < 2) rpg = 2;
  q.rows_per_group = rpg;
  const int n_groups = (b.B + rpg - 1) / rpg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_groups < p.num_sms ? n_groups : p.num_sms);
  cfg.blockDim = dim3(kRtThreads);
  cfg.dynamicSmemBytes = din_rt_smem_bytes();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see the kernel prologue
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, din_rt_kernel, q, b);
}

cudaError_t setup_din_rt_attributes() {
  return cudaFuncSetAttribute(din_rt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)din_rt_smem_bytes());
}

}  // namespace srs
