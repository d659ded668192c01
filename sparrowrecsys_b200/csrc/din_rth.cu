// din_rth.cu - DIN forward, "row tile" kernel at half an SM per CTA (SRS_DIN_IMPL=rth).
//
// EXPERIMENTAL, opt-in, NOT YET RUN ON A GPU (written after the round's GPU budget was spent;
// tests/test_gpu_parity.py::test_din_rth_kernel is skipped unless SRS_TEST_RTH=1).
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167.  Same math,
// operand forms and weight images as din_rt.cu (read that header first).  What changes is the
// footprint: 256 threads, 102 KB of shared memory and 256 TMEM columns per CTA, so that TWO CTAs
// share an SM.  A din_rt launch spends 60 % of its time in serial latency chains (prologue, id
// fetch, first tile, top-MLP round trips) during which its SM idles; with two resident CTAs -
// of one launch (ctas_per_sm = 2: groups of 14 rows) or of two launches on different streams
// (ctas_per_sm = 1, bench.py --streams 2 without an SM limit) - one CTA's chains overlap the
// other's tile phase, which runs at the shared-memory bandwidth either way.
//
// Differences from din_rt_kernel:
//   * ring of 3 slots (24 KB each), gatherers 2 tiles ahead; warps 0-1 gather, warp 2 issues every
//     tcgen05.mma, warp 3 builds the per-row weight operand, warps 4-7 are ONE consumer that takes
//     every tile (TMEM lane quarter = warp & 3);
//   * ONE activation-unit accumulator buffer (TMEM columns 0..127): the consumer releases it
//     (`d1_free`) as soon as its tcgen05.ld of the tile has completed, half way through the gate,
//     so MMA1 of tile K+1 overlaps the gate arithmetic of tile K; pooled accumulators in columns
//     128..159 (two buffers), top-MLP accumulators alias columns 0..127;
//   * the 128 KB top-MLP weight image does not fit: its eight 16 KB K-block pieces stream through
//     three buffers behind the X operand (`pfull` / `pfree` mbarriers, bulk copies by warp 3),
//     each consumed by the MMAs of its K block;
//   * no register prefetch of the next group's ids (the other CTA covers that latency);
//   * optional (SRS_DIN_RTH_BG=1 -> p.nch = 1): the builder warp also gathers a third of the tile's
//     rows (96 gathering threads x 11 copies instead of 64 x 16) - cp.async issue is what bounds
//     the tile phase, and the builder is idle most of the time.
#include <climits>

#include "rt_common.cuh"

namespace srs {

namespace {

constexpr int kHThreads = 256;
constexpr int kHRows = 32;                  // row slots per group = N/2 of the top-MLP MMAs
constexpr int kHSlots = 3;
constexpr int kHAhead = 2;                  // tiles in flight ahead of the one being delivered
constexpr int kHGatherThreads = 64;         // warps 0-1
constexpr int kHBuilderThreads = 32;        // warp 3
constexpr int kHCopies = 16;                // 2 rows * 64 positions * 8 chunks / 64 threads
constexpr int kHIdsLd = 64;

constexpr uint32_t HS_A = 16384, HS_B = 8192, HS_SLOT = HS_A + HS_B;
constexpr uint32_t HRING = kHSlots * HS_SLOT;           // 73728
// phase 2 view of the ring
constexpr uint32_t H2_XB = 0;                           // 3 K blocks x [32 rows hi | 32 rows lo][64 k]  (24 KB)
constexpr uint32_t H2_H1 = 0;                           // 2 K blocks, after layer 1 (over the X operand)
constexpr uint32_t H2_PIECE = 24576;                    // 3 x 16 KB weight pieces
constexpr uint32_t H2_PIECE_BYTES = 16384;
static_assert(H2_PIECE + 3 * H2_PIECE_BYTES <= HRING, "pieces must fit behind the X operand");
// weight image in global memory (model.cu::build_din_rt): W2 | W1 hi | W1 lo
constexpr uint32_t HI_W2 = 0, HI_W1_HI = 32768, HI_W1_LO = 81920;
// scratch behind the ring
constexpr uint32_t HX_IDS = 0;                          // int [32][64]; phase 2: f32 red[64][32]
constexpr uint32_t HX_CAND = 8192;                      // f32 [32][32]; phase 2: f32 zp[8][32]
constexpr uint32_t HX_CST = 12288;                      // f32 [32][32]
constexpr uint32_t HX_POOL = 16384;                     // f32 [32][hi 32 | lo 32]
constexpr uint32_t HX_B2 = 24576;                       // [buffer] x 2 K blocks x [8 n][64 positions] bf16, SW128
constexpr uint32_t HX_NUMS = 28672;                     // f32 [32][8]
constexpr uint32_t HX_SID = 29696;                      // int [32][4]: checked candidate, user, userGenre1, movieGenre1 ids
constexpr uint32_t HX_BYTES = 30208;
// tensor memory columns (256 allocated)
constexpr uint32_t HT_D1 = 0;                           // [128 x 128] activation-unit accumulators
constexpr uint32_t HT_D2 = 128;                         // buffer u, tile row r: 128 + 16 u + 8 r
constexpr uint32_t HT_TOP1 = 0, HT_TOP2 = 64;
constexpr uint32_t HT_COLS = 256;

__device__ unsigned long long g_din_rth_trace[40];
// phase timestamps of CTA 0 (srs_debug_din_trace; slot meaning as in din_rt.cu / profiles/trace_din_rt.py)
#define RTH_TRACE(slot, cond)                                                     \
  do {                                                                            \
    if (p.trace && blockIdx.x == 0 && (cond)) g_din_rth_trace[slot] = clock64();  \
  } while (0)

__device__ __forceinline__ void rth_store_x4(uint8_t* tile, int block, int row, int col, float4 v) {
  const uint32_t off = block * 8192u + sw128_offset(row, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(tile + off + 4096u) = make_uint2(s0.lo, s1.lo);   // row + 32: same swizzle phase
}

}  // namespace

__global__ void __launch_bounds__(kHThreads, 2) din_rth_kernel(const __grid_constant__ DinRtParams p,
                                                               BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t full[kHSlots];        // tile operands in place (64 gatherer + 32 builder arrivals)
  __shared__ uint64_t empty[kHSlots];       // both MMAs of the tile in the slot have completed
  __shared__ uint64_t d1_full;              // tile K: activation-unit accumulators ready
  __shared__ uint64_t d1_free;              // tile K: the consumer has read them (128 arrivals)
  __shared__ uint64_t w_ready[2];           // tile K -> [K & 1]: pooling weights written (128 arrivals).  Two,
                                            // because MMA1 of tile K + 1 is issued before the issuer waits for
                                            // tile K here: one barrier could run a phase ahead of its waiter
  __shared__ uint64_t d2_full[2];           // buffer u: pooled accumulators ready
  __shared__ uint64_t pfull[3];             // weight piece landed in buffer j
  __shared__ uint64_t pfree[3];             // MMAs reading buffer j have completed
  __shared__ uint64_t cbar;                 // top-MLP layer complete
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  RTH_TRACE(0, tid == 0);
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // warp-uniform by construction
  const int wg = warp >> 2;                                      // 0: warps 0-3, 1: warps 4-7
  const int warp_w = warp & 3;                                   // TMEM lane quarter of this warp
  const int tw = tid & 127;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  uint8_t* xs = ring + HRING;
  int* ids_s = reinterpret_cast<int*>(xs + HX_IDS);
  float* cand = reinterpret_cast<float*>(xs + HX_CAND);
  float* cst = reinterpret_cast<float*>(xs + HX_CST);
  float* pooled = reinterpret_cast<float*>(xs + HX_POOL);
  uint8_t* b2s = xs + HX_B2;
  float* nums = reinterpret_cast<float*>(xs + HX_NUMS);
  int* sid = reinterpret_cast<int*>(xs + HX_SID);
  const int T = p.T;
  const int RPG = p.rows_per_group;
  const int n_groups = (b.B + RPG - 1) / RPG;
  const bool is_gather = warp < 2, is_issuer = warp == 2, is_builder = warp == 3, is_consumer = wg == 1;

  // ---- prologue ---------------------------------------------------------------------------
  // No programmatic dependent launch here: a dependent CTA would take the SM's second CTA slot
  // and sit in griddepcontrol.wait, which is exactly the slot another stream's launch should get.
  if (warp == 0) tmem_alloc(&tmem_slot, HT_COLS);
  if (warp == 1) {                                       // one mbarrier per lane
    if (lane < 3) mbar_init(&full[lane], (p.nch ? 96 : kHGatherThreads) + kHBuilderThreads);
    else if (lane < 6) mbar_init(&empty[lane - 3], 1);
    else if (lane == 6) mbar_init(&d1_full, 1);
    else if (lane == 7) mbar_init(&d1_free, 128);
    else if (lane < 10) mbar_init(&w_ready[lane - 8], 128);
    else if (lane < 12) mbar_init(&d2_full[lane - 10], 1);
    else if (lane < 15) mbar_init(&pfull[lane - 12], 1);
    else if (lane < 18) mbar_init(&pfree[lane - 15], 1);
    else if (lane == 18) mbar_init(&cbar, 1);
    fence_mbar_init();
  }
  // per-thread constants of the roles
  //   builder : rc[16 cq + 0..7] = (Wsub+Wh)[8 cq .. 8 cq + 7][j], rc[16 cq + 8..15] = Wp[..][j], j = lane
  //   consumer: rc[0..31] = P_t[j], rc[32..63] = Q_t[j] of position t = (32 warp_w + lane) & 63
  float rc[64];
  if (is_consumer) {
    const float* src = p.pq + (size_t)min((32 * warp_w + lane) & 63, T - 1) * 64;   // positions >= T: w is forced to 0
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = ldg4(src + 4 * i);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
  } else if (is_builder) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {                                    // index 4 i = 16 cq + 8 part + 4 h
      const int cq = i >> 2, part = (i >> 1) & 1, h = i & 1;
      const float4 v = ldg4((part ? p.wpT : p.waT) + lane * 32 + 8 * cq + 4 * h);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) rc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RTH_TRACE(1, tid == 0);
  const uint32_t tbase = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t s_ring = smem_u32(ring);
  const uint32_t idesc_top = idesc_bf16(128, 2 * kHRows);
  uint32_t cphase = 0;
  int kbase = 0;                                        // tiles of earlier groups of this CTA
  int pbase = 0;                                        // weight pieces of earlier groups

  // gatherer constants: copy n of this thread moves chunk (i & 7) of history cell (i >> 3), i = gt + G n,
  // cells counted row 0 positions 0..T-1, then row 1; G = 64 gathering threads (warps 0-1), or 96 when
  // the builder warp gathers too (gt = 64 + lane there): 16 resp. 11 copies per thread
  const bool builder_gathers = p.nch != 0;
  const int gthreads = builder_gathers ? 96 : kHGatherThreads;
  const int ncopies = builder_gathers ? 11 : kHCopies;
  const int gt = is_builder ? 64 + lane : (tid & 63);
  const uint32_t g_chunk = (uint32_t)(gt & 7);          // gthreads % 8 == 0: every copy of a thread moves the same chunk
  auto gather = [&](int k) {                            // local tile k -> slot of global tile kbase + k
    const int K = kbase + k, slot = K % kHSlots;
    if (K >= kHSlots) mbar_wait(&empty[slot], ((K / kHSlots) + 1) & 1);
    uint8_t* A = ring + slot * HS_SLOT;
    const int* idrow = ids_s + 2 * k * kHIdsLd;
#pragma unroll
    for (int n = 0; n < kHCopies; ++n) {
      const int cell = (gt >> 3) + (gthreads >> 3) * n;   // addresses on the fly: 32 registers less than tables
      if (n < ncopies && cell < 2 * T) {
        const int r = cell >= T ? 1 : 0, pos = cell - r * T;
        cp_async16(A + (uint32_t)(r * 64 + pos) * 128u + ((g_chunk ^ (uint32_t)(pos & 7)) << 4),
                   p.movie_split + (size_t)idrow[r * kHIdsLd + pos] * 128 + g_chunk * 16u);
      }
    }
  };

  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int row0 = g * RPG;
    const int nrows = min(RPG, b.B - row0);
    const int n_tiles = (nrows + 1) >> 1;

    // ================= phase 0: ids, side rows, candidate rows =============================
    // side features: threads 0..127 own 4 items each; item -> (row slot, feature pair, float4).
    // Only the candidate row and the numerics are needed before phase 2; the user / genre rows are
    // prefetched into L2 here and read in phase 2 (holding them in registers across the tile phase
    // would cost the consumer warps 32 registers they do not have).
    {
      // history ids of the whole group: float32 round trip, range check
      int hraw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * kHThreads, r = i >> 6, t = i & 63;
        hraw[u] = (r < nrows && t < T) ? __ldg(b.hist + (size_t)(row0 + r) * b.hist_stride + t) : 0;
      }
      if (tid < 128) {
        int ia[4], ib[4];
        float nv[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int item = tid + 128 * it, xr = item >> 4, which = (item >> 3) & 1, sq = item & 7;
          ia[it] = ib[it] = -1; nv[it] = 0.f;
          if (xr < nrows) {
            const int row = row0 + xr;
            if (which == 0) {
              ia[it] = __ldg(b.movie_id + row);
              ib[it] = __ldg(b.user_id + row);
            } else {
              ia[it] = __ldg(b.user_genre + row * 5);
              ib[it] = __ldg(b.movie_genre + row * 3);
              if (sq < kNumNumerics) nv[it] = __ldg(b.numerics + row * kNumNumerics + sq);
            }
          }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int item = tid + 128 * it, xr = item >> 4, which = (item >> 3) & 1, sq = item & 7;
          int id_a = -1, id_b = -1;                              // -1: zero vector
          if (xr < nrows) {
            if (which == 0) {
              id_a = checked_id(rt_f32_roundtrip_id(ia[it]), p.n_movies, b.err_flag);
              id_b = checked_id(ib[it], p.n_users, b.err_flag);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.user + (size_t)id_b * 32 + 4 * sq));
            } else {
              id_a = ia[it]; id_b = ib[it];
              if (id_a >= p.n_genres) { atomicExch(b.err_flag, 1); id_a = -1; }
              if (id_b >= p.n_genres) { atomicExch(b.err_flag, 1); id_b = -1; }
              if (id_a < 0) id_a = -1;
              if (id_b < 0) id_b = -1;
            }
          }
          if (which == 0) {
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (id_a >= 0) c4 = ldg4(p.movie + (size_t)id_a * 32 + 4 * sq);
            *reinterpret_cast<float4*>(cand + xr * 32 + 4 * sq) = c4;
          } else {
            nums[xr * 8 + sq] = nv[it];
          }
          if (sq == 0) { sid[xr * 4 + 2 * which] = id_a; sid[xr * 4 + 2 * which + 1] = id_b; }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        ids_s[tid + u * kHThreads] = checked_id(rt_f32_roundtrip_id(hraw[u]), p.n_movies, b.err_flag);
      // tile rows of positions >= T are read by both MMAs: keep them zero (phase 2 of the previous
      // group used the ring for the X operand and the weight pieces)
      const int pad = 64 - T;                                     // item -> (pad row i >> 3, chunk i & 7)
      for (int i = tid; i < pad * 8; i += kHThreads) {
#pragma unroll
        for (int sr = 0; sr < kHSlots * 2; ++sr)                  // sr = slot * 2 + row
          *reinterpret_cast<uint4*>(ring + (sr >> 1) * HS_SLOT + ((sr & 1) * 64 + T + (i >> 3)) * 128 +
                                    ((i & 7) << 4)) = make_uint4(0, 0, 0, 0);
      }
    }
    fence_async_smem();                                     // the zeroed pad rows are MMA operand bytes
    __syncthreads();                                        // history ids, candidate rows staged
    RTH_TRACE(2, tid == 0);

    // ================= phase 1: tiles ====================================================
    if (is_gather) {
#pragma unroll
      for (int a = 0; a < kHAhead; ++a) {
        if (a < n_tiles) gather(a);
        cp_async_commit();
      }
      for (int k = 0; k < n_tiles; ++k) {
        const int slot = (kbase + k) % kHSlots;
        cp_async_wait<kHAhead - 1>();                       // this tile's rows have landed
        fence_async_smem();
        mbar_arrive(&full[slot]);
        if (k == 0) RTH_TRACE(20, tid == 0);
        if (k == 2) RTH_TRACE(21, tid == 0);
        if (k == 4) RTH_TRACE(22, tid == 0);
        if (k + kHAhead < n_tiles) gather(k + kHAhead);     // its slot frees when tile K + kHAhead - kHSlots retires
        cp_async_commit();
      }
      cp_async_wait<0>();
    } else if (is_builder) {
      // ---- B operand of every tile: W_r = (Wsub+Wh) + diag(c_r) Wp, bf16 hi / lo; lane = unit
      auto build_b = [&](int k) {                           // the slot of tile k is free (caller waited)
        const int slot = (kbase + k) % kHSlots;
        uint8_t* Bt = ring + slot * HS_SLOT + HS_A;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int cq = 0; cq < 4; ++cq) {
            const float* cv = cand + (2 * k + r) * 32 + 8 * cq;
            const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
            const float* wa = rc + 16 * cq;
            const float* wp = rc + 16 * cq + 8;
            const float2 v0 = fma2(make_float2(c0.x, c0.y), make_float2(wp[0], wp[1]), make_float2(wa[0], wa[1]));
            const float2 v1 = fma2(make_float2(c0.z, c0.w), make_float2(wp[2], wp[3]), make_float2(wa[2], wa[3]));
            const float2 v2 = fma2(make_float2(c1.x, c1.y), make_float2(wp[4], wp[5]), make_float2(wa[4], wa[5]));
            const float2 v3 = fma2(make_float2(c1.z, c1.w), make_float2(wp[6], wp[7]), make_float2(wa[6], wa[7]));
            const Split2 s0 = split_pack(v0.x, v0.y), s1 = split_pack(v1.x, v1.y);
            const Split2 s2 = split_pack(v2.x, v2.y), s3 = split_pack(v3.x, v3.y);
            const uint32_t n = r * 32 + lane;
            *reinterpret_cast<uint4*>(Bt + sw64_offset(n, cq)) = make_uint4(s0.hi, s1.hi, s2.hi, s3.hi);
            *reinterpret_cast<uint4*>(Bt + sw64_offset(64 + n, cq)) = make_uint4(s0.lo, s1.lo, s2.lo, s3.lo);
          }
        fence_async_smem();
        mbar_arrive(&full[slot]);
      };
      if (!builder_gathers) {
        for (int k = 0; k < n_tiles; ++k) {
          const int K = kbase + k, slot = K % kHSlots;
          if (K >= kHSlots) mbar_wait(&empty[slot], ((K / kHSlots) + 1) & 1);
          build_b(k);
        }
      } else {
        // this warp is also the third gatherer: per tile it requests its share of the rows (gather()
        // waits for the slot), builds B while they fly (first arrival on `full`), and arrives a
        // second time when its copies have landed - same schedule as warps 0-1
#pragma unroll
        for (int a = 0; a < kHAhead; ++a) {
          if (a < n_tiles) gather(a);
          cp_async_commit();
          if (a < n_tiles) build_b(a);
        }
        for (int k = 0; k < n_tiles; ++k) {
          const int slot = (kbase + k) % kHSlots;
          cp_async_wait<kHAhead - 1>();
          fence_async_smem();
          mbar_arrive(&full[slot]);
          if (k + kHAhead < n_tiles) gather(k + kHAhead);
          cp_async_commit();
          if (k + kHAhead < n_tiles) build_b(k + kHAhead);
        }
        cp_async_wait<0>();
      }
    } else if (is_issuer) {
      // ---- every MMA of the tile phase, in the order the operands become ready
      auto mma1 = [&](int k) {
        const int K = kbase + k, slot = K % kHSlots;
        mbar_wait(&full[slot], (K / kHSlots) & 1);
        if (K >= 1) mbar_wait(&d1_free, (K - 1) & 1);       // the accumulators of tile K - 1 have been read
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD1 = tbase + HT_D1;
          const uint64_t ad = smem_desc_sw128(s_ring + slot * HS_SLOT);
          const uint64_t bd = smem_desc_sw64(s_ring + slot * HS_SLOT + HS_A);
          mma_ss(tD1, ad + 0, bd + 0, idesc_bf16(128, 128), 0);     // H_hi . [W_hi | W_lo]
          mma_ss(tD1, ad + 2, bd + 2, idesc_bf16(128, 128), 1);
          mma_ss(tD1, ad + 4, bd + 0, idesc_bf16(128, 64), 1);      // H_lo . W_hi
          mma_ss(tD1, ad + 6, bd + 2, idesc_bf16(128, 64), 1);
          mma_commit(&d1_full);
        }
        __syncwarp();
      };
      if (0 < n_tiles) mma1(0);
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, slot = K % kHSlots, u = K & 1;
        if (k + 1 < n_tiles) mma1(k + 1);                   // overlaps the gate arithmetic of tile K
        mbar_wait(&w_ready[u], (K >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD2 = tbase + HT_D2 + 16u * u;
          const uint32_t s_b2 = smem_u32(b2s) + u * 2048;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ss(tD2 + 8 * r, smem_desc_mn_sw128(s_ring + slot * HS_SLOT + r * 8192 + ks * 2048),
                     smem_desc_sw128(s_b2 + r * 1024) + 2 * ks, idesc_mn(64, 8, 1), ks > 0);
          mma_commit(&d2_full[u]);
          mma_commit(&empty[slot]);
        }
        __syncwarp();
      }
    } else if (is_consumer) {
      const int m = 32 * warp_w + lane;                     // D1 row = TMEM lane of this thread
      const int r_t = m >> 6, t = m & 63;                   // its tile row and position
      // cst[row][j] = au_b[j] + sum_e cand[row][e] (Wc - Wsub)[e][j]: 128 threads x 8 outputs
      {
        const int ct = tid - 128, cr = ct >> 2, j0 = (ct & 3) * 8;
        float4 a0 = ldg4(p.au_b + j0), a1 = ldg4(p.au_b + j0 + 4);
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
          const float cv = cand[cr * 32 + e];
          const float4 w0 = ldg4(p.au_wc + e * 32 + j0), w1 = ldg4(p.au_wc + e * 32 + j0 + 4);
          a0.x = fmaf(cv, w0.x, a0.x); a0.y = fmaf(cv, w0.y, a0.y);
          a0.z = fmaf(cv, w0.z, a0.z); a0.w = fmaf(cv, w0.w, a0.w);
          a1.x = fmaf(cv, w1.x, a1.x); a1.y = fmaf(cv, w1.y, a1.y);
          a1.z = fmaf(cv, w1.z, a1.z); a1.w = fmaf(cv, w1.w, a1.w);
        }
        *reinterpret_cast<float4*>(cst + cr * 32 + j0) = a0;
        *reinterpret_cast<float4*>(cst + cr * 32 + j0 + 4) = a1;
      }
      named_sync(5, 128);
      // pooled accumulators of local tile k -> shared memory
      auto pool_out = [&](int k) {
        const int K = kbase + k, u = K & 1;
        mbar_wait(&d2_full[u], (K >> 1) & 1);
        tc_fence_after();
        // D2 row mm = 16 warp_w + lane (lane < 16): mm < 32 -> hi e = mm, else lo e = mm - 32;
        // columns 8 r + {0: . w_hi, 1: . w_lo}
        uint32_t d[16];
        tmem_ld16(tbase + HT_D2 + 16u * u + lane_base, d);
        tmem_ld_wait();
        if (lane < 16) {
          const int mm = 16 * warp_w + lane;
          const bool hi = warp_w < 2;
          pooled[(2 * k) * 64 + mm] = hi ? __uint_as_float(d[0]) + __uint_as_float(d[1]) : __uint_as_float(d[0]);
          pooled[(2 * k + 1) * 64 + mm] = hi ? __uint_as_float(d[8]) + __uint_as_float(d[9]) : __uint_as_float(d[8]);
        }
        tc_fence_before();
      };
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, u = K & 1;
        const uint32_t tD1 = tbase + HT_D1;
        mbar_wait(&d1_full, K & 1);
        tc_fence_after();
        if (k == 0) RTH_TRACE(3, tid == 128);
        if (k == 2) RTH_TRACE(10, tid == 128);
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
            if (half == 1) {                                // every value of the tile is in registers
              tc_fence_before();
              mbar_arrive(&d1_free);
            }
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
          uint8_t* dstw = b2s + u * 2048 + r_t * 1024 + (t & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(0, t >> 3)) = wh;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(1, t >> 3)) = wl;
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(&w_ready[u]);
        if (k == 2) RTH_TRACE(11, tid == 128);
        if (k >= 1) pool_out(k - 1);                        // the other D2 buffer: its MMAs finished long ago
        if (k == 2) RTH_TRACE(16, tid == 128);
      }
      if (n_tiles > 0) pool_out(n_tiles - 1);
      RTH_TRACE(4, tid == 128);
    }
    kbase += n_tiles;
    tc_fence_before();
    __syncthreads();                                        // every MMA of the group has completed: the ring is free
    RTH_TRACE(5, tid == 0);

    // ================= phase 2: top MLP on the group's 32 row slots, whole CTA ===============
    // warp 3 lane 0 streams the 8 weight pieces (W1 hi/lo per K block, then W2) through 3 buffers.
    // Pieces 0..2 go out now (their buffers were released by the previous group); pieces 3..7 wait
    // for layer-1 MMAs and are issued after the X operand is in place (below), before this warp
    // joins the layer-1 epilogue - every pfree they wait for is committed during layer 1.
    auto stream_pieces = [&](int i0, int i1) {
      for (int i = i0; i < i1; ++i) {
        const int P = pbase + i, j = P % 3;
        if (P >= 3) mbar_wait(&pfree[j], ((P / 3) + 1) & 1);
        const uint32_t src = i < 6 ? ((i & 1) ? HI_W1_LO : HI_W1_HI) + (uint32_t)(i >> 1) * H2_PIECE_BYTES
                                   : HI_W2 + (uint32_t)(i - 6) * H2_PIECE_BYTES;
        mbar_arrive_expect_tx(&pfull[j], H2_PIECE_BYTES);
        bulk_g2s(ring + H2_PIECE + j * H2_PIECE_BYTES, p.image + src, H2_PIECE_BYTES, &pfull[j]);
      }
    };
    if (is_builder) {
      if (lane == 0) stream_pieces(0, 3);
      __syncwarp();
    }
    if (tid < 128) {
      uint8_t* xb = ring + H2_XB;
      float4 fa[4], fb[4];                    // which 0: (candidate, user); 1: (userGenre1, movieGenre1)
#pragma unroll
      for (int it = 0; it < 4; ++it) {        // every load first: the user rows may still be on their way to L2
        const int item = tid + 128 * it, xr = item >> 4, which = (item >> 3) & 1, sq = item & 7;
        const int id_a = sid[xr * 4 + 2 * which], id_b = sid[xr * 4 + 2 * which + 1];
        fa[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        fb[it] = fa[it];
        if (which == 0) {
          fa[it] = *reinterpret_cast<const float4*>(cand + xr * 32 + 4 * sq);
          if (id_b >= 0) fb[it] = ldg4(p.user + (size_t)id_b * 32 + 4 * sq);
        } else {
          if (id_a >= 0) fa[it] = ldg4(p.ugenre + id_a * 32 + 4 * sq);
          if (id_b >= 0) fb[it] = ldg4(p.mgenre + id_b * 32 + 4 * sq);
        }
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int item = tid + 128 * it, xr = item >> 4, which = (item >> 3) & 1, sq = item & 7;
        if (which == 0) {
          rth_store_x4(xb, 0, xr, 32 + 4 * sq, fb[it]);          // K block 0: [userGenre1 | userId]
          rth_store_x4(xb, 1, xr, 32 + 4 * sq, fa[it]);          // K block 1: [pooled | candidate]
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
          rth_store_x4(xb, 0, xr, 4 * sq, fa[it]);
          rth_store_x4(xb, 1, xr, 4 * sq, pl);
          rth_store_x4(xb, 2, xr, 4 * sq, fb[it]);
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tT1 = tbase + HT_TOP1, tT2 = tbase + HT_TOP2;
    if (is_builder) {
      if (lane == 0) stream_pieces(3, 8);
      __syncwarp();
    }
    if (is_issuer) {
      // layer 1: pieces 0..5 = (K block kb, hi | lo); D1[128 units x (32 rows hi | 32 rows lo)]
      for (int i = 0; i < 6; ++i) {
        const int P = pbase + i, j = P % 3, kb = i >> 1;
        mbar_wait(&pfull[j], (P / 3) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wd = smem_desc_sw128(s_ring + H2_PIECE + j * H2_PIECE_BYTES);
          const uint64_t xd = smem_desc_sw128(s_ring + H2_XB + kb * 8192);      // [X hi | X lo], N = 64
          const int nks = kb == 2 ? 2 : 4;                                    // K block 2: columns 32..63 are zero
          for (int ks = 0; ks < nks; ++ks)
            mma_ss(tT1, wd + 2 * ks, xd + 2 * ks, idesc_top, (i > 0 || ks > 0) ? 1u : 0u);
          mma_commit(&pfree[j]);
          if (i == 5) mma_commit(&cbar);
        }
        __syncwarp();
      }
    }
    // layer-1 epilogue: this thread is unit `tw` for row slots 16 wg .. 16 wg + 15
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    RTH_TRACE(6, tid == 0);
    {
      uint32_t d[16], d2[16];
      tmem_ld16(tT1 + 16 * wg + lane_base, d);             // W1 . X hi
      tmem_ld16(tT1 + 32 + 16 * wg + lane_base, d2);       // W1 . X lo
      tmem_ld_wait();
      const uint32_t koff = (uint32_t)(tw >> 6) * 8192u;
      const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int sr = wg * 16 + r;
        const float4 n0 = *reinterpret_cast<const float4*>(nums + sr * 8);
        const float4 n1 = *reinterpret_cast<const float4*>(nums + sr * 8 + 4);
        float v = (__uint_as_float(d[r]) + __uint_as_float(d2[r])) + b1;
        v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
        v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
        v = fmaf(n1.z, w1n[6], v);
        v = v > 0.f ? v : a1 * v;
        const uint32_t off = koff + sw128_offset(sr, chunk) + within;
        const __nv_bfloat16 vh = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(ring + H2_H1 + off) = vh;
        *reinterpret_cast<__nv_bfloat16*>(ring + H2_H1 + off + 4096u) = __float2bfloat16_rn(v - __bfloat162float(vh));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (is_issuer) {
      // layer 2: pieces 6, 7 = W2 K blocks; D2[(64 hi | 64 lo units) x (32 rows hi | 32 rows lo)]
      for (int i = 6; i < 8; ++i) {
        const int P = pbase + i, j = P % 3, kb = i - 6;
        mbar_wait(&pfull[j], (P / 3) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wd = smem_desc_sw128(s_ring + H2_PIECE + j * H2_PIECE_BYTES);
          const uint64_t hs = smem_desc_sw128(s_ring + H2_H1 + kb * 8192);      // [H1 hi | H1 lo], N = 64
          for (int ks = 0; ks < 4; ++ks)
            mma_ss(tT2, wd + 2 * ks, hs + 2 * ks, idesc_top, (i > 6 || ks > 0) ? 1u : 0u);
          mma_commit(&pfree[j]);
          if (i == 7) mma_commit(&cbar);
        }
        __syncwarp();
      }
    }
    pbase += 8;
    const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    {
      uint32_t d[16], d2[16];
      tmem_ld16(tT2 + 16 * wg + lane_base, d);
      tmem_ld16(tT2 + 32 + 16 * wg + lane_base, d2);
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 16; ++r) d[r] = __float_as_uint(__uint_as_float(d[r]) + __uint_as_float(d2[r]));
      float* red = reinterpret_cast<float*>(xs + HX_IDS);      // [64 units][32 rows] (the ids are consumed)
      float* zp = reinterpret_cast<float*>(xs + HX_CAND);      // [8][32]            (so are the candidate rows)
      if (tw >= 64) {                                          // lo halves of W2 -> smem
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4)
          *reinterpret_cast<float4*>(red + (tw - 64) * 32 + 16 * wg + 4 * r4) =
              make_float4(__uint_as_float(d[4 * r4]), __uint_as_float(d[4 * r4 + 1]),
                          __uint_as_float(d[4 * r4 + 2]), __uint_as_float(d[4 * r4 + 3]));
      }
      __syncthreads();
      if (tw < 64) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float v = __uint_as_float(d[r]) + red[tw * 32 + 16 * wg + r] + b2;   // (W2hi + W2lo) . (H1hi + H1lo)
          v = v > 0.f ? v : a2 * v;
          red[tw * 32 + 16 * wg + r] = v * w3;
        }
      }
      __syncthreads();
      {  // 32 rows x 8 partial sums of 8 units
        const int r = tid & 31, pt = tid >> 5;
        float sum = 0.f;
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) sum += red[(pt * 8 + uu) * 32 + r];
        zp[pt * 32 + r] = sum;
      }
      __syncthreads();
      if (tid < kHRows) {
        float z = p.b3;
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) z += zp[pt * 32 + tid];
        if (tid < nrows) {
          store_score(b, row0 + tid, sigmoidf_acc(z));
          if (b.logits) b.logits[row0 + tid] = z;
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                     // ring and scratch are reused by the next group
    RTH_TRACE(7, tid == 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, HT_COLS);
  RTH_TRACE(8, tid == 0);
}

cudaError_t read_din_rth_trace(unsigned long long* out40) {
  return cudaMemcpyFromSymbol(out40, g_din_rth_trace, sizeof(unsigned long long) * 40);
}

size_t din_rth_smem_bytes() { return 1024 + HRING + HX_BYTES; }

cudaError_t launch_din_rth(const DinRtParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  DinRtParams q = p;
  // rows per group: as even as possible over the CTA slots (ctas_per_sm per SM), at most 32, even
  const int cps = p.ctas_per_sm == 2 ? 2 : 1;
  const int slots = p.num_sms * cps;
  const int waves = (b.B + kHRows * slots - 1) / (kHRows * slots);
  int rpg = (b.B + waves * slots - 1) / (waves * slots);
  rpg = (rpg + 1) & ~1;
  if (rpg > kHRows) rpg = kHRows;
  if (rpg < 2) rpg = 2;
  q.rows_per_group = rpg;
  const int n_groups = (b.B + rpg - 1) / rpg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_groups < slots ? n_groups : slots);
  cfg.blockDim = dim3(kHThreads);
  cfg.dynamicSmemBytes = din_rth_smem_bytes();
  cfg.stream = s;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, din_rth_kernel, q, b);
}

cudaError_t setup_din_rth_attributes() {
  cudaError_t e = cudaFuncSetAttribute(din_rth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)din_rth_smem_bytes());
  if (e != cudaSuccess) return e;
  // two CTAs per SM need the whole shared-memory carve-out (2 x 103 KB), not the smallest one
  // that fits a single CTA
  return cudaFuncSetAttribute(din_rth_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

}  // namespace srs
