// din_rt64.cu - the row-tile DIN kernel (din_rt.cu) for wide embeddings and long histories:
// 32 < E <= 64, T <= 256 (BASELINE cfg 5: E = 64, T = 200, 10^8 movies = 25.6 GB of table in HBM).
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167.
// Same math and the same operand tricks as din_rt.cu; what changes with E = 64:
//   * a pre-split table row is 256 bytes [64 x bf16 hi | 64 x bf16 lo]; a tile is 128 positions of
//     ONE batch row ("chunk" c = positions 128 c ..), stored as two K blocks (hi, lo) of 16 KB;
//   * W_r = (Wsub+Wh) + diag(c_r) Wp is 64 x 32: B = [W_r hi (32 units) | W_r lo] x 64 k, 8 KB;
//     D[128 x 64] = H_hi . [W_hi | W_lo] (4 MMAs, N = 64) + H_lo . W_hi (4 MMAs, N = 32);
//   * pooling: two M = 64 MMAs per 16 positions (hi block^T, lo block^T) against [w_hi | w_lo];
//   * the history ids do not fit in shared memory (32 x 200 ints): the gather warps read them
//     straight from global memory, two tiles ahead, into registers;
//   * consumer q owns chunk q of every row, so the P/Q gate tables of its positions stay in
//     registers; the pooled halves of a row's two chunks meet in shared memory;
//   * top MLP: K = 5 x 64 (one K block per feature); W1 hi / lo (160 KB) land in the ring as its
//     slots retire, W2 follows into the space of W1 hi after layer 1.
// Warp roles as in din_rt.cu (warps 0-3, 7 gather; 4 issues MMAs; 5-6 build W_r and stream the
// weight images; 8-15 two consumers).  Ring: 4 slots of 40 KB.
#include <climits>

#include "rt_common.cuh"

namespace srs {

// ---- debugging aid (-DRT64_WATCHDOG, profiles/exp/rt64_hang_probe.py): every mbarrier wait of this file gives up
// after ~2^22 polls, records {source line, block, thread, parity} and lets the launch run to its end (with
// invalid scores); srs_model_status() then reports the records instead of the process hanging.
__device__ unsigned int g_rt64_abort;
__device__ unsigned int g_rt64_nrec;
__device__ unsigned long long g_rt64_rec[64];
#ifdef RT64_WATCHDOG
__device__ __noinline__ void rt64_record(int line, uint32_t parity) {
  const unsigned int m = __activemask();
  if ((threadIdx.x & 31) != __ffs(m) - 1) return;        // one record per waiting warp
  atomicExch(&g_rt64_abort, 1u);
  const unsigned int i = atomicAdd(&g_rt64_nrec, 1u);
  if (i < 16u) {
    g_rt64_rec[4 * i] = (unsigned long long)line;
    g_rt64_rec[4 * i + 1] = (unsigned long long)blockIdx.x;
    g_rt64_rec[4 * i + 2] = (unsigned long long)threadIdx.x | ((unsigned long long)m << 32);
    g_rt64_rec[4 * i + 3] = (unsigned long long)parity;
  }
  __threadfence();
}
__device__ __forceinline__ void rt64_wait_dbg(uint64_t* bar, uint32_t parity, int line) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const bool aborted = *reinterpret_cast<volatile unsigned int*>(&g_rt64_abort) != 0;
      if (spins >= (1u << 22) || (aborted && spins >= (1u << 20))) { rt64_record(line, parity); return; }
      if (aborted) return;
    }
  }
}
#define mbar_wait(bar, par) rt64_wait_dbg(bar, par, __LINE__)
#elif defined(RT64_SLEEP_NS)
__device__ __forceinline__ void rt64_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(RT64_SLEEP_NS);
}
#define mbar_wait(bar, par) rt64_wait_sleep(bar, par)
#endif

constexpr int k64Threads = 512;
constexpr int k64Rows = 32;                 // row slots per group = N/2 of the top-MLP MMAs
constexpr int k64Slots = 4;
constexpr int k64Ahead = 2;                 // tiles in flight ahead of the one being delivered
constexpr int k64GatherThreads = 160;       // warps 0-3 and 7
constexpr int k64Copies = 13;               // ceil(128 positions * 16 chunks / 160)

// ring slot: A hi [128 positions][64 bf16] SW128 (16 KB) | A lo (16 KB) | B [32 hi | 32 lo units][64 k] SW128 (8 KB)
constexpr uint32_t S64_ALO = 16384, S64_B = 32768, S64_SLOT = 40960;
constexpr uint32_t RING64 = k64Slots * S64_SLOT;     // 163840
// weight image in global memory: W1 hi (5 K blocks x 16 KB) | W1 lo (5 x 16 KB) | W2 (2 x 16 KB)
constexpr uint32_t I64_W1 = 163840, I64_W2 = 32768;
constexpr uint32_t R64_W1HI = 0, R64_W1LO = 81920, R64_W2 = 0;
// behind the ring: layer-2 scratch, then the scratch that the X operand overlays in phase 2
constexpr uint32_t Q64_RED = 0;                      // f32 [64][32]
constexpr uint32_t Q64_ZP = 8192;                    // f32 [16][32]
constexpr uint32_t Q64_X = 10240;                    // phase 2: 5 K blocks x [32 rows hi | 32 rows lo][64 k] (40 KB)
constexpr uint32_t Q64_CAND = Q64_X;                 // f32 [32][64]                       (phase 0/1)
constexpr uint32_t Q64_CST = Q64_X + 8192;           // f32 [32][32]
constexpr uint32_t Q64_POOL = Q64_X + 12288;         // f32 [32][2 chunks][64]
constexpr uint32_t Q64_B2 = Q64_X + 28672;           // [consumer][buffer] x 2 K blocks x [8 n][64 positions] bf16
constexpr uint32_t Q64_NUMS = Q64_X + 40960;         // f32 [32][8]
constexpr uint32_t Q64_BYTES = Q64_NUMS + 1024;
// tensor memory columns
constexpr uint32_t T64_D1 = 0;                       // tile K: [64 (K % 3), + 64)
constexpr uint32_t T64_D2 = 192;                     // consumer q, buffer u: 192 + 32 q + 16 u; hi part +0, lo part +8
constexpr uint32_t T64_TOP1 = 0, T64_TOP2 = 64;      // phase 2

__device__ __forceinline__ void st64_x4(uint8_t* tile, int block, int row, int col, float4 v) {
  const uint32_t off = block * 8192u + sw128_offset(row, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(tile + off + 4096u) = make_uint2(s0.lo, s1.lo);   // row + 32: same swizzle phase
}

struct Rt64GroupLoads {
  int cid, uid, ug, mg;            // raw ids of row slot tid >> 4
  float nv;                        // numeric (tid & 15) of the row
};

__global__ void __launch_bounds__(k64Threads, 1) din_rt64_kernel(const __grid_constant__ DinRtParams p,
                                                                 BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t wbar;                 // W1 images landed (once per group)
  __shared__ uint64_t w2bar;                // W2 image landed (once per group)
  __shared__ uint64_t cbar;                 // top-MLP MMAs complete
  __shared__ uint64_t full[k64Slots];       // tile operands in place (160 gatherer + 64 builder arrivals)
  __shared__ uint64_t empty[k64Slots];      // both MMA groups of the tile in the slot have completed
  __shared__ uint64_t d1_full[3];           // tile K: gate accumulators in buffer K % 3 ready
  __shared__ uint64_t w_ready[2];           // consumer q: pooling weights written (128 arrivals)
#ifdef SRS_WREADY_SPLIT
  // One barrier per (consumer, pooled buffer).  The four warps of a consumer are not synchronised with each
  // other: a warp that finds the accumulators of its next tile ready can arrive for tile K + 2 while a sibling is
  // still gating tile K, and with ONE barrier per consumer that arrival completes tile K's phase for the sibling
  // (profiles/exp/rt_protocol_sim.py --single-wready: mixed phases, then parity aliasing on d1_full and a
  // deadlock).  With the split, the early arrival lands on the other buffer's barrier.
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
  const int lane = tid & 31;
  const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);          // warp-uniform by construction
  const int warp_w = __shfl_sync(0xffffffffu, (tid >> 5) & 3, 0);
  const int tw = tid & 127;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  uint8_t* xs = ring + RING64;
  float* cand = reinterpret_cast<float*>(xs + Q64_CAND);
  float* cst = reinterpret_cast<float*>(xs + Q64_CST);
  float* pooled = reinterpret_cast<float*>(xs + Q64_POOL);
  uint8_t* b2s = xs + Q64_B2;
  float* nums = reinterpret_cast<float*>(xs + Q64_NUMS);
  const int T = p.T;
  const int NCH = p.nch;                                 // 128-position chunks per row: 1 or 2
  const int RPG = p.rows_per_group;
  const int n_groups = (b.B + RPG - 1) / RPG;
  const bool is_gather = wg == 0 || (wg == 1 && warp_w == 3), is_issuer = wg == 1 && warp_w == 0;
  const bool is_builder = wg == 1 && (warp_w == 1 || warp_w == 2), is_consumer = wg >= 2;
  // phase-0 / phase-2 role: row slot and float4 index of every 64-float feature row
  const int xr = tid >> 4, sq = tid & 15;

  auto issue_group_loads = [&](int g) -> Rt64GroupLoads {
    Rt64GroupLoads L;
    L.cid = L.uid = -1; L.ug = L.mg = -1; L.nv = 0.f;
    if (g >= n_groups) return L;
    const int row0 = g * RPG;
    const int nrows = min(RPG, b.B - row0);
    if (xr < nrows) {
      const int row = row0 + xr;
      L.cid = __ldg(b.movie_id + row);
      L.uid = __ldg(b.user_id + row);
      L.ug = __ldg(b.user_genre + row * 5);
      L.mg = __ldg(b.movie_genre + row * 3);
      if (sq < kNumNumerics) L.nv = __ldg(b.numerics + row * kNumNumerics + sq);
    }
    return L;
  };

  // ---- prologue -----------------------------------------------------------------------------
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid < 32) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&w2bar, 1);
    mbar_init(&cbar, 1);
    for (int i = 0; i < k64Slots; ++i) { mbar_init(&full[i], k64GatherThreads + 64); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 3; ++i) mbar_init(&d1_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_ready[i], 128);
#ifdef SRS_WREADY_SPLIT
      mbar_init(&w_ready_u1[i], 128);
#endif
      mbar_init(&d2_full[i][0], 1); mbar_init(&d2_full[i][1], 1);
    }
    fence_mbar_init();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");    // inputs may come from the previous kernel
  Rt64GroupLoads pre = issue_group_loads(blockIdx.x);
  // per-thread constants of the roles
  //   builders : unit j = bt >> 1, e half = bt & 1; rc[16 c + 0..7] = (Wsub+Wh)[8 (4 half + c) ..][j],
  //              rc[16 c + 8..15] = Wp[..][j], c = 0..3
  //   consumers: rc[0..31] = P_t[j], rc[32..63] = Q_t[j] of position t = 128 chunk + tw, chunk = consumer
  float rc[64];
  if (is_consumer) {
    const int chunk = NCH == 2 ? wg - 2 : 0;
    const float* src = p.pq + (size_t)min(128 * chunk + tw, T - 1) * 64;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = ldg4(src + 4 * i);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
  } else if (is_builder) {
    const int bt = tid & 63, j = bt >> 1, half = bt & 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {                                    // index 4 i = 16 c + 8 part + 4 h
      const int c = i >> 2, part = (i >> 1) & 1, h = i & 1;
      const float4 v = ldg4((part ? p.wpT : p.waT) + j * 64 + 8 * (4 * half + c) + 4 * h);
      rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) rc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t s_ring = smem_u32(ring);
  const uint32_t idesc_top = idesc_bf16(128, 2 * k64Rows);
  uint32_t cphase = 0, wphase = 0;
  int kbase = 0;                                        // tiles of earlier groups of this CTA (even when NCH == 2)

  // gatherer constants: copy n moves 16-byte chunk c16 of position cell0 + 10 n
  const int gt = tid < 128 ? tid : tid - 96;            // warp 7 (tid 224..255) -> 128..159
  const int c16 = gt & 15, cell0 = gt >> 4;
  const uint32_t g_part = (uint32_t)(c16 >> 3) * S64_ALO;

  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int row0 = g * RPG;
    const int nrows = min(RPG, b.B - row0);
    const int n_tiles = nrows * NCH;

    // ================= phase 0: side rows, candidate rows, zero padding ======================
    float4 f_cand = make_float4(0.f, 0.f, 0.f, 0.f), f_user = f_cand, f_ug = f_cand, f_mg = f_cand;
    if (xr < nrows) {
      const int cid = checked_id(rt_f32_roundtrip_id(pre.cid), p.n_movies, b.err_flag);
      const int uid = checked_id(pre.uid, p.n_users, b.err_flag);
      int ug = pre.ug, mg = pre.mg;
      if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
      if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
      f_cand = ldg4(p.movie + (size_t)cid * 64 + 4 * sq);
      f_user = ldg4(p.user + (size_t)uid * 64 + 4 * sq);
      if (ug >= 0) f_ug = ldg4(p.ugenre + ug * 64 + 4 * sq);
      if (mg >= 0) f_mg = ldg4(p.mgenre + mg * 64 + 4 * sq);
    }
    {
      // tile rows of positions >= T are read by both MMA groups: keep them zero (phase 2 of the
      // previous group used the ring for the weight images); slot s holds chunk s & 1 when NCH == 2
      const int zr = tid >> 3, zc = (tid & 7) << 4;
#pragma unroll
      for (int sl = 0; sl < k64Slots; ++sl) {
        const int chunk = NCH == 2 ? (sl & 1) : 0;
        const int valid = min(128, T - 128 * chunk);
        for (int r = valid + zr; r < 128; r += 64) {
          *reinterpret_cast<uint4*>(ring + sl * S64_SLOT + r * 128 + zc) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(ring + sl * S64_SLOT + S64_ALO + r * 128 + zc) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    *reinterpret_cast<float4*>(cand + xr * 64 + 4 * sq) = f_cand;
    if (sq < 8) nums[xr * 8 + sq] = pre.nv;
    __syncthreads();                                        // candidate rows staged, padding zeroed

    // ================= phase 1: tiles ====================================================
    if (is_gather) {
      // raw ids of local tile k for this thread's 13 copies
      auto load_ids = [&](int k, int (&ids)[k64Copies]) {
        const int row = NCH == 2 ? (k >> 1) : k, chunk = NCH == 2 ? (k & 1) : 0;
        const bool live = k < n_tiles;
        const int* h = b.hist + (size_t)(row0 + row) * b.hist_stride + 128 * chunk;
        const int valid = live ? min(128, T - 128 * chunk) : 0;
#pragma unroll
        for (int n = 0; n < k64Copies; ++n) {
          const int cell = cell0 + 10 * n;
          ids[n] = cell < valid ? __ldg(h + cell) : 0;       // liveness is decided from the cell index, never from the id value
        }
      };
      auto gather = [&](int k, const int (&ids)[k64Copies]) {
        const int K = kbase + k, slot = K % k64Slots;
        if (K >= k64Slots) mbar_wait(&empty[slot], ((K / k64Slots) + 1) & 1);
        uint8_t* A = ring + slot * S64_SLOT + g_part;
        const int valid = min(128, T - 128 * (NCH == 2 ? (k & 1) : 0));     // gather() is only called for live tiles
#pragma unroll
        for (int n = 0; n < k64Copies; ++n) {
          const int cell = cell0 + 10 * n;
          if (cell < valid) {
            const int id = checked_id(rt_f32_roundtrip_id(ids[n]), p.n_movies, b.err_flag);
            cp_async16(A + cell * 128 + (((c16 & 7) ^ (cell & 7)) << 4), p.movie_split + (size_t)id * 256 + c16 * 16);
          }
        }
      };
      int ida[k64Copies], idb[k64Copies];                   // ids of the next two tiles to gather
      load_ids(0, ida);
      load_ids(1, idb);
#pragma unroll
      for (int a = 0; a < k64Ahead; ++a) {                  // k64Ahead == 2: tiles 0 and 1
        if (a == 0) { if (0 < n_tiles) gather(0, ida); load_ids(2, ida); }
        else { if (1 < n_tiles) gather(1, idb); load_ids(3, idb); }
        cp_async_commit();
      }
      for (int k = 0; k < n_tiles; k += 2) {                // two tiles per iteration: register sets alternate
        cp_async_wait<k64Ahead - 1>();
        fence_async_smem();
        mbar_arrive(&full[(kbase + k) % k64Slots]);
        if (k + 2 < n_tiles) gather(k + 2, ida);
        load_ids(k + 4, ida);
        cp_async_commit();
        if (k + 1 < n_tiles) {
          cp_async_wait<k64Ahead - 1>();
          fence_async_smem();
          mbar_arrive(&full[(kbase + k + 1) % k64Slots]);
        }
        if (k + 3 < n_tiles) gather(k + 3, idb);
        load_ids(k + 5, idb);
        cp_async_commit();
      }
      cp_async_wait<0>();
    } else if (is_builder) {
      // ---- B operand of every tile: W_r = (Wsub+Wh) + diag(c_r) Wp, bf16 hi / lo
      const int bt = tid & 63, pj = bt >> 1, half = bt & 1;
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, slot = K % k64Slots;
        const int row = NCH == 2 ? (k >> 1) : k;
        if (K >= k64Slots) mbar_wait(&empty[slot], ((K / k64Slots) + 1) & 1);
        uint8_t* Bt = ring + slot * S64_SLOT + S64_B;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cq = 4 * half + c;
          const float* cv = cand + row * 64 + 8 * cq;
          const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
          const float* wa = rc + 16 * c;
          const float* wp = rc + 16 * c + 8;
          const float2 v0 = fma2(make_float2(c0.x, c0.y), make_float2(wp[0], wp[1]), make_float2(wa[0], wa[1]));
          const float2 v1 = fma2(make_float2(c0.z, c0.w), make_float2(wp[2], wp[3]), make_float2(wa[2], wa[3]));
          const float2 v2 = fma2(make_float2(c1.x, c1.y), make_float2(wp[4], wp[5]), make_float2(wa[4], wa[5]));
          const float2 v3 = fma2(make_float2(c1.z, c1.w), make_float2(wp[6], wp[7]), make_float2(wa[6], wa[7]));
          const Split2 s0 = split_pack(v0.x, v0.y), s1 = split_pack(v1.x, v1.y);
          const Split2 s2 = split_pack(v2.x, v2.y), s3 = split_pack(v3.x, v3.y);
          *reinterpret_cast<uint4*>(Bt + sw128_offset(pj, cq)) = make_uint4(s0.hi, s1.hi, s2.hi, s3.hi);
          *reinterpret_cast<uint4*>(Bt + sw128_offset(32 + pj, cq)) = make_uint4(s0.lo, s1.lo, s2.lo, s3.lo);
        }
        fence_async_smem();
        mbar_arrive(&full[slot]);
      }
      // ---- W1 images: each ring slot receives its 40 KB as soon as its last tile retires
      if (warp_w == 1) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&wbar, I64_W1);
          const int tail = n_tiles < k64Slots ? n_tiles : k64Slots;
          auto load_slot = [&](int slot) {
            const uint32_t off = slot * S64_SLOT;
            bulk_g2s(ring + off, p.image + off, 32768u, &wbar);
            bulk_g2s(ring + off + 32768u, p.image + off + 32768u, S64_SLOT - 32768u, &wbar);
          };
          for (int sl = 0; sl < k64Slots; ++sl) {
            bool used = false;
            for (int j = 0; j < tail; ++j) used |= ((kbase + n_tiles - tail + j) % k64Slots) == sl;
            if (!used) load_slot(sl);
          }
          for (int j = 0; j < tail; ++j) {
            const int K = kbase + n_tiles - tail + j, slot = K % k64Slots;
            mbar_wait(&empty[slot], (K / k64Slots) & 1);
            load_slot(slot);
          }
        }
        __syncwarp();
      }
    } else if (is_issuer) {
      auto mma1 = [&](int k) {
        const int K = kbase + k, slot = K % k64Slots, db = K % 3;
        mbar_wait(&full[slot], (K / k64Slots) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD1 = tbase + T64_D1 + 64u * db;
          const uint64_t ah = smem_desc_sw128(s_ring + slot * S64_SLOT);
          const uint64_t al = smem_desc_sw128(s_ring + slot * S64_SLOT + S64_ALO);
          const uint64_t bd = smem_desc_sw128(s_ring + slot * S64_SLOT + S64_B);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss(tD1, ah + 2 * ks, bd + 2 * ks, idesc_bf16(128, 64), ks > 0);   // H_hi . [W_hi | W_lo]
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss(tD1, al + 2 * ks, bd + 2 * ks, idesc_bf16(128, 32), 1);        // H_lo . W_hi
          mma_commit(&d1_full[db]);
        }
        __syncwarp();
      };
      if (0 < n_tiles) mma1(0);
      if (1 < n_tiles) mma1(1);
      if (2 < n_tiles) mma1(2);
      for (int k = 0; k < n_tiles; ++k) {
        const int K = kbase + k, slot = K % k64Slots, q = K & 1, u = (K >> 1) & 1;
        mbar_wait(W_READY(q, u), W_READY_PAR(K));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD2 = tbase + T64_D2 + 32u * q + 16u * u;
          const uint32_t s_b2 = smem_u32(b2s) + (q * 2 + u) * 2048;
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_ss(tD2 + 8 * part, smem_desc_mn_sw128(s_ring + slot * S64_SLOT + part * S64_ALO + ks * 2048),
                     smem_desc_sw128(s_b2 + (ks >> 2) * 1024) + 2 * (ks & 3), idesc_mn(64, 8, 1), ks > 0);
          mma_commit(&d2_full[q][u]);
          mma_commit(&empty[slot]);
        }
        __syncwarp();
        if (k + 3 < n_tiles) mma1(k + 3);                   // accumulator buffer K % 3 was read before w_ready
      }
    } else if (is_consumer) {
      const int q = wg - 2;
      // cst[xr][j] = au_b[j] + sum_e cand[xr][e] (Wc - Wsub)[e][j]: 256 threads x 4 outputs
      {
        const int ct = tid - 256, cr = ct >> 3, j0 = (ct & 7) * 4;
        float4 acc = ldg4(p.au_b + j0);
#pragma unroll 8
        for (int e = 0; e < 64; ++e) {
          const float cv = cand[cr * 64 + e];
          const float4 w4 = ldg4(p.au_wc + e * 32 + j0);
          acc.x = fmaf(cv, w4.x, acc.x); acc.y = fmaf(cv, w4.y, acc.y);
          acc.z = fmaf(cv, w4.z, acc.z); acc.w = fmaf(cv, w4.w, acc.w);
        }
        *reinterpret_cast<float4*>(cst + cr * 32 + j0) = acc;
      }
      named_sync(5, 256);
      auto pool_out = [&](int k) {
        const int K = kbase + k, u = (K >> 1) & 1;
        mbar_wait(&d2_full[q][u], (K >> 2) & 1);
        tc_fence_after();
        // D2 row m = e = 16 warp_w + lane (lane < 16); columns 0: hi . w_hi, 1: hi . w_lo, 8: lo . w_hi
        uint32_t d[16];
        tmem_ld16(tbase + T64_D2 + 32u * q + 16u * u + lane_base, d);
        tmem_ld_wait();
        if (lane < 16)
          pooled[k * 64 + 16 * warp_w + lane] =
              (__uint_as_float(d[0]) + __uint_as_float(d[1])) + __uint_as_float(d[8]);
        tc_fence_before();
      };
      const int first = (q - kbase) & 1;                    // local tiles k with (kbase + k) & 1 == q
      for (int k = first; k < n_tiles; k += 2) {
        const int K = kbase + k, u = (K >> 1) & 1;
        const int row = NCH == 2 ? (k >> 1) : k, chunk = NCH == 2 ? (k & 1) : 0;
        const uint32_t tD1 = tbase + T64_D1 + 64u * (K % 3);
        mbar_wait(&d1_full[K % 3], (K / 3) & 1);
        tc_fence_after();
        // ---- gate: v = D_hi + D_lo + cst; s = sum_j v_j P_tj + |v_j| Q_tj
        float2 sa = make_float2(p.au_bout, 0.f), sb = make_float2(0.f, 0.f);
        {
          const float* cs = cst + row * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t dh[16], dl[16];
            tmem_ld16(tD1 + 16 * half + lane_base, dh);
            tmem_ld16(tD1 + 32 + 16 * half + lane_base, dl);
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
        const float w = (128 * chunk + tw < T) ? 1.f / (1.f + __expf(-s)) : 0.f;
        {
          // pooling weights operand: K block tw >> 6, row 0 = w hi, row 1 = w lo, column = position
          const __nv_bfloat16 wh = __float2bfloat16_rn(w);
          const __nv_bfloat16 wl = __float2bfloat16_rn(w - __bfloat162float(wh));
          uint8_t* dstw = b2s + (q * 2 + u) * 2048 + (tw >> 6) * 1024 + (tw & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(0, (tw & 63) >> 3)) = wh;
          *reinterpret_cast<__nv_bfloat16*>(dstw + sw128_offset(1, (tw & 63) >> 3)) = wl;
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(W_READY(q, u));
        if (k - 2 >= 0) pool_out(k - 2);
      }
      {
        const int last = first + ((n_tiles - 1 - first) & ~1);
        if (first < n_tiles) pool_out(last);
      }
    }
    kbase += n_tiles;
    tc_fence_before();
    __syncthreads();
    pre = issue_group_loads(g + gridDim.x);               // next group's ids: request from HBM now

    // ================= phase 2: top MLP on the group's 32 row slots, whole CTA ===============
    float4 f_pool = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xr < nrows) {
      f_pool = *reinterpret_cast<const float4*>(pooled + (xr * NCH) * 64 + 4 * sq);
      if (NCH == 2) {
        const float4 p1 = *reinterpret_cast<const float4*>(pooled + (xr * 2 + 1) * 64 + 4 * sq);
        f_pool = make_float4(f_pool.x + p1.x, f_pool.y + p1.y, f_pool.z + p1.z, f_pool.w + p1.w);
      }
    }
    __syncthreads();                                        // pooled / candidate scratch is overlaid by the X operand
    {
      uint8_t* xb = xs + Q64_X;                             // K blocks: userGenre1, userId, pooled, candidate, movieGenre1
      st64_x4(xb, 0, xr, 4 * sq, f_ug);
      st64_x4(xb, 1, xr, 4 * sq, f_user);
      st64_x4(xb, 2, xr, 4 * sq, f_pool);
      st64_x4(xb, 3, xr, 4 * sq, f_cand);
      st64_x4(xb, 4, xr, 4 * sq, f_mg);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tT1 = tbase + T64_TOP1, tT2 = tbase + T64_TOP2;
    const uint32_t s_x = smem_u32(xs + Q64_X);
    if (wg == 0 && warp_w == 0) {
      mbar_wait(&wbar, wphase);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {
          const uint64_t ah = smem_desc_sw128(s_ring + R64_W1HI + kb * 16384);
          const uint64_t al = smem_desc_sw128(s_ring + R64_W1LO + kb * 16384);
          const uint64_t xd = smem_desc_sw128(s_x + kb * 8192);                // [X hi | X lo], N = 64
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ss(tT1, ah + 2 * ks, xd + 2 * ks, idesc_top, acc);             // W1hi.(Xhi | Xlo)
            acc = 1;
            mma_ss(tT1, al + 2 * ks, xd + 2 * ks, idesc_top, 1);               // W1lo.(Xhi | Xlo)
          }
        }
        mma_commit(&cbar);
      }
      __syncwarp();
    }
    // this thread is unit `tw` of layer 1 for row slots 8*wg .. 8*wg+7
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    mbar_wait(&cbar, cphase);
    cphase ^= 1;
    __syncwarp();
    tc_fence_after();
    if (tid == 0) {     // layer-1 MMAs are done with the W1 hi images: W2 takes their place
      mbar_arrive_expect_tx(&w2bar, I64_W2);
      bulk_g2s(ring + R64_W2, p.image + I64_W1, I64_W2, &w2bar);
    }
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
        *reinterpret_cast<__nv_bfloat16*>(xs + Q64_X + off) = vh;        // H1 operand overlays X blocks 0, 1
        *reinterpret_cast<__nv_bfloat16*>(xs + Q64_X + off + 4096u) = __float2bfloat16_rn(v - __bfloat162float(vh));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (wg == 0 && warp_w == 0) {
      mbar_wait(&w2bar, wphase);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a = smem_desc_sw128(s_ring + R64_W2 + kb * 16384);
          const uint64_t hs = smem_desc_sw128(s_x + kb * 8192);               // [H1 hi | H1 lo], N = 64
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ss(tT2, a + 2 * ks, hs + 2 * ks, idesc_top, acc);             // (W2hi ; W2lo).(H1hi | H1lo)
            acc = 1;
          }
        }
        mma_commit(&cbar);
      }
      __syncwarp();
    }
    wphase ^= 1;
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
      float* red = reinterpret_cast<float*>(xs + Q64_RED);     // [64 units][32 rows]
      float* zp = reinterpret_cast<float*>(xs + Q64_ZP);       // [16][32]
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
      if (tid < k64Rows) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_slot, 512);
}

// fp32 table [rows][64] -> [rows][64 x bf16 hi | 64 x bf16 lo]
__global__ void split_table64_kernel(const float* __restrict__ src, uint32_t* __restrict__ dst, int64_t n_pairs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // pair index: row * 32 + pair
  if (i >= n_pairs) return;
  const int64_t row = i >> 5;
  const int pr = (int)(i & 31);
  const float2 v = *reinterpret_cast<const float2*>(src + row * 64 + 2 * pr);
  const Split2 s = split_pack(v.x, v.y);
  dst[row * 64 + pr] = s.hi;
  dst[row * 64 + 32 + pr] = s.lo;
}

cudaError_t launch_split_table64(const float* src, void* dst, int64_t rows, cudaStream_t s) {
  const int64_t n_pairs = rows * 32;
  const int threads = 256;
  const int64_t blocks = (n_pairs + threads - 1) / threads;
  split_table64_kernel<<<(unsigned)blocks, threads, 0, s>>>(src, reinterpret_cast<uint32_t*>(dst), n_pairs);
  ++g_launch_count;
  return cudaGetLastError();
}

#if defined(RT64_WATCHDOG) || defined(RT64_SLEEP_NS)
#undef mbar_wait
#endif

// *n = number of timed-out waits recorded since the last call (always 0 unless built with -DRT64_WATCHDOG)
cudaError_t take_din_rt64_abort(int* n, unsigned long long* rec64) {
  unsigned int flag = 0, cnt = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&flag, g_rt64_abort, sizeof(flag));
  if (e != cudaSuccess) return e;
  *n = 0;
  if (!flag) return cudaSuccess;
  e = cudaMemcpyFromSymbol(&cnt, g_rt64_nrec, sizeof(cnt));
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(rec64, g_rt64_rec, sizeof(unsigned long long) * 64);
  if (e != cudaSuccess) return e;
  *n = (int)(cnt < 16u ? cnt : 16u);
  flag = 0;
  e = cudaMemcpyToSymbol(g_rt64_abort, &flag, sizeof(flag));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_rt64_nrec, &flag, sizeof(flag));
  return e;
}

size_t din_rt64_smem_bytes() { return 1024 + RING64 + Q64_BYTES; }

static cudaError_t launch_din_rt64_once(const DinRtParams& p, const BatchView& b, cudaStream_t s) {
  DinRtParams q = p;
  const int waves = (b.B + k64Rows * p.num_sms - 1) / (k64Rows * p.num_sms);
  int rpg = (b.B + waves * p.num_sms - 1) / (waves * p.num_sms);
  if (rpg > k64Rows) rpg = k64Rows;
  if (rpg < 1) rpg = 1;
  q.rows_per_group = rpg;
  const int n_groups = (b.B + rpg - 1) / rpg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_groups < p.num_sms ? n_groups : p.num_sms);
  cfg.blockDim = dim3(k64Threads);
  cfg.dynamicSmemBytes = din_rt64_smem_bytes();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef RT64_NO_PDL                     // (experiment switch: profiles/r02/rt64_hang/README.md)
  cfg.numAttrs = 0;
#else
  cfg.numAttrs = 1;
#endif
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, din_rt64_kernel, q, b);
}

// RT64_CHUNK_GROUPS > 0: a call is split into launches of at most that many row groups per CTA
#ifndef RT64_CHUNK_GROUPS
#define RT64_CHUNK_GROUPS 0
#endif

cudaError_t launch_din_rt64(const DinRtParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  const int64_t max_rows = RT64_CHUNK_GROUPS > 0 ? (int64_t)RT64_CHUNK_GROUPS * p.num_sms * k64Rows : (int64_t)b.B;
  if (b.B <= max_rows) return launch_din_rt64_once(p, b, s);
  for (int64_t lo = 0; lo < b.B; lo += max_rows) {
    BatchView c = b;
    c.B = (int)((b.B - lo) < max_rows ? (b.B - lo) : max_rows);
    c.movie_id = b.movie_id + lo;
    c.user_id = b.user_id + lo;
    c.hist = b.hist + lo * b.hist_stride;
    c.movie_genre = b.movie_genre + lo * 3;
    c.user_genre = b.user_genre + lo * 5;
    c.numerics = b.numerics + lo * kNumNumerics;
    c.probs = b.probs + lo;
    if (b.logits) c.logits = b.logits + lo;
    for (int k = 0; k < b.n_peers; ++k) c.peer_probs[k] = b.peer_probs[k] + lo;
    cudaError_t e = launch_din_rt64_once(p, c, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t setup_din_rt64_attributes() {
  return cudaFuncSetAttribute(din_rt64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)din_rt64_smem_bytes());
}

}  // namespace srs
