// din_rtp.cu - DIN forward, row-tile kernel with the phases of consecutive row groups PIPELINED.
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167.  Same math and
// operand forms as din_rt.cu (read that header first): pre-split movie table rows gathered by
// cp.async straight into tcgen05 operand tiles, per-row weight W_r = (Wsub+Wh) + diag(c_r) Wp as
// the B operand, gate on CUDA cores, pooling as a second MMA on the same tile read MN-major, top
// MLP transposed with the rows of a group as the MMA N, bf16x3 everywhere.
//
// What din_rt_kernel loses: a CTA runs its phases one after the other (ids -> tiles -> top MLP,
// __syncthreads between them, the ring reused for the top-MLP weight images), so 60 % of a launch
// is latency chains with an idle SM.  Here every role is a persistent loop
// over the CTA's groups and nothing but mbarriers orders them:
//
//   warps 0-3, 20-23  gatherers (8 warps: the cp.async issue rate of 5 bounded the tile phase):
//                     history rows of tile K -> A ring slot K % 6, 3 tiles in flight, continuous
//                     across group boundaries
//   warp  4           issuer of every tile-phase tcgen05.mma / commit
//   warp  5           loader: ids, numerics and candidate rows of group j+2 -> staging buffer j & 1
//                     while groups j, j+1 are in flight
//   warps 6-7         builders: per-row weight operand W_r -> B ring slot K % 2
//   warps 8-15        two consumers (tile K -> consumer K & 1): gate, pooling weights, pooled rows
//   warps 16-19       top MLP of group j while the tiles of group j+1 run
//
// To make room for that overlap the top MLP no longer streams a 128 KB weight image through the
// ring: the userGenre1 / movieGenre1 columns of Dense(128) are folded at model build into fp32
// tables G[genre][unit] (19 values each: exact, added in the epilogue like the numerics), the
// remaining K = 96 columns of W1^T live in TENSOR MEMORY as the A operand (96 columns, loaded
// once per launch), W2 stays resident in shared memory (32 KB).  Registers are re-divided with
// setmaxnreg (gatherers 40, consumers 120).
//
// A launch limited to num_sms / S CTAs (srs_model_set_sm_limit) walks S groups per CTA: prologue,
// first ids and the top-MLP tail are then paid once per S groups, and S launches of consecutive
// batches share the machine.
#include <climits>

#include "rt_common.cuh"

namespace srs {

namespace {

constexpr int kPThreads = 768;
constexpr int kPRows = 32;                  // row slots per group = N/2 of the top-MLP MMAs
constexpr int kPSlotsA = 6;                 // history tiles in flight or being consumed
constexpr int kPSlotsB = 3;                 // per-row weight operands
constexpr int kPCstSlots = 8;               // per-tile constants of the gate (see the builders)
#ifndef RTP_AHEAD
#define RTP_AHEAD 0
#endif
constexpr int kPAhead = RTP_AHEAD;          // OWN tiles a gather team keeps in flight before it delivers one
constexpr int kPMaxCopies = 16;             // copies per thread and tile: 2 x 64 cells x 8 chunks / 64 threads
constexpr int kPGatherThreads = 128;        // warps 0-3: two teams of two warps
constexpr int kPBuilderThreads = 128;       // warps 20-23
constexpr int kPIdsLd = 64;                 // ints per row of the staged history ids

// shared memory (offsets from the 1024-aligned base)
constexpr uint32_t PA_SLOT = 16384;                          // [2 rows x 64 positions][hi 32 | lo 32] bf16, SW128 K-major
constexpr uint32_t PB_SLOT = 8192;                           // [hi: row0 32 units, row1 32 units | lo: same][32 k] bf16, SW64
constexpr uint32_t PO_A = 0;
constexpr uint32_t PO_B = PO_A + kPSlotsA * PA_SLOT;         // 98304
constexpr uint32_t PO_W2 = PO_B + kPSlotsB * PB_SLOT;        // 114688: 2 K blocks x [64 hi | 64 lo units][64 k]
constexpr uint32_t PO_XB = PO_W2 + 32768;                    // 147456: X operand, 2 K blocks x [32 rows hi | 32 rows lo][64 k]
constexpr uint32_t PO_H1 = PO_XB;                            //         H1 operand over it after layer 1
constexpr uint32_t PO_X = PO_XB + 16384;                     // 163840: scratch
constexpr uint32_t PX_IDS = 0;                               // [2] int [32][64]
constexpr uint32_t PX_CAND = PX_IDS + 2 * 8192;              // [2] f32 [32][32]
constexpr uint32_t PX_POOL = PX_CAND + 2 * 4096;             // [2] f32 [32][hi 32 | lo 32]
constexpr uint32_t PX_NUMS = PX_POOL + 2 * 8192;             // [2] f32 [32][8]
constexpr uint32_t PX_SID = PX_NUMS + 2 * 1024;              // [2] int [32][4]: userGenre1, movieGenre1, user id, -
constexpr uint32_t PX_B2 = PX_SID + 2 * 512;                 // [consumer][buffer] x 2 K blocks x [8 n][64 positions] bf16, SW128
constexpr uint32_t PX_CST = PX_B2 + 8192;                    // [tile K % 8] f32 [2 rows][32]
constexpr uint32_t PX_BYTES = PX_CST + kPCstSlots * 256;
// layer-2 scratch lies over the X / H1 operand tile (dead once the layer-2 MMAs have completed)
constexpr uint32_t PO_RED = PO_XB;                           // f32 [64][32]
constexpr uint32_t PO_ZP = PO_XB + 8192;                     // f32 [4][32]
static_assert((kPCstSlots & (kPCstSlots - 1)) == 0, "constants ring is indexed with a mask");
static_assert(PX_B2 % 1024 == 0 && (PO_X + PX_B2) % 1024 == 0, "pooling-weight operand tiles are 1024-byte aligned");
constexpr uint32_t P_SMEM = PO_X + PX_BYTES;
static_assert(P_SMEM + 1024 <= 232448, "does not fit the 227 KB of one CTA");
// tensor memory columns (512 allocated)
constexpr uint32_t PT_D1 = 0;                                // consumer q: [128 q, + 128)
constexpr uint32_t PT_D2 = 256;                              // consumer q, buffer u, tile row r: 256 + 32 q + 16 u + 8 r
constexpr uint32_t PT_TOP = 320;                             // top-MLP accumulators, 64 columns (layer 1, then layer 2)
constexpr uint32_t PT_W1HI = 384, PT_W1LO = 432;             // W1^T as A operand: 48 + 48 columns (96 bf16 of K each)

__device__ unsigned long long g_din_rtp_trace[40];
// per-tile timeline of CTA 0 (debug): [kind][tile K < 64]; kinds: 0 gather issued, 1 delivered, 2 B built,
// 3 MMA1 issued, 4 consumer sees D1, 5 gate done (w_ready), 6 pooling MMAs issued, 7 pooled read back
__device__ unsigned long long g_din_rtp_tl[12 * 64];   // + 8 issuer: pool wait passed, 9 pool MMAs issued, 10 pool commits done, 11 iteration start
#ifdef RTP_TIMELINE
#define RTP_TL(kind, K, cond)                                                                   \
  do {                                                                                          \
    if (p.trace && blockIdx.x == 0 && (K) < 64 && (cond)) g_din_rtp_tl[(kind) * 64 + (K)] = clock64(); \
  } while (0)
#else
#define RTP_TL(kind, K, cond) do { } while (0)
#endif
// Tracing is compiled in only with -DRTP_TIMELINE (profiles/exp/build_variants.py): every probe costs the
// single-warp roles (issuers, builders) a few dependent instructions per tile, and those roles are latency-
// bound on their own instruction stream (~8 cycles per instruction: ncu / timeline of round 2).
#ifdef RTP_TIMELINE
#define RTP_TRACE(slot, cond)                                                     \
  do {                                                                            \
    if (p.trace && blockIdx.x == 0 && (cond)) g_din_rtp_trace[slot] = clock64();  \
  } while (0)
#else
#define RTP_TRACE(slot, cond) do { } while (0)
#endif

// Every mbarrier wait of this kernel goes through rtp_wait: a wait that lasts longer than any
// legitimate one (2^28 cycles = 0.14 s) records who waited for what and raises g_din_rtp_abort, after
// which every wait in the grid returns at once - a protocol error ends the launch with wrong scores
// and a diagnosis (srs_model_status / srs_debug_din_trace slots 32..36) instead of hanging the GPU.
__device__ unsigned int g_din_rtp_abort;
// SLEEP_NS > 0: back off between polls (waits that are not latency-critical - the top MLP waiting for a whole
// group of tiles, the loader - would otherwise keep the SM's barrier unit and issue slots busy for nothing).
__device__ __forceinline__ void rtp_record_timeout(int code, uint32_t parity) {
  if (atomicCAS(&g_din_rtp_abort, 0u, 1u) == 0u) {
    g_din_rtp_trace[32] = 1ull;
    g_din_rtp_trace[33] = (unsigned long long)code;
    g_din_rtp_trace[34] = (unsigned long long)blockIdx.x;
    g_din_rtp_trace[35] = (unsigned long long)threadIdx.x;
    g_din_rtp_trace[36] = (unsigned long long)parity;
    __threadfence();
  }
}
template <int SLEEP_NS>
__device__ __forceinline__ void rtp_wait_loop(uint64_t* bar, uint32_t parity, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if ((++spins & 1023u) == 0) {
      if (*reinterpret_cast<volatile unsigned int*>(&g_din_rtp_abort)) return;
      if (spins >= (SLEEP_NS > 0 ? (1u << 19) : (1u << 24))) {       // far beyond any legitimate wait
        rtp_record_timeout(code, parity);
        return;
      }
    }
  }
}
// (an out-of-line watchdog loop would keep the hot loops smaller, but one call anywhere in the kernel makes
// ptxas fail to allocate the consumers' 120 registers)
__device__ __forceinline__ void rtp_wait(uint64_t* bar, uint32_t parity, int code) { rtp_wait_loop<0>(bar, parity, code); }
__device__ __forceinline__ void rtp_wait_lazy(uint64_t* bar, uint32_t parity, int code) { rtp_wait_loop<200>(bar, parity, code); }
__device__ __forceinline__ void rtp_wait_inl(uint64_t* bar, uint32_t parity, int code) { rtp_wait_loop<0>(bar, parity, code); }

__device__ __forceinline__ void rtp_store_x4(uint8_t* tile, int block, int row, int col, float4 v) {
  const uint32_t off = block * 8192u + sw128_offset(row, col >> 3) + ((col & 4) ? 8u : 0u);
  const Split2 s0 = split_pack(v.x, v.y), s1 = split_pack(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = make_uint2(s0.hi, s1.hi);
  *reinterpret_cast<uint2*>(tile + off + 4096u) = make_uint2(s0.lo, s1.lo);   // row + 32: same swizzle phase
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

struct GroupGeom {
  int row0, nrows, n_tiles;
};

}  // namespace

__global__ void __launch_bounds__(kPThreads, 1) din_rtp_kernel(const __grid_constant__ DinRtParams p,
                                                               BatchView b) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t a_full[kPSlotsA];     // history rows of the tile have landed (224 gatherer arrivals)
  __shared__ uint64_t a_empty[kPSlotsA];    // the pooling MMAs reading the slot have completed
  __shared__ uint64_t b_full[kPSlotsB];     // weight operand built (64 builder arrivals)
  __shared__ uint64_t b_empty[kPSlotsB];    // the activation-unit MMAs reading it have completed
  __shared__ uint64_t d1_full[2];           // consumer q: accumulators of its next tile ready
  __shared__ uint64_t d1_free[2];           // consumer q: it has read them into registers (128 arrivals)
  __shared__ uint64_t w_ready[2][2];        // consumer q, buffer u: pooling weights written (128 arrivals)
  __shared__ uint64_t d2_full[2][2];        // consumer q, buffer u: pooled accumulators ready
  __shared__ uint64_t staged[2];            // staging buffer s holds the ids / candidate rows of a group (32 arrivals)
  __shared__ uint64_t stage_free[2];        // every reader of staging buffer s is done with it (672 arrivals)
  __shared__ uint64_t pooled_ready[2];      // pooled rows of the group in buffer s are complete (256 arrivals)
  __shared__ uint64_t pooled_free[2];       // the top MLP has read them (128 arrivals)
  __shared__ uint64_t started;              // the gatherers have requested their first tile (256 arrivals, once)
  __shared__ uint64_t wbar;                 // W2 image landed (once per launch)
  __shared__ uint64_t cbar;                 // top-MLP MMAs complete
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  RTP_TRACE(0, tid == 0);
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // warp-uniform by construction
  const int wg = warp >> 2;
  const int warp_w = warp & 3;                                   // TMEM lane quarter of this warp
  const int tw = tid & 127;
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* ringA = base + PO_A;
  uint8_t* ringB = base + PO_B;
  uint8_t* xs = base + PO_X;
  int* ids_all = reinterpret_cast<int*>(xs + PX_IDS);
  float* cand_all = reinterpret_cast<float*>(xs + PX_CAND);
  float* pooled_all = reinterpret_cast<float*>(xs + PX_POOL);
  float* nums_all = reinterpret_cast<float*>(xs + PX_NUMS);
  int* sid_all = reinterpret_cast<int*>(xs + PX_SID);
  uint8_t* b2s = xs + PX_B2;
  float* cst_all = reinterpret_cast<float*>(xs + PX_CST);
  const int T = p.T;
  const int RPG = p.rows_per_group;
  const int n_groups = (b.B + RPG - 1) / RPG;
  const int n_my = blockIdx.x < n_groups ? (n_groups - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  auto geom = [&](int j) -> GroupGeom {
    GroupGeom g;
    g.row0 = (blockIdx.x + j * gridDim.x) * RPG;
    g.nrows = min(RPG, b.B - g.row0);
    g.n_tiles = (g.nrows + 1) >> 1;
    return g;
  };
  // ---- staging of one group's inputs (history ids: `team` threads; side rows: one warp, lane = row) ----
  auto stage_hist = [&](int j, int ti, int team) {
    const GroupGeom g = geom(j);
    int* ids = ids_all + (j & 1) * (kPRows * kPIdsLd);
    const int cells = 2 * g.n_tiles * kPIdsLd;
    for (int i = ti; i < cells; i += team) {
      const int r = i >> 6, t = i & 63;
      if (r < g.nrows && t < T) cp_async4(ids + i, b.hist + (size_t)(g.row0 + r) * b.hist_stride + t);
      else ids[i] = 0;
    }
    cp_async_commit();
    cp_async_wait<0>();
    for (int i = ti; i < cells; i += team) {                  // float32 round trip, range check (own cells)
      const int r = i >> 6, t = i & 63;
      if (r < g.nrows && t < T) ids[i] = checked_id(rt_f32_roundtrip_id(ids[i]), p.n_movies, b.err_flag);
    }
  };
  auto stage_rows = [&](int j) {                              // one warp, lane = row slot
    const GroupGeom g = geom(j);
    const int s = j & 1;
    float* cand = cand_all + s * (kPRows * 32);
    float* nums = nums_all + s * (kPRows * 8);
    int* sid = sid_all + s * (kPRows * 4);
    int cid = 0, uid = -1, ug = -1, mg = -1;
    float nv[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) nv[n] = 0.f;
    const bool live = lane < g.nrows;
    if (live) {
      const int row = g.row0 + lane;
      cid = __ldg(b.movie_id + row);
      uid = __ldg(b.user_id + row);
      ug = __ldg(b.user_genre + row * 5);
      mg = __ldg(b.movie_genre + row * 3);
#pragma unroll
      for (int n = 0; n < kNumNumerics; ++n) nv[n] = __ldg(b.numerics + row * kNumNumerics + n);
      cid = checked_id(rt_f32_roundtrip_id(cid), p.n_movies, b.err_flag);
      uid = checked_id(uid, p.n_users, b.err_flag);
      if (ug >= p.n_genres) { atomicExch(b.err_flag, 1); ug = -1; }
      if (mg >= p.n_genres) { atomicExch(b.err_flag, 1); mg = -1; }
      if (ug < 0) ug = -1;
      if (mg < 0) mg = -1;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.user + (size_t)uid * 32));
    }
    // candidate row: 8 x 16-byte cp.async (no registers; the chunk order is rotated per lane so that the
    // 32 rows, 128 bytes apart, do not hit the same banks); the caller commits / waits
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const int qq = (q4 + lane) & 7;
      if (live) cp_async16(cand + lane * 32 + 4 * qq, p.movie + (size_t)cid * 32 + 4 * qq);
      else *reinterpret_cast<float4*>(cand + lane * 32 + 4 * qq) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
    cp_async_wait<0>();
    *reinterpret_cast<float4*>(nums + lane * 8) = make_float4(nv[0], nv[1], nv[2], nv[3]);
    *reinterpret_cast<float4*>(nums + lane * 8 + 4) = make_float4(nv[4], nv[5], nv[6], 0.f);
    *reinterpret_cast<int4*>(sid + lane * 4) = make_int4(ug, mg, uid, 0);
  };

  // ---- prologue ---------------------------------------------------------------------------
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");    // inputs may come from the previous kernel
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (warp == 1) {                                       // one mbarrier per lane
    if (lane < 6) mbar_init(&a_full[lane], 64);            // one gather team (two warps) per tile parity
    else if (lane < 12) mbar_init(&a_empty[lane - 6], 1);
    else if (lane < 14) mbar_init(&b_full[lane - 12], kPBuilderThreads);
    else if (lane < 16) mbar_init(&b_empty[lane - 14], 1);
    else if (lane < 18) mbar_init(&d1_full[lane - 16], 1);
    else if (lane < 20) mbar_init(&d1_free[lane - 18], 128);
    else if (lane < 24) mbar_init(&w_ready[(lane - 20) >> 1][(lane - 20) & 1], 128);
    else if (lane < 28) mbar_init(&d2_full[(lane - 24) >> 1][(lane - 24) & 1], 1);
    else if (lane < 30) mbar_init(&staged[lane - 28], 32);
    else if (lane < 32) mbar_init(&stage_free[lane - 30], kPGatherThreads + kPBuilderThreads + 256 + 128);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (lane < 2) mbar_init(&pooled_ready[lane], 256);
    else if (lane < 4) mbar_init(&pooled_free[lane - 2], 128);
    else if (lane == 4) mbar_init(&wbar, 1);
    else if (lane == 5) mbar_init(&cbar, 1);
    else if (lane == 6) mbar_init(&b_full[2], kPBuilderThreads);
    else if (lane == 7) mbar_init(&b_empty[2], 1);
    else if (lane == 8) mbar_init(&started, kPGatherThreads);
    fence_mbar_init();
  }
  {
    // tile rows of positions >= T are read by both MMAs and never written by a gather: zero them once
    const int pad = 64 - T;
    if (pad > 0)
    for (int i = tid; i < pad * 8 * kPSlotsA * 2; i += kPThreads) {
      const int sr = i / (pad * 8), w = i - sr * (pad * 8);     // sr = slot * 2 + row
      *reinterpret_cast<uint4*>(ringA + (sr >> 1) * PA_SLOT + ((sr & 1) * 64 + T + (w >> 3)) * 128 + ((w & 7) << 4)) =
          make_uint4(0, 0, 0, 0);
    }
  }
  // The gate tables P_t | Q_t (T x 256 bytes) go through shared memory - the X operand tile, unused until the
  // first top MLP - so that an SM reads them from L2 once, not once per consumer thread (4x): at launch every SM
  // asks for the same lines, and whatever the L2 has to serve then delays the first history tile.
  for (int i = tid; i < T * 16; i += kPThreads)
    cp_async16(base + PO_XB + i * 16, reinterpret_cast<const uint8_t*>(p.pq) + i * 16);
  cp_async_commit();
  // groups 0 and 1 are staged by everybody (nobody has anything else to do yet); warps 6 / 7 take the rows
  if (n_my > 0) {
    if (warp == 6) stage_rows(0);
    else if (warp == 7) { if (n_my > 1) stage_rows(1); }
    else {
      const int ti = warp < 6 ? tid : tid - 64;
      stage_hist(0, ti, kPThreads - 64);
      if (n_my > 1) stage_hist(1, ti, kPThreads - 64);
    }
  }
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RTP_TRACE(1, tid == 0);
  const uint32_t tbase = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp_w * 32) << 16;
  const uint32_t s_ringA = smem_u32(ringA), s_ringB = smem_u32(ringB);
  const int first_loader_group = 2;
  // parity helpers: the n-th completion (n = 0, 1, ...) of an mbarrier is observed with parity n & 1
  auto staged_wait = [&](int j) { if (j >= first_loader_group) rtp_wait(&staged[j & 1], ((j >> 1) - 1) & 1, 1); };
  auto staged_wait_inl = [&](int j) { if (j >= first_loader_group) rtp_wait_inl(&staged[j & 1], ((j >> 1) - 1) & 1, 1); };

  if (wg == 0) {
    reg_dec<40>();
    // =================================== gatherers (warps 0-3) ==================================
    // Two TEAMS, one per tile parity: team 0 = warps 0-1 (even tiles), team 1 = warps 2-3 (odd tiles).
    // A gather warp is latency-bound on its own instruction stream (~10 cycles per instruction, timeline of
    // round 2): what counts is instructions per tile PER WARP, so the tiles are split between the teams instead
    // of the copies of every tile among all warps, and each thread keeps the shared-memory offset of its copies
    // in registers (the id index is offset >> 7): LDS id, IMAD.WIDE, LDGSTS per copy.
    RTP_TRACE(21, tid == 0);
    const int team = warp >> 1;
    const int lt = tid & 63;                               // thread of the team
    const int tsize = 64;
    const uint32_t c16 = (uint32_t)(lt & 7) * 16u;
    uint32_t dofs[kPMaxCopies];                            // byte offset in the A tile, 0xFFFFFFFF: no copy
#pragma unroll
    for (int n = 0; n < kPMaxCopies; ++n) {
      const int cell = (lt >> 3) + (tsize >> 3) * n;       // cells: row 0 positions 0..T-1, then row 1
      const int r = cell >= T ? 1 : 0, pos = cell - r * T;
      dofs[n] = cell < 2 * T ? (uint32_t)(r * 64 + pos) * 128u + ((((uint32_t)(lt & 7)) ^ (uint32_t)(pos & 7)) << 4)
                             : 0xFFFFFFFFu;
    }
    int Kg = 0, Dg = 0;                                    // OWN tiles issued / delivered (tile index = 2 * n + team)
    auto slot_of = [&](int n_own) { return (2 * n_own + team) % kPSlotsA; };
    auto deliver_one = [&]() {                             // the oldest outstanding own tile HAS landed: publish it
      fence_async_smem();
      RTP_TL(1, 2 * Dg + team, lt == 0);
      mbar_arrive(&a_full[slot_of(Dg)]);
      ++Dg;
    };
    auto deliver_oldest = [&]() {                          // wait for the oldest outstanding tile, then publish it
      const int pending = Kg - Dg;                         // 1 .. kPAhead + 1 commit groups in flight
      if (pending >= 3) cp_async_wait<2>();
      else if (pending == 2) cp_async_wait<1>();
      else cp_async_wait<0>();
      deliver_one();
    };
    int kbase = 0;
    for (int j = 0; j < n_my; ++j) {
      const GroupGeom g = geom(j);
      if (j >= first_loader_group && !mbar_test_wait(&staged[j & 1], ((j >> 1) - 1) & 1)) {
        cp_async_wait<0>();                                // the loader is late: do not sit on landed tiles
        while (Dg < Kg) deliver_one();
        staged_wait(j);
      }
      const int* ids = ids_all + (j & 1) * (kPRows * kPIdsLd);
      for (int k = (team - kbase) & 1; k < g.n_tiles; k += 2) {      // this team's tiles of the group
        const int K = kbase + k;
        const int slot = K % kPSlotsA;
        if (K >= kPSlotsA) {
          // No free slot yet: do not sit on tiles that have landed (the consumers that will free the slot may
          // be waiting for exactly those) - publish them first, then wait.
          const uint32_t par = ((K / kPSlotsA) + 1) & 1;
          while (!mbar_test_wait(&a_empty[slot], par)) {
            if (Dg < Kg) deliver_oldest();
            else { rtp_wait(&a_empty[slot], par, 2); break; }
          }
        }
        uint8_t* A = ringA + slot * PA_SLOT;
        const int* idrow = ids + 2 * k * kPIdsLd;          // ids of the tile's two rows: [64 | 64]
#pragma unroll
        for (int n = 0; n < kPMaxCopies; ++n) {
          if (dofs[n] != 0xFFFFFFFFu) {
            const uint32_t d = dofs[n];
            cp_async16(A + d, p.movie_split + (size_t)idrow[d >> 7] * 128 + c16);
          }
        }
        cp_async_commit();
        RTP_TL(0, K, lt == 0);
        ++Kg;
        if (Kg - Dg > kPAhead) deliver_oldest();
        if (K == 0) RTP_TRACE(20, tid == 0);
      }
      if (j == 0) mbar_arrive(&started);                   // the top MLP may now load its weights
      mbar_arrive(&stage_free[j & 1]);                     // the ids were read when the copies were issued
      kbase += g.n_tiles;
    }
    cp_async_wait<0>();
    while (Dg < Kg) deliver_one();
    cp_async_wait<0>();
    while (Dg < Kg) deliver_one();
  } else if (wg == 1) {
    // registers: the CTA's pool is what it was LAUNCHED with (768 x 80): 480 per thread slot =
    // 40 (gatherers) + 40 (here: issuers, loader) + 120 + 120 (consumers) + 80 (top MLP) + 80 (builders)
    reg_dec<40>();
    if (warp == 6) {
      // =================================== loader ==========================================
      for (int j = first_loader_group; j < n_my; ++j) {
        rtp_wait_lazy(&stage_free[j & 1], ((j >> 1) - 1) & 1, 7);  // every reader of group j - 2 is done
        stage_rows(j);
        stage_hist(j, lane, 32);
        mbar_arrive(&staged[j & 1]);
      }
    } else if (warp == 4) {
      // ============================ issuer of the activation-unit MMAs ==========================
      // Two issuer warps (this one and warp 5): one warp issuing both MMA groups of every tile in order needed
      // ~1.4 K cycles per tile - it is latency-bound on its own dependent instruction stream (~8 cycles per
      // instruction), not on the tensor pipe - and, being in order, it could deadlock one-tile groups.  The
      // two streams are independent: this one needs the tile, its weight operand and the accumulator buffer.
      // Slot counters and phase bits are kept incrementally: no division in the loop.
      int NT = 0;
      for (int j = 0; j < n_my; ++j) NT += geom(j).n_tiles;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int K = 0; K < NT; ++K) {
        const int q = K & 1;
        RTP_TL(11, K, lane == 0);
        rtp_wait(&a_full[sa], pa, 3);
        RTP_TL(8, K, lane == 0);
        rtp_wait(&b_full[sb], pb, 4);
        if (K >= 2) rtp_wait(&d1_free[q], ((K >> 1) - 1) & 1, 5);   // the accumulators of tile K - 2 are in registers
        RTP_TL(9, K, lane == 0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD1 = tbase + PT_D1 + 128u * q;
          const uint64_t ad = smem_desc_sw128(s_ringA + sa * PA_SLOT);
          const uint64_t bd = smem_desc_sw64(s_ringB + sb * PB_SLOT);
          mma_ss(tD1, ad + 0, bd + 0, idesc_bf16(128, 128), 0);     // H_hi . [W_hi | W_lo]
          mma_ss(tD1, ad + 2, bd + 2, idesc_bf16(128, 128), 1);
          mma_ss(tD1, ad + 4, bd + 0, idesc_bf16(128, 64), 1);      // H_lo . W_hi
          mma_ss(tD1, ad + 6, bd + 2, idesc_bf16(128, 64), 1);
          RTP_TL(10, K, true);
          mma_commit(&d1_full[q]);
          mma_commit(&b_empty[sb]);
        }
        __syncwarp();
        RTP_TL(3, K, lane == 0);
        if (++sa == kPSlotsA) { sa = 0; pa ^= 1u; }
        if (++sb == kPSlotsB) { sb = 0; pb ^= 1u; }
      }
    } else if (warp == 5) {
      // ============================ issuer of the pooling MMAs ==================================
      int NT = 0;
      for (int j = 0; j < n_my; ++j) NT += geom(j).n_tiles;
      int sa = 0;
      for (int K = 0; K < NT; ++K) {
        const int q = K & 1, u = (K >> 1) & 1;
        rtp_wait(&w_ready[q][u], (K >> 2) & 1, 6);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tD2 = tbase + PT_D2 + 32u * q + 16u * u;
          const uint32_t s_b2 = smem_u32(b2s) + (q * 2 + u) * 2048;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ss(tD2 + 8 * r, smem_desc_mn_sw128(s_ringA + sa * PA_SLOT + r * 8192 + ks * 2048),
                     smem_desc_sw128(s_b2 + r * 1024) + 2 * ks, idesc_mn(64, 8, 1), ks > 0);
          mma_commit(&d2_full[q][u]);
          mma_commit(&a_empty[sa]);
        }
        __syncwarp();
        RTP_TL(6, K, lane == 0);
        if (++sa == kPSlotsA) sa = 0;
      }
    }                                                      // warp 7 has no role
  } else if (wg == 2 || wg == 3) {
    // =================================== consumers ===========================================
    reg_inc<120>();
    const int q = wg - 2;
    const int r_t = warp_w >> 1, t = tw & 63;             // this thread's tile row and position
    float rc[64];                                         // P_t[0..31] | Q_t[0..31] of its position
    {
      // from the copy the prologue left in the X tile region (the top MLP overwrites it after the first group's
      // tiles, long after this); positions >= T: any finite values do (w is forced to 0)
      const float* src = reinterpret_cast<const float*>(base + PO_XB) + (size_t)min(t, T - 1) * 64;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
        rc[4 * i] = v.x; rc[4 * i + 1] = v.y; rc[4 * i + 2] = v.z; rc[4 * i + 3] = v.w;
      }
    }
    const uint32_t tD1 = tbase + PT_D1 + 128u * q;
    // a tile whose pooled accumulators are still to be read back
    bool pend = false, pend_last = false;
    int pend_K = 0, pend_j = 0, pend_k = 0;
    int pool_group = -1;                                  // group this thread last wrote pooled rows of
    auto pooled_buffer_wait = [&](int j) {                // before the first pooled write of group j
      if (pool_group != j) {
        if (j >= 2) rtp_wait_inl(&pooled_free[j & 1], ((j >> 1) - 1) & 1, 9);
        pool_group = j;
      }
    };
    auto pool_out = [&]() {
      const int u = (pend_K >> 1) & 1;
      rtp_wait_inl(&d2_full[q][u], (pend_K >> 2) & 1, 10);
      tc_fence_after();
      pooled_buffer_wait(pend_j);
      // D2 row m = 16 warp_w + lane (lane < 16): m < 32 -> hi e = m, else lo e = m - 32;
      // columns 8 r + {0: . w_hi, 1: . w_lo}
      uint32_t d[16];
      tmem_ld16(tbase + PT_D2 + 32u * q + 16u * u + lane_base, d);
      tmem_ld_wait();
      float* pooled = pooled_all + (pend_j & 1) * (kPRows * 64);
      if (lane < 16) {
        const int m = 16 * warp_w + lane;
        const bool hi = warp_w < 2;
        pooled[(2 * pend_k) * 64 + m] = hi ? __uint_as_float(d[0]) + __uint_as_float(d[1]) : __uint_as_float(d[0]);
        pooled[(2 * pend_k + 1) * 64 + m] = hi ? __uint_as_float(d[8]) + __uint_as_float(d[9]) : __uint_as_float(d[8]);
      }
      tc_fence_before();
      RTP_TL(7, pend_K, tw == 0);
      if (pend_last) mbar_arrive(&pooled_ready[pend_j & 1]);
      pend = false;
    };
    int kbase = 0;
    for (int j = 0; j < n_my; ++j) {
      const GroupGeom g = geom(j);
      staged_wait_inl(j);
      const float* cand = cand_all + (j & 1) * (kPRows * 32);
      bool any = false;
      for (int k = (q - kbase) & 1; k < g.n_tiles; k += 2) {
        any = true;
        const int K = kbase + k, u = (K >> 1) & 1;
        const float* cs_buf = cst_all + (K % kPCstSlots) * 64;      // written by the builders before MMA1(K) was issued
        rtp_wait_inl(&d1_full[q], (K >> 1) & 1, 11);
        tc_fence_after();
        if (K == q) RTP_TRACE(3 + 7 * q, tw == 0);
        if (K == 4) RTP_TRACE(12, tw == 0);
        RTP_TL(4, K, tw == 0);
        // ---- gate: v = D_hi + D_lo + cst; s = sum_j v_j P_tj + |v_j| Q_tj
        float2 sa = make_float2(p.au_bout, 0.f), sb = make_float2(0.f, 0.f);
        {
          const float* cs = cs_buf + r_t * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t dh[16], dl[16];
            tmem_ld16(tD1 + r_t * 32 + 16 * half + lane_base, dh);
            tmem_ld16(tD1 + 64 + r_t * 32 + 16 * half + lane_base, dl);
            tmem_ld_wait();
            if (half == 1) {                                // every value of the tile is in registers
              tc_fence_before();
              mbar_arrive(&d1_free[q]);
            }
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const int jx = 16 * half + jj;
              const float4 c4 = *reinterpret_cast<const float4*>(cs + jx);
              float2 v01 = add2(make_float2(__uint_as_float(dh[jj]), __uint_as_float(dh[jj + 1])),
                                make_float2(__uint_as_float(dl[jj]), __uint_as_float(dl[jj + 1])));
              float2 v23 = add2(make_float2(__uint_as_float(dh[jj + 2]), __uint_as_float(dh[jj + 3])),
                                make_float2(__uint_as_float(dl[jj + 2]), __uint_as_float(dl[jj + 3])));
              v01 = add2(v01, make_float2(c4.x, c4.y));
              v23 = add2(v23, make_float2(c4.z, c4.w));
              sa = fma2(v01, make_float2(rc[jx], rc[jx + 1]), sa);
              sb = fma2(v23, make_float2(rc[jx + 2], rc[jx + 3]), sb);
              sa = fma2(make_float2(fabsf(v01.x), fabsf(v01.y)), make_float2(rc[32 + jx], rc[32 + jx + 1]), sa);
              sb = fma2(make_float2(fabsf(v23.x), fabsf(v23.y)), make_float2(rc[32 + jx + 2], rc[32 + jx + 3]), sb);
            }
          }
        }
        const float sg = (sa.x + sa.y) + (sb.x + sb.y);
        const float w = (t < T) ? 1.f / (1.f + __expf(-sg)) : 0.f;
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
        mbar_arrive(&w_ready[q][u]);
        RTP_TL(5, K, tw == 0);
        if (K == 4) RTP_TRACE(13, tw == 0);
        if (pend) pool_out();                               // the previous own tile's pooling MMAs finished long ago
        if (K == 4) RTP_TRACE(14, tw == 0);
        pend = true; pend_K = K; pend_j = j; pend_k = k; pend_last = k + 2 >= g.n_tiles;
      }
      if (!any) {                                           // a one-tile group of the other consumer
        if (pend) pool_out();
        pooled_buffer_wait(j);
        mbar_arrive(&pooled_ready[j & 1]);
      }
      mbar_arrive(&stage_free[j & 1]);                      // candidate rows no longer needed here
      kbase += g.n_tiles;
    }
    if (pend) pool_out();
    RTP_TRACE(4 + 7 * q, tw == 0);
  } else if (wg == 5) {
    // =================================== builders (warps 20-23) ==============================
    // B operand of every tile: W_r = (Wsub+Wh) + diag(c_r) Wp, bf16 hi / lo, and the gate constants of its two
    // rows.  Four warps: a builder warp is latency-bound on its own instruction stream like every single-warp
    // role here (two warps needed 1.4 K cycles per tile).  Thread -> (unit pj, 8-wide chunk cq of e):
    //   rc[0..7] = (Wsub+Wh)[8 cq ..][pj], rc[8..15] = Wp[8 cq ..][pj], wcst[0..7] = (Wc - Wsub)[8 cq ..][pj]
    const int bt = tid - 640, pj = bt >> 2, cq = bt & 3;
    float rc[16], wcst[8];
    {
      const float4 a0 = ldg4(p.waT + pj * 32 + 8 * cq), a1 = ldg4(p.waT + pj * 32 + 8 * cq + 4);
      const float4 p0 = ldg4(p.wpT + pj * 32 + 8 * cq), p1 = ldg4(p.wpT + pj * 32 + 8 * cq + 4);
      rc[0] = a0.x; rc[1] = a0.y; rc[2] = a0.z; rc[3] = a0.w; rc[4] = a1.x; rc[5] = a1.y; rc[6] = a1.z; rc[7] = a1.w;
      rc[8] = p0.x; rc[9] = p0.y; rc[10] = p0.z; rc[11] = p0.w; rc[12] = p1.x; rc[13] = p1.y; rc[14] = p1.z; rc[15] = p1.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) wcst[e] = __ldg(p.au_wc + (8 * cq + e) * 32 + pj);
    }
    const float cst_bias = __ldg(p.au_b + pj);
    int Kb = 0, slot = 0;
    uint32_t pe = 1;                                       // parity to wait with on b_empty: ((Kb / 3) + 1) & 1, kept incrementally
    for (int j = 0; j < n_my; ++j) {
      const GroupGeom g = geom(j);
      staged_wait(j);
      const float* cand = cand_all + (j & 1) * (kPRows * 32);
      for (int k = 0; k < g.n_tiles; ++k, ++Kb) {
        if (Kb >= kPSlotsB) rtp_wait(&b_empty[slot], pe, 8);
        uint8_t* Bt = ringB + slot * PB_SLOT;
        float part[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float* cv = cand + (2 * k + r) * 32 + 8 * cq;
          const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
          const float2 v0 = fma2(make_float2(c0.x, c0.y), make_float2(rc[8], rc[9]), make_float2(rc[0], rc[1]));
          const float2 v1 = fma2(make_float2(c0.z, c0.w), make_float2(rc[10], rc[11]), make_float2(rc[2], rc[3]));
          const float2 v2 = fma2(make_float2(c1.x, c1.y), make_float2(rc[12], rc[13]), make_float2(rc[4], rc[5]));
          const float2 v3 = fma2(make_float2(c1.z, c1.w), make_float2(rc[14], rc[15]), make_float2(rc[6], rc[7]));
          const Split2 s0 = split_pack(v0.x, v0.y), s1 = split_pack(v1.x, v1.y);
          const Split2 s2 = split_pack(v2.x, v2.y), s3 = split_pack(v3.x, v3.y);
          const uint32_t n = r * 32 + pj;
          *reinterpret_cast<uint4*>(Bt + sw64_offset(n, cq)) = make_uint4(s0.hi, s1.hi, s2.hi, s3.hi);
          *reinterpret_cast<uint4*>(Bt + sw64_offset(64 + n, cq)) = make_uint4(s0.lo, s1.lo, s2.lo, s3.lo);
          // gate constant of the row: this thread's 8 of the 32 terms of cst[r][pj]
          float a = c0.x * wcst[0];
          a = fmaf(c0.y, wcst[1], a); a = fmaf(c0.z, wcst[2], a); a = fmaf(c0.w, wcst[3], a);
          a = fmaf(c1.x, wcst[4], a); a = fmaf(c1.y, wcst[5], a); a = fmaf(c1.z, wcst[6], a); a = fmaf(c1.w, wcst[7], a);
          part[r] = a;
        }
        // the four chunks of a unit sit in adjacent lanes.  Slot Kb % 8 of the constants ring: the gate of tile
        // Kb - 8 finished before MMA1(Kb - 3) completed (b_empty above).
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float a = part[r];
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          if (cq == 0) cst_all[(Kb & (kPCstSlots - 1)) * 64 + r * 32 + pj] = a + cst_bias;
        }
        fence_async_smem();
        mbar_arrive(&b_full[slot]);
        RTP_TL(2, Kb, bt == 0);
        if (++slot == kPSlotsB) { slot = 0; pe ^= 1u; }
      }
      mbar_arrive(&stage_free[j & 1]);
    }
  } else {
    // =================================== top MLP (wg == 4) ===================================
    // Its weights (48 KB + 32 KB per SM, the same lines for every SM) are not needed before the first group's
    // tiles are done: ask for them only once the first history tile has been requested.
    rtp_wait_lazy(&started, 0, 15);
    // W1^T -> tensor memory (A operand): this thread's lane = unit tw, 96 packed bf16 pairs
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.w1_tmem + (size_t)tw * 96);
#pragma unroll
      for (int cch = 0; cch < 6; ++cch) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 x = __ldg(src + cch * 4 + i);
          v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
        }
        tmem_st16(tbase + PT_W1HI + 16 * cch + lane_base, v);
      }
      tmem_st_wait();
      tc_fence_before();
    }
    if (warp == 16) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&wbar, 32768);
        bulk_g2s(base + PO_W2, p.image, 32768, &wbar);      // W2 image = first 32 KB of the din_rt image
      }
      __syncwarp();
    }
    named_sync(3, 128);
    const float b1 = __ldg(p.b1 + tw), a1 = __ldg(p.a1 + tw);
    float w1n[kNumNumerics];
#pragma unroll
    for (int n = 0; n < kNumNumerics; ++n) w1n[n] = __ldg(p.w1num + n * 128 + tw);
    const float b2 = __ldg(p.b2 + (tw & 63)), a2 = __ldg(p.a2 + (tw & 63)), w3 = __ldg(p.w3 + (tw & 63));
    const uint32_t idesc_top = idesc_bf16(128, 2 * kPRows);
    const uint32_t tTop = tbase + PT_TOP;
    const uint32_t s_xb = smem_u32(base + PO_XB), s_w2 = smem_u32(base + PO_W2);
    uint8_t* xb = base + PO_XB;
    float* red = reinterpret_cast<float*>(base + PO_RED);
    float* zp = reinterpret_cast<float*>(base + PO_ZP);
    uint32_t cphase = 0;
    bool w2_ready = false;
    for (int j = 0; j < n_my; ++j) {
      const GroupGeom g = geom(j);
      const int s = j & 1;
      staged_wait(j);
      rtp_wait_lazy(&pooled_ready[s], (j >> 1) & 1, 12);
      if (j == 0) RTP_TRACE(5, tw == 0);
      const float* cand = cand_all + s * (kPRows * 32);
      const float* pooled = pooled_all + s * (kPRows * 64);
      const float* nums = nums_all + s * (kPRows * 8);
      const int* sid = sid_all + s * (kPRows * 4);
      // ---- X operand: K block 0 = [userId | pooled], K block 1 = [candidate | -]; thread -> (row, 8 floats)
      {
        const int xr = tw >> 2, c8 = (tw & 3) * 8;
        float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0, p0 = u0, p1 = u0, c0 = u0, c1 = u0;
        if (xr < g.nrows) {
          const int uid = sid[xr * 4 + 2];
          u0 = ldg4(p.user + (size_t)uid * 32 + c8);
          u1 = ldg4(p.user + (size_t)uid * 32 + c8 + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(pooled + xr * 64 + c8);
          const float4 h1 = *reinterpret_cast<const float4*>(pooled + xr * 64 + c8 + 4);
          const float4 l0 = *reinterpret_cast<const float4*>(pooled + xr * 64 + 32 + c8);
          const float4 l1 = *reinterpret_cast<const float4*>(pooled + xr * 64 + 32 + c8 + 4);
          p0 = make_float4(h0.x + l0.x, h0.y + l0.y, h0.z + l0.z, h0.w + l0.w);
          p1 = make_float4(h1.x + l1.x, h1.y + l1.y, h1.z + l1.z, h1.w + l1.w);
          c0 = *reinterpret_cast<const float4*>(cand + xr * 32 + c8);
          c1 = *reinterpret_cast<const float4*>(cand + xr * 32 + c8 + 4);
        }
        mbar_arrive(&pooled_free[s]);                         // pooled rows are in registers
        rtp_store_x4(xb, 0, xr, c8, u0);
        rtp_store_x4(xb, 0, xr, c8 + 4, u1);
        rtp_store_x4(xb, 0, xr, 32 + c8, p0);
        rtp_store_x4(xb, 0, xr, 32 + c8 + 4, p1);
        rtp_store_x4(xb, 1, xr, c8, c0);
        rtp_store_x4(xb, 1, xr, c8 + 4, c1);
      }
      fence_async_smem();
      tc_fence_before();
      named_sync(3, 128);
      if (warp == 16) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {                    // K = 96: block 0 steps 0..3, block 1 steps 0..1
            const uint64_t xd = smem_desc_sw128(s_xb + (ks >> 2) * 8192) + 2 * (ks & 3);   // [X hi | X lo], N = 64
            mma_ts(tTop, tbase + PT_W1HI + 8 * ks, xd, idesc_top, ks > 0);       // W1hi.(Xhi | Xlo)
            mma_ts(tTop, tbase + PT_W1LO + 8 * ks, xd, idesc_top, 1);            // W1lo.(Xhi | Xlo)
          }
          mma_commit(&cbar);
        }
        __syncwarp();
      }
      // genre columns of Dense(128): G_u[userGenre1][unit] + G_m[movieGenre1][unit] per row slot.  The loads are L2
      // round trips: chunk 0 is requested before the MMA wait, chunk c + 1 while chunk c is processed.
      auto genre_sums = [&](int r8, float (&gs)[8]) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int2 gid = *reinterpret_cast<const int2*>(sid + (r8 * 8 + r) * 4);       // userGenre1, movieGenre1
          const float gu = gid.x >= 0 ? __ldg(p.gtab_u + gid.x * 128 + tw) : 0.f;
          const float gm = gid.y >= 0 ? __ldg(p.gtab_m + gid.y * 128 + tw) : 0.f;
          gs[r] = gu + gm;
        }
      };
      float gnext[8];
      genre_sums(0, gnext);
      rtp_wait(&cbar, cphase, 13);
      cphase ^= 1;
      __syncwarp();
      tc_fence_after();
      if (j == 0) RTP_TRACE(6, tw == 0);
      // ---- layer-1 epilogue: this thread is unit tw for all 32 row slots
      {
        const uint32_t koff = (uint32_t)(tw >> 6) * 8192u;
        const uint32_t chunk = (tw & 63) >> 3, within = (tw & 7) * 2;
#pragma unroll
        for (int r8 = 0; r8 < 4; ++r8) {
          float gsum[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) gsum[r] = gnext[r];
          if (r8 < 3) genre_sums(r8 + 1, gnext);
          uint32_t d[8], d2[8];
          tmem_ld8(tTop + 8 * r8 + lane_base, d);              // W1 . X hi
          tmem_ld8(tTop + 32 + 8 * r8 + lane_base, d2);        // W1 . X lo
          tmem_ld_wait();
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int sr = r8 * 8 + r;
            const float4 n0 = *reinterpret_cast<const float4*>(nums + sr * 8);
            const float4 n1 = *reinterpret_cast<const float4*>(nums + sr * 8 + 4);
            float v = (__uint_as_float(d[r]) + __uint_as_float(d2[r])) + b1;
            v += gsum[r];
            v = fmaf(n0.x, w1n[0], v); v = fmaf(n0.y, w1n[1], v); v = fmaf(n0.z, w1n[2], v);
            v = fmaf(n0.w, w1n[3], v); v = fmaf(n1.x, w1n[4], v); v = fmaf(n1.y, w1n[5], v);
            v = fmaf(n1.z, w1n[6], v);
            v = v > 0.f ? v : a1 * v;
            const uint32_t off = koff + sw128_offset(sr, chunk) + within;
            const __nv_bfloat16 vh = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16*>(base + PO_H1 + off) = vh;
            *reinterpret_cast<__nv_bfloat16*>(base + PO_H1 + off + 4096u) = __float2bfloat16_rn(v - __bfloat162float(vh));
          }
        }
      }
      if (j == 0) RTP_TRACE(22, tw == 0);
      mbar_arrive(&stage_free[s]);                            // numerics / ids / candidate rows consumed
      fence_async_smem();
      tc_fence_before();
      named_sync(3, 128);
      if (warp == 16) {
        if (!w2_ready) { rtp_wait(&wbar, 0, 14); w2_ready = true; }
        tc_fence_after();
        if (elect_one()) {
          uint32_t acc = 0;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t a = smem_desc_sw128(s_w2 + kb * 16384);
            const uint64_t hs = smem_desc_sw128(s_xb + kb * 8192);               // [H1 hi | H1 lo], N = 64
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_ss(tTop, a + 2 * ks, hs + 2 * ks, idesc_top, acc);             // (W2hi ; W2lo).(H1hi | H1lo)
              acc = 1;
            }
          }
          mma_commit(&cbar);
        }
        __syncwarp();
      }
      rtp_wait(&cbar, cphase, 13);
      cphase ^= 1;
      __syncwarp();
      tc_fence_after();
      if (j == 0) RTP_TRACE(23, tw == 0);
      // ---- layer-2 epilogue: rows 0..63 of D hold W2hi . (H1hi | H1lo), rows 64..127 W2lo . (...)
      {
        float dsum[32];
#pragma unroll
        for (int r8 = 0; r8 < 4; ++r8) {
          uint32_t d[8], d2[8];
          tmem_ld8(tTop + 8 * r8 + lane_base, d);
          tmem_ld8(tTop + 32 + 8 * r8 + lane_base, d2);
          tmem_ld_wait();
#pragma unroll
          for (int r = 0; r < 8; ++r) dsum[8 * r8 + r] = __uint_as_float(d[r]) + __uint_as_float(d2[r]);
        }
        tc_fence_before();
        if (tw >= 64) {                                          // lo halves of W2 -> smem
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4)
            *reinterpret_cast<float4*>(red + (tw - 64) * 32 + 4 * r4) =
                make_float4(dsum[4 * r4], dsum[4 * r4 + 1], dsum[4 * r4 + 2], dsum[4 * r4 + 3]);
        }
        named_sync(3, 128);
        if (tw < 64) {
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            float v = dsum[r] + red[tw * 32 + r] + b2;           // (W2hi + W2lo) . (H1hi + H1lo)
            v = v > 0.f ? v : a2 * v;
            red[tw * 32 + r] = v * w3;
          }
        }
        named_sync(3, 128);
        {  // 32 rows x 4 partial sums of 16 units
          const int r = tw & 31, pt = tw >> 5;
          float sum = 0.f;
#pragma unroll
          for (int uu = 0; uu < 16; ++uu) sum += red[(pt * 16 + uu) * 32 + r];
          zp[pt * 32 + r] = sum;
        }
        named_sync(3, 128);
        if (tw < kPRows) {
          const float z = p.b3 + ((zp[tw] + zp[32 + tw]) + (zp[64 + tw] + zp[96 + tw]));
          if (tw < g.nrows) {
            store_score(b, g.row0 + tw, sigmoidf_acc(z));
            if (b.logits) b.logits[g.row0 + tw] = z;
          }
        }
      }
      named_sync(3, 128);                                     // X / H1 tile, red and the accumulators are reused
      if (j == 0) RTP_TRACE(7, tw == 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
  RTP_TRACE(8, tid == 0);
  gather_signal_tail(b);                                  // spanning ranking call: publish "slice complete"
}

cudaError_t read_din_rtp_trace(unsigned long long* out40) {
  return cudaMemcpyFromSymbol(out40, g_din_rtp_trace, sizeof(unsigned long long) * 40);
}
cudaError_t read_din_rtp_timeline(unsigned long long* out768) {
  return cudaMemcpyFromSymbol(out768, g_din_rtp_tl, sizeof(unsigned long long) * 768);
}

// Did a wait of an earlier launch time out (see rtp_wait)?  Copies the record {code, block, thread,
// parity} and clears the flag.
cudaError_t take_din_rtp_abort(int* aborted, unsigned long long* rec4) {
  unsigned int flag = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&flag, g_din_rtp_abort, sizeof(flag));
  if (e != cudaSuccess) return e;
  *aborted = flag != 0;
  if (flag) {
    unsigned long long t[40];
    e = cudaMemcpyFromSymbol(t, g_din_rtp_trace, sizeof(t));
    if (e != cudaSuccess) return e;
    for (int i = 0; i < 4; ++i) rec4[i] = t[33 + i];
    flag = 0;
    e = cudaMemcpyToSymbol(g_din_rtp_abort, &flag, sizeof(flag));
  }
  return e;
}

static size_t din_rtp_smem_bytes() { return 1024 + P_SMEM; }

cudaError_t launch_din_rtp(const DinRtParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  DinRtParams q = p;
  // rows per group: as even as possible over the SMs, at most 32, even
  const int waves = (b.B + kPRows * p.num_sms - 1) / (kPRows * p.num_sms);
  int rpg = (b.B + waves * p.num_sms - 1) / (waves * p.num_sms);
  rpg = (rpg + 1) & ~1;
  if (rpg > kPRows) rpg = kPRows;
  if (rpg < 2) rpg = 2;
  q.rows_per_group = rpg;
  const int n_groups = (b.B + rpg - 1) / rpg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_groups < p.num_sms ? n_groups : p.num_sms);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = din_rtp_smem_bytes();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see the kernel prologue
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, din_rtp_kernel, q, b);
}

cudaError_t setup_din_rtp_attributes() {
  return cudaFuncSetAttribute(din_rtp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)din_rtp_smem_bytes());
}

}  // namespace srs
