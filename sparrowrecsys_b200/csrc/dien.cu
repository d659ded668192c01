// dien.cu - DIEN forward (y_pred): gather + GRU + attention + AUGRU + top MLP in one kernel.
//
// Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIEN.py:154-256.
//   X = Emb[hist] [T,E], C = Emb[cand] [E]                  (one shared table, :161-167)
//   g_t = GRU(X)_t     Keras GRU, gates z|r|h, reset_after, masked where hist_t == 0 (:169)
//   s_t = sigmoid(Dense1(sigmoid(Dense32(g_t * C))))         attention score (:172-195)
//   u_t = AUGRU(g_t, s_t, u_{t-1})                           three two-layer gates (:204-245)
//   y = sigmoid(Dense1(PReLU(Dense64(PReLU(Dense128([u_T | C | profile | context]))))))
// The AUGRU initial state, a fresh GlorotUniform draw per call in the reference (:235-236),
// is the stored vector `augru_h0` (oracle/ctr_oracle.py::dien_forward states the semantics).
//
// The recurrence is sequential over T and E x E small, so it runs on CUDA cores: one warp
// owns a row, lane e owns element e of every state vector; a matrix-vector product is EP
// shuffle-broadcasts against weight rows held in shared memory (the whole sequence part,
// 15 EP^2 + 45 EP + 68 floats, is staged once per CTA).  GRU step, attention and AUGRU step
// of position t are fused in one loop, so no [T,E] intermediate exists.  The top MLP is the
// 32-row tile code DIN uses (common.cuh::dense_layer).
#include "kernels.h"

namespace srs {

constexpr int kDienRows = 32;     // rows per CTA tile

template <int EP>
struct DienBlob {                 // float offsets inside DienParams::seq (see model.cu::build_dien)
  static constexpr int GW = 0;                       // gru kernel            [EP k][3][EP]
  static constexpr int GU = GW + 3 * EP * EP;        // gru recurrent kernel  [EP k][3][EP]
  static constexpr int AW = GU + 3 * EP * EP;        // attention Dense32     [EP k][32]
  static constexpr int IW = AW + 32 * EP;            // augru input kernels   [3 g][EP k][EP]
  static constexpr int HW = IW + 3 * EP * EP;        // augru hidden kernels  [3 g][EP k][EP]
  static constexpr int SW = HW + 3 * EP * EP;        // augru act kernels     [3 g][EP k][EP]
  static constexpr int BX = SW + 3 * EP * EP;        // gru input bias        [3][EP]
  static constexpr int BH = BX + 3 * EP;             // gru recurrent bias    [3][EP]
  static constexpr int BI = BH + 3 * EP;             // augru input bias      [3][EP]
  static constexpr int BA = BI + 3 * EP;             // augru act bias        [3][EP]
  static constexpr int H0 = BA + 3 * EP;             // augru initial state   [EP]
  static constexpr int AB = H0 + EP;                 // attention Dense32 bias [32]
  static constexpr int AO = AB + 32;                 // attention Dense1 kernel [32]
  static constexpr int ABO = AO + 32;                // attention Dense1 bias  [1] (+3 pad)
  static constexpr int TOTAL = ABO + 4;
};

template <int EP>
__global__ void __launch_bounds__(kThreads) dien_kernel(DienParams p, BatchView b) {
  static_assert(EP <= 32, "one lane per state element");
  using L = DienBlob<EP>;
  constexpr int R = kDienRows;
  constexpr int KP = 5 * EP + kNumPad;
  constexpr int LDX = KP + 4;
  constexpr int LDH1 = 128 + 4;
  constexpr int LDH2 = 64 + 4;
  // tile column offsets (the order model.cu permutes dense/kernel to)
  constexpr int OFF_UG = 0, OFF_U = EP, OFF_ST = 2 * EP, OFF_C = 3 * EP, OFF_MG = 4 * EP,
                OFF_NUM = 5 * EP;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                      // [R][LDX]
  float* H1 = Xs + R * LDX;              // [R][LDH1]
  float* H2 = H1 + R * LDH1;             // [R][LDH2]
  float* Sq = H2 + R * LDH2;             // [L::TOTAL] sequence-part weights
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * R;

  stage_weights(Sq, p.seq, L::TOTAL);

  // ---- side features: user genre, user, movie genre rows and numerics ------------
  tile_side_features<EP, R>(Xs, LDX, row0, b, p.user, p.ugenre, p.mgenre, p.n_users, p.n_genres,
                            OFF_UG, OFF_U, OFF_MG, OFF_NUM);
  stage_wait();
  __syncthreads();

  // ---- interest evolution: one warp per row, lane = state element --------------------
  const int le = lane < EP ? lane : EP - 1;          // lanes >= EP mirror lane EP-1 (unused)
  const float* gw = Sq + L::GW + le;
  const float* gu = Sq + L::GU + le;
  const float* aw = Sq + L::AW + lane;
  const float* iw = Sq + L::IW + le;
  const float* hw = Sq + L::HW + le;
  const float* sw = Sq + L::SW + le;
  const float att_b = Sq[L::AB + lane], att_wo = Sq[L::AO + lane], att_bo = Sq[L::ABO];
  for (int r = warp; r < R; r += kThreads / 32) {
    const int row = row0 + r;
    float* xrow = Xs + r * LDX;
    if (row >= b.B) {                                      // warp-uniform
      if (lane < EP) { xrow[OFF_C + lane] = 0.f; xrow[OFF_ST + lane] = 0.f; }
      continue;
    }
    // ids pass through float32 numeric columns before the Embedding layer (DIEN.py:96-105)
    int cid = __float2int_rz(__int2float_rn(__ldg(b.movie_id + row)));
    cid = checked_id(cid, p.n_movies, b.err_flag);
    const float c = lane < EP ? __ldg(p.movie + (size_t)cid * EP + lane) : 0.f;
    const int32_t* hrow = b.hist + (size_t)row * b.hist_stride;
    float h = 0.f;                                         // GRU state = its output g_t
    float u = Sq[L::H0 + le];                              // AUGRU state
    int hid_next = __ldg(hrow);
    for (int t = 0; t < p.T; ++t) {
      const int raw = hid_next;
      if (t + 1 < p.T) hid_next = __ldg(hrow + t + 1);
      int hid = __float2int_rz(__int2float_rn(raw));
      const bool valid = hid != 0;                         // Embedding(mask_zero=True) mask
      hid = checked_id(hid, p.n_movies, b.err_flag);
      const float x = lane < EP ? __ldg(p.movie + (size_t)hid * EP + lane) : 0.f;
      // -- GRU step (Keras: z | r | h, reset_after)
      float xz = Sq[L::BX + le], xr = Sq[L::BX + EP + le], xh = Sq[L::BX + 2 * EP + le];
      float rz = Sq[L::BH + le], rr = Sq[L::BH + EP + le], rh = Sq[L::BH + 2 * EP + le];
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        const float xk = __shfl_sync(0xffffffffu, x, k);
        const float hk = __shfl_sync(0xffffffffu, h, k);
        xz = fmaf(xk, gw[k * 3 * EP], xz);
        xr = fmaf(xk, gw[k * 3 * EP + EP], xr);
        xh = fmaf(xk, gw[k * 3 * EP + 2 * EP], xh);
        rz = fmaf(hk, gu[k * 3 * EP], rz);
        rr = fmaf(hk, gu[k * 3 * EP + EP], rr);
        rh = fmaf(hk, gu[k * 3 * EP + 2 * EP], rh);
      }
      {
        const float z = sigmoidf_acc(xz + rz);
        const float rg = sigmoidf_acc(xr + rr);
        const float hh = tanhf(xh + rg * rh);
        const float hn = z * h + (1.f - z) * hh;
        if (valid) h = hn;                                 // masked step: state and output carried
      }
      // -- attention score of position t
      const float pc = h * c;
      float a = att_b;
#pragma unroll
      for (int k = 0; k < EP; ++k) a = fmaf(__shfl_sync(0xffffffffu, pc, k), aw[k * 32], a);
      a = sigmoidf_acc(a);
      const float s = sigmoidf_acc(warp_sum(a * att_wo) + att_bo);
      // -- AUGRU step: x = g_t (= h), state u
      float pr = Sq[L::BI + le], pz = Sq[L::BI + EP + le], ph = Sq[L::BI + 2 * EP + le];
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        const float hk = __shfl_sync(0xffffffffu, h, k);
        const float uk = __shfl_sync(0xffffffffu, u, k);
        pr = fmaf(hk, iw[k * EP], pr);
        pz = fmaf(hk, iw[EP * EP + k * EP], pz);
        ph = fmaf(hk, iw[2 * EP * EP + k * EP], ph);
        pr = fmaf(uk, hw[k * EP], pr);
        pz = fmaf(uk, hw[EP * EP + k * EP], pz);
      }
      float ar = Sq[L::BA + le], az = Sq[L::BA + EP + le];
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        ar = fmaf(__shfl_sync(0xffffffffu, pr, k), sw[k * EP], ar);
        az = fmaf(__shfl_sync(0xffffffffu, pz, k), sw[EP * EP + k * EP], az);
      }
      const float rg = sigmoidf_acc(ar), zg = sigmoidf_acc(az);
      const float uz = u * zg;
#pragma unroll
      for (int k = 0; k < EP; ++k)
        ph = fmaf(__shfl_sync(0xffffffffu, uz, k), hw[2 * EP * EP + k * EP], ph);
      float ah = Sq[L::BA + 2 * EP + le];
#pragma unroll
      for (int k = 0; k < EP; ++k)
        ah = fmaf(__shfl_sync(0xffffffffu, ph, k), sw[2 * EP * EP + k * EP], ah);
      const float hn = tanhf(ah);
      const float ra = s * rg;
      u = (1.f - ra) * u + ra * hn;
    }
    if (lane < EP) { xrow[OFF_C + lane] = c; xrow[OFF_ST + lane] = u; }
  }
  __syncthreads();

  // ---- top MLP on the tile ----------------------------------------------------------
  dense_layer<R, 128, 2, 8>(Xs, LDX, KP, p.W1, p.b1, ACT_PRELU, p.a1, H1, LDH1);
  __syncthreads();
  dense_layer<R, 64, 1, 8>(H1, LDH1, 128, p.W2, p.b2, ACT_PRELU, p.a2, H2, LDH2);
  __syncthreads();
  row_dot<R>(H2, LDH2, 64, p.w3, [&](int r, float s) {
    const int row = row0 + r;
    if (row >= b.B) return;
    const float z = s + p.b3;
    store_score(b, row, sigmoidf_acc(z));
    if (b.logits) b.logits[row] = z;
  });
}

template <int EP>
static size_t dien_smem() {
  return (size_t)(kDienRows * ((5 * EP + kNumPad + 4) + 132 + 68) + DienBlob<EP>::TOTAL) *
         sizeof(float);
}

int dien_seq_floats(int EP) {
  switch (EP) {
    case 12: return DienBlob<12>::TOTAL;
    case 16: return DienBlob<16>::TOTAL;
    case 32: return DienBlob<32>::TOTAL;
  }
  return -1;
}

template <int EP>
static cudaError_t launch_dien_t(const DienParams& p, const BatchView& b, cudaStream_t s) {
  const int blocks = (b.B + kDienRows - 1) / kDienRows;
  dien_kernel<EP><<<blocks, kThreads, dien_smem<EP>(), s>>>(p, b);
  ++g_launch_count;
  return cudaGetLastError();
}

cudaError_t launch_dien(const DienParams& p, const BatchView& b, cudaStream_t s) {
  if (b.B <= 0) return cudaSuccess;
  switch (p.EP) {
    case 12: return launch_dien_t<12>(p, b, s);
    case 16: return launch_dien_t<16>(p, b, s);
    case 32: return launch_dien_t<32>(p, b, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t setup_dien_attributes() {
  cudaError_t e;
#define SRS_ATTR(E_)                                                                     \
  e = cudaFuncSetAttribute(dien_kernel<E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)dien_smem<E_>());                                        \
  if (e != cudaSuccess) return e;
  SRS_ATTR(12) SRS_ATTR(16) SRS_ATTR(32)
#undef SRS_ATTR
  return cudaSuccess;
}

}  // namespace srs
