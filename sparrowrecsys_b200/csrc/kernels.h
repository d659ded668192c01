// kernels.h - parameter blocks (device pointers into the model's private weight
// layout) and launch entry points of the fused per-model forward kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace srs {

// ---- NeuralCF / two towers (NeuralCF.py:45-70) ------------------------------------
// One thread per row.  All Dense weights live in one small blob that each CTA copies
// to shared memory; offsets are in floats.  Hidden widths are zero-padded to HP.
struct NcfParams {
  const float* movie;      // [n_movies][EP]
  const float* user;       // [n_users][EP]
  const float* blob;       // dense weights, layout below
  int blob_floats;
  int n_movies, n_users;
  int EP, HP;
  int n_layers;            // hidden layers (1..3)
  int two_towers;          // 0: neural_cf_model_1, 1: neural_cf_model_2
  int final_dense;         // two towers only
  // neuralcf:  L0 kernel [2EP][HP] @w_off[0], bias @b_off[0]; Ll kernel [HP][HP] @w_off[l]
  //            out kernel [HP] @out_w, bias @out_b
  // twotowers: item tower @w_off[l]/b_off[l], user tower @w_off[3+l]/b_off[3+l];
  //            out kernel [1] @out_w, bias @out_b
  int w_off[6], b_off[6];
  int out_w, out_b;
};

// ---- EmbeddingMLP / Wide&Deep (EmbeddingMLP.py:72-77, WideNDeep.py:101-107) --------
struct EmbMlpParams {
  const float* genre[8];   // movieGenre1..3, userGenre1..5 tables [19][EP]
  const float* movie;      // [n_movies][EP]
  const float* user;       // [n_users][EP]
  const float* W1;         // [KP = 10*EP + 8][128] rows in tile order, zero padded
  const float* b1;         // [128]
  const float* W2;         // [128][128]
  const float* b2;         // [128]
  const float* w3;         // [128] deep rows of dense_2
  const float* wide;       // [cross_buckets] wide rows of dense_2 (nullptr for EmbeddingMLP)
  float b3;
  int n_movies, n_users, n_genres, cross_buckets;
  int EP;
};

// ---- EmbeddingMLP / Wide&Deep on tensor cores (embmlp_tc.cu): E <= 12 ------------------------
struct EmbMlpTcParams {
  const float* genre[8];   // [19][12]
  const float* movie;      // [n_movies][12]
  const float* user;       // [n_users][12]
  const uint8_t* image;    // 128 KB: W1^T hi/lo, W2^T hi/lo as bf16 SW128 operand tiles
  const float* b1;         // [128]
  const float* w1num;      // [8][128] rows of dense/kernel that multiply the 7 numerics
  const float* b2;         // [128]
  const float* w3;         // [128]
  const float* wide;       // [cross_buckets] or nullptr
  float b3;
  int n_movies, n_users, n_genres, cross_buckets;
  int num_sms;
};

// ---- DeepFM (DeepFM.py:91-113) -----------------------------------------------------
struct DeepFmParams {
  const float* fm_movie;   // [n_movies][EP]
  const float* fm_user;    // [n_users][EP]
  const float* fm_mgenre;  // [19][EP]
  const float* fm_ugenre;  // [19][EP]
  const float* deep_movie; // [n_movies][EP]
  const float* deep_user;  // [n_users][EP]
  const float* W1;         // [KP = 2*EP + 8][64]
  const float* b1;
  const float* W2;         // [64][64]
  const float* b2;
  const float* first;      // [fm1_width] one-hot rows of dense_2 (movieGenre1|movieId|userGenre1|userId)
  const float* wdeep;      // [64]
  float wdot[4];
  float bout;
  int n_movies, n_users, n_genres;
  int EP;
};

// ---- DeepFM with the deep MLP on tensor cores (deepfm_tc.cu): emb_dim 13..16 --------------------
struct DeepFmTcParams {
  const float* fm_movie;   // [n_movies][16]
  const float* fm_user;
  const float* fm_mgenre;
  const float* fm_ugenre;
  const float* deep_movie;
  const float* deep_user;
  const uint8_t* image;    // 64 KB: W1^T hi/lo, W2^T hi/lo as [128][64] bf16 SW128 tiles
  const float* b1;         // [64]
  const float* w1num;      // [8][64]
  const float* b2;         // [64]
  const float* first;      // [fm1_width]
  const float* wdeep;      // [64]
  float wdot[4];
  float bout;
  int n_movies, n_users, n_genres;
  int num_sms;
};

// ---- DeepFM_v2 (DeepFM_v2.py:98-155) -----------------------------------------------
struct DeepFm2Params {
  const float* mgenre;     // [19][EP]
  const float* movie;      // [n_movies][EP]
  const float* ugenre;     // [19][EP]
  const float* user;       // [n_users][EP]
  const float* first;      // [fm1_width] first_cat kernel
  const float* first_num;  // [8]
  float first_bias;        // first_cat bias + first_num bias
  const float* proj[4];    // [EP][64] each, field order movieGenre1, movieId, userGenre1, userId
  const float* proj_b[4];  // [64]
  const float* proj_num;   // [8][64]
  const float* proj_num_b; // [64]
  const float* Wd;         // [320][32]
  const float* bd;         // [32]
  const float* Wd1;        // [32][16]
  const float* bd1;        // [16]
  const float* wout;       // [1 + 64 + 16]
  float bout;
  int n_movies, n_users, n_genres;
  int EP;
};

// ---- DIN (DIN.py:125-167) ------------------------------------------------------------
struct DinParams {
  const float* movie;      // shared candidate/history table [n_movies][EP]
  const float* user;       // [n_users][EP]
  const float* ugenre;     // [19][EP]
  const float* mgenre;     // [19][EP]
  // activation unit, algebraically folded (DESIGN.md "DIN activation unit"):
  //   Dense32([h-c, h, c, h*c]) = h.(W_sub+W_h) + (h*c).W_prod + c.(W_c-W_sub) + b
  const float* au_wh;      // [EP][32]  W_sub + W_h
  const float* au_wp;      // [EP][32]  W_prod
  const float* au_wc;      // [EP][32]  W_c - W_sub
  const float* au_b;       // [32]
  const float* au_alpha;   // [T][32]   per-position PReLU
  const float* au_wout;    // [32]
  float au_bout;
  // top MLP, first kernel permuted to the tile order
  //   [userGenre1 | userId | pooled | candidate | movieGenre1] x EP, then 7 numerics + pad
  const float* W1;         // [KP = 5*EP + 8][128]
  const float* b1;         // [128]
  const float* a1;         // [128] PReLU alpha
  const float* W2;         // [128][64]
  const float* b2;         // [64]
  const float* a2;         // [64]
  const float* w3;         // [64]
  float b3;
  int n_movies, n_users, n_genres;
  int T;
  int EP;
};

// ---- DIEN (DIEN.py:154-256), CUDA-core kernel for E <= 32 -------------------------------------
struct DienParams {
  const float* movie;      // shared candidate/history table [n_movies][EP]
  const float* user;       // [n_users][EP]
  const float* ugenre;     // [19][EP]
  const float* mgenre;     // [19][EP]
  const float* seq;        // GRU + attention + AUGRU weights, layout dien.cu::DienBlob<EP>
  // top MLP, first kernel permuted to the tile order
  //   [userGenre1 | userId | augru state | candidate | movieGenre1] x EP, then 7 numerics + pad
  const float* W1;         // [KP = 5*EP + 8][128]
  const float* b1;         // [128]
  const float* a1;         // [128] PReLU alpha
  const float* W2;         // [128][64]
  const float* b2;         // [64]
  const float* a2;         // [64]
  const float* w3;         // [64]
  float b3;
  int n_movies, n_users, n_genres;
  int T;
  int EP;
};

// ---- DIN on tensor cores (din_tc.cu): E padded to 32, T <= 128 -----------------------------
struct DinTcParams {
  const float* movie;      // [n_movies][32]
  const float* user;       // [n_users][32]
  const float* ugenre;     // [19][32]
  const float* mgenre;     // [19][32]
  const uint8_t* image;    // shared-memory image: bf16 hi/lo SW128 operand tiles + P/Q epilogue tables
  const float* au_wc;      // [32][32]  W_c - W_sub
  const float* au_b;       // [32]
  const float* b1;         // [128]
  const float* a1;         // [128]
  const float* w1num;      // [8][128] rows of dense/kernel that multiply the 7 numerics
  const float* b2;         // [64]
  const float* a2;         // [64]
  const float* w3;         // [64]
  float au_wout[32];
  float au_bout;
  float b3;
  int n_movies, n_users, n_genres;
  int T;
  int CPR;                 // 32-position chunks per row = ceil(T / 32)
  int num_sms;
  int trace;               // debug: record phase timestamps of worker 0 (srs_debug_din_trace)
};

// din_rt.cu (E padded to 32, T <= 64) and din_rt64.cu (E padded to 64, T <= 256): history rows gathered
// by cp.async into tcgen05 operand tiles; table pitches and the P/Q row length follow the padded E
struct DinRtParams {
  const float* movie;        // [n_movies][32] fp32 (candidate rows)
  const uint8_t* movie_split;// [n_movies][32 bf16 hi | 32 bf16 lo]  (history rows)
  const float* user;         // [n_users][32]
  const float* ugenre;       // [19][32]
  const float* mgenre;       // [19][32]
  const uint8_t* image;      // W2 | W1 hi | W1 lo operand images (131072 bytes)
  const float* waT;          // [32 units][32 e]  (Wsub + Wh)^T
  const float* wpT;          // [32 units][32 e]  Wp^T
  const float* pq;           // [T][64]: P_t[0..31] | Q_t[0..31]
  const float* au_wc;        // [32][32]  W_c - W_sub
  const float* au_b;         // [32]
  const float* b1;           // [128]
  const float* a1;           // [128]
  const float* w1num;        // [8][128]
  const float* b2;           // [64]
  const float* a2;           // [64]
  const float* w3;           // [64]
  float au_bout;
  float b3;
  int n_movies, n_users, n_genres;
  int T;
  int rows_per_group;        // set by the launcher
  int nch;                   // din_rt64: 128-position chunks per row (1 or 2)
  int num_sms;
  int trace;
  // din_rtp (pipelined row-tile kernel): layer-1 weights as a tensor-memory A operand and the genre
  // columns of the top MLP folded into fp32 tables
  const uint32_t* w1_tmem;   // [128 units][48 words hi | 48 words lo]: packed bf16 pairs of W1^T over
                             // K = [userId 32 | pooled 32 | candidate 32]
  const float* gtab_u;       // [n_genres][128]: userGenre1 embedding row . its rows of dense/kernel
  const float* gtab_m;       // [n_genres][128]: movieGenre1 likewise
};

// launchers (defined next to their kernels); return cudaGetLastError()
cudaError_t launch_din_rt(const DinRtParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_split_table(const float* src, void* dst, int64_t rows, cudaStream_t s);
cudaError_t read_din_rt_trace(unsigned long long* out40);
cudaError_t setup_din_rt_attributes();
cudaError_t launch_din_rtp(const DinRtParams& p, const BatchView& b, cudaStream_t s);
cudaError_t setup_din_rtp_attributes();
cudaError_t read_din_rtp_trace(unsigned long long* out40);
cudaError_t take_din_rtp_abort(int* aborted, unsigned long long* rec4);
cudaError_t read_din_rtp_timeline(unsigned long long* out768);
cudaError_t launch_din_rt64(const DinRtParams& p, const BatchView& b, cudaStream_t s);
cudaError_t take_din_rt64_abort(int* n, unsigned long long* rec64);   // debugging builds (-DRT64_WATCHDOG) only
cudaError_t launch_split_table64(const float* src, void* dst, int64_t rows, cudaStream_t s);
cudaError_t setup_din_rt64_attributes();
cudaError_t launch_din_tc(const DinTcParams& p, const BatchView& b, cudaStream_t s);
cudaError_t read_din_tc_trace(unsigned long long* out40);
cudaError_t launch_ncf(const NcfParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_embmlp(const EmbMlpParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_embmlp_tc(const EmbMlpTcParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_deepfm(const DeepFmParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_deepfm_tc(const DeepFmTcParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_deepfm2(const DeepFm2Params& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_din(const DinParams& p, const BatchView& b, cudaStream_t s);
cudaError_t launch_dien(const DienParams& p, const BatchView& b, cudaStream_t s);
int dien_seq_floats(int EP);     // size of DienParams::seq for a padded width, -1 if unsupported
cudaError_t setup_dien_attributes();
cudaError_t launch_fill_uniform(float* x, int64_t n, uint64_t seed, float lo, float hi,
                                cudaStream_t s);
cudaError_t launch_umma_selftest(const float* A, const float* B, float* D, int N, int KB,
                                 int a_in_tmem, cudaStream_t s);
cudaError_t launch_umma_bench(unsigned long long* out, int N, int n_mma, int a_in_tmem, int two_acc,
                              int uniform, cudaStream_t s);
cudaError_t launch_cosine(const float* q, const float* c, int n, int dim, float* out,
                          cudaStream_t s);
cudaError_t launch_widen_u16(const uint16_t* src, int32_t* dst, int64_t n, cudaStream_t s);
cudaError_t launch_assemble_request(const int32_t* req, const void* movie_feats, int n_table, int n, int hc,
                                    int dense, int32_t* movie_id, int32_t* user_id, int32_t* hist,
                                    int32_t* movie_genre, int32_t* user_genre, float* numerics, int* err_flag,
                                    cudaStream_t s);
// topk.cu: ranking = descending score, ties by position; min(k, n) results
size_t topk_scratch_bytes(int n);
cudaError_t launch_topk(const float* scores, int n, int k, int32_t* top_idx, float* top_scores,
                        void* scratch, cudaStream_t s);
// latency path: the last kernel of the call publishes {seq, error word (cleared)} to a host-mapped record
cudaError_t launch_topk_done(const float* scores, int n, int k, int32_t* top_idx, float* top_scores,
                             void* scratch, int* err_flag, uint32_t* done, uint32_t seq, cudaStream_t s);
cudaError_t launch_finish(int* err_flag, uint32_t* done, uint32_t seq, cudaStream_t s);

// one-time per-device kernel attribute setup (dynamic shared memory opt-in)
cudaError_t setup_kernel_attributes();

extern int64_t g_launch_count;   // kernels launched by this library

}  // namespace srs
