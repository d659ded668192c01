/*
 * ctr_oracle_c.c - plain-C restatement of the reference's DIN forward graph, threaded over
 * batch rows with OpenMP.
 *
 * THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.  Same rule as ctr_oracle.py: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load
 * it.  PARITY UNPINNED in the same sense as ctr_oracle.py (TensorFlow cannot run here); this
 * file is held to the numpy oracle by tests/test_oracle_c.py.
 *
 * Reference: TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167, statement by
 * statement (same order as oracle/ctr_oracle.py::din_forward):
 *   :134-137  H = Emb[hist], C = Emb[cand]          (one shared table, id 0 an ordinary row)
 *   :139-147  A = [H - C | H | C | H * C]           [T, 4E] per row
 *   :149-150  a = PReLU(A . au_dense + b), alpha [T, 32]
 *   :151-152  w = sigmoid(a . au_out + b)            [T]
 *   :153-158  pooled = sum_t w_t H_t                 (sigmoid gate, sum pooling, padding included)
 *   :108-128  user_profile / context in DenseFeatures' sorted column order
 *   :161-167  x = [user_profile | pooled | C | context] -> Dense128 PReLU -> Dense64 PReLU -> Dense1
 *
 * Why it exists next to the numpy oracle: as a *timing* baseline the numpy/torch restatements
 * materialise [B, T, 4E] tensors and swing 15x between hosts with the thread-pool's mood; a
 * row-parallel C loop is what an optimised CPU executor of this graph (TF's Eigen/oneDNN
 * kernels on an intra-op pool) amounts to, and it is reproducible across boxes.
 *
 * Inputs are the encoded batch of the C ABI (include/srs_ctr.h srs_batch): ids as int32
 * (history ids already passed through the float32 round trip by the caller - exact below 2^24,
 * rounded above, as numeric_column does), genre vocabulary indices with -1 = zero vector.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct srs_oracle_din {
  int32_t T, E, AU, H1, H2;
  int32_t n_movies, n_users, n_genres;
  const float* emb;        /* [n_movies][E]  Keras Embedding, shared candidate / history    */
  const float* user_emb;   /* [n_users][E]                                                 */
  const float* ugenre_emb; /* [n_genres][E]                                                */
  const float* mgenre_emb; /* [n_genres][E]                                                */
  const float* au_w;       /* [4E][AU]                                                     */
  const float* au_b;       /* [AU]                                                         */
  const float* au_alpha;   /* [T][AU]                                                      */
  const float* au_out_w;   /* [AU]                                                         */
  float au_out_b;
  const float* w1;         /* [5E+7][H1] rows in the sorted concat order of DIN.py:161-162  */
  const float* b1;
  const float* a1;         /* PReLU alpha [H1]                                             */
  const float* w2;         /* [H1][H2]                                                     */
  const float* b2;
  const float* a2;
  const float* w3;         /* [H2]                                                         */
  float b3;
} srs_oracle_din;

static inline float prelu_f(float x, float a) { return x > 0.f ? x : a * x; }
static inline float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

/* y[n] = x[k] . W[k][n] + b, plain loops; the compiler vectorises over n */
static void dense_row(const float* x, int K, const float* W, const float* b, int N, float* y) {
  for (int n = 0; n < N; ++n) y[n] = b ? b[n] : 0.f;
  for (int k = 0; k < K; ++k) {
    const float xv = x[k];
    const float* w = W + (size_t)k * N;
    for (int n = 0; n < N; ++n) y[n] += xv * w[n];
  }
}

/* numerics order of srs_batch: movieAvgRating, movieRatingCount, movieRatingStddev, releaseYear,
 * userAvgRating, userRatingCount, userRatingStddev */
int srs_oracle_din_forward(const srs_oracle_din* m, int32_t B, const int32_t* movie_id,
                           const int32_t* user_id, const int32_t* hist, int32_t hist_stride,
                           const int32_t* user_genre1, const int32_t* movie_genre1,
                           const float* numerics, float* prob, float* logit, int32_t threads) {
  const int T = m->T, E = m->E, AU = m->AU, H1 = m->H1, H2 = m->H2;
  const int K1 = 5 * E + 7;
  int bad = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    float* A = (float*)malloc(sizeof(float) * (size_t)T * 4 * E);
    float* a = (float*)malloc(sizeof(float) * (size_t)(AU > H1 ? AU : H1));
    float* x = (float*)malloc(sizeof(float) * (size_t)K1);
    float* h1 = (float*)malloc(sizeof(float) * (size_t)H1);
    float* h2 = (float*)malloc(sizeof(float) * (size_t)H2);
    float* pooled = (float*)malloc(sizeof(float) * (size_t)E);
#pragma omp for schedule(static)
    for (int r = 0; r < B; ++r) {
      const int cid = movie_id[r];
      const int uid = user_id[r];
      if (cid < 0 || cid >= m->n_movies || uid < 0 || uid >= m->n_users) { bad = 1; prob[r] = 0.f; continue; }
      const float* C = m->emb + (size_t)cid * E;                                   /* :136-137 */
      int ok = 1;
      for (int t = 0; t < T; ++t) {
        const int hid = hist[(size_t)r * hist_stride + t];
        if (hid < 0 || hid >= m->n_movies) { ok = 0; break; }
        const float* H = m->emb + (size_t)hid * E;                                 /* :134 */
        float* At = A + (size_t)t * 4 * E;
        for (int e = 0; e < E; ++e) {                                              /* :141-147 */
          At[e] = H[e] - C[e];
          At[E + e] = H[e];
          At[2 * E + e] = C[e];
          At[3 * E + e] = H[e] * C[e];
        }
      }
      if (!ok) { bad = 1; prob[r] = 0.f; continue; }
      for (int e = 0; e < E; ++e) pooled[e] = 0.f;
      for (int t = 0; t < T; ++t) {
        const float* At = A + (size_t)t * 4 * E;
        dense_row(At, 4 * E, m->au_w, m->au_b, AU, a);                             /* :149 */
        float s = m->au_out_b;
        for (int j = 0; j < AU; ++j) s += prelu_f(a[j], m->au_alpha[t * AU + j]) * m->au_out_w[j];  /* :150-151 */
        const float w = sigmoid_f(s);                                              /* :152 */
        const float* H = At + E;
        for (int e = 0; e < E; ++e) pooled[e] += w * H[e];                         /* :153-158 */
      }
      const float* nv = numerics + (size_t)r * 7;
      const int ug = user_genre1[r], mg = movie_genre1[r];
      int o = 0;                                                                   /* :108-114 sorted */
      x[o++] = nv[4];                                                              /* userAvgRating */
      for (int e = 0; e < E; ++e) x[o++] = (ug >= 0 && ug < m->n_genres) ? m->ugenre_emb[(size_t)ug * E + e] : 0.f;
      for (int e = 0; e < E; ++e) x[o++] = m->user_emb[(size_t)uid * E + e];
      x[o++] = nv[5];                                                              /* userRatingCount */
      x[o++] = nv[6];                                                              /* userRatingStddev */
      for (int e = 0; e < E; ++e) x[o++] = pooled[e];                              /* :161 */
      for (int e = 0; e < E; ++e) x[o++] = C[e];
      x[o++] = nv[0];                                                              /* :117-123 sorted */
      for (int e = 0; e < E; ++e) x[o++] = (mg >= 0 && mg < m->n_genres) ? m->mgenre_emb[(size_t)mg * E + e] : 0.f;
      x[o++] = nv[1];
      x[o++] = nv[2];
      x[o++] = nv[3];
      dense_row(x, K1, m->w1, m->b1, H1, h1);                                      /* :163-164 */
      for (int j = 0; j < H1; ++j) h1[j] = prelu_f(h1[j], m->a1[j]);
      dense_row(h1, H1, m->w2, m->b2, H2, h2);                                     /* :165-166 */
      float z = m->b3;
      for (int j = 0; j < H2; ++j) z += prelu_f(h2[j], m->a2[j]) * m->w3[j];       /* :167 */
      if (logit) logit[r] = z;
      prob[r] = sigmoid_f(z);
    }
    free(A); free(a); free(x); free(h1); free(h2); free(pooled);
  }
  return bad ? -5 : 0;   /* SRS_ERR_RANGE */
}

int srs_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
