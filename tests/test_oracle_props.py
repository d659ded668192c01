"""Structural checks of the oracle against slow, literal re-derivations of the Keras
graphs (pure-Python loops on tiny inputs) and algebraic properties of the domain."""
import math

import numpy as np
import pytest

from oracle import ctr_oracle as O
from sparrowrecsys_b200.features import synthetic_features
from sparrowrecsys_b200.spec import GENRE_VOCAB, default_spec, history_keys
from sparrowrecsys_b200.weights import init_weights, weight_shapes


def _sig(x):
    return 1.0 / (1.0 + math.exp(-x))


def test_din_matches_literal_loops():
    spec = default_spec("din", emb_dim=4, hist_len=3, n_movies=50, n_users=40)
    W = init_weights(spec, 7)
    f = synthetic_features(spec, 6, seed=3)
    p, z = O.din_forward(spec, W, f, dtype=np.float64)
    E, T = 4, 3
    Wd = {k: v.astype(np.float64) for k, v in W.items()}
    for b in range(6):
        c = Wd["embedding"][int(f["movieId"][b])]
        pooled = np.zeros(E)
        for t, key in enumerate(history_keys(T)):
            h = Wd["embedding"][int(f[key][b])]
            a_in = np.concatenate([h - c, h, c, h * c])               # DIN.py:141-147
            a = a_in @ Wd["au_dense/kernel"] + Wd["au_dense/bias"]
            a = np.where(a > 0, a, Wd["au_prelu/alpha"][t] * a)        # alpha per position
            w = _sig(float(a @ Wd["au_out/kernel"][:, 0] + Wd["au_out/bias"][0]))
            pooled += w * h
        gi = lambda key: GENRE_VOCAB.index(f[key][b]) if f[key][b] in GENRE_VOCAB else -1
        ug = Wd["userGenre1_embedding"][gi("userGenre1")] if gi("userGenre1") >= 0 else np.zeros(E)
        mg = Wd["movieGenre1_embedding"][gi("movieGenre1")] if gi("movieGenre1") >= 0 else np.zeros(E)
        up = np.concatenate([[f["userAvgRating"][b]], ug, Wd["userId_embedding"][int(f["userId"][b])],
                             [f["userRatingCount"][b]], [f["userRatingStddev"][b]]])
        ctx = np.concatenate([[f["movieAvgRating"][b]], mg, [f["movieRatingCount"][b]],
                              [f["movieRatingStddev"][b]], [f["releaseYear"][b]]])
        x = np.concatenate([up, pooled, c, ctx]).astype(np.float64)
        x = x @ Wd["dense/kernel"] + Wd["dense/bias"]
        x = np.where(x > 0, x, Wd["prelu/alpha"] * x)
        x = x @ Wd["dense_1/kernel"] + Wd["dense_1/bias"]
        x = np.where(x > 0, x, Wd["prelu_1/alpha"] * x)
        zz = float(x @ Wd["dense_2/kernel"][:, 0] + Wd["dense_2/bias"][0])
        assert abs(zz - z[b, 0]) < 1e-9
        assert abs(_sig(zz) - p[b, 0]) < 1e-12


def test_din_padding_is_not_masked():
    """id 0 is an ordinary row (mask_zero has no numerical effect):
    changing row 0 of the table changes the output of a padded sample."""
    spec = default_spec("din", emb_dim=4, hist_len=3, n_movies=20, n_users=10)
    W = init_weights(spec, 1)
    f = synthetic_features(spec, 4, seed=0)
    f["userRatedMovie3"][:] = 0
    p0, _ = O.din_forward(spec, W, f)
    W2 = dict(W)
    W2["embedding"] = W["embedding"].copy()
    W2["embedding"][0] += 1.0
    p1, _ = O.din_forward(spec, W2, f)
    assert np.abs(p0 - p1).max() > 1e-6


def test_din_activation_unit_fold_identity():
    """Dense32([h-c,h,c,h*c]) == h.(Wsub+Wh) + (h*c).Wp + c.(Wc-Wsub) + b - the algebra
    the CUDA kernel relies on (csrc/din.cu)."""
    rng = np.random.default_rng(0)
    E = 8
    K = rng.standard_normal((4 * E, 32))
    b = rng.standard_normal(32)
    h, c = rng.standard_normal(E), rng.standard_normal(E)
    ref = np.concatenate([h - c, h, c, h * c]) @ K + b
    Wsub, Wh, Wc, Wp = K[:E], K[E:2 * E], K[2 * E:3 * E], K[3 * E:]
    fold = h @ (Wsub + Wh) + (h * c) @ Wp + c @ (Wc - Wsub) + b
    np.testing.assert_allclose(fold, ref, rtol=1e-12, atol=1e-12)


def test_deepfm_first_order_is_one_hot_matmul():
    spec = default_spec("deepfm", emb_dim=4, n_movies=30, n_users=25)
    W = init_weights(spec, 2)
    f = synthetic_features(spec, 16, seed=5)
    p, z = O.deepfm_forward(spec, W, f, dtype=np.float64)
    # literal: build the [B, fm1+4+64] concat with explicit one-hots
    G = spec.n_genres
    Wd = {k: v.astype(np.float64) for k, v in W.items()}
    for b in range(16):
        onehot = np.zeros(spec.fm1_width)
        gi = lambda key: GENRE_VOCAB.index(f[key][b]) if f[key][b] in GENRE_VOCAB else -1
        if gi("movieGenre1") >= 0:
            onehot[gi("movieGenre1")] = 1
        onehot[G + int(f["movieId"][b])] = 1
        if gi("userGenre1") >= 0:
            onehot[G + spec.n_movies + gi("userGenre1")] = 1
        onehot[2 * G + spec.n_movies + int(f["userId"][b])] = 1
        zero = np.zeros(4)
        item = Wd["fm_movieId_embedding"][int(f["movieId"][b])]
        user = Wd["fm_userId_embedding"][int(f["userId"][b])]
        ig = Wd["fm_movieGenre1_embedding"][gi("movieGenre1")] if gi("movieGenre1") >= 0 else zero
        ug = Wd["fm_userGenre1_embedding"][gi("userGenre1")] if gi("userGenre1") >= 0 else zero
        dots = [item @ user, ig @ ug, ig @ user, item @ ug]
        deep = np.concatenate([[f["movieAvgRating"][b]], Wd["deep_movieId_embedding"][int(f["movieId"][b])],
                               [f["movieRatingCount"][b]], [f["movieRatingStddev"][b]], [f["releaseYear"][b]],
                               [f["userAvgRating"][b]], Wd["deep_userId_embedding"][int(f["userId"][b])],
                               [f["userRatingCount"][b]], [f["userRatingStddev"][b]]]).astype(np.float64)
        deep = np.maximum(deep @ Wd["dense/kernel"] + Wd["dense/bias"], 0)
        deep = np.maximum(deep @ Wd["dense_1/kernel"] + Wd["dense_1/bias"], 0)
        concat = np.concatenate([onehot, dots, deep])
        zz = float(concat @ Wd["dense_2/kernel"][:, 0] + Wd["dense_2/bias"][0])
        assert abs(zz - z[b, 0]) < 1e-9


def test_deepfm_v2_fm_identity():
    """(sum v)^2 - sum v^2 == 2 * sum_{i<j} v_i v_j elementwise (no 1/2 in the reference)."""
    spec = default_spec("deepfm_v2", emb_dim=4, n_movies=30, n_users=25)
    W = init_weights(spec, 3)
    f = synthetic_features(spec, 8, seed=6)
    _, z = O.deepfm_v2_forward(spec, W, f, dtype=np.float64)
    assert np.isfinite(z).all()
    rng = np.random.default_rng(1)
    F = rng.standard_normal((5, 64))
    lhs = F.sum(0) ** 2 - (F * F).sum(0)
    rhs = 2 * sum(F[i] * F[j] for i in range(5) for j in range(i + 1, 5))
    np.testing.assert_allclose(lhs, rhs, atol=1e-12)


def test_embeddingmlp_sorted_concat_order():
    """DenseFeatures sorts by column name: the first Dense sees movieAvgRating first and
    userRatingStddev last."""
    spec = default_spec("embeddingmlp", emb_dim=2, n_movies=10, n_users=10)
    W = init_weights(spec, 0)
    f = synthetic_features(spec, 3, seed=0)
    x = O._embmlp_input(spec, W, f, np.float64)
    assert x.shape == (3, 7 + 10 * 2)
    np.testing.assert_allclose(x[:, 0], f["movieAvgRating"])
    np.testing.assert_allclose(x[:, -1], f["userRatingStddev"])
    np.testing.assert_allclose(x[:, 1 + 4 * 2], f["movieRatingCount"])


def test_missing_genre_is_zero_vector():
    spec = default_spec("embeddingmlp", emb_dim=2, n_movies=10, n_users=10)
    W = init_weights(spec, 0)
    f = synthetic_features(spec, 4, seed=0)
    f["movieGenre2"] = np.array(["", "NotAGenre", "Action", b"Drama"], dtype=object)
    x = O._embmlp_input(spec, W, f, np.float64)
    cols = slice(1 + 2, 1 + 4)
    assert np.all(x[0, cols] == 0) and np.all(x[1, cols] == 0)
    np.testing.assert_allclose(x[2, cols], W["movieGenre2_embedding"][1])
    np.testing.assert_allclose(x[3, cols], W["movieGenre2_embedding"][10])


def test_crossed_bucket_properties():
    b = O.crossed_bucket_array(np.arange(1, 200), np.arange(200, 1, -1), 10000)
    assert b.min() >= 0 and b.max() < 10000
    assert len(set(b.tolist())) > 150                   # spreads
    assert O.crossed_bucket(5, 7) != O.crossed_bucket(7, 5)   # leaf order matters
    assert O.crossed_bucket(5, 0) == O.crossed_bucket(5, 0)
    # FingerprintCat64 stays inside 64 bits
    assert 0 <= O.fingerprint_cat64(0xDECAFCAFFE, (1 << 64) - 1) < (1 << 64)


def test_identity_column_asserts_range():
    spec = default_spec("neuralcf")
    W = {n: np.zeros(s, np.float32) for n, s in weight_shapes(spec)}
    with pytest.raises(ValueError):
        O.neuralcf_forward(spec, W, {"movieId": np.array([1001]), "userId": np.array([1])})
    with pytest.raises(ValueError):
        O.neuralcf_forward(spec, W, {"movieId": np.array([1]), "userId": np.array([-1])})


@pytest.mark.parametrize("model", ["embeddingmlp", "widendeep", "neuralcf", "twotowers",
                                   "deepfm", "deepfm_v2", "din"])
def test_fp32_close_to_fp64(model, head_rows):
    """float32 restatement stays within 1e-5 of the float64 graph on real rows."""
    spec = default_spec(model)
    W = init_weights(spec, 11)
    sub = {k: v[:128] for k, v in head_rows.items()}
    p32, z32 = O.forward(spec, W, sub, np.float32)
    p64, z64 = O.forward(spec, W, sub, np.float64)
    assert p32.dtype == np.float32 and p32.shape == (128, 1)
    assert np.abs(p32 - p64).max() < 1e-5
    assert np.abs(z64).max() < 30 and np.abs(z64).std() > 0.05     # logits O(1): test has teeth


def test_predict_batches_like_keras(head_rows):
    spec = default_spec("neuralcf")
    W = init_weights(spec, 1)
    a = O.predict(spec, W, head_rows)
    b = O.predict(spec, W, head_rows, batch_size=12)     # make_csv_dataset(batch_size=12)
    np.testing.assert_allclose(a, b, atol=1e-7)
    assert a.shape == (512, 1) and a.dtype == np.float32


def test_fill_uniform_and_cosine_helpers():
    x = O.fill_uniform(np.arange(1000), seed=4, lo=-0.05, hi=0.05)
    assert x.dtype == np.float32 and x.min() >= -0.05 and x.max() < 0.05
    assert abs(float(x.mean())) < 0.01
    y = O.fill_uniform(np.array([17, 999]), seed=4, lo=-0.05, hi=0.05)
    assert y[0] == x[17] and y[1] == x[999]                 # counter based
    q = np.array([1.0, 0.0], np.float32)
    c = np.array([[1.0, 0.0], [0.0, 2.0], [-3.0, 0.0]], np.float32)
    np.testing.assert_allclose(O.cosine_similarity(q, c), [1.0, 0.0, -1.0])


def test_threaded_cpu_baselines_equal_the_oracle():
    """bench.py times `oracle/ctr_oracle_torch.py` as the CPU side; it must compute what the
    parity oracle computes."""
    from oracle import ctr_oracle_torch as OT
    from sparrowrecsys_b200.features import synthetic_features
    from sparrowrecsys_b200.spec import default_spec
    from sparrowrecsys_b200.weights import init_weights
    for model, kw in (("din", dict(emb_dim=16, hist_len=12, n_movies=500, n_users=300)),
                      ("deepfm", dict(n_movies=500, n_users=300)), ("neuralcf", {})):
        spec = default_spec(model, **kw)
        W = init_weights(spec, 4)
        f = synthetic_features(spec, 700, seed=5)
        fwd, how = OT.cpu_predictor(spec, W, threads=4)
        p, z = fwd(f)
        po, zo = O.forward(spec, W, f)
        assert p.shape == po.shape and np.abs(z - zo).max() < 2e-5, (model, how)
        assert ("torch" in how) == (model == "din")
