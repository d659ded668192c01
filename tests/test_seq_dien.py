"""DIEN forward (TFRecModel/.../DIEN.py:154-256): the oracle's
restatement against a literal per-row loop (CPU), and `dien_kernel` against the oracle
through the C ABI (GPU).  PARITY UNPINNED like every graph here (no TensorFlow, no shipped
DIEN weights); on top of that the reference's own forward is not reproducible - its AUGRU
starts from a fresh GlorotUniform draw per call (DIEN.py:235-236) - so the initial state is
the stored tensor `augru_h0` in both the oracle and the kernel."""
import math

import numpy as np
import pytest

from oracle import ctr_oracle as O
from sparrowrecsys_b200.features import synthetic_features
from sparrowrecsys_b200.spec import NUMERIC_KEYS, default_spec
from sparrowrecsys_b200.weights import init_weights, numeric_rows, weight_shapes

PROB_ATOL = 2e-5
LOGIT_ATOL = 2e-4


def _stress_weights(spec, seed):
    """Reference initialisers keep the behaviour embeddings in +-0.05, where the recurrent
    part barely moves the logit; x10 (and a steeper attention unit) makes GRU / attention /
    AUGRU errors visible."""
    W = init_weights(spec, seed)
    W["embedding"] = W["embedding"] * np.float32(10.0)
    W["att_dense/kernel"] = W["att_dense/kernel"] * np.float32(4.0)
    W["att_out/kernel"] = W["att_out/kernel"] * np.float32(4.0)
    return W


# ---- oracle (CPU) --------------------------------------------------------------------------
def _sig(x):
    return 1.0 / (1.0 + math.exp(-x))


def _literal_dien_row(spec, W, f, i):
    """DIEN.py:154-256 for one row, scalar Python loops in float64."""
    E, T = spec.emb_dim, spec.hist_len
    w = {k: v.astype(np.float64) for k, v in W.items()}
    keys = sorted("userRatedMovie%d" % k for k in range(1, T + 1))
    hist = [int(f[k][i]) for k in keys]
    c = w["embedding"][int(f["movieId"][i])]
    h = [0.0] * E
    G = []
    for t in range(T):
        x = w["embedding"][hist[t]]
        mx = [sum(x[k] * w["gru/kernel"][k][j] for k in range(E)) + w["gru/bias"][0][j] for j in range(3 * E)]
        mh = [sum(h[k] * w["gru_recurrent/kernel"][k][j] for k in range(E)) + w["gru/bias"][1][j]
              for j in range(3 * E)]
        hn = []
        for e in range(E):
            z = _sig(mx[e] + mh[e])
            r = _sig(mx[E + e] + mh[E + e])
            hh = math.tanh(mx[2 * E + e] + r * mh[2 * E + e])
            hn.append(z * h[e] + (1 - z) * hh)
        if hist[t] != 0:                                   # masked step keeps state and output
            h = hn
        G.append(list(h))
    att = []
    for t in range(T):
        a = [_sig(sum(G[t][e] * c[e] * w["att_dense/kernel"][e][j] for e in range(E)) + w["att_dense/bias"][j])
             for j in range(32)]
        att.append(_sig(sum(a[j] * w["att_out/kernel"][j][0] for j in range(32)) + w["att_out/bias"][0]))

    def gate(g, x, hid):
        pre = [sum(x[k] * w["augru_%s_input/kernel" % g][k][e] for k in range(E)) + w["augru_%s_input/bias" % g][e]
               + sum(hid[k] * w["augru_%s_hidden/kernel" % g][k][e] for k in range(E)) for e in range(E)]
        return [sum(pre[k] * w["augru_%s_act/kernel" % g][k][e] for k in range(E)) + w["augru_%s_act/bias" % g][e]
                for e in range(E)]
    u = list(w["augru_h0"][0])
    for t in range(T):
        r = [_sig(v) for v in gate("r", G[t], u)]
        z = [_sig(v) for v in gate("z", G[t], u)]
        hn = [math.tanh(v) for v in gate("h", G[t], [u[e] * z[e] for e in range(E)])]
        u = [(1 - att[t] * r[e]) * u[e] + att[t] * r[e] * hn[e] for e in range(E)]
    gi = lambda key: O._GIDX.get(f[key][i], -1)
    emb = lambda name, idx: list(w[name][idx]) if idx >= 0 else [0.0] * E
    profile = [float(f["userAvgRating"][i])] + emb("userGenre1_embedding", gi("userGenre1")) \
        + emb("userId_embedding", int(f["userId"][i])) + [float(f["userRatingCount"][i]), float(f["userRatingStddev"][i])]
    context = [float(f["movieAvgRating"][i])] + emb("movieGenre1_embedding", gi("movieGenre1")) \
        + [float(f["movieRatingCount"][i]), float(f["movieRatingStddev"][i]), float(f["releaseYear"][i])]
    x = np.array(u + list(c) + profile + context)
    for d, a in (("dense", "prelu"), ("dense_1", "prelu_1")):
        x = x @ w[d + "/kernel"] + w[d + "/bias"]
        x = np.maximum(x, 0) - w[a + "/alpha"] * np.maximum(-x, 0)
    return float((x @ w["dense_2/kernel"] + w["dense_2/bias"])[0])


def _with_masks(f, T):
    """Histories with a zero in the middle, all zeros, and no zeros."""
    keys = sorted("userRatedMovie%d" % k for k in range(1, T + 1))
    f = {k: np.array(v, copy=True) for k, v in f.items()}
    for k in keys:
        f[k][0] = 0                                        # row 0: nothing valid
        f[k][1] = max(1, int(f[k][1]))                     # row 1: everything valid
    if T >= 3:
        f[keys[0]][2], f[keys[1]][2], f[keys[2]][2] = 7, 0, 9   # row 2: a hole in the middle
    return f


def test_oracle_matches_literal_loops_including_masked_steps():
    spec = default_spec("dien", emb_dim=3, hist_len=4, n_movies=40, n_users=30)
    W = _stress_weights(spec, 7)
    f = _with_masks(synthetic_features(spec, 6, seed=1), 4)
    _, z = O.forward(spec, W, f, dtype=np.float64)
    for i in range(6):
        assert abs(z[i, 0] - _literal_dien_row(spec, W, f, i)) < 1e-10, i
    _, z32 = O.forward(spec, W, f)
    assert np.abs(z32 - z).max() < 5e-6


def test_sequence_part_reaches_the_logit():
    """The parity tests below have teeth only if GRU / attention / AUGRU errors move z."""
    spec = default_spec("dien")
    W = _stress_weights(spec, 3)
    f = synthetic_features(spec, 128, seed=2)
    _, z = O.forward(spec, W, f)
    for name in ("gru/kernel", "gru_recurrent/kernel", "att_dense/kernel", "att_out/kernel",
                 "augru_r_input/kernel", "augru_z_hidden/kernel", "augru_h_act/kernel", "augru_h0"):
        W2 = dict(W)
        W2[name] = W[name] * np.float32(1.05)
        assert np.abs(O.forward(spec, W2, f)[1] - z).max() > 2e-4, name


def test_weight_inventory_and_numeric_rows():
    spec = default_spec("dien")
    shapes = dict(weight_shapes(spec))
    assert shapes["gru/bias"] == (2, 30) and shapes["augru_h0"] == (1, 10)
    assert shapes["dense/kernel"] == (57, 128) and shapes["att_dense/kernel"] == (10, 32)
    assert spec.kind == 7 and spec.required_keys() == default_spec("din").required_keys()
    assert spec.bytes_per_inference() == default_spec("din").bytes_per_inference()
    # numeric_rows must name the dense/kernel rows the oracle's concat puts the numerics in
    W = init_weights(spec, 1)
    rows = numeric_rows(spec)["dense/kernel"]
    f = synthetic_features(spec, 4, seed=5)
    for j, key in enumerate(NUMERIC_KEYS):
        W2 = {k: (np.zeros_like(v) if k == "dense/kernel" else v) for k, v in W.items()}
        W2["dense/bias"] = np.zeros_like(W["dense/bias"])
        W2["dense/kernel"][rows[j], 0] = 1.0
        hidden0 = f[key].astype(np.float32)                # unit 0 of Dense128 == that numeric
        # read it back through prelu (positive numerics pass unchanged) and a one-hot Dense64/Dense1
        W2["dense_1/kernel"] = np.zeros_like(W["dense_1/kernel"]); W2["dense_1/kernel"][0, 0] = 1.0
        W2["dense_1/bias"] = np.zeros_like(W["dense_1/bias"])
        W2["dense_2/kernel"] = np.zeros_like(W["dense_2/kernel"]); W2["dense_2/kernel"][0, 0] = 1.0
        W2["dense_2/bias"] = np.zeros_like(W["dense_2/bias"])
        _, z = O.forward(spec, W2, f)
        np.testing.assert_allclose(z[:, 0], hidden0, rtol=1e-6), key


# ---- CUDA kernel (GPU) -------------------------------------------------------------------------
def _compare(spec, W, feats, prob_atol=PROB_ATOL, logit_atol=LOGIT_ATOL):
    from sparrowrecsys_b200.model import CTRModel
    with CTRModel(spec, W, device=0) as m:
        assert m.kernel_name == "dien_kernel"
        p, z = m.predict_with_logits(feats)
    po, zo = O.forward(spec, W, feats)
    assert p.shape == po.shape == (len(feats["movieId"]), 1) and p.dtype == np.float32
    assert np.abs(z - zo).max() <= logit_atol, "logit err %g" % np.abs(z - zo).max()
    assert np.abs(p - po).max() <= prob_atol, "prob err %g" % np.abs(p - po).max()
    return p, z


@pytest.mark.gpu
def test_dien_reference_shape_on_bundled_rows(head_rows):
    spec = default_spec("dien")
    _compare(spec, init_weights(spec, 107), head_rows)
    _compare(spec, _stress_weights(spec, 108), head_rows)


@pytest.mark.gpu
@pytest.mark.parametrize("E,T", [(10, 5), (16, 7), (32, 33), (12, 64), (8, 1), (4, 3), (20, 50)])
def test_dien_shapes(E, T):
    spec = default_spec("dien", emb_dim=E, hist_len=T, n_movies=5000, n_users=3000)
    feats = _with_masks(synthetic_features(spec, 333, seed=T), T)
    _compare(spec, _stress_weights(spec, E * 100 + T), feats, logit_atol=5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 2, 31, 33, 65, 129])
def test_dien_ragged_batch_sizes(B, head_rows):
    spec = default_spec("dien")
    _compare(spec, _stress_weights(spec, 9), {k: v[:B] for k, v in head_rows.items()})


@pytest.mark.gpu
def test_dien_rows_are_independent_and_ids_checked():
    from sparrowrecsys_b200.model import CTRModel
    spec = default_spec("dien", emb_dim=16, hist_len=12, n_movies=2000, n_users=500)
    W = _stress_weights(spec, 21)
    f = synthetic_features(spec, 300, seed=8)
    perm = np.random.default_rng(0).permutation(300)
    with CTRModel(spec, W) as m:
        p = m.predict(f)
        q = m.predict({k: v[perm] for k, v in f.items()})
        assert np.array_equal(p[perm], q)                  # bit exact wherever a row sits
        bad = {k: np.array(v, copy=True) for k, v in f.items()}
        bad["userRatedMovie3"][5] = spec.n_movies
        with pytest.raises(ValueError):
            m.predict(bad)
        assert np.array_equal(m.predict(f), p)             # the error flag does not stick


@pytest.mark.gpu
def test_dien_through_tfrecmodel_surface_and_rank(head_rows):
    from tfrecmodel import dien
    spec = dien.spec()
    W = _stress_weights(spec, 4)
    dien.load(weights=W)
    sub = {k: v[:200] for k, v in head_rows.items()}
    p = dien.predict(sub)
    po, _ = O.forward(spec, W, sub)
    assert np.abs(p - po).max() <= PROB_ATOL
    idx, top = dien.model.rank(sub, 10)
    ridx, rtop = O.rank_topk(p[:, 0], 10)
    assert np.array_equal(idx, ridx) and np.array_equal(top, rtop)
    dien.model.close()


def test_oracle_gru_agrees_with_torch_gru_cell():
    """Independent check of the GRU restatement: Keras `GRU(reset_after=True)` and `torch.nn.GRU`
    are the same cell up to gate order (Keras z | r | h, torch r | z | n) and weight transposition;
    with every history id valid (no masked step) the oracle's hidden states must match torch's."""
    import torch
    E, T, B = 6, 5, 9
    spec = default_spec("dien", emb_dim=E, hist_len=T, n_movies=60, n_users=30)
    W = _stress_weights(spec, 12)
    rng = np.random.default_rng(0)
    hist = rng.integers(1, 60, size=(B, T))
    X = W["embedding"][hist].astype(np.float64)                       # [B, T, E]
    perm = np.r_[E:2 * E, 0:E, 2 * E:3 * E]                           # z|r|h -> r|z|n
    gru = torch.nn.GRU(E, E, batch_first=True).double()
    with torch.no_grad():
        gru.weight_ih_l0.copy_(torch.from_numpy(W["gru/kernel"].astype(np.float64)[:, perm].T.copy()))
        gru.weight_hh_l0.copy_(torch.from_numpy(W["gru_recurrent/kernel"].astype(np.float64)[:, perm].T.copy()))
        gru.bias_ih_l0.copy_(torch.from_numpy(W["gru/bias"].astype(np.float64)[0, perm].copy()))
        gru.bias_hh_l0.copy_(torch.from_numpy(W["gru/bias"].astype(np.float64)[1, perm].copy()))
        G_torch = gru(torch.from_numpy(X))[0].numpy()
    # the oracle's recurrence, same code path as dien_forward (mask all true)
    K, U = W["gru/kernel"].astype(np.float64), W["gru_recurrent/kernel"].astype(np.float64)
    bx, bh = W["gru/bias"].astype(np.float64)
    h = np.zeros((B, E))
    for t in range(T):
        mx, mh = X[:, t] @ K + bx, h @ U + bh
        z = O.sigmoid(mx[:, :E] + mh[:, :E])
        r = O.sigmoid(mx[:, E:2 * E] + mh[:, E:2 * E])
        hh = np.tanh(mx[:, 2 * E:] + r * mh[:, 2 * E:])
        h = z * h + (1 - z) * hh
        assert np.abs(h - G_torch[:, t]).max() < 1e-12, t
    # and dien_forward itself uses that recurrence: perturbing one GRU weight moves its logits
    f = synthetic_features(spec, B, seed=1)
    for k in range(1, T + 1):
        f["userRatedMovie%d" % k] = hist[:, k - 1].astype(np.int32)
    _, z0 = O.forward(spec, W, f, dtype=np.float64)
    for i in range(B):
        assert abs(z0[i, 0] - _literal_dien_row(spec, W, f, i)) < 1e-10
