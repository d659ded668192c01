"""The plain-C restatement of the DIN graph (oracle/ctr_oracle_c.c, the CPU timing baseline of
bench.py) against the numpy oracle: same graph (DIN.py:125-167), so they agree to float32
reassociation error on every shape, including the reference's own (E=10, T=5)."""
import numpy as np
import pytest

from oracle import ctr_oracle as O
from oracle import ctr_oracle_cext as OC
from sparrowrecsys_b200.features import synthetic_features
from sparrowrecsys_b200.spec import default_spec
from sparrowrecsys_b200.weights import init_weights


@pytest.mark.parametrize("E,T,B", [(10, 5, 37), (32, 50, 300), (64, 200, 9), (20, 33, 1)])
def test_c_oracle_matches_numpy_oracle(E, T, B):
    spec = default_spec("din", emb_dim=E, hist_len=T, n_movies=3000, n_users=500)
    W = init_weights(spec, E + T)
    feats = synthetic_features(spec, B, seed=B)
    p, z = OC.din_predictor(spec, W, threads=3)(feats)
    po, zo = O.forward(spec, W, feats)
    assert p.shape == po.shape == (B, 1)
    assert np.abs(z - zo).max() <= 2e-5
    assert np.abs(p - po).max() <= 2e-6
    z64 = O.forward(spec, W, feats, dtype=np.float64)[1]
    assert np.abs(z - z64).max() <= 2e-5


def test_c_oracle_thread_count_does_not_change_results():
    spec = default_spec("din", emb_dim=32, hist_len=50, n_movies=3000, n_users=500)
    W = init_weights(spec, 3)
    feats = synthetic_features(spec, 257, seed=5)
    fwd = OC.din_predictor(spec, W)
    a = fwd(feats, 1)
    b = fwd(feats, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_c_oracle_rejects_out_of_range_ids():
    spec = default_spec("din", emb_dim=10, hist_len=5, n_movies=100, n_users=50)
    W = init_weights(spec, 0)
    feats = synthetic_features(spec, 8, seed=1)
    feats["userRatedMovie2"] = np.asarray(feats["userRatedMovie2"]).copy()
    feats["userRatedMovie2"][3] = 100
    with pytest.raises(ValueError):
        OC.din_predictor(spec, W)(feats)
