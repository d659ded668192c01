"""History ids as uint16 on the host side of the C ABI (`srs_batch::hist16`): the packed
layout (CPU) and bit-identical scores to the int32 form through every host entry point (GPU)."""
import ctypes as C

import numpy as np
import pytest

from sparrowrecsys_b200.features import encode_batch, synthetic_features
from sparrowrecsys_b200.spec import baseline_spec, default_spec
from sparrowrecsys_b200.weights import init_weights


def test_narrow_batch_layout_is_the_packed_order():
    spec = baseline_spec("cfg3_din")
    f = synthetic_features(spec, 37, seed=1)               # odd row count: 2-byte tail padding
    wide, nar = encode_batch(spec, f), encode_batch(spec, f, narrow_ids=True)
    assert nar.hist.dtype == np.uint16 and np.array_equal(wide.hist, nar.hist)
    for name in ("movie_id", "user_id", "movie_genre", "user_genre", "numerics"):
        assert np.array_equal(getattr(wide, name), getattr(nar, name))
    base = nar.movie_id.ctypes.data
    hist_bytes = (37 * 50 * 2 + 3) & ~3
    assert nar.user_id.ctypes.data - base == 37 * 4
    assert nar.hist.ctypes.data - base == 37 * 8
    assert nar.movie_genre.ctypes.data - base == 37 * 8 + hist_bytes
    assert nar.numerics.ctypes.data - base == 37 * 8 + hist_bytes + 37 * 32
    half = nar.slice(10, 20)
    assert half.hist.dtype == np.uint16 and half.hist.shape == (10, 50)
    with pytest.raises(ValueError):
        big = default_spec("din", n_movies=70000)
        encode_batch(big, synthetic_features(big, 4, seed=0), narrow_ids=True)
    ncf = default_spec("neuralcf")                         # no history: the flag is a no-op
    assert encode_batch(ncf, {"movieId": np.array([1]), "userId": np.array([2])}, narrow_ids=True).hist is None


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,B", [("cfg3_din", 4096), ("cfg3_din", 37), ("cfg4_widendeep", 1001),
                                   ("ref_dien", 333)])
def test_hist16_scores_equal_int32_scores(cfg, B):
    from sparrowrecsys_b200.model import CTRModel
    spec = baseline_spec(cfg)
    W = init_weights(spec, 5)
    f = synthetic_features(spec, B, seed=6)
    with CTRModel(spec, W) as wide, CTRModel(spec, W, narrow_ids=True) as nar:
        assert nar.narrow_ids and not wide.narrow_ids
        p, z = wide.predict_with_logits(f)
        q, y = nar.predict_with_logits(f)
        assert np.array_equal(p, q) and np.array_equal(z, y)
        assert np.array_equal(nar.predict(f, batch_size=max(B // 5, 1)), p)    # pipelined slots
        i0, t0 = wide.rank(f, 17)
        i1, t1 = nar.rank(f, 17)
        assert np.array_equal(i0, i1) and np.array_equal(t0, t1)
        bad = {k: np.array(v, copy=True) for k, v in f.items()}
        bad["userRatedMovie1"][B // 2] = spec.n_movies                          # still range-checked
        with pytest.raises(ValueError):
            nar.predict(bad)


@pytest.mark.gpu
def test_hist16_strided_host_arrays_and_rejections():
    """Not-packed host arrays with a row stride, and the two refusals: a device batch and a
    vocabulary that does not fit 16 bits."""
    import torch
    from sparrowrecsys_b200 import _lib
    from sparrowrecsys_b200.model import CTRModel, _host_struct
    spec = default_spec("din", emb_dim=32, hist_len=20, n_movies=3000, n_users=2000)
    W = init_weights(spec, 3)
    f = synthetic_features(spec, 200, seed=4)
    enc = encode_batch(spec, f)
    lib = _lib.load()
    with CTRModel(spec, W) as m:
        ref = m.predict(f)[:, 0]
        wide16 = np.zeros((200, 32), np.uint16)             # stride 32 > T = 20
        wide16[:, :20] = enc.hist
        keep = []
        b = _host_struct(enc, keep)
        b.hist, b.hist16, b.hist_stride = None, wide16.ctypes.data, 32
        out = np.empty(200, np.float32)
        _lib.check(lib.srs_predict_host(m._h, C.byref(b), out.ctypes.data, None))
        assert np.array_equal(out, ref)
        probs = torch.empty(200, dtype=torch.float32, device="cuda:0")
        assert lib.srs_predict_device(m._h, C.byref(b), probs.data_ptr(), None, None) == _lib.SRS_ERR_INVALID
    big = default_spec("din", emb_dim=32, hist_len=20, n_movies=70000, n_users=2000)
    fb = synthetic_features(big, 16, seed=1)
    eb = encode_batch(big, fb)
    with CTRModel(big, init_weights(big, 1), narrow_ids=True) as m:
        assert not m.narrow_ids                             # the flag is dropped, not an error
        keep = []
        b = _host_struct(eb, keep)
        h16 = eb.hist.astype(np.uint16)
        b.hist, b.hist16 = None, h16.ctypes.data
        out = np.empty(16, np.float32)
        assert lib.srs_predict_host(m._h, C.byref(b), out.ctypes.data, None) == _lib.SRS_ERR_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,B", [("cfg3_din", 4096), ("cfg3_din", 4736 + 33), ("cfg4_widendeep", 8192),
                                   ("cfg2_deepfm", 4096)])
def test_sm_limit_does_not_change_scores(cfg, B):
    """`srs_model_set_sm_limit`: fewer CTAs per launch, each walking more row groups - the
    scores must be the same bits (bench.py --streams relies on it)."""
    from sparrowrecsys_b200.model import CTRModel
    spec = baseline_spec(cfg)
    f = synthetic_features(spec, B, seed=11)
    with CTRModel(spec, init_weights(spec, 5)) as m:
        ref = m.predict(f)
        for n in (74, 37, 5, 1, 0, 1000):
            m.set_sm_limit(n)
            assert np.array_equal(m.predict(f), ref), n


@pytest.mark.gpu
@pytest.mark.skipif(not __import__("os").environ.get("SRS_TEST_ZERO_COPY"),
                    reason="SRS_ZERO_COPY_SCORES is experimental: opt in with SRS_TEST_ZERO_COPY=1")
def test_zero_copy_scores_equal_copied_scores(monkeypatch):
    """Kernels writing the scores of a host batch straight into the caller's pinned buffer
    (SRS_ZERO_COPY_SCORES=1, read at model creation) must give the bytes the copy gives."""
    import torch
    from sparrowrecsys_b200 import _lib
    from sparrowrecsys_b200.model import CTRModel, _host_struct
    spec = baseline_spec("cfg3_din")
    W = init_weights(spec, 5)
    f = synthetic_features(spec, 4096, seed=6)
    enc = encode_batch(spec, f)
    with CTRModel(spec, W) as m:
        ref = m.predict(f)[:, 0]
    monkeypatch.setenv("SRS_ZERO_COPY_SCORES", "1")
    with CTRModel(spec, W) as m:
        assert np.array_equal(m.predict(f)[:, 0], ref)            # pageable numpy buffer: falls back
        out = torch.zeros(4096, dtype=torch.float32).pin_memory()
        keep = []
        b = _host_struct(enc, keep)
        _lib.check(_lib.load().srs_predict_host(m._h, C.byref(b), out.data_ptr(), None))
        assert np.array_equal(out.numpy(), ref)
        outs = [torch.zeros(1024, dtype=torch.float32).pin_memory() for _ in range(4)]
        m.predict_batches([enc.slice(i * 1024, (i + 1) * 1024) for i in range(4)], [o.numpy() for o in outs])
        assert np.array_equal(np.concatenate([o.numpy() for o in outs]), ref)
