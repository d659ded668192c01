"""Latency of ONE ranking request (800 candidates: RecForYouProcess.java:34) through the C ABI, host
clock around the synchronous ctypes call only (the batch is encoded beforehand, host buffers pinned):

  rank_host       srs_rank_host: H2D of 800 full feature rows, forward, device sort-and-cut, k results
  rank_user_host  srs_rank_user_host: the user's row + 800 candidate ids, movie features gathered on the
                  device from the table in HBM, forward, sort-and-cut
  predict+sort    srs_predict_host into a pinned buffer + numpy argsort on the host (the reference's split)

for NeuralCF on the reference's shipped weights and for DIN at the BASELINE cfg 3 shape (T = 50, E = 32).
Also `srs_topk_device` alone (CUDA events).  Writes gpurun_out/rank_latency_r02.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden_weights                      # noqa: E402
from sparrowrecsys_b200 import _lib                           # noqa: E402
from sparrowrecsys_b200.features import encode_batch, synthetic_features   # noqa: E402
from sparrowrecsys_b200.model import CTRModel                 # noqa: E402
from sparrowrecsys_b200.ranking import topk_device            # noqa: E402
from sparrowrecsys_b200.spec import baseline_spec, default_spec   # noqa: E402
from sparrowrecsys_b200.weights import init_weights           # noqa: E402

lib = _lib.load()
out = {"topk_device_us": {}, "gpu": torch.cuda.get_device_name(0)}
for n, k in ((800, 10), (1024, 100), (4096, 100), (65536, 100)):
    s = torch.rand(n, device="cuda")
    for _ in range(5):
        topk_device(s, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        topk_device(s, k)
    e1.record()
    torch.cuda.synchronize()
    out["topk_device_us"]["n=%d,k=%d" % (n, k)] = round(e0.elapsed_time(e1) * 1e3 / reps, 2)


def timed(fn, reps=400, warm=40):
    for _ in range(warm):
        fn()
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.array(lat)
    return {"median": round(float(np.median(lat)), 1), "p10": round(float(np.percentile(lat, 10)), 1),
            "p99": round(float(np.percentile(lat, 99)), 1)}


def pinned(nbytes):
    t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    keep.append(t)
    return t.numpy()


keep = []
N, K = 800, 10
for name, spec, W in (("neuralcf", default_spec("neuralcf"), load_golden_weights("neuralcf_002")),
                      ("din_cfg3", baseline_spec("cfg3_din"), None)):
    if W is None:
        W = init_weights(spec, 2)
    with CTRModel(spec, W) as m:
        feats = synthetic_features(spec, N, seed=4)
        feats["userId"] = np.full(N, 10351, np.int32)
        enc = encode_batch(spec, feats, arena_alloc=pinned)
        hp = lambda a: None if a is None else a.ctypes.data
        T = m.hist_cols
        b = _lib.SrsBatch(N, T, hp(enc.movie_id), hp(enc.user_id), hp(enc.hist), hp(enc.movie_genre),
                          hp(enc.user_genre), hp(enc.numerics), None)
        idx = torch.empty(K, dtype=torch.int32).pin_memory()
        top = torch.empty(K, dtype=torch.float32).pin_memory()
        probs = torch.empty(N, dtype=torch.float32).pin_memory()
        h = m._h

        def rank_host():
            rc = lib.srs_rank_host(h, C.byref(b), K, idx.data_ptr(), top.data_ptr())
            assert rc == 0, rc

        def predict_sort():
            rc = lib.srs_predict_host(h, C.byref(b), probs.data_ptr(), None)
            assert rc == 0, rc
            np.argsort(-probs.numpy(), kind="stable")[:K]

        res = {"kernel": m.kernel_name, "rank_host": timed(rank_host), "predict_then_host_sort": timed(predict_sort)}
        os.environ["SRS_ZERO_COPY_SCORES"] = "0"
        with CTRModel(spec, W) as m0:                      # the general path (D2H copies + stream synchronise)
            h0 = m0._h

            def predict_sort_general():
                rc = lib.srs_predict_host(h0, C.byref(b), probs.data_ptr(), None)
                assert rc == 0, rc
                np.argsort(-probs.numpy(), kind="stable")[:K]
            res["predict_then_host_sort_general_path"] = timed(predict_sort_general)
        del os.environ["SRS_ZERO_COPY_SCORES"]
        # the same request as (user row, candidate ids): movie features resident in HBM
        rng = np.random.default_rng(1)
        V = spec.n_movies
        genres = rng.integers(-1, spec.n_genres, size=(V, 3)).astype(np.int32)
        nums = rng.random((V, 4)).astype(np.float32)
        _lib.check(lib.srs_model_set_movie_features(h, V, genres.ctypes.data, nums.ctypes.data))
        row = _lib.SrsUserRow()
        row.user_id = 10351
        for g in range(5):
            row.user_genre[g] = g
        row.user_numerics[0], row.user_numerics[1], row.user_numerics[2] = 3.5, 40.0, 0.9
        hist = np.ascontiguousarray(np.asarray(enc.hist[0] if T else np.zeros(0), np.int32))
        row.n_hist = T
        row.hist = hist.ctypes.data if T else None
        cand = np.ascontiguousarray(enc.movie_id)

        def rank_user():
            rc = lib.srs_rank_user_host(h, C.byref(row), cand.ctypes.data, N, K, idx.data_ptr(), top.data_ptr(), None)
            assert rc == 0, rc
        res["rank_user_host"] = timed(rank_user)
        out[name + "_800_candidates_us"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "rank_latency_r02.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
