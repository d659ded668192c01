"""Ranking tail of the online rankers: the oracle's restatement
of `sorted(comparingByValue(reverseOrder()))` + `subList(0, size)`
(online/recprocess/RecForYouProcess.java:56-59,92-94) on CPU, and the device ranking
(`srs_topk_device`, `srs_rank_host`) against it, bit exact, on the GPU."""
import numpy as np
import pytest

from oracle import ctr_oracle as O


# ---- oracle (CPU) ------------------------------------------------------------------------
def test_java_double_compare_total_order():
    inf, nan = float("inf"), float("nan")
    chain = [-inf, -1.5, -0.0, 0.0, 1e-30, 2.0, inf, nan]
    for i, a in enumerate(chain):
        for j, b in enumerate(chain):
            assert O.java_double_compare(a, b) == (i > j) - (i < j), (a, b)


def test_rank_topk_matches_default_ranker_scores():
    # RecForYouProcess.java:86-90: the default branch scores candidate i with (n - i),
    # so the ranked list is the candidate list itself.
    n = 800
    idx, top = O.rank_topk(np.arange(n, 0, -1, dtype=np.float32), 10)
    assert idx.tolist() == list(range(10)) and top.tolist() == list(range(800, 790, -1))


def test_rank_topk_ties_nan_and_cut():
    s = np.array([0.5, np.nan, 0.5, -0.0, 0.0, 1.0, -np.inf, 0.5], np.float32)
    idx, top = O.rank_topk(s, 100)                  # size > n: the whole list (:56-59)
    assert idx.tolist() == [1, 5, 0, 2, 7, 4, 3, 6]
    assert np.isnan(top[0]) and top[1] == 1.0
    assert O.rank_topk(s, 3)[0].tolist() == [1, 5, 0]
    assert O.rank_topk(s, 0)[0].size == 0 and O.rank_topk(np.zeros(0, np.float32), 5)[0].size == 0


def test_rank_topk_agrees_with_numpy_on_distinct_scores():
    rng = np.random.default_rng(5)
    s = rng.standard_normal(3000).astype(np.float32)
    s = np.unique(s)
    rng.shuffle(s)
    idx, _ = O.rank_topk(s, 200)
    assert np.array_equal(idx, np.argsort(-s.astype(np.float64), kind="stable")[:200])


# ---- device (GPU) --------------------------------------------------------------------------
def _device_topk(s, k):
    import torch
    from sparrowrecsys_b200.ranking import topk_device
    idx, top = topk_device(torch.from_numpy(s).cuda(), k)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), top.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1, 1), (2, 5), (31, 31), (33, 7), (800, 10), (800, 800),
                                 (1000, 1), (4096, 100), (4097, 50), (10000, 10000),
                                 (70000, 300)])
def test_topk_device_bit_exact(n, k):
    rng = np.random.default_rng(n * 7 + k)
    s = rng.standard_normal(n).astype(np.float32)
    s[rng.integers(0, n, n // 3)] = np.float32(0.25)          # ties
    if n > 8:
        s[3], s[5], s[6], s[7] = np.nan, -0.0, 0.0, np.inf
        s[n - 1] = -np.inf
    idx, top = _device_topk(s, k)
    ridx, rtop = O.rank_topk(s, k)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(top.view(np.uint32), rtop.view(np.uint32))


@pytest.mark.gpu
def test_topk_device_large_properties():
    """Size-independent properties at a size the Python oracle would be slow for."""
    n, k = 1 << 20, 5000
    rng = np.random.default_rng(11)
    s = rng.random(n, dtype=np.float32)                        # many exact ties at 2^-24 grid
    idx, top = _device_topk(s, k)
    assert len(np.unique(idx)) == k and np.array_equal(s[idx], top)
    assert np.all(np.diff(top) <= 0)
    same = np.diff(top) == 0
    assert np.all(np.diff(idx.astype(np.int64))[same] > 0)     # ties by position
    assert top[-1] >= np.partition(s, n - k)[n - k]            # nothing better left out
    assert np.array_equal(idx, np.lexsort((np.arange(n), -s.astype(np.float64)))[:k])


@pytest.mark.gpu
def test_rank_host_neuralcf_golden_and_oracle():
    """RecForYouProcess.getRecList with model "nerualcf": one user, 800 candidates, the
    reference's shipped weights; positions bit exact against the oracle's ranking of the
    device scores, scores against the oracle forward."""
    from conftest import load_golden_weights
    from sparrowrecsys_b200.model import CTRModel
    from sparrowrecsys_b200.spec import default_spec
    W = load_golden_weights("neuralcf_002")
    spec = default_spec("neuralcf")
    f = {"movieId": np.arange(1, 801, dtype=np.int32), "userId": np.full(800, 10351, np.int32)}
    with CTRModel(spec, W) as m:
        p = m.predict(f)[:, 0]
        idx, top = m.rank(f, 10)
        idx_all, _ = m.rank(f, 5000)
    ridx, rtop = O.rank_topk(p, 10)
    assert np.array_equal(idx, ridx) and np.array_equal(top, rtop)
    assert np.array_equal(idx_all, O.rank_topk(p, 800)[0])
    po, _ = O.forward(spec, W, f)
    assert np.abs(top - po[idx, 0]).max() < 1e-6
    assert set(idx.tolist()) == set(O.rank_topk(po[:, 0], 10)[0].tolist())


@pytest.mark.gpu
def test_rank_host_din_large_batch():
    from sparrowrecsys_b200.features import synthetic_features
    from sparrowrecsys_b200.model import CTRModel
    from sparrowrecsys_b200.spec import baseline_spec
    from sparrowrecsys_b200.weights import init_weights
    spec = baseline_spec("cfg3_din")
    feats = synthetic_features(spec, 6000, seed=9)            # > 4096: the multi-CTA sort
    with CTRModel(spec, init_weights(spec, 2)) as m:
        p = m.predict(feats)[:, 0]
        idx, top = m.rank(feats, 64)
    ridx, rtop = O.rank_topk(p, 64)
    assert np.array_equal(idx, ridx) and np.array_equal(top, rtop)


@pytest.mark.gpu
def test_rank_by_embedding_matches_reference_emb_ranker():
    """ranker(..., "emb"): Embedding.calculateSimilarity + sort (RecForYouProcess.java:74-78)."""
    from sparrowrecsys_b200.ranking import rank_by_embedding
    rng = np.random.default_rng(3)
    q = rng.standard_normal(10).astype(np.float32)
    c = rng.standard_normal((800, 10)).astype(np.float32)
    idx, top = rank_by_embedding(q, c, 20)
    ref = O.cosine_similarity(q, c)
    ridx, _ = O.rank_topk(ref.astype(np.float32), 20)
    assert np.array_equal(idx, ridx)
    assert np.abs(top - ref[idx]).max() < 1e-6


@pytest.mark.gpu
def test_rank_sharded_on_nccl_single_rank_group():
    """`sharding.rank_sharded` with the real pieces (CUDA model, device top-k, NCCL) on a
    one-rank group; the two-rank exchange logic is covered on CPU (tests/test_sharding_gloo.py)."""
    import os
    import torch
    import torch.distributed as dist
    from conftest import load_golden_weights
    from sparrowrecsys_b200.model import CTRModel
    from sparrowrecsys_b200.ranking import topk_device
    from sparrowrecsys_b200.sharding import rank_sharded
    from sparrowrecsys_b200.spec import default_spec
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29700 + os.getpid() % 200))
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        spec = default_spec("neuralcf")
        f = {"movieId": np.arange(1, 901, dtype=np.int32), "userId": np.full(900, 10351, np.int32)}
        with CTRModel(spec, load_golden_weights("neuralcf_002")) as m:
            def local_rank(shard, k):
                batch = m.to_device(shard)
                probs = torch.empty(batch.B, dtype=torch.float32, device="cuda:0")
                m.predict_device(batch, probs)
                return topk_device(probs, k)
            pos, top = rank_sharded(local_rank, f, 25)
            torch.cuda.synchronize()
            p = m.predict(f)[:, 0]
        ridx, rtop = O.rank_topk(p, 25)
        assert np.array_equal(pos.cpu().numpy(), ridx) and np.array_equal(top.cpu().numpy(), rtop)
    finally:
        dist.destroy_process_group()


# ---- the sort network of csrc/topk.cu, re-derived on the CPU ----------------------------------
def _rank_keys(s):
    """rank_key() of topk.cu: ascending key order = the ranking."""
    u = s.view(np.uint32).astype(np.uint64)
    mono = np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000)
    mono = np.where(np.isnan(s), 0xFFFFFFFF, mono)
    return ((~mono & 0xFFFFFFFF) << np.uint64(32)) | np.arange(len(s), dtype=np.uint64)


def _bitonic(keys, chunk):
    """The launch sequence of launch_topk for NP keys: per-chunk shared-memory stages, global
    compare-exchange steps for strides >= chunk (pair_lo / direction rule copied from the kernels)."""
    NP = len(keys)
    pair_lo = lambda t, j: ((t & ~(j - 1)) << 1) | (t & (j - 1))

    def step(a, lo, j, w, base=0):
        hi = lo | j
        x, y = a[lo].copy(), a[hi].copy()
        swap = (x > y) == (((base + lo) & w) == 0)
        a[lo], a[hi] = np.where(swap, y, x), np.where(swap, x, y)

    def chunk_sort(first_w):
        for c in range(NP // chunk):
            a, base = keys[c * chunk:(c + 1) * chunk], c * chunk
            t = np.arange(chunk // 2)
            w_end = chunk if first_w == 2 else first_w
            w = first_w
            while True:
                j = min(w, chunk) >> 1
                while j > 0:
                    step(a, pair_lo(t, j), j, w, base)
                    j >>= 1
                if w == w_end:
                    break
                w <<= 1
    chunk_sort(2)
    w = 2 * chunk
    while w <= NP:
        j = w >> 1
        while j >= chunk:
            step(keys, pair_lo(np.arange(NP // 2), j), j, w)
            j >>= 1
        chunk_sort(w)
        w <<= 1
    return keys


@pytest.mark.parametrize("n,chunk", [(5, 32), (800, 1024), (100, 16), (1000, 64), (4097, 256)])
def test_sort_network_and_key_order_match_the_java_comparator(n, chunk):
    rng = np.random.default_rng(n)
    s = rng.standard_normal(n).astype(np.float32)
    s[rng.integers(0, n, n // 3)] = np.float32(0.5)
    if n > 8:
        s[1], s[2], s[3], s[4] = np.nan, -0.0, 0.0, -np.inf
    NP = chunk
    while NP < n:
        NP <<= 1
    keys = np.full(NP, np.iinfo(np.uint64).max, np.uint64)
    keys[:n] = _rank_keys(s)
    order = (_bitonic(keys, chunk)[:n] & np.uint64(0xFFFFFFFF)).astype(np.int32)
    assert np.array_equal(order, O.rank_topk(s, n)[0])


# ---- the "emb" ranker on the reference's shipped embeddings -----------------------------------
# Known answers: top 5 of all 881 shipped movie embeddings for the first three users of
# userEmb.csv, by a literal restatement of Embedding.calculateSimilarity (float products,
# double sums; tests/golden/make_golden.py copies the two files from the reference).
EMB_KNOWN = {
    10292: ([875, 15, 424, 433, 387], [0.9197663, 0.854994, 0.8533278, 0.8480263, 0.8424886]),
    19125: ([293, 555, 868, 288, 16], [0.7498215, 0.7454202, 0.718767, 0.7139385, 0.7110246]),
    26985: ([15, 415, 433, 170, 23], [0.8829988, 0.8794499, 0.8699859, 0.8569337, 0.8543974]),
}


def _shipped_embeddings():
    import os
    from conftest import GOLDEN
    from sparrowrecsys_b200.ranking import load_embeddings_csv
    mid, M = load_embeddings_csv(os.path.join(GOLDEN, "item2vecEmb.csv"))
    uid, U = load_embeddings_csv(os.path.join(GOLDEN, "userEmb_head.csv"))
    return mid, M, uid, U


def _java_cosine(a, b):
    """online/model/Embedding.java:33-47, statement for statement."""
    import math
    dot = n1 = n2 = 0.0
    for x, y in zip(a, b):
        dot += float(np.float32(x) * np.float32(y))
        n1 += float(np.float32(x) * np.float32(x))
        n2 += float(np.float32(y) * np.float32(y))
    return dot / (math.sqrt(n1) * math.sqrt(n2))


def test_emb_ranker_oracle_on_shipped_embeddings():
    mid, M, uid, U = _shipped_embeddings()
    assert M.shape == (881, 10) and U.shape == (20, 10) and mid[0] == 710 and uid[0] == 10292
    for k in range(3):
        ref = O.cosine_similarity(U[k], M)
        lit = np.array([_java_cosine(U[k], m) for m in M[:60]])
        assert np.abs(lit - ref[:60]).max() < 1e-15
        idx, _ = O.rank_topk(ref.astype(np.float32), 5)
        ids, sims = EMB_KNOWN[int(uid[k])]
        assert mid[idx].tolist() == ids
        np.testing.assert_allclose(ref[idx], sims, atol=5e-8)


@pytest.mark.gpu
def test_emb_ranker_device_on_shipped_embeddings():
    from sparrowrecsys_b200.ranking import rank_by_embedding
    mid, M, uid, U = _shipped_embeddings()
    for k in range(3):
        idx, top = rank_by_embedding(U[k], M, 5)
        ids, sims = EMB_KNOWN[int(uid[k])]
        assert mid[idx].tolist() == ids
        np.testing.assert_allclose(top, sims, atol=1e-6)
