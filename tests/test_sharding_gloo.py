"""N>1 host logic on CPU: world_size-2 gloo run of the row-shard + score all-gather."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import ctr_oracle as O
    from sparrowrecsys_b200.features import synthetic_features
    from sparrowrecsys_b200.sharding import predict_sharded, shard_bounds
    from sparrowrecsys_b200.spec import default_spec
    from sparrowrecsys_b200.weights import init_weights
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = default_spec("neuralcf")
        W = init_weights(spec, 3)
        feats = synthetic_features(spec, 1001, seed=3)       # odd size: ragged shards
        # stand-in scorer (the oracle): the CUDA model is exercised in the -m gpu tests
        score = lambda f: torch.from_numpy(O.forward(spec, W, f)[0][:, 0].copy())
        full = predict_sharded(score, feats)
        ref = O.forward(spec, W, feats)[0][:, 0]
        lo, hi = shard_bounds(1001, world, rank)
        # one ranking call spanning both ranks: local top-k, one all-gather, merge.  Scores are
        # quantised so that ties straddle the shard boundary; stand-ins: the oracle's ranking.
        from sparrowrecsys_b200.sharding import rank_sharded
        coarse = np.round(ref * 50).astype(np.float32) / 50

        def local_rank(f, k):
            s = np.round(O.forward(spec, W, f)[0][:, 0] * 50).astype(np.float32) / 50
            i, t = O.rank_topk(s, k)
            return torch.from_numpy(i), torch.from_numpy(t)

        def merge(scores, k):
            i, t = O.rank_topk(scores.numpy(), k)
            return torch.from_numpy(i), torch.from_numpy(t)
        ok_rank = True
        for size in (1, 50, 600, 5000):
            pos, top = rank_sharded(local_rank, feats, size, merge_fn=merge)
            ridx, rtop = O.rank_topk(coarse, size)
            ok_rank &= bool(np.array_equal(pos.numpy(), ridx) and np.array_equal(top.numpy(), rtop))
        q.put((rank, bool(np.array_equal(full.numpy(), ref)), hi - lo, ok_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]         # bit-identical to the unsharded scores
    assert sorted(r[2] for r in res) == [500, 501]
    assert [r[3] for r in res] == [True, True]         # sharded ranking == single-rank ranking
