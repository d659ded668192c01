"""Worker of tests/test_multi_gpu.py: run under torchrun with one rank per GPU (NCCL).

Checks, on every rank, that a ranking call sharded by rows over the ranks gives exactly the
single-GPU result: (1) the fused score exchange (srs_gather_*: scores stored into every rank's
buffer by the forward kernel) over several steps, for a kernel with the in-kernel signal (DIN) and
one with the separate signal kernel (NeuralCF); (2) sharding.gather_scores (NCCL all-gather);
(3) sharding.rank_sharded (local top-k, one all-gather of size x N pairs, merge)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sparrowrecsys_b200 import sharding                      # noqa: E402
from sparrowrecsys_b200.features import encode_batch, synthetic_features   # noqa: E402
from sparrowrecsys_b200.model import CTRModel                # noqa: E402
from sparrowrecsys_b200.ranking import topk_device           # noqa: E402
from sparrowrecsys_b200.spec import baseline_spec, default_spec   # noqa: E402
from sparrowrecsys_b200.weights import init_weights          # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for name, spec in (("din", baseline_spec("cfg3_din")), ("neuralcf", default_spec("neuralcf"))):
        W = init_weights(spec, 2)
        with CTRModel(spec, W, device=local) as m:
            slice_rows = 1000
            n = world * slice_rows - 37                           # the last rank scores fewer rows
            fg = sharding.FusedScoreGather(m, slice_rows, dev)
            for step in range(5):                                  # both buffers, several parities
                feats = synthetic_features(spec, n, seed=100 + step)     # same on every rank
                full = m.predict(feats)[:, 0]                      # single-GPU answer for the whole list
                lo, hi = rank * slice_rows, min(n, (rank + 1) * slice_rows)
                d = m.to_device({k: np.asarray(v)[lo:hi] for k, v in feats.items()})
                fg.predict(d.struct(), torch.cuda.current_stream().cuda_stream)
                got = fg.scores().cpu().numpy()
                for r in range(world):
                    a, b = r * slice_rows, min(n, (r + 1) * slice_rows)
                    assert np.array_equal(got[a:b], full[a:b]), (name, step, rank, r)
            m.status()
            fg.close()
            # NCCL paths
            feats = synthetic_features(spec, 801, seed=7)
            full = m.predict(feats)[:, 0]

            def score(shard):
                return torch.from_numpy(m.predict(shard)[:, 0]).to(dev)
            g = sharding.predict_sharded(score, feats).cpu().numpy()
            assert np.array_equal(g, full), (name, "gather_scores")

            def rank_fn(shard, k):
                s = score(shard)
                idx, top = topk_device(s, k)
                return idx, top
            pos, top = sharding.rank_sharded(rank_fn, feats, 25)
            ridx, rtop = topk_device(torch.from_numpy(full).to(dev), 25)
            assert torch.equal(pos.cpu(), ridx.cpu()) and torch.equal(top.cpu(), rtop.cpu()), (name, "rank_sharded")
    dist.barrier()
    if rank == 0:
        print("mgpu worker ok: world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
