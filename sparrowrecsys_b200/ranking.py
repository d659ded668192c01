"""Device-side ranking tail of the online rankers.

`RecForYouProcess.ranker` / `SimilarMovieProcess.ranker` score every candidate, sort the
`HashMap<Movie, Double>` by value in reverse order and keep the first `size`
(`online/recprocess/RecForYouProcess.java:56-59,69-95`,
`SimilarMovieProcess.java:26-31,121-137`).  Here the scores stay in HBM: the "emb" model is
`srs_cosine_scores_device` (`online/model/Embedding.java:33-47`), the sort + cut is
`srs_topk_device`, and for the CTR models `CTRModel.rank` does forward + ranking in one
library call (`srs_rank_host`).  torch tensors are the device containers, nothing more.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def topk_device(scores, k: int, stream=None):
    """`scores`: float32 CUDA tensor [n].  Returns (positions int32 [min(k,n)], scores
    float32 [min(k,n)]) as CUDA tensors, best first; ties by position, NaN first (the order
    of Java's `Double.compareTo` reversed)."""
    import torch
    if scores.dtype != torch.float32 or not scores.is_cuda or scores.dim() != 1 \
            or not scores.is_contiguous():
        raise ValueError("scores must be a contiguous 1-D float32 CUDA tensor")
    n = scores.shape[0]
    k = max(0, min(int(k), n))
    dev = scores.device.index or 0
    idx = torch.empty(k, dtype=torch.int32, device=scores.device)
    top = torch.empty(k, dtype=torch.float32, device=scores.device)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    _lib.check(_lib.load().srs_topk_device(scores.data_ptr(), n, k, idx.data_ptr(),
                                           top.data_ptr(), dev, stream.cuda_stream))
    return idx, top


def rank_by_embedding(query, cands, size: int, device: int = 0):
    """The "emb" ranker: cosine similarity of `query` [dim] against `cands` [n, dim]
    (numpy float32 or CUDA tensors), best `size` positions and similarities as numpy."""
    import torch
    dev = torch.device("cuda:%d" % device)
    to = lambda a: a.to(dev, torch.float32).contiguous() if isinstance(a, torch.Tensor) \
        else torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
    q, c = to(query), to(cands)
    if c.dim() != 2 or q.dim() != 1 or c.shape[1] != q.shape[0]:
        raise ValueError("query [dim] and cands [n, dim] expected")
    n, dim = c.shape
    with torch.cuda.device(dev):
        scores = torch.empty(n, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev)
        _lib.check(_lib.load().srs_cosine_scores_device(q.data_ptr(), c.data_ptr(), n, dim,
                                                        scores.data_ptr(), device,
                                                        stream.cuda_stream))
        idx, top = topk_device(scores, size, stream)
    return idx.cpu().numpy(), top.cpu().numpy()


def load_embeddings_csv(path: str):
    """The reference's embedding files (`webroot/modeldata/item2vecEmb.csv`, `userEmb.csv`):
    one `id:f f f ...` line per entity, read as `DataManager.loadMovieEmb` / `loadUserEmb` do
    (`online/datamanager/DataManager.java:88-108,143-163`: split on ":", two parts or the line is
    skipped; values split on whitespace, `Float.parseFloat` - `online/util/Utility.java:6-13`).
    Returns (ids int32 [n], embeddings float32 [n, dim]) in file order."""
    ids, rows = [], []
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split(":")
            if len(parts) != 2:
                continue
            ids.append(int(parts[0]))
            rows.append([float(x) for x in parts[1].split()])
    return np.asarray(ids, np.int32), np.asarray(rows, np.float32)
