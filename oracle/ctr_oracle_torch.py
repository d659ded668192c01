"""Multi-threaded CPU stand-in for "TF2 CPU model.predict" in bench.py's reference legs.

THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT (same rule as ctr_oracle.py: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it).

The numpy oracle (ctr_oracle.py) is the parity checker; as a *timing* baseline it is unfair to
the reference: apart from the GEMMs it runs on one core, whereas TensorFlow executes every op of
the Keras graph on an intra-op thread pool.  This module restates the DIN graph
(TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:125-167, statement for statement as
in ctr_oracle.din_forward) on torch CPU tensors, whose ops are threaded the same way, and gives
the other models a row-chunked thread pool over the numpy oracle.  `cpu_predictor` returns the
fastest of the two that applies; tests/test_oracle_props.py holds it to the numpy oracle."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import ctr_oracle as O


def _din_torch(spec, W, threads):
    import torch
    torch.set_num_threads(max(1, int(threads)))
    Wt = {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)) for k, v in W.items()}
    keys = O.din_history_keys(spec.hist_len)

    def prelu(x, a):
        return torch.relu(x) - a * torch.relu(-x)

    def emb(tab, idx):                                   # embedding_column: -1 -> zero vector
        out = tab[idx.clamp(min=0)]
        out[idx < 0] = 0
        return out

    def ids(col):                                        # numeric_column float32 -> Embedding int32 cast
        return torch.from_numpy(np.asarray(col).astype(np.float32).astype(np.int64))

    def num(feats, k):
        return torch.from_numpy(np.asarray(feats[k]).astype(np.float32))[:, None]

    @torch.no_grad()
    def forward(feats):
        cand = ids(feats["movieId"])
        hist = torch.stack([ids(feats[k]) for k in keys], dim=1)
        if int(cand.min()) < 0 or int(max(cand.max(), hist.max())) >= spec.n_movies or int(hist.min()) < 0:
            raise ValueError("movie id out of range")
        tab = Wt["embedding"]
        H, C = tab[hist], tab[cand]                                                      # :134-137
        Cr = C[:, None, :].expand(-1, H.shape[1], -1)                                    # :139
        A = torch.cat([H - Cr, H, Cr, H * Cr], dim=-1)                                   # :141-147
        a = prelu(A @ Wt["au_dense/kernel"] + Wt["au_dense/bias"], Wt["au_prelu/alpha"])  # :149-150
        w = torch.sigmoid(a @ Wt["au_out/kernel"] + Wt["au_out/bias"])[..., 0]           # :151-152
        pooled = (H * w[:, :, None]).sum(dim=1)                                          # :153-158
        uid = torch.from_numpy(O.identity_ids(feats, "userId", spec.n_users))
        ug = torch.from_numpy(O.genre_index(feats, "userGenre1"))
        mg = torch.from_numpy(O.genre_index(feats, "movieGenre1"))
        profile = torch.cat([num(feats, "userAvgRating"), emb(Wt["userGenre1_embedding"], ug),
                             Wt["userId_embedding"][uid], num(feats, "userRatingCount"),
                             num(feats, "userRatingStddev")], dim=1)                     # :108-114 sorted
        context = torch.cat([num(feats, "movieAvgRating"), emb(Wt["movieGenre1_embedding"], mg),
                             num(feats, "movieRatingCount"), num(feats, "movieRatingStddev"),
                             num(feats, "releaseYear")], dim=1)                          # :117-123 sorted
        x = torch.cat([profile, pooled, C, context], dim=1)                              # :161-162
        x = prelu(x @ Wt["dense/kernel"] + Wt["dense/bias"], Wt["prelu/alpha"])
        x = prelu(x @ Wt["dense_1/kernel"] + Wt["dense_1/bias"], Wt["prelu_1/alpha"])
        z = x @ Wt["dense_2/kernel"] + Wt["dense_2/bias"]
        return torch.sigmoid(z).numpy(), z.numpy()

    return forward


def _chunked_numpy(spec, W, threads, chunk_rows=256):
    """Rows are independent in every graph: score row chunks on a thread pool (numpy releases
    the GIL inside its kernels), one BLAS thread per chunk."""
    pool = ThreadPoolExecutor(max(1, int(threads)))

    def forward(feats):
        from threadpoolctl import threadpool_limits
        n = len(np.asarray(feats["movieId"]))
        cols = {k: np.asarray(v) for k, v in feats.items()}
        bounds = [(lo, min(lo + chunk_rows, n)) for lo in range(0, n, chunk_rows)]
        with threadpool_limits(limits=1):
            parts = list(pool.map(lambda b: O.forward(spec, W, {k: v[b[0]:b[1]] for k, v in cols.items()}),
                                  bounds))
        return (np.concatenate([p for p, _ in parts], axis=0), np.concatenate([z for _, z in parts], axis=0))

    return forward


def cpu_predictor(spec, W, threads=None):
    """(forward(feats) -> (prob [B,1], logit [B,1]), description) using every host thread."""
    threads = threads or os.cpu_count() or 1
    if spec.model == "din":
        try:
            return _din_torch(spec, W, threads), "torch CPU restatement of DIN.py, %d intra-op threads" % threads
        except ImportError:
            pass
    return (_chunked_numpy(spec, W, threads),
            "numpy oracle over 256-row chunks on %d threads (1 BLAS thread each)" % threads)
