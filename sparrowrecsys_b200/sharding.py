"""Multi-GPU: shard ranking instances by row, replicate the model, gather scores.

Rows of every reference graph are independent (no batch statistics, no cross-row
op), so the path shards by user-batch with no data-path
collective: rank r scores the contiguous slice [r*B/N, (r+1)*B/N).  The only
exchange is an all-gather of float32 scores, and only when one ranking call needs
the whole vector on every rank (`gather_scores`) - or, when the call ends in the rankers'
sort-and-cut (`RecForYouProcess.java:56-59,92-94`), an all-gather of each rank's best `size`
(position, score) pairs followed by a merge (`rank_sharded`): size x N values cross NVLink
instead of B.  One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)
is the plumbing.
"""
from __future__ import annotations

from typing import Callable, Dict, Mapping, Tuple

import numpy as np


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split; the first `n_rows % world_size` ranks get one
    extra row."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_features(features: Mapping[str, object], world_size: int, rank: int) -> Dict[str, np.ndarray]:
    n = len(np.asarray(features["movieId"]))
    lo, hi = shard_bounds(n, world_size, rank)
    return {k: np.asarray(v)[lo:hi] for k, v in features.items()}


def gather_scores(local_scores, n_rows: int, group=None):
    """All-gather per-rank score slices (torch tensor [n_local] on the rank's device)
    into the full [n_rows] vector on every rank.  Slices may differ by one row, so each
    rank pads to the largest slice; the kernel's output can be written straight into
    the rank's slot of the gather buffer."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_rows, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(world * width, dtype=local_scores.dtype, device=local_scores.device)
    mine = buf[rank * width: rank * width + local_scores.numel()]
    mine.copy_(local_scores.reshape(-1))
    dist.all_gather_into_tensor(buf, buf[rank * width:(rank + 1) * width].clone(), group=group)
    out = torch.cat([buf[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
    return out


def predict_sharded(score_fn: Callable[[Dict[str, np.ndarray]], object], features, group=None,
                    gather: bool = True):
    """Score this rank's row shard with `score_fn(shard_features) -> torch tensor [n_local]`
    and (optionally) all-gather the full score vector."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(np.asarray(features["movieId"]))
    local = score_fn(shard_features(features, world, rank))
    if not gather:
        return local
    return gather_scores(local, n, group)


def rank_sharded(rank_fn: Callable[[Dict[str, np.ndarray], int], tuple], features, size: int,
                 group=None, merge_fn: Callable = None):
    """One ranking call spanning the group's GPUs.  `rank_fn(shard_features, k)` ranks this
    rank's row shard and returns (local positions int32 [k'], scores float32 [k']) as torch
    tensors on the rank's device, best first (e.g. `CTRModel` forward + `ranking.topk_device`);
    every rank receives the global result (positions into the unsharded candidate list, scores).

    Each rank contributes its best min(size, shard rows) candidates - a global top-`size` entry
    is necessarily in its own shard's top-`size` - in ONE all-gather of `[scores | positions]`;
    `merge_fn(scores, k)` (default: `ranking.topk_device`) then ranks the gathered candidates.
    Shards are contiguous and gathered in rank order, so "ties by position in the gathered list"
    is "ties by global position": the result equals the single-GPU ranking exactly."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(np.asarray(features["movieId"]))
    size = max(0, min(int(size), n))
    bounds = [shard_bounds(n, world, r) for r in range(world)]
    counts = [min(size, hi - lo) for lo, hi in bounds]
    width = max(max(counts), 1)
    lo, hi = bounds[rank]
    pos, sc = rank_fn(shard_features(features, world, rank), counts[rank])
    if pos.numel() != counts[rank] or sc.numel() != counts[rank]:
        raise ValueError("rank_fn returned %d candidates, expected %d" % (pos.numel(), counts[rank]))
    mine = torch.zeros(2 * width, dtype=torch.float32, device=sc.device)
    mine[:counts[rank]] = sc.reshape(-1).to(torch.float32)
    mine[width:width + counts[rank]] = (pos.reshape(-1).to(torch.int32) + lo).view(torch.float32)
    buf = torch.empty(world * 2 * width, dtype=torch.float32, device=sc.device)
    dist.all_gather_into_tensor(buf, mine, group=group)
    buf = buf.view(world, 2, width)
    scores = torch.cat([buf[r, 0, :counts[r]] for r in range(world)]).contiguous()
    positions = torch.cat([buf[r, 1, :counts[r]] for r in range(world)]).contiguous().view(torch.int32)
    if merge_fn is None:
        from .ranking import topk_device
        merge_fn = topk_device
    idx, top = merge_fn(scores, size)
    return positions[idx.long()], top


class FusedScoreGather:
    """The score exchange of a ranking call that spans GPUs, fused into the forward kernel
    (`include/srs_ctr.h`: srs_gather_*): every rank's kernel stores its scores into its slice of
    every rank's gather buffer over NVLink (CUDA IPC peer mappings), one flag word per rank follows,
    and `wait()` makes the consuming stream wait for all N slices.  Replaces
    `srs_predict_device` + `all_gather_into_tensor` (the call site it serves:
    RecForYouProcess.java:56-59,92-94 with the candidate list sharded by rows).

    One process per GPU; `torch.distributed` only carries the 64-byte IPC handles at set-up."""

    def __init__(self, model, slice_rows: int, device, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.load()
        self._check = _lib.check
        self.model = model
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.slice_rows = int(slice_rows)
        self.device = torch.device(device)
        h = C.c_void_p()
        _lib.check(self._lib.srs_gather_create(self.device.index or 0, self.world, self.rank, self.slice_rows,
                                               C.byref(h)))
        self._h = h
        mine = (C.c_uint8 * 64)()
        _lib.check(self._lib.srs_gather_export(self._h, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine), group=group)
        blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
        _lib.check(self._lib.srs_gather_connect(self._h, blob))
        dist.barrier(group=group)

    def describe(self) -> str:
        return ("scores stored by the forward kernel's epilogue into every rank's gather buffer over NVLink "
                "(CUDA IPC peer mappings, %d B per rank per launch to each of %d peers), one flag word per rank, "
                "a one-warp wait kernel per launch; no collective" % (4 * self.slice_rows, self.world - 1))

    def predict(self, batch_struct, stream_ptr=None, wait: bool = True):
        """Forward pass of this rank's slice with the fused exchange; asynchronous on the stream."""
        import ctypes as C
        rc = self._lib.srs_predict_device_gather(self.model._h, C.byref(batch_struct), self._h, stream_ptr)
        if rc != 0:
            self._check(rc)
        if wait:
            self.wait(stream_ptr)

    def wait(self, stream_ptr=None):
        rc = self._lib.srs_gather_wait(self._h, stream_ptr)
        if rc != 0:
            self._check(rc)

    def scores(self):
        """The gathered [world * slice_rows] float32 scores of the latest call as a torch tensor
        (copied out of the gather buffer on the current stream; call after `wait`)."""
        import ctypes as C
        import torch
        ptr, rows = C.c_void_p(), C.c_int64()
        self._check(self._lib.srs_gather_scores(self._h, C.byref(ptr), C.byref(rows)))
        out = torch.empty(rows.value, dtype=torch.float32, device=self.device)
        self._check(self._lib.srs_gather_copy_scores(self._h, out.data_ptr(), 0,
                                                     torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.srs_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
