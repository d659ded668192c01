"""Host-side mirror of the reference's model-call surface over the C ABI.

The reference's callable is a module-level Keras `model` and the call is
`model.predict(feature_dict) -> float32[N,1]` (e.g.
`TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:169,185`); at serve
time the same graph answers TF-Serving's `:predict`
(`online/recprocess/RecForYouProcess.java:113-138`).  `CTRModel` keeps that
surface - same key names, same dtypes, unknown keys ignored, `KeyError` for a
missing key, `ValueError` for an out-of-range id (TF's identity-column assert) -
and forwards to `libsrs_ctr.so` through ctypes.  There is no fallback: without the
CUDA library or a GPU, construction raises.

PyTorch appears only in the `*_device` helpers, as the on-device container.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import numpy as np

from . import _lib
from .features import EncodedBatch, encode_batch
from .spec import ModelSpec, default_spec
from .weights import check_weights, init_weights, weight_shapes


def _spec_struct(spec: ModelSpec) -> _lib.SrsSpec:
    s = _lib.SrsSpec()
    s.kind = spec.kind
    s.emb_dim = spec.emb_dim
    s.n_movies = spec.n_movies
    s.n_users = spec.n_users
    s.n_genres = spec.n_genres
    s.hist_len = spec.hist_len
    s.n_hidden = len(spec.hidden)
    for i, h in enumerate(spec.hidden):
        s.hidden[i] = h
    s.au_hidden = spec.au_hidden
    s.cross_buckets = spec.cross_buckets
    s.proj_dim = spec.proj_dim
    s.final_dense = 1 if spec.final_dense else 0
    return s


class DeviceBatch:
    """An encoded batch resident in HBM (torch tensors as containers)."""

    def __init__(self, enc: EncodedBatch, device, hist_cols: int):
        import torch
        dev = torch.device(device)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.B = enc.B
        self.movie_id = t(enc.movie_id)
        self.user_id = t(enc.user_id)
        self.hist = t(None if enc.hist is None else enc.hist.astype(np.int32, copy=False))
        self.movie_genre = t(enc.movie_genre)
        self.user_genre = t(enc.user_genre)
        self.numerics = t(enc.numerics)
        self.hist_stride = 0 if enc.hist is None else enc.hist.shape[1]

    def struct(self) -> _lib.SrsBatch:
        p = lambda x: None if x is None else x.data_ptr()
        return _lib.SrsBatch(self.B, self.hist_stride, p(self.movie_id), p(self.user_id),
                             p(self.hist), p(self.movie_genre), p(self.user_genre),
                             p(self.numerics))


def _host_struct(enc: EncodedBatch, keep: list) -> _lib.SrsBatch:
    def p(a, dtype):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a.ctypes.data
    hs = 0 if enc.hist is None else enc.hist.shape[1]
    narrow = enc.hist is not None and enc.hist.dtype == np.uint16      # srs_batch::hist16
    return _lib.SrsBatch(enc.B, hs, p(enc.movie_id, np.int32), p(enc.user_id, np.int32),
                         None if narrow else p(enc.hist, np.int32), p(enc.movie_genre, np.int32),
                         p(enc.user_genre, np.int32), p(enc.numerics, np.float32),
                         p(enc.hist, np.uint16) if narrow else None)


class CTRModel:
    """One CTR ranking model resident on one GPU."""

    def __init__(self, spec: ModelSpec, weights: Mapping[str, object], device: int = 0,
                 narrow_ids: bool = False, options: Optional[Mapping[str, object]] = None):
        """`weights`: canonical name -> float32 numpy array (reference shapes), or a
        torch CUDA tensor for an embedding table that is already in HBM (used in
        place, see SRS_DEVICE_BORROWED in include/srs_ctr.h).  `narrow_ids`: host batches
        carry the history ids as uint16 (`srs_batch::hist16`, n_movies <= 65536).  `options`:
        kernel-variant choices for `srs_model_create_ex`, e.g. {"din_impl": "rt"}."""
        self.spec = spec
        self.narrow_ids = bool(narrow_ids) and spec.n_movies <= 65536
        self.device = int(device)
        self._h = None
        lib = _lib.load()
        self._lib = lib
        borrowed = {k for k, v in weights.items() if not isinstance(v, np.ndarray)}
        check_weights(spec, {k: v for k, v in weights.items() if k not in borrowed},
                      skip=tuple(borrowed))
        shapes = dict(weight_shapes(spec))
        tensors = (_lib.SrsTensor * len(shapes))()
        self._keep = []
        for i, (name, shape) in enumerate(shapes.items()):
            if name not in weights:
                raise KeyError("missing weight tensor %r" % name)
            w = weights[name]
            rows = shape[0]
            cols = shape[1] if len(shape) > 1 else 1
            if name in borrowed:
                if tuple(w.shape) != tuple(shape) or str(w.dtype) != "torch.float32" or not w.is_contiguous():
                    raise ValueError("device tensor %r must be contiguous float32 %s" % (name, shape))
                self._keep.append(w)
                tensors[i] = _lib.SrsTensor(name.encode(), w.data_ptr(), rows, cols,
                                            _lib.SRS_DEVICE_BORROWED)
            else:
                a = np.ascontiguousarray(w, dtype=np.float32)
                self._keep.append(a)
                tensors[i] = _lib.SrsTensor(name.encode(), a.ctypes.data, rows, cols, _lib.SRS_HOST)
        handle = C.c_void_p()
        sp = _spec_struct(spec)
        opts = ";".join("%s=%s" % (k, v) for k, v in (options or {}).items()).encode()
        _lib.check(lib.srs_model_create_ex(C.byref(sp), tensors, len(shapes), self.device, opts or None,
                                           C.byref(handle)))
        self._h = handle
        self._keep = [w for w in self._keep if not isinstance(w, np.ndarray)]  # host copies done
        self.hist_cols = spec.hist_len if spec.model in ("din", "dien") \
            else (1 if spec.model == "widendeep" else 0)

    # ---- constructors ------------------------------------------------------------
    @classmethod
    def from_spec(cls, spec: ModelSpec, seed: int = 0, device: int = 0, for_test: bool = True):
        return cls(spec, init_weights(spec, seed, for_test=for_test), device)

    @classmethod
    def from_savedmodel(cls, savedmodel_dir: str, model: str = "neuralcf", device: int = 0):
        """Load one of the reference's shipped exports (`webroot/modeldata/neuralcf/<v>`,
        `webroot/modeldata/MLPRec/005`) without TensorFlow."""
        from . import bundle
        if model == "neuralcf":
            return cls(default_spec("neuralcf"), bundle.load_neuralcf(savedmodel_dir), device)
        if model == "twotowers":
            return cls(default_spec("twotowers", hidden=(10,), final_dense=False),
                       bundle.load_twotowers(savedmodel_dir), device)
        raise ValueError("no shipped SavedModel layout known for %r" % model)

    # ---- lifetime ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.srs_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def kernel_name(self) -> str:
        return self._lib.srs_model_kernel_name(self._h).decode()

    def set_sm_limit(self, n_sms: int):
        """At most `n_sms` CTAs per launch of the persistent tensor-core kernels (<= 0: all
        SMs), so that launches on other streams run beside it (`srs_model_set_sm_limit`)."""
        _lib.check(self._lib.srs_model_set_sm_limit(self._h, int(n_sms)))

    @property
    def bytes_per_inference(self) -> int:
        return int(self._lib.srs_model_bytes_per_inference(self._h))

    # ---- the reference's call surface -----------------------------------------------
    def predict(self, features: Mapping[str, object], batch_size: Optional[int] = None) -> np.ndarray:
        """`model.predict(x)`: feature dict of 1-D columns -> float32 [N,1]."""
        return self._predict(features, batch_size, want_logits=False)[0]

    def predict_with_logits(self, features, batch_size: Optional[int] = None):
        return self._predict(features, batch_size, want_logits=True)

    def _predict(self, features, batch_size, want_logits):
        enc = encode_batch(self.spec, features, narrow_ids=self.narrow_ids)
        n = enc.B
        probs = np.empty(n, np.float32)
        logits = np.empty(n, np.float32) if want_logits else None
        step = n if not batch_size else max(int(batch_size), 1)
        if n <= step:
            self.predict_encoded(enc, probs, logits)
        else:   # the Keras predict loop over dataset batches -> one pipelined library call
            bounds = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
            self.predict_batches([enc.slice(lo, hi) for lo, hi in bounds],
                                 [probs[lo:hi] for lo, hi in bounds],
                                 None if logits is None else [logits[lo:hi] for lo, hi in bounds])
        return probs.reshape(n, 1), (None if logits is None else logits.reshape(n, 1))

    def predict_batches(self, encs, probs_list, logits_list=None):
        """`srs_predict_host_batches`: score a list of encoded batches, H2D / kernel / D2H of
        successive batches overlapped over the library's slots."""
        n = len(encs)
        keep = []
        structs = (_lib.SrsBatch * n)(*[_host_struct(e, keep) for e in encs])
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in probs_list])
        lp = None
        if logits_list is not None:
            lp = (C.c_void_p * n)(*[l.ctypes.data for l in logits_list])
        _lib.check(self._lib.srs_predict_host_batches(self._h, n, structs, pp, lp))

    def predict_encoded(self, enc: EncodedBatch, probs: np.ndarray,
                        logits: Optional[np.ndarray] = None) -> np.ndarray:
        """Host arrays in, host scores out (`srs_predict_host`)."""
        if enc.B == 0:
            return probs
        keep = []
        b = _host_struct(enc, keep)
        assert probs.dtype == np.float32 and probs.flags.c_contiguous and probs.shape[0] == enc.B
        lp = None
        if logits is not None:
            assert logits.dtype == np.float32 and logits.flags.c_contiguous
            lp = logits.ctypes.data
        _lib.check(self._lib.srs_predict_host(self._h, C.byref(b), probs.ctypes.data, lp))
        return probs

    def rank(self, features: Mapping[str, object], size: int):
        """`RecForYouProcess.getRecList`'s tail for one request: score the candidate rows
        and return the best `size` (positions int32 [k], scores float32 [k], best first;
        equal scores by position) - `srs_rank_host`, only k results leave the device."""
        enc = encode_batch(self.spec, features, narrow_ids=self.narrow_ids)
        k = max(0, min(int(size), enc.B))
        idx = np.empty(k, np.int32)
        top = np.empty(k, np.float32)
        if k == 0:
            return idx, top
        keep = []
        b = _host_struct(enc, keep)
        _lib.check(self._lib.srs_rank_host(self._h, C.byref(b), k, idx.ctypes.data,
                                           top.ctypes.data))
        return idx, top

    # ---- one user x n candidates, movie features resident in HBM ----------------------------
    def set_movie_table(self, table):
        """Upload the movie side of the serving feature store (`featurestore.MovieFeatureTable`,
        i.e. the `mf:<movieId>` hashes of FeatureEngForRecModel.scala:130-174) to HBM once;
        `rank_user` requests then carry only the user's row and the candidate ids."""
        n = table.n_movies
        genres = np.ascontiguousarray(np.stack([table.idx_cols[k] for k in
                                                ("movieGenre1", "movieGenre2", "movieGenre3")], axis=1), np.int32)
        nums = np.ascontiguousarray(np.stack([
            table.float_cols["movieAvgRating"], table.int_cols["movieRatingCount"].astype(np.float32),
            table.float_cols["movieRatingStddev"], table.int_cols["releaseYear"].astype(np.float32)], axis=1),
            np.float32)
        _lib.check(self._lib.srs_model_set_movie_features(self._h, n, genres.ctypes.data, nums.ctypes.data))

    def rank_user(self, user_id: int, user_fields: Mapping[str, object], candidate_ids, size: int,
                  return_scores: bool = False):
        """`RecForYouProcess.getRecList` for one request (`:40-59`): `user_fields` is the user's
        `uf:` hash (strings, as `FeatureStore.user_features` returns it, or already typed values
        under the model input names), `candidate_ids` the n candidate movies.  Ships the user's
        row and the ids only (`srs_rank_user_host`); the movie features are gathered on the
        device from the table uploaded by `set_movie_table`.  Returns (positions int32 [k],
        scores float32 [k]) best first (+ all n scores with `return_scores`)."""
        from .featurestore import parse_user_features
        from .features import genre_to_index
        typed = parse_user_features(user_fields, max(self.hist_cols, 1))
        cand = np.ascontiguousarray(np.asarray(candidate_ids, np.int32).reshape(-1))
        n = cand.shape[0]
        k = max(0, min(int(size), n))
        from .spec import history_keys
        hkeys = history_keys(self.spec.hist_len) if self.spec.model in ("din", "dien") else ["userRatedMovie1"]
        hist = np.ascontiguousarray(np.array([typed[k] for k in hkeys[:self.hist_cols]], np.int32))   # graph position order
        row = _lib.SrsUserRow()
        row.user_id = int(user_id)
        for g in range(5):
            v = typed["userGenre%d" % (g + 1)]
            row.user_genre[g] = int(genre_to_index([v])[0]) if isinstance(v, (str, bytes)) else int(v)
        row.user_numerics[0] = float(typed["userAvgRating"])
        row.user_numerics[1] = float(np.float32(typed["userRatingCount"]))
        row.user_numerics[2] = float(typed["userRatingStddev"])
        row.n_hist = self.hist_cols
        row.hist = hist.ctypes.data if self.hist_cols else None
        idx = np.empty(k, np.int32)
        top = np.empty(k, np.float32)
        probs = np.empty(n, np.float32) if return_scores else None
        _lib.check(self._lib.srs_rank_user_host(self._h, C.byref(row), cand.ctypes.data, n, k,
                                                idx.ctypes.data if k else None, top.ctypes.data if k else None,
                                                probs.ctypes.data if return_scores else None))
        return (idx, top, probs) if return_scores else (idx, top)

    # ---- pipelined host path -----------------------------------------------------------
    def num_slots(self) -> int:
        return int(self._lib.srs_num_slots())

    def submit_host(self, slot: int, batch_struct: _lib.SrsBatch, probs_ptr: int,
                    logits_ptr: Optional[int] = None):
        _lib.check(self._lib.srs_predict_host_async(self._h, slot, C.byref(batch_struct),
                                                    probs_ptr, logits_ptr))

    def wait(self, slot: int):
        _lib.check(self._lib.srs_wait_slot(self._h, slot))

    # ---- device-resident path -----------------------------------------------------------
    def to_device(self, features_or_enc) -> DeviceBatch:
        enc = features_or_enc if isinstance(features_or_enc, EncodedBatch) \
            else encode_batch(self.spec, features_or_enc)
        return DeviceBatch(enc, "cuda:%d" % self.device, self.hist_cols)

    def predict_device(self, batch: DeviceBatch, probs, logits=None, stream=None):
        """Everything in HBM: `probs` / `logits` are float32 CUDA tensors [B]; asynchronous
        on `stream` (a torch stream; default: torch's current stream)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        b = batch.struct()
        _lib.check(self._lib.srs_predict_device(
            self._h, C.byref(b), probs.data_ptr(), None if logits is None else logits.data_ptr(),
            stream.cuda_stream))
        return probs

    def status(self):
        """Synchronise; raises ValueError if a device batch carried an out-of-range id."""
        _lib.check(self._lib.srs_model_status(self._h))


def launch_count() -> int:
    return int(_lib.load().srs_launch_count())
