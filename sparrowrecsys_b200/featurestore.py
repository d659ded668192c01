"""Serving-side feature assembly.

The reference defines - and leaves commented out in `main` - an online feature store for
the models that need more than `(userId, movieId)`:

* `extractAndSaveUserFeaturesToRedis` (`offline/spark/featureeng/FeatureEngForRecModel.scala:
  208-259`): per user the *latest* sample row (row_number over `timestamp desc` = 1), written
  as the Redis hash `uf:<userId>` with 15 string fields (`userRatedMovie1..5`,
  `userGenre1..5`, `userRatingCount`, `userAvgReleaseYear`, `userReleaseYearStddev`,
  `userAvgRating`, `userRatingStddev`), nulls as `""`;
* `extractAndSaveMovieFeaturesToRedis` (`:130-174`): the same per movie, hash `mf:<movieId>`
  with `movieGenre1..3`, `movieRatingCount`, `releaseYear`, `movieAvgRating`,
  `movieRatingStddev`;
* the readers: `RecForYouProcess.java:46-52` (`hgetAll("uf:" + userId)` per request) and
  `DataManager.java:127-140` (`mf:*` into every `Movie` at start-up).

This module is the host logic between those hashes and `model.predict`: `FeatureStore` reads
the hashes (any backend with `hgetall`, e.g. a `redis.Redis` client; `DictBackend` in process),
`FeatureStore.from_samples` writes them the way the Spark job does, `MovieFeatureTable`
compiles the movie side once into arrays indexed by movie id, and `assemble` builds the
feature dict of one ranking request (one user x n candidates) with a gather instead of n
hash reads.  Empty / missing fields take the `make_csv_dataset(na_value="0")` defaults that
the models were trained with: 0, 0.0, "".
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Optional, Sequence

import numpy as np

from .features import genre_to_index

USER_PREFIX = "uf:"          # FeatureEngForRecModel.scala:222
MOVIE_PREFIX = "mf:"         # FeatureEngForRecModel.scala:144

USER_INT_FIELDS = ("userRatedMovie1", "userRatedMovie2", "userRatedMovie3", "userRatedMovie4",
                   "userRatedMovie5", "userRatingCount", "userAvgReleaseYear")
USER_FLOAT_FIELDS = ("userReleaseYearStddev", "userAvgRating", "userRatingStddev")
USER_STR_FIELDS = ("userGenre1", "userGenre2", "userGenre3", "userGenre4", "userGenre5")
USER_FIELDS = USER_INT_FIELDS[:5] + USER_STR_FIELDS + USER_INT_FIELDS[5:] + USER_FLOAT_FIELDS

MOVIE_STR_FIELDS = ("movieGenre1", "movieGenre2", "movieGenre3")
MOVIE_INT_FIELDS = ("movieRatingCount", "releaseYear")
MOVIE_FLOAT_FIELDS = ("movieAvgRating", "movieRatingStddev")
MOVIE_FIELDS = MOVIE_STR_FIELDS + MOVIE_INT_FIELDS + MOVIE_FLOAT_FIELDS


def _text(v) -> str:
    return v.decode("utf-8", "replace") if isinstance(v, (bytes, bytearray)) else str(v)


def parse_int(v) -> int:
    """A hash field that the CSV loader would type as int32 ("" -> 0; "1995.0" tolerated)."""
    s = _text(v).strip()
    if not s:
        return 0
    try:
        return int(s)
    except ValueError:
        return int(float(s))


def parse_float(v) -> float:
    s = _text(v).strip()
    return float(s) if s else 0.0


class DictBackend:
    """In-process stand-in for the Redis server: key -> {field: value}, all strings."""

    def __init__(self, hashes: Optional[Mapping[str, Mapping[str, str]]] = None):
        self.hashes: Dict[str, Dict[str, str]] = {k: dict(v) for k, v in (hashes or {}).items()}

    def hset(self, key: str, mapping: Mapping[str, str]):
        self.hashes.setdefault(key, {}).update({k: _text(v) for k, v in mapping.items()})

    def hgetall(self, key: str) -> Dict[str, str]:
        return dict(self.hashes.get(key, {}))          # Jedis: empty map when the key is absent

    def keys(self, pattern: str):
        assert pattern.endswith("*")
        return [k for k in self.hashes if k.startswith(pattern[:-1])]


class FeatureStore:
    """The `uf:` / `mf:` hashes behind one `hgetall`-capable backend."""

    def __init__(self, backend=None):
        self.backend = backend if backend is not None else DictBackend()

    # ---- readers (RecForYouProcess.java:46-52, DataManager.java:127-140) ---------------
    def user_features(self, user_id: int) -> Dict[str, str]:
        return {_text(k): _text(v) for k, v in self.backend.hgetall(USER_PREFIX + str(int(user_id))).items()}

    def movie_features(self, movie_id: int) -> Dict[str, str]:
        return {_text(k): _text(v) for k, v in self.backend.hgetall(MOVIE_PREFIX + str(int(movie_id))).items()}

    def movie_ids(self):
        return sorted(int(_text(k).split(":")[1]) for k in self.backend.keys(MOVIE_PREFIX + "*"))

    # ---- writer (the Spark job's two extractAndSave* functions) -------------------------
    @classmethod
    def from_samples(cls, samples, backend=None) -> "FeatureStore":
        """`samples`: path of a sample CSV (trainingSamples.csv layout) or its columns as
        *strings* (`read_sample_strings`).  Per user / per movie the row with the greatest
        timestamp wins (first in file order among equals; Spark's row_number leaves that tie
        unspecified), and its fields are stored verbatim as strings."""
        cols = read_sample_strings(samples) if isinstance(samples, str) else samples
        store = cls(backend)
        ts = np.array([parse_int(t) for t in cols["timestamp"]], dtype=np.int64)
        for id_key, prefix, fields in (("userId", USER_PREFIX, USER_FIELDS),
                                       ("movieId", MOVIE_PREFIX, MOVIE_FIELDS)):
            best: Dict[str, int] = {}
            for row, ident in enumerate(cols[id_key]):
                j = best.get(ident)
                if j is None or ts[row] > ts[j]:
                    best[ident] = row
            for ident, row in best.items():
                store.backend.hset(prefix + ident, {f: cols[f][row] for f in fields})
        return store


def read_sample_strings(path: str) -> Dict[str, list]:
    """A sample CSV as columns of raw strings (what Spark's CSV reader hands the job)."""
    import csv
    with open(path, newline="") as f:
        reader = csv.reader(f)
        header = next(reader)
        cols = {h: [] for h in header}
        for row in reader:
            if len(row) != len(header):
                continue
            for h, v in zip(header, row):
                cols[h].append(v)
    return cols


class MovieFeatureTable:
    """Movie-side features compiled once into arrays indexed by movie id (what
    `DataManager.loadMovieFeatures` keeps per `Movie` object), so that the n candidates of a
    request are one gather.  Ids without a hash keep the defaults ("" / 0 / 0.0)."""

    def __init__(self, n_movies: int):
        self.n_movies = int(n_movies)
        self.present = np.zeros(n_movies, bool)
        self.str_cols = {k: np.full(n_movies, "", dtype=object) for k in MOVIE_STR_FIELDS}
        self.idx_cols = {k: np.full(n_movies, -1, np.int32) for k in MOVIE_STR_FIELDS}   # vocabulary indices
        self.int_cols = {k: np.zeros(n_movies, np.int32) for k in MOVIE_INT_FIELDS}
        self.float_cols = {k: np.zeros(n_movies, np.float32) for k in MOVIE_FLOAT_FIELDS}

    @classmethod
    def from_store(cls, store: FeatureStore, n_movies: int,
                   movie_ids: Optional[Iterable[int]] = None) -> "MovieFeatureTable":
        t = cls(n_movies)
        for mid in (store.movie_ids() if movie_ids is None else movie_ids):
            if not 0 <= mid < n_movies:
                continue
            h = store.movie_features(mid)
            if not h:
                continue
            t.present[mid] = True
            for k in MOVIE_STR_FIELDS:
                t.str_cols[k][mid] = h.get(k, "")
                t.idx_cols[k][mid] = genre_to_index([h.get(k, "")])[0]
            for k in MOVIE_INT_FIELDS:
                t.int_cols[k][mid] = parse_int(h.get(k, ""))
            for k in MOVIE_FLOAT_FIELDS:
                t.float_cols[k][mid] = np.float32(parse_float(h.get(k, "")))
        return t

    def gather(self, movie_ids: np.ndarray, encoded: bool = False) -> Dict[str, np.ndarray]:
        """Columns of the given movies; `encoded`: genres as vocabulary indices (-1 = missing /
        OOV), which `predict` accepts in place of the strings and skips the lookup for."""
        ids = np.asarray(movie_ids, np.int64)
        if ids.size and (ids.min() < 0 or ids.max() >= self.n_movies):
            raise ValueError("movie id out of range [0, %d)" % self.n_movies)
        out: Dict[str, np.ndarray] = {}
        for cols in (self.idx_cols if encoded else self.str_cols, self.int_cols, self.float_cols):
            for k, a in cols.items():
                out[k] = a[ids]
        return out


def parse_user_features(fields: Mapping[str, str], hist_len: int = 5) -> Dict[str, object]:
    """One `uf:` hash -> typed scalars under the model input names; history slots beyond
    the five the store carries are 0 (the padding id)."""
    out: Dict[str, object] = {}
    for k in range(1, max(hist_len, 5) + 1):
        out["userRatedMovie%d" % k] = parse_int(fields.get("userRatedMovie%d" % k, ""))
    for k in USER_INT_FIELDS[5:]:
        out[k] = parse_int(fields.get(k, ""))
    for k in USER_FLOAT_FIELDS:
        out[k] = np.float32(parse_float(fields.get(k, "")))
    for k in USER_STR_FIELDS:
        out[k] = _text(fields.get(k, ""))
    return out


def assemble(user_id: int, user_fields: Mapping[str, str], candidate_ids: Sequence[int],
             movies: MovieFeatureTable, hist_len: int = 5, encoded: bool = False) -> Dict[str, np.ndarray]:
    """Feature dict of one ranking request - the instances `callNeuralCFTFServing` would
    post (`RecForYouProcess.java:118-127`) widened by the stored features: user columns
    broadcast over the n candidates, movie columns gathered by candidate id.  Keys and dtypes
    are those of the Keras `inputs` dicts (e.g. `DIN.py:34-59`), so the result goes straight
    into `predict` / `rank` of any of the models.  `encoded`: genre columns as vocabulary indices
    instead of strings (same scores; saves the per-row string lookup on the request path)."""
    cand = np.asarray(candidate_ids, np.int32).reshape(-1)
    n = cand.shape[0]
    feats: Dict[str, np.ndarray] = {"movieId": cand, "userId": np.full(n, int(user_id), np.int32)}
    for k, v in parse_user_features(user_fields, hist_len).items():
        if isinstance(v, str) and encoded:
            col = np.full(n, genre_to_index([v])[0], np.int32)
        elif isinstance(v, str):
            col = np.empty(n, dtype=object)
            col[:] = v
        elif isinstance(v, np.floating):
            col = np.full(n, v, np.float32)
        else:
            col = np.full(n, v, np.int32)
        feats[k] = col
    feats.update(movies.gather(cand, encoded))
    return feats


def fill_instances(features: Dict[str, np.ndarray], missing: Sequence[str], store: FeatureStore,
                   hist_len: int = 5) -> Dict[str, np.ndarray]:
    """Row-format request carrying only some keys (the Java server posts `userId` and
    `movieId`): fill the `missing` model inputs from the store, one hash read per distinct
    user / movie in the request."""
    n = features["movieId"].shape[0]
    need_user = [k for k in missing if k.startswith("user")]
    need_movie = [k for k in missing if not k.startswith("user")]
    if need_user:
        uniq, inv = np.unique(np.asarray(features["userId"], np.int64), return_inverse=True)
        parsed = [parse_user_features(store.user_features(int(u)), hist_len) for u in uniq]
        for k in need_user:
            vals = [p[k] for p in parsed]
            dtype = object if k in USER_STR_FIELDS else (np.float32 if k in USER_FLOAT_FIELDS else np.int32)
            features[k] = np.array(vals, dtype=dtype)[inv.reshape(-1)]
    if need_movie:
        uniq, inv = np.unique(np.asarray(features["movieId"], np.int64), return_inverse=True)
        hashes = [store.movie_features(int(m)) for m in uniq]
        for k in need_movie:
            if k in MOVIE_STR_FIELDS:
                vals = np.array([h.get(k, "") for h in hashes], dtype=object)
            elif k in MOVIE_INT_FIELDS:
                vals = np.array([parse_int(h.get(k, "")) for h in hashes], np.int32)
            elif k in MOVIE_FLOAT_FIELDS:
                vals = np.array([parse_float(h.get(k, "")) for h in hashes], np.float32)
            else:
                raise KeyError("no stored feature %r" % k)
            features[k] = vals[inv.reshape(-1)]
    assert all(features[k].shape[0] == n for k in missing)
    return features
