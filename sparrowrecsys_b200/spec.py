"""Model specifications for the SparrowRecSys CTR ranking forward path.

The reference hard-codes every hyper-parameter as a module constant
(`TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DIN.py:30-31,66,132`,
`EmbeddingMLP.py:50-58`, `NeuralCF.py:30-35,74`).  `ModelSpec` lifts them into
one frozen dataclass so that the reference shapes (E=10, T=5, 1001/30001
buckets) and the BASELINE.json shapes (E=16/32/64, T=50/200, ML-20M vocab,
10^8-item vocab) run through the same code path.

Names below follow the reference's domain: movies, users, genres, history.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, List, Tuple

# Model ids shared with include/srs_ctr.h (enum srs_model_kind).
MODEL_KINDS: Dict[str, int] = {
    "embeddingmlp": 0,
    "widendeep": 1,
    "neuralcf": 2,
    "twotowers": 3,
    "deepfm": 4,
    "deepfm_v2": 5,
    "din": 6,
    "dien": 7,
}

# Reference genre vocabulary, identical in every model script
# (EmbeddingMLP.py:30-32, DIN.py:70-72 ...).  Index = position; OOV / "" -> -1.
GENRE_VOCAB: Tuple[str, ...] = (
    "Film-Noir", "Action", "Adventure", "Horror", "Romance", "War", "Comedy",
    "Western", "Documentary", "Sci-Fi", "Drama", "Thriller", "Crime", "Fantasy",
    "Animation", "IMAX", "Mystery", "Children", "Musical",
)

# The 7 numeric columns every dense model consumes, in DenseFeatures order
# (ASCII-sorted by column name).  This is the order
# of the `numerics` [B,7] array at the C-ABI.
NUMERIC_KEYS: Tuple[str, ...] = (
    "movieAvgRating", "movieRatingCount", "movieRatingStddev", "releaseYear",
    "userAvgRating", "userRatingCount", "userRatingStddev",
)
MOVIE_GENRE_KEYS: Tuple[str, ...] = ("movieGenre1", "movieGenre2", "movieGenre3")
USER_GENRE_KEYS: Tuple[str, ...] = ("userGenre1", "userGenre2", "userGenre3",
                                    "userGenre4", "userGenre5")


def history_keys(hist_len: int) -> List[str]:
    """Input keys of the DIN behaviour sequence, in *graph position* order.

    The reference builds `DenseFeatures(recent_rate_col)` (DIN.py:97-103,126);
    DenseFeatures concatenates its columns sorted by name, so for T > 9 the
    position order is the ASCII order of `userRatedMovie<k>` (1,10,11,...,2,...).
    Position matters only through the per-position PReLU alpha (DIN.py:150).
    """
    return sorted("userRatedMovie%d" % k for k in range(1, hist_len + 1))


@dataclass(frozen=True)
class ModelSpec:
    model: str
    emb_dim: int = 10            # E: `embedding_column(..., 10)` / EMBEDDING_SIZE
    n_movies: int = 1001         # num_buckets of movieId (ids 0..n-1 valid)
    n_users: int = 30001         # num_buckets of userId
    n_genres: int = len(GENRE_VOCAB)
    hist_len: int = 5            # T: RECENT_MOVIES (DIN); W&D uses slot 1 only
    hidden: Tuple[int, ...] = ()  # model dependent, see default_spec()
    au_hidden: int = 32          # DIN activation-unit / DIEN attention width (DIN.py:149, DIEN.py:178)
    cross_buckets: int = 10000   # W&D crossed_column hash_bucket_size (WideNDeep.py:73)
    proj_dim: int = 64           # DeepFM_v2 per-field projection width (DeepFM_v2.py:114)
    final_dense: bool = True     # two towers: Dense(1,sigmoid) after Dot (NeuralCF.py:67);
                                 # the shipped MLPRec/005 export has none (raw dot)

    def __post_init__(self):
        if self.model not in MODEL_KINDS:
            raise ValueError("unknown model %r (one of %s)" % (self.model, sorted(MODEL_KINDS)))
        if self.emb_dim < 1 or self.emb_dim > 64:
            raise ValueError("emb_dim must be in 1..64")
        if self.model == "dien" and self.emb_dim > 32:
            raise ValueError("dien: emb_dim must be <= 32 (one warp lane per state element)")
        if self.n_genres != len(GENRE_VOCAB):
            raise ValueError("n_genres is fixed by the reference vocabulary (19)")
        if self.model in ("din", "dien", "widendeep") and self.hist_len < 1:
            raise ValueError("hist_len must be >= 1")
        object.__setattr__(self, "hidden", tuple(int(h) for h in self.hidden))

    @property
    def kind(self) -> int:
        return MODEL_KINDS[self.model]

    def to_dict(self) -> dict:
        d = asdict(self)
        d["hidden"] = list(self.hidden)
        return d

    # ---- derived sizes (reference formulation) --------------------------------
    @property
    def fm1_width(self) -> int:
        """Width of the one-hot first-order block [movieGenre1|movieId|userGenre1|userId]."""
        return self.n_genres + self.n_movies + self.n_genres + self.n_users

    def required_keys(self) -> List[str]:
        m = self.model
        if m in ("neuralcf", "twotowers"):
            return ["movieId", "userId"]
        if m == "embeddingmlp":
            return ["movieId", "userId", *NUMERIC_KEYS, *MOVIE_GENRE_KEYS, *USER_GENRE_KEYS]
        if m == "widendeep":
            return ["movieId", "userId", "userRatedMovie1", *NUMERIC_KEYS,
                    *MOVIE_GENRE_KEYS, *USER_GENRE_KEYS]
        if m in ("deepfm", "deepfm_v2"):
            return ["movieId", "userId", *NUMERIC_KEYS, "movieGenre1", "userGenre1"]
        if m in ("din", "dien"):
            # DIEN's Keras `inputs` also list label and negtive_userRatedMovie2..5
            # (DIEN.py:82-87); they feed the auxiliary loss only, not y_pred
            return ["movieId", "userId", *history_keys(self.hist_len), *NUMERIC_KEYS,
                    "movieGenre1", "userGenre1"]
        raise AssertionError(m)

    # ---- algorithmic work per inference (SURVEY.md section 8d definition) -----
    def bytes_per_inference(self) -> int:
        """Every embedding-row lookup once at 4E bytes + 4 B per int32 id + 4 B per
        numeric + 4 B per scalar-weight gather + 4 B output; dense-layer weights
        amortised to 0."""
        E, T = self.emb_dim, self.hist_len
        m = self.model
        if m == "embeddingmlp":
            return 10 * 4 * E + 10 * 4 + 7 * 4 + 4
        if m == "widendeep":
            return 10 * 4 * E + 11 * 4 + 7 * 4 + 4 + 4
        if m in ("neuralcf", "twotowers"):
            return 2 * 4 * E + 2 * 4 + 4
        if m == "deepfm":
            return 6 * 4 * E + 4 * 4 + 4 * 4 + 7 * 4 + 4
        if m == "deepfm_v2":
            return 4 * 4 * E + 4 * 4 + 4 * 4 + 7 * 4 + 4
        if m in ("din", "dien"):
            return (T + 1) * 4 * E + 3 * 4 * E + 28 + 4 * (T + 4) + 4
        raise AssertionError(m)

    def flops_per_inference(self) -> int:
        """2 x MACs of the reference formulation."""
        E, T = self.emb_dim, self.hist_len
        m = self.model
        h = self.hidden
        if m in ("embeddingmlp", "widendeep"):
            k = 7 + 10 * E
            f = 2 * k * h[0] + 2 * h[0] * h[1] + 2 * h[1]
            return f + (2 if m == "widendeep" else 0)
        if m == "neuralcf":
            dims = [2 * E, *h, 1]
            return sum(2 * a * b for a, b in zip(dims[:-1], dims[1:]))
        if m == "twotowers":
            dims = [E, *h]
            f = 2 * sum(2 * a * b for a, b in zip(dims[:-1], dims[1:])) + 2 * dims[-1]
            return f + (2 if self.final_dense else 0)
        if m == "deepfm":
            k = 7 + 2 * E
            return 4 * 2 * E + 2 * k * h[0] + 2 * h[0] * h[1] + 2 * (4 + h[1]) + 8
        if m == "deepfm_v2":
            P = self.proj_dim
            return (4 * 2 * E * P + 2 * 7 * P + 2 * 5 * P * h[0] + 2 * h[0] * h[1]
                    + 5 * P * 3 + 2 * (1 + P + h[1]) + 2 * 7 + 8)
        if m == "din":
            A = self.au_hidden
            return (T * (2 * 4 * E * A + 2 * A + 2 * A + 3 * E) + 2 * T * E
                    + 2 * (5 * E + 7) * h[0] + 2 * h[0] * h[1] + 2 * h[1])
        if m == "dien":
            A = self.au_hidden
            gru = 2 * 2 * E * 3 * E + 10 * E              # two [E,3E] products + gate math
            att = E + 2 * E * A + 2 * A                   # product, Dense32, Dense1
            augru = 9 * 2 * E * E + 12 * E                # 3 gates x (input, hidden, act) Dense
            return (T * (gru + att + augru)
                    + 2 * (5 * E + 7) * h[0] + 2 * h[0] * h[1] + 2 * h[1])
        raise AssertionError(m)


_DEFAULT_HIDDEN = {
    "embeddingmlp": (128, 128),   # EmbeddingMLP.py:74-75
    "widendeep": (128, 128),      # WideNDeep.py:102-103
    "neuralcf": (10, 10),         # NeuralCF.py:74
    "twotowers": (10,),           # shipped MLPRec/005: one Dense(10,relu) per tower
    "deepfm": (64, 64),           # DeepFM.py:107-108
    "deepfm_v2": (32, 16),        # DeepFM_v2.py:125-126
    "din": (128, 64),             # DIN.py:163,165
    "dien": (128, 64),            # DIEN.py:252,254
}


def default_spec(model: str, **overrides) -> ModelSpec:
    """Spec with the reference script's own constants unless overridden."""
    kw = dict(model=model, hidden=_DEFAULT_HIDDEN[model])
    kw.update(overrides)
    return ModelSpec(**kw)


# BASELINE.json configs -> specs.
ML20M_MOVIES = 27279   # 27 278 ML-20M movies remapped to dense ids + id 0
ML20M_USERS = 138494   # 138 493 users + id 0


def baseline_spec(cfg: str) -> ModelSpec:
    if cfg == "cfg1_embeddingmlp":
        return default_spec("embeddingmlp")
    if cfg == "cfg2_deepfm":
        return default_spec("deepfm", emb_dim=16, n_movies=ML20M_MOVIES, n_users=ML20M_USERS)
    if cfg == "cfg2_deepfm_v2":
        return default_spec("deepfm_v2", emb_dim=16, n_movies=ML20M_MOVIES, n_users=ML20M_USERS)
    if cfg == "cfg3_din":
        return default_spec("din", emb_dim=32, hist_len=50,
                            n_movies=ML20M_MOVIES, n_users=ML20M_USERS)
    if cfg == "cfg4_widendeep":
        return default_spec("widendeep")
    if cfg == "cfg4_neuralcf":
        return default_spec("neuralcf")
    if cfg == "cfg4_twotowers":
        return default_spec("twotowers", hidden=(10,), final_dense=True)
    if cfg == "cfg5_din":
        return default_spec("din", emb_dim=64, hist_len=200,
                            n_movies=100_000_000, n_users=ML20M_USERS)
    if cfg == "ref_dien":
        return default_spec("dien")
    raise KeyError(cfg)
