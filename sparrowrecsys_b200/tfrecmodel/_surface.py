"""Shared implementation of the per-model `tfrecmodel.<name>` modules.

Each reference model script is a flat module whose public surface is the
module-level Keras `model` and the call `model.predict(feature_dict)`
(`TFRecModel/src/com/sparrowrecsys/offline/tensorflow/<Name>.py`).  The modules in
this package keep that surface: `load(...)` builds the module-level `model` (a
`sparrowrecsys_b200.model.CTRModel` living on one GPU), `predict(features)` is
`model.predict(features)`.
"""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np

from ..model import CTRModel
from ..spec import ModelSpec, default_spec
from ..weights import init_weights


class Surface:
    def __init__(self, name: str):
        self.name = name
        self.model: Optional[CTRModel] = None

    def spec(self, **overrides) -> ModelSpec:
        """The reference script's own constants unless overridden."""
        return default_spec(self.name, **overrides)

    def load(self, weights: Optional[Mapping[str, np.ndarray]] = None,
             spec: Optional[ModelSpec] = None, seed: Optional[int] = None,
             savedmodel: Optional[str] = None, device: int = 0) -> CTRModel:
        if self.model is not None:
            self.model.close()
            self.model = None
        if savedmodel is not None:
            self.model = CTRModel.from_savedmodel(savedmodel, self.name, device)
            return self.model
        spec = spec or self.spec()
        if spec.model != self.name:
            raise ValueError("spec is for %r, this module is %r" % (spec.model, self.name))
        if weights is None:
            # untrained model: the reference's initialisers (what `model` holds before fit)
            weights = init_weights(spec, 0 if seed is None else seed, for_test=False)
        self.model = CTRModel(spec, weights, device)
        return self.model

    def predict(self, features, batch_size: Optional[int] = None) -> np.ndarray:
        if self.model is None:
            raise RuntimeError("tfrecmodel.%s: call load() before predict()" % self.name)
        return self.model.predict(features, batch_size)
