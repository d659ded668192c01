"""tfrecmodel.deepfm_v2 - B200 drop-in for the reference's `DeepFM_v2.py` model
(TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DeepFM_v2.py:98-173).

    from tfrecmodel import deepfm_v2
    deepfm_v2.load(weights)            # or load(savedmodel=...), load(spec=..., seed=...)
    p = deepfm_v2.predict(features)    # dict of 1-D columns -> float32 [N,1]
"""
from ._surface import Surface

_surface = Surface("deepfm_v2")
model = None          # the module-level model, as in the reference script
spec = _surface.spec


def load(weights=None, spec=None, seed=None, savedmodel=None, device=0):
    global model
    model = _surface.load(weights, spec, seed, savedmodel, device)
    return model


def predict(features, batch_size=None):
    return _surface.predict(features, batch_size)
