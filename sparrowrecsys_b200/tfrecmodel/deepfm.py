"""tfrecmodel.deepfm - B200 drop-in for the reference's `DeepFM.py` model
(TFRecModel/src/com/sparrowrecsys/offline/tensorflow/DeepFM.py:91-131).

    from tfrecmodel import deepfm
    deepfm.load(weights)            # or load(savedmodel=...), load(spec=..., seed=...)
    p = deepfm.predict(features)    # dict of 1-D columns -> float32 [N,1]
"""
from ._surface import Surface

_surface = Surface("deepfm")
model = None          # the module-level model, as in the reference script
spec = _surface.spec


def load(weights=None, spec=None, seed=None, savedmodel=None, device=0):
    global model
    model = _surface.load(weights, spec, seed, savedmodel, device)
    return model


def predict(features, batch_size=None):
    return _surface.predict(features, batch_size)
