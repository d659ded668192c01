"""`tfrecmodel` - the reference's Python model-call surface, one module per model
script of TFRecModel/src/com/sparrowrecsys/offline/tensorflow/."""
from . import deepfm, deepfm_v2, dien, din, embeddingmlp, neuralcf, twotowers, widendeep  # noqa: F401

__all__ = ["embeddingmlp", "widendeep", "neuralcf", "twotowers", "deepfm", "deepfm_v2", "din",
           "dien"]
