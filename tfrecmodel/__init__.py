"""Top-level alias so that `import tfrecmodel.din` works from the repo root; the
implementation lives in `sparrowrecsys_b200.tfrecmodel`."""
import sys

from sparrowrecsys_b200 import tfrecmodel as _impl
from sparrowrecsys_b200.tfrecmodel import (deepfm, deepfm_v2, dien, din, embeddingmlp,  # noqa: F401
                                           neuralcf, twotowers, widendeep)

__all__ = list(_impl.__all__)
for _name in __all__:
    sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
