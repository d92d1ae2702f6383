"""u2gnn_b200 — B200-native U2GNN train-step engine behind the reference's module surface.

Importing this package loads libu2gnn_b200.so (built by graph-transformer_b200/build.py); there is
no CPU or stock-PyTorch fallback for the hot path.
"""
from ._lib import LIB, LIB_PATH, SIGNATURES, require_device  # noqa: F401
from .engine import DropoutCfg  # noqa: F401
from .model import (  # noqa: F401
    LogUniformSampler,
    SampledSoftmax,
    TransformerU2GNN,
    TransformerU2GNNUnSup,
    label_smoothing,
)

__all__ = ["TransformerU2GNN", "TransformerU2GNNUnSup", "SampledSoftmax", "LogUniformSampler", "label_smoothing",
           "DropoutCfg", "LIB", "require_device"]
