"""efficientlo-net_b200 -- B200-native (sm_100a) implementation of EfficientLO-Net's projection-aware
point-cloud hot path.  The directory name is not a Python identifier; import it as ``elo_b200``
(the alias module at the repository root) or with ``importlib.import_module("efficientlo-net_b200")``.
"""
from . import _lib, synth  # noqa: F401
from .fused_conv import fused_conv_indices, fused_conv_random_k, fused_conv_select_k  # noqa: F401

__version__ = "0.1.0"
