"""efficientlo-net_b200 -- B200-native (sm_100a) implementation of EfficientLO-Net's projection-aware
point-cloud hot path.  The directory name is not a Python identifier; import it as ``elo_b200``
(the alias module at the repository root) or with ``importlib.import_module("efficientlo-net_b200")``.
"""
from . import (_lib, dist, kitti, model_util, params, pointnet_util, pwclo_model, rowband, store, synth,  # noqa: F401
               tf_checkpoint, train_graph)
from .engine import PWCLOEngine, PWCLOPipeline  # noqa: F401
from .fused_conv import fused_conv_indices, fused_conv_random_k, fused_conv_select_k  # noqa: F401
from .model_util import (ProjectPC2SphericalRing, PreProcess, get_selected_idx, inv_q, mul_point_q,  # noqa: F401
                         mul_q_point, softmax_valid)
from .pointnet_util import cost_volume, down_conv, flow_predictor, get_hw_idx, up_conv  # noqa: F401
from .pwclo_model import RowBand, get_loss, get_model, placeholder_inputs  # noqa: F401
from .store import ParamStore, use_store, variable_scope  # noqa: F401

__version__ = "0.1.0"
