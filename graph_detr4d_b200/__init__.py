"""graph_detr4d_b200 -- B200-native (sm_100a) cross-view 3D->2D feature-sampling
attention for Graph-DETR4D, behind the reference's mmcv ATTENTION module API.

(The directory is ``graph_detr4d_b200`` rather than ``graph-detr4d_b200`` because
a hyphen cannot appear in a Python package name.)
"""
from . import _lib, ops, synthetic  # noqa: F401
from .modules import (ATTENTION, Deform3DCrossAttn, Detr3DCrossAtten, Detr3DCrossAttenV2, build_attention,  # noqa: F401
                      clear_caches, inverse_sigmoid, lidar2img_device)
from .ops import (MODE_A, MODE_C, PackedFeatures, XViewConfig, pack_features,  # noqa: F401
                  xview_attention, xview_backward, xview_forward)
from . import assign, fpe, fused, frustum, loss_sync, optim  # noqa: F401,E402   (rows f2-f4 of SURVEY 8f)
from .assign import BatchedHungarianAssigner3D  # noqa: F401,E402
from .frustum import position_embeding  # noqa: F401,E402
from .fpe import position_embed_features  # noqa: F401,E402
from .loss_sync import packed_avg_factors  # noqa: F401,E402
from .optim import MultiTensorAdamW  # noqa: F401,E402

__version__ = "0.1.0"
