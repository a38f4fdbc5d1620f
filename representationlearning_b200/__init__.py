"""representationlearning_b200 — B200-native (sm_100a) RSSFormer training hot path.

Drop-in for the `RSSFormer` model path of Rongtao-Xu/RepresentationLearning (RSSFormer-TIP2023):
same module names / state_dict keys / forward contract, arithmetic in hand-written CUDA behind the
C ABI of include/rss_b200.h.  No CPU fallback: importing is cheap, but every op raises unless
librss_b200.so is built and an sm_100 GPU is present.
"""
import torch

from . import _lib  # noqa: F401

# strict fp32 parity runs must not silently drop to TF32 inside library convolutions
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
# RSS_CUDNN_BENCHMARK=1: let the library pick its convolution engines by measurement during the eager warm-up steps (every shape of
# the step is static, and the warm-up runs before the CUDA-graph capture) instead of by heuristic
import os as _os
if _os.environ.get("RSS_CUDNN_BENCHMARK", "0") != "0":
    torch.backends.cudnn.benchmark = True
from .model import HRNetFusion, MODEL, RSSFORMER_CONFIG, build_rssformer  # noqa: F401
from .modules import (FusedBNAct, GeneralTransformerBlock, InterlacedPoolAttention2, Mhca, MlpDWBN,  # noqa: F401
                      SimpleFusion8, SpatialAttention)
from .trainer import FlatSGD, GraphedTrainStep, poly_lr, train_step  # noqa: F401
from . import evalops  # noqa: F401
from .evalops import PixelMetric, Scale, TestTimeAugmentation, tta  # noqa: F401

__all__ = ["HRNetFusion", "MODEL", "RSSFORMER_CONFIG", "build_rssformer", "GeneralTransformerBlock",
           "InterlacedPoolAttention2", "Mhca", "MlpDWBN", "SimpleFusion8", "SpatialAttention", "FusedBNAct",
           "FlatSGD", "GraphedTrainStep", "poly_lr", "train_step", "PixelMetric", "Scale", "TestTimeAugmentation", "tta"]
