"""qlidar -- B200-native (sm_100a) quantized sparse-3D-conv backbone path behind the Q-LiDAR / OpenPCDet module API.

Python here is the host-side mirror of the reference's plugin interface; all compute is in libqlidar_b200.so
(include/qlidar.h).  Importing this package never falls back to a CPU implementation."""
from . import ops
from ._lib import QlidarError, lib
from .sparse import (SparseConvTensor, SparseModule, SparseSequential, SparseConvolution, SubMConv3d, SparseConv3d,
                     SubMConv2d, SparseConv2d, SparseInverseConv3d, replace_feature)
from .tensor_quant import QuantDescriptor, TensorQuantizer, MaxCalibrator, HistogramCalibrator
from .quant import QConvNd, QConv3d, QConv2d, GQConv3d, SQConv3d, q_conv3d, gq_conv3d, sq_conv3d, collect_stats, compute_amax
from .backbones import (Cfg, post_act_block, SparseBasicBlock, VoxelBackBone8x, VoxelResBackBone8x,
                        VoxelResBackBone8xVoxelNeXt, MeanVFE, DynamicMeanVFE, VoxelizeMeanVFE, VoxelGeneratorWrapper, HeightCompression)
from .engine import BackboneEngine
from . import shard
from .center_head import CenterHeadPostProcessor
from .bev_backbone import BaseBEVBackbone
from .voxelnext_head import VoxelNeXtHead, SeparateHead
from . import smoothquant as _smoothquant_mod
from .smoothquant import (SQConv2d, SQConv1d, SQConvT2d, SQLinear, SQSubM2d, SparseSQConv2d, smoothquant_layer, smoothquant)

# `import qlidar as spconv` at pcdet/utils/spconv_utils.py:3-10 keeps the reference's attribute paths working:
# spconv.__version__[2:] (:4), spconv.constants.SPCONV_USE_DIRECT_TABLE (:5), spconv.pytorch (:8), spconv.conv.SparseConvolution (:23),
# spconv.pytorch.modules.SparseModule (quant/quant.py:3)
import sys as _sys
import types as _types

# the reference's quant/quant_voxelnext.py names: QConv3d, QConv2d and the SPARSE SQConv2d(sqsubm2d, subm2d)
quant_voxelnext = _types.SimpleNamespace(QConv3d=QConv3d, QConv2d=QConv2d, SQConv2d=SparseSQConv2d)

__version__ = "2.3.6"
constants = _types.SimpleNamespace(SPCONV_USE_DIRECT_TABLE=False)
conv = _types.SimpleNamespace(SparseConvolution=SparseConvolution)
modules = _types.SimpleNamespace(SparseModule=SparseModule, SparseSequential=SparseSequential)
pytorch = _sys.modules[__name__]

__all__ = [n for n in dir() if not n.startswith("_")]
