"""CPU stand-ins for the two un-vendored third-party packages the reference imports -- spconv 2.x and NVIDIA
pytorch_quantization -- so that the reference's OWN Python sources (quant/quant.py, quant/quantize.py,
pcdet/models/backbones_3d/spconv_backbone.py, .../vfe/mean_vfe.py, .../map_to_bev/height_compression.py) can be imported
and executed, unmodified, from /root/reference in the build container.

TEST INFRASTRUCTURE, not product code: only tests/golden/make_golden.py uses it, to generate the committed golden vectors
under tests/golden/.  The stubs implement the packages' *published* behaviour with the oracle's restatements
(qlidar_oracle.py: rulebooks, gather-GEMM-scatter, TensorQuantizer arithmetic); everything above them -- network
topology, indice_key sharing, the QConvNd weight permute / fake-quant / restore dance, BatchNorm/ReLU/residual order,
MeanVFE, HeightCompression -- is the reference's code, which is what the golden vectors pin.

API surface mirrored (only what the reference touches):
  spconv.__version__, spconv.constants, spconv.pytorch.{SparseConvTensor, SparseModule, SparseSequential, SubMConv3d,
  SparseConv3d, SparseInverseConv3d (constructor only), conv.SparseConvolution}, spconv.pytorch.modules.SparseModule
  pytorch_quantization.tensor_quant.QuantDescriptor, pytorch_quantization.nn.modules.tensor_quantizer.TensorQuantizer,
  pytorch_quantization.{calib, nn, nn.modules._utils} (names only; quantize.py imports them at module level)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

import qlidar_oracle as O


# ------------------------------------------------------------------------------------------------ spconv 2.x
class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(v) for v in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, indice_dict=self.indice_dict)

    def dense(self, channels_first=True):
        d = O.to_dense(self.features, self.indices.numpy(), self.spatial_shape, self.batch_size)   # (B, C, *spatial)
        return d if channels_first else d.permute(0, 2, 3, 4, 1).contiguous()


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 subm=False, indice_key=None, **kw):
        super().__init__()
        assert ndim == 3 and dilation == 1 and groups == 1
        self.ndim, self.in_channels, self.out_channels = ndim, in_channels, out_channels
        self.kernel_size = list(O._triple(kernel_size))
        self.stride = list(O._triple(stride))
        self.padding = list(O._triple(padding))
        self.subm, self.indice_key = subm, indice_key
        self.weight = nn.Parameter(torch.zeros((out_channels, *self.kernel_size, in_channels)))    # spconv-2 layout
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x: SparseConvTensor):
        coords = x.indices.numpy().astype(np.int32)
        hit = x.indice_dict.get(self.indice_key) if self.indice_key is not None else None
        if hit is None:
            if self.subm:
                hit = (coords, x.spatial_shape, O.rulebook_subm(coords, x.spatial_shape, self.kernel_size))
            else:
                hit = O.rulebook_strided(coords, x.spatial_shape, self.kernel_size, self.stride, self.padding)
            if self.indice_key is not None:
                x.indice_dict[self.indice_key] = hit
        out_coords, out_shape, nbr = hit
        y = O.sparse_conv(x.features.float(), nbr, self.weight.data.float(), None if self.bias is None else self.bias.data.float())
        return SparseConvTensor(y, torch.from_numpy(np.ascontiguousarray(out_coords)), list(out_shape), x.batch_size,
                                indice_dict=x.indice_dict)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, 1, padding, dilation, groups, bias, True, indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False, indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, 1, 0, 1, 1, bias, False, indice_key)

    def forward(self, x):
        raise NotImplementedError("UNet-only layer; not on the hot path")


# ------------------------------------------------------------------------------------------------ pytorch_quantization
class QuantDescriptor:
    def __init__(self, num_bits=8, name=None, fake_quant=True, axis=None, amax=None, learn_amax=False, scale_amax=None,
                 calib_method="max", unsigned=False, narrow_range=True, **kw):
        assert fake_quant and not unsigned and narrow_range, "the reference only uses the defaults (SURVEY.md 8a-Q)"
        self.num_bits, self.axis, self.amax, self.calib_method = num_bits, axis, amax, calib_method


class MaxCalibrator:
    """calib.MaxCalibrator: running max of the dynamic amax (kept inside TensorQuantizer below)."""


class TensorQuantizer(nn.Module):
    def __init__(self, quant_desc=QuantDescriptor(), disabled=False, if_quant=True, if_clip=False, if_calib=False):
        super().__init__()
        self._num_bits, self._axis = quant_desc.num_bits, quant_desc.axis
        self._calibrator = MaxCalibrator() if quant_desc.calib_method == "max" else None
        self._disabled, self._if_quant, self._if_calib = disabled, if_quant, if_calib
        self._calib_amax = None
        if quant_desc.amax is not None:
            self.register_buffer("_amax", torch.as_tensor(quant_desc.amax, dtype=torch.float32))

    num_bits = property(lambda self: self._num_bits)
    axis = property(lambda self: self._axis)
    amax = property(lambda self: getattr(self, "_amax", None))

    def enable_calib(self): self._if_calib = True
    def disable_calib(self): self._if_calib = False
    def enable_quant(self): self._if_quant = True
    def disable_quant(self): self._if_quant = False

    def load_calib_amax(self, *a, **kw):
        if self._calib_amax is not None:
            self.register_buffer("_amax", self._calib_amax.clone())

    def forward(self, inputs):
        if self._disabled:
            return inputs
        if self._if_calib:                                      # MaxCalibrator: running max of the dynamic amax
            am = O.dynamic_amax(inputs.detach().float(), self._axis)
            self._calib_amax = am if self._calib_amax is None else torch.maximum(self._calib_amax, am)
        if not self._if_quant:
            return inputs
        return O.fake_quant(inputs.float(), self._num_bits, self._axis, self.amax).to(inputs.dtype)


# ------------------------------------------------------------------------------------------------ installation
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install() -> None:
    """Register the stand-in packages in sys.modules (idempotent)."""
    if "spconv" in sys.modules and getattr(sys.modules["spconv"], "_qlidar_stub", False):
        return
    conv = _mod("spconv.pytorch.conv", SparseConvolution=SparseConvolution)
    modules = _mod("spconv.pytorch.modules", SparseModule=SparseModule, SparseSequential=SparseSequential)
    sp = _mod("spconv.pytorch", SparseConvTensor=SparseConvTensor, SparseModule=SparseModule, SparseSequential=SparseSequential,
              SubMConv3d=SubMConv3d, SparseConv3d=SparseConv3d, SparseInverseConv3d=SparseInverseConv3d, conv=conv, modules=modules)
    consts = _mod("spconv.constants", SPCONV_USE_DIRECT_TABLE=True)
    root = _mod("spconv", __version__="2.3.6", pytorch=sp, constants=consts, _qlidar_stub=True)
    root.__path__ = []
    sp.__path__ = []
    tq = _mod("pytorch_quantization.nn.modules.tensor_quantizer", TensorQuantizer=TensorQuantizer)
    utils = _mod("pytorch_quantization.nn.modules._utils")
    nnmods = _mod("pytorch_quantization.nn.modules", tensor_quantizer=tq, _utils=utils)
    nnmods.__path__ = []
    pqnn = _mod("pytorch_quantization.nn", TensorQuantizer=TensorQuantizer, modules=nnmods)
    pqnn.__path__ = []
    tquant = _mod("pytorch_quantization.tensor_quant", QuantDescriptor=QuantDescriptor)
    calib = _mod("pytorch_quantization.calib", MaxCalibrator=MaxCalibrator)
    pq = _mod("pytorch_quantization", nn=pqnn, tensor_quant=tquant, calib=calib)
    pq.__path__ = []


def load_reference(ref_root: str = "/root/reference") -> dict:
    """Import the reference's own source files for the hot path under the stand-ins.  Returns the loaded modules."""
    install()

    def pkg(name):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        return sys.modules[name]

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    for n in ("pcdet", "pcdet.utils", "pcdet.models", "pcdet.models.backbones_3d", "pcdet.models.backbones_3d.vfe",
              "pcdet.models.backbones_2d", "pcdet.models.backbones_2d.map_to_bev"):
        pkg(n)
    sys.modules["pcdet.models"].load_data_to_gpu = lambda batch_dict: batch_dict      # quantize.py imports the name only
    out = {}
    out["spconv_utils"] = load("pcdet.utils.spconv_utils", "pcdet/utils/spconv_utils.py")
    out["spconv_backbone"] = load("pcdet.models.backbones_3d.spconv_backbone", "pcdet/models/backbones_3d/spconv_backbone.py")
    out["vfe_template"] = load("pcdet.models.backbones_3d.vfe.vfe_template", "pcdet/models/backbones_3d/vfe/vfe_template.py")
    out["mean_vfe"] = load("pcdet.models.backbones_3d.vfe.mean_vfe", "pcdet/models/backbones_3d/vfe/mean_vfe.py")
    out["height_compression"] = load("pcdet.models.backbones_2d.map_to_bev.height_compression",
                                     "pcdet/models/backbones_2d/map_to_bev/height_compression.py")
    out["quant"] = load("quant", "quant/quant.py")                                     # quantize.py does `from quant import QConvNd`
    out["quantize"] = load("quantize", "quant/quantize.py")
    return out
