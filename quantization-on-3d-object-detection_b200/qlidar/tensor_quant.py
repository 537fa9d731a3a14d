"""Mirror of the slice of NVIDIA pytorch_quantization that quant/*.py uses (QuantDescriptor, TensorQuantizer,
calib.MaxCalibrator): symmetric, narrow-range, round-half-even fake quantisation with a dynamic or calibrated amax.
Reference call sites: quant/quant.py:1-2,14-32; quant/quantize.py:3-7,175-207.  Semantics: SURVEY.md 8a-Q.

The sparse-conv hot path does NOT run `TensorQuantizer.forward`: QConvNd reads the descriptor (bits, axis, _amax)
and dispatches the fused sm_100a kernels.  `forward` is the torch restatement used for weights at wrap time, for
the dense SQ* layers, and when a user calls the quantizer directly."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class QuantDescriptor:
    def __init__(self, num_bits: int = 8, name=None, fake_quant: bool = True, axis=None, amax=None, learn_amax=False,
                 scale_amax=None, calib_method: str = "max", unsigned: bool = False, narrow_range: bool = True):
        if unsigned:
            raise NotImplementedError("unsigned quantisation is commented out at every reference call site (quant/quant.py:24,30)")
        self.num_bits = int(num_bits)
        self.fake_quant = fake_quant
        if axis is not None and not isinstance(axis, (tuple, list)):
            axis = (int(axis),)
        self.axis = None if axis is None else tuple(int(a) for a in axis)
        self.amax = amax
        self.calib_method = calib_method
        self.narrow_range = narrow_range
        self.unsigned = unsigned


class MaxCalibrator:
    """Running max of |x| over the reduction axes ([EXT] calib.MaxCalibrator)."""

    def __init__(self, num_bits, axis, unsigned=False):
        self._axis = axis
        self._calib_amax = None

    def collect(self, x: torch.Tensor):
        a = reduce_amax(x, self._axis)
        self._calib_amax = a if self._calib_amax is None else torch.maximum(self._calib_amax, a)

    def compute_amax(self):
        return self._calib_amax

    def reset(self):
        self._calib_amax = None


def reduce_amax(x: torch.Tensor, axis) -> torch.Tensor:
    a = x.detach().abs()
    if axis is None:
        return a.max() if a.numel() else a.new_zeros(())
    keep = [ax % x.dim() for ax in axis]
    red = [d for d in range(x.dim()) if d not in keep]
    return a.amax(dim=red, keepdim=True) if red else a


def quant_scale(amax: torch.Tensor, bound: float) -> torch.Tensor:
    """scale = bound / amax (0 where amax <= 2^-24) as ONE correctly rounded fp32 division per element, like [EXT]
    pytorch_quantization (`max_bound / amax`, both tensors) and like the device (__fdiv_rn).  Python's `bound / tensor` is
    tensor.reciprocal() * bound -- two roundings, one ulp off in a quarter of the cases -- and CUDA's `tensor / scalar`
    multiplies by the reciprocal; an ulp in a scale flips int8 codes that sit on a rounding boundary."""
    amax = amax.to(torch.float32)
    tiny = amax <= (1.0 / (1 << 24))
    return torch.where(tiny, torch.zeros_like(amax), torch.full_like(amax, float(bound)) / torch.where(tiny, torch.ones_like(amax), amax))


def fake_quant(x: torch.Tensor, amax: torch.Tensor, num_bits: int) -> torch.Tensor:
    bound = float(2 ** (num_bits - 1) - 1)
    scale = quant_scale(amax, bound)
    xf = x.to(torch.float32)
    q = torch.round(xf * scale).clamp_(-bound, bound)
    out = torch.where(scale == 0, torch.zeros_like(q), q / torch.where(scale == 0, torch.ones_like(scale), scale))
    return out.to(x.dtype)


class TensorQuantizer(nn.Module):
    def __init__(self, quant_desc: Optional[QuantDescriptor] = None, disabled=False, if_quant=True, if_calib=False):
        super().__init__()
        quant_desc = quant_desc or QuantDescriptor()
        self._num_bits = quant_desc.num_bits
        self._axis = quant_desc.axis
        self._fake_quant = quant_desc.fake_quant
        self._narrow_range = quant_desc.narrow_range
        self._unsigned = quant_desc.unsigned
        self._disabled = disabled
        self._if_quant = if_quant
        self._if_calib = if_calib
        if quant_desc.amax is not None:
            self.register_buffer("_amax", torch.as_tensor(quant_desc.amax, dtype=torch.float32))
        self._calibrator = MaxCalibrator(self._num_bits, self._axis) if quant_desc.calib_method == "max" else None
        if quant_desc.calib_method not in ("max",):
            raise NotImplementedError("histogram/entropy calibration belongs to the 2-D head (SURVEY.md 8f rank 3)")

    # --- descriptor ---
    @property
    def num_bits(self):
        return self._num_bits

    @property
    def axis(self):
        return self._axis

    @property
    def amax(self):
        return getattr(self, "_amax", None)

    @amax.setter
    def amax(self, value):
        if value is None:
            if hasattr(self, "_amax"):
                delattr(self, "_amax")
            return
        v = torch.as_tensor(value, dtype=torch.float32)
        if hasattr(self, "_amax"):
            self._amax = v.to(self._amax.device)
        else:
            self.register_buffer("_amax", v)

    # --- switches used by collect_stats (quant/quantize.py:177-194) ---
    def enable_calib(self):
        if self._calibrator is None:
            raise RuntimeError("calibrator was not created")
        self._if_calib = True

    def disable_calib(self):
        self._if_calib = False

    def enable_quant(self):
        self._if_quant = True

    def disable_quant(self):
        self._if_quant = False

    def enable(self):
        self._disabled = False

    def disable(self):
        self._disabled = True

    def load_calib_amax(self, *args, strict=True, **kwargs):
        amax = self._calibrator.compute_amax() if self._calibrator is not None else None
        if amax is None:
            if strict:
                raise RuntimeError("calibrator returned None (no data was collected)")
            return
        self.amax = amax.detach().clone()

    # --- torch fake-quant (not the sparse hot path) ---
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._disabled:
            return x
        if self._if_calib:
            self._calibrator.collect(x)
        if not self._if_quant:
            return x
        amax = self.amax if self.amax is not None else reduce_amax(x, self._axis)
        return fake_quant(x, amax.to(x.device), self._num_bits)

    def extra_repr(self):
        return f"{self._num_bits} bit fake per-{'channel axis=' + str(self._axis) if self._axis else 'tensor'} amax={'dynamic' if self.amax is None else 'calibrated'}"
