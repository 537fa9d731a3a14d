"""Drop-in mirror of the reference's SmoothQuant wrapper family, executing REAL int8 kernels:

  SQConv2d / SQConv1d / SQConvT2d / SQLinear     quant/smoothquant.py:6-99, 102-176, 179-270, 273-322   (dense 2-D backbone / heads)
  SQSubM2d                                       quant/SQSubM2d.py:7-91          forward(dense) -> (weight, x)
  SparseSQConv2d  (the reference calls it SQConv2d too: quant/quant_voxelnext.py:118-135, a SparseModule over SQSubM2d + SubMConv2d)
  smoothquant_layer / smoothquant                quant/quantize.py:48-115        (__new__ + attribute copy construction, surgery walk)

Same constructors, attribute names (`_weight_quantizer`, `_input_quantizer`, `scaling_factor`, `weight`, `bias`), errors (ValueError
without a scaling_factor) and results as the reference's forward:

    cols = unfold(x);  s = max|cols|^a / max|w|^(1-a) per column, zeros -> 1;  w' = w * s, cols' = cols / s
    y = fake_quant_per_tensor(cols') @ fake_quant_per_out_channel(w').T  (+ bias)

but the unfolded matrix only ever exists as int8 codes (ql_unfold_absmax / ql_unfold_quantize), the smoothed weights are quantised
into the conv kernel's packed image on the device (ql_sq_prepare_weights) and the product is ONE tcgen05 kind::i8 launch with
INT32 accumulation (ql_spconv_mma, kernel volume 1): y = acc * (amax_x/127) * (amax_w[oc]/127) + bias, i.e. the reference's
fake-quant product up to fp32 summation order.  The sparse wrapper never densifies the (1504 x 1504) BEV grid: on a sparse tensor
every active site appears at every kernel position of the unfold, so the per-column statistics reduce to per-input-channel ones
(SQConv3d's arithmetic, qlidar/quant.py)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import QL_S8, lib
from .sparse import SparseConvTensor, SparseModule, SubMConv2d, _round_up
from .tensor_quant import QuantDescriptor, TensorQuantizer


def _one(v):
    return int(v[0]) if isinstance(v, (list, tuple)) else int(v)


def _bits(q, default=8) -> int:
    return int(getattr(q, "num_bits", default)) if q is not None else default


_IDENTITY = {}


def _identity_rulebook(m: int, dev) -> torch.Tensor:
    """nbr [tiles, 1, 128] = row r reads row r (kernel volume 1): the rulebook of a plain GEMM."""
    key = (m, str(dev))
    t = _IDENTITY.get(key)
    if t is None:
        tiles = ops.num_tiles(max(m, 1))
        t = torch.arange(tiles * ops.TILE_M, dtype=torch.int32, device=dev)
        t[m:] = -1
        t = t.view(tiles, 1, ops.TILE_M).contiguous()
        if len(_IDENTITY) > 16:
            _IDENTITY.clear()
        _IDENTITY[key] = t
    return t


def _int8_gemm(codes: torch.Tensor, act_scale: torch.Tensor, w_rows: torch.Tensor, act_absmax_cols: torch.Tensor, alpha: float,
               bias: Optional[torch.Tensor], smooth: torch.Tensor, out_dtype, post_a: Optional[torch.Tensor] = None,
               post_b: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """codes [M, Kp] int8 (already smoothed + quantised with `smooth`), w_rows [N, Kp] fp32 (columns padded like the codes):
    returns [M, N] = dequantised product with the smoothed, per-row (output channel) quantised weights.  N is cut into blocks of
    256 output channels (the kernel's widest accumulator)."""
    M, Kp = codes.shape
    N = w_rows.shape[0]
    dev = codes.device
    nbr = _identity_rulebook(M, dev)
    w_col = w_rows.abs().amax(dim=0).contiguous()
    outs = []
    for n0 in range(0, N, 256):
        wb = w_rows[n0:n0 + 256]
        nb = wb.shape[0]
        nb_p = _round_up(nb, 16)
        if nb_p != nb:
            wb = F.pad(wb, (0, 0, 0, nb_p - nb))
        sm = torch.empty(Kp, dtype=torch.float32, device=dev)
        _, packed, scale = ops.sq_prepare_weights(wb.reshape(nb_p, 1, Kp).contiguous(), w_col, act_absmax_cols, alpha, smooth=sm)
        shift = torch.zeros(nb_p, dtype=torch.float32, device=dev)
        if bias is not None:
            shift[:nb] = bias[n0:n0 + nb].float()
        if post_a is not None:
            # an affine per-output-channel op that follows the layer (eval BatchNorm: a * y + b) rides in the kernel's one FMA:
            # a * (acc * s + bias) + b = acc * (s * a) + (a * bias + b)
            a = torch.ones(nb_p, dtype=torch.float32, device=dev)
            a[:nb] = post_a[n0:n0 + nb].float()
            scale = scale * a
            shift = shift * a
            shift[:nb] += post_b[n0:n0 + nb].float()
        y = ops.spconv_mma(codes, nbr, M, None, nb_p, packed, scale, shift, act_scale=act_scale, out_dtype=out_dtype, relu=relu)
        outs.append(y[:, :nb])
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def _smooth_of(act_absmax: torch.Tensor, w_col_absmax: torch.Tensor, alpha: float) -> torch.Tensor:
    """Device-side smoothing vector (the same kernel the weight preparation uses, so codes and weights see identical values)."""
    dev = act_absmax.device
    k = act_absmax.numel()
    dummy_w = torch.zeros((16, 1, k), dtype=torch.float32, device=dev)
    sm, _, _ = ops.sq_prepare_weights(dummy_w, w_col_absmax, act_absmax, alpha)
    return sm


class _SQDense(nn.Module):
    """Shared machinery: y[M, N] from an NCHW input through the int8 unfold + GEMM."""

    def _check(self):
        if self.scaling_factor is None:
            raise ValueError("Please specify the scaling_factor parameter!")

    def _conv_rows(self, x: torch.Tensor, w2d: torch.Tensor, kernel, stride, pad, dil, bias, post_a=None, post_b=None, relu=False) -> torch.Tensor:
        if not x.is_cuda:
            raise ops.QlidarError("qlidar SmoothQuant wrappers run on CUDA tensors (there is no CPU fallback)")
        x = x.contiguous()
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        B, C, H, W = x.shape
        n_cols = C * kernel[0] * kernel[1]
        kp = _round_up(n_cols, 16)
        dev = x.device
        i32 = lambda v: (ops.C.c_int32 * 2)(int(v[0]), int(v[1]))
        absmax = torch.zeros(kp, dtype=torch.float32, device=dev)
        ops.check(lib().ql_unfold_absmax(ops._ptr(x), ops._DT[x.dtype], B, C, H, W, i32(kernel), i32(stride), i32(pad), i32(dil), ops._ptr(absmax),
                                         ops._stream()), "ql_unfold_absmax")
        w_rows = w2d.float()
        if kp != n_cols:
            w_rows = F.pad(w_rows, (0, kp - n_cols))
        w_rows = w_rows.contiguous()
        smooth = _smooth_of(absmax, w_rows.abs().amax(dim=0).contiguous(), float(self.scaling_factor))
        Ho = (H + 2 * pad[0] - dil[0] * (kernel[0] - 1) - 1) // stride[0] + 1
        Wo = (W + 2 * pad[1] - dil[1] * (kernel[1] - 1) - 1) // stride[1] + 1
        M = B * Ho * Wo
        codes = ops.zero_led_rows(M, kp, torch.int8, dev, fill=False)         # ql_unfold_quantize writes every row (padding columns too)
        scales = torch.empty(2, dtype=torch.float32, device=dev)
        ops.check(lib().ql_unfold_quantize(ops._ptr(x), ops._DT[x.dtype], B, C, H, W, i32(kernel), i32(stride), i32(pad), i32(dil), ops._ptr(absmax),
                                           ops._ptr(smooth), _bits(self._input_quantizer), kp, ops._ptr(codes), ops._ptr(scales), ops._stream()),
                  "ql_unfold_quantize")
        y = _int8_gemm(codes, scales[1:2], w_rows, absmax, float(self.scaling_factor), bias, smooth, torch.float32, post_a, post_b, relu)
        return y, (B, Ho, Wo)


class SQConv2d(_SQDense):
    """quant/smoothquant.py:6-99."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, input_quantizer=None,
                 weight_quantizer=None, device='cuda', scaling_factor=None) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self._weight_quantizer, self._input_quantizer = weight_quantizer, input_quantizer
        self.scaling_factor = scaling_factor
        k = _one(kernel_size)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, k, k, device=device))
        self.bias = nn.Parameter(torch.empty(out_channels, device=device))

    def forward(self, x):
        return self.forward_fused(x)

    def forward_fused(self, x, extra_pad: int = 0, post_a=None, post_b=None, relu: bool = False):
        """forward() with what surrounds the layer in BaseBEVBackbone's blocks folded in: a ZeroPad2d in front (`extra_pad`, added to
        the unfold's padding: zero columns either way), an eval-mode BatchNorm2d behind it (y * post_a + post_b per output channel) and
        a ReLU, both in the conv kernel's epilogue."""
        self._check()
        k, s, p, d = _one(self.kernel_size), _one(self.stride), _one(self.padding) + int(extra_pad), _one(self.dilation)
        w2d = self.weight.detach().reshape(self.out_channels, -1)
        y, (B, Ho, Wo) = self._conv_rows(x, w2d, (k, k), (s, s), (p, p), (d, d), None if self.bias is None else self.bias.detach(),
                                         post_a, post_b, relu)
        return y.view(B, Ho * Wo, self.out_channels).permute(0, 2, 1).reshape(B, self.out_channels, Ho, Wo).to(x.dtype)


class SQConv1d(_SQDense):
    """quant/smoothquant.py:102-176 (unfold with a (1, k) kernel)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, weight_quantizer=None,
                 input_quantizer=None, device="cuda", scaling_factor=None) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self._weight_quantizer, self._input_quantizer = weight_quantizer, input_quantizer
        self.scaling_factor = scaling_factor
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, _one(kernel_size), device=device))
        self.bias = nn.Parameter(torch.empty(out_channels, device=device))

    def forward(self, x):
        self._check()
        k, s, p, d = _one(self.kernel_size), _one(self.stride), _one(self.padding), _one(self.dilation)
        bs, ic, ln = x.shape
        w2d = self.weight.detach().reshape(self.out_channels, -1)
        y, (B, Ho, Wo) = self._conv_rows(x.unsqueeze(2), w2d, (1, k), (1, s), (0, p), (1, d), None if self.bias is None else self.bias.detach())
        return y.view(bs, Wo, self.out_channels).permute(0, 2, 1).contiguous().to(x.dtype)


class SQConvT2d(_SQDense):
    """quant/smoothquant.py:179-270: x [B*ih*iw, ic] @ w[ic, oc*kh*kw], per-input-channel smoothing, then col2im (F.fold)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, input_quantizer=None,
                 weight_quantizer=None, device='cuda', scaling_factor=None) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.output_padding, self.dilation = stride, padding, output_padding, dilation
        self._weight_quantizer, self._input_quantizer = weight_quantizer, input_quantizer
        self.scaling_factor = scaling_factor
        k = _one(kernel_size)
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels, k, k, device=device))
        self.bias = nn.Parameter(torch.empty(out_channels, device=device))

    def forward(self, x):
        self._check()
        k, s, p, d, op = _one(self.kernel_size), _one(self.stride), _one(self.padding), _one(self.dilation), _one(self.output_padding)
        bs, ic, ih, iw = x.shape
        w2d = self.weight.detach().reshape(self.in_channels, -1).t()                  # [oc*kh*kw, ic]: quantised per row, as the reference
        y, _ = self._conv_rows(x, w2d, (1, 1), (1, 1), (0, 0), (1, 1), None)          # a 1x1 "conv": rows = pixels, columns = ic
        y = y.view(bs, ih * iw, -1).permute(0, 2, 1)
        h_out = (ih - 1) * s - 2 * p + d * (k - 1) + op + 1
        w_out = (iw - 1) * s - 2 * p + d * (k - 1) + op + 1
        y = F.fold(y, output_size=(h_out, w_out), kernel_size=k, dilation=d, padding=p, stride=s)
        if self.bias is not None:
            y = y + self.bias.detach().view(1, self.out_channels, 1, 1)
        return y.to(x.dtype)


class SQLinear(_SQDense):
    """quant/smoothquant.py:273-322 (input [seq, bs, in_features])."""

    def __init__(self, in_features, out_features, weight_quantizer=None, input_quantizer=None, device="cuda", scaling_factor=None) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self._weight_quantizer, self._input_quantizer = weight_quantizer, input_quantizer
        self.scaling_factor = scaling_factor
        self.weight = nn.Parameter(torch.empty(out_features, in_features, device=device))
        self.bias = nn.Parameter(torch.empty(out_features, device=device))

    def forward(self, x):
        self._check()
        _, bs, _ = x.shape
        rows = x.reshape(-1, self.in_features)
        # rows x in_features is a 1x1 "image" per row: NCHW [M, C, 1, 1]
        y, _ = self._conv_rows(rows.reshape(-1, self.in_features, 1, 1), self.weight.detach(), (1, 1), (1, 1), (0, 0), (1, 1),
                               None if self.bias is None else self.bias.detach())
        return y.view(-1, bs, self.out_features).to(x.dtype)


class SQSubM2d(nn.Module):
    """quant/SQSubM2d.py:7-91 with its constructor's `super(MyQuantConv2d, self)` NameError fixed (the class cannot be
    instantiated as shipped).  forward(dense NCHW) -> (weight, x): the smoothed + fake-quantised weight in the sparse conv's layout
    (oc, kh, kw, ic) and the smoothed + fake-quantised activations folded back to (B, H, W, C) -- literally the reference's
    unfold -> scale -> quantise -> fold sequence (its fold SUMS the overlapping patches; kept, it is what the file computes).
    This dense form exists for API parity; the sparse wrapper below does not go through it."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, input_quantizer=None,
                 weight_quantizer=None, device='cuda', scaling_factor=None):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self._input_quantizer, self._weight_quantizer = input_quantizer, weight_quantizer
        self.scaling_factor = scaling_factor
        k = _one(kernel_size)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, k, k, device=device))

    def forward(self, input):
        if self.scaling_factor is None:
            raise ValueError("Please specify the scaling_factor parameter!")
        k, s, p, d = _one(self.kernel_size), _one(self.stride), _one(self.padding), _one(self.dilation)
        h_in, w_in = input.shape[2:]
        bs = input.shape[0]
        ksize = self.in_channels * k * k
        x = F.unfold(input, kernel_size=k, padding=p, stride=s)
        x = torch.transpose(x, 1, 2).reshape(-1, ksize)
        w_flat = self.weight.data.clone().view(self.out_channels, ksize)
        act_scale = x.abs().detach().max(dim=0)[0]
        w_scale = w_flat.abs().detach().max(dim=0)[0]
        scale = act_scale ** self.scaling_factor / w_scale ** (1 - self.scaling_factor)
        scale[scale == 0] = 1
        x = x / scale
        w_flat = w_flat * scale
        x = self._input_quantizer(x)
        w_flat = self._weight_quantizer(w_flat)
        x = x.reshape(bs, -1, ksize).transpose(1, 2)
        x = F.fold(x, (h_in, w_in), kernel_size=k, dilation=d, padding=p, stride=s)
        x = x.permute(0, 2, 3, 1)
        weight = w_flat.view(self.out_channels, self.in_channels, k, k).permute(0, 2, 3, 1).contiguous()
        return weight, x


class SparseSQConv2d(SparseModule):
    """quant/quant_voxelnext.py:118-135 -- `SQConv2d(sqsubm2d, subm2d)`, the SparseModule that wraps a SubMConv2d with SmoothQuant
    (exported as qlidar.quant_voxelnext_SQConv2d and, inside this module, under the reference's own name via `sparse_SQConv2d`).
    The reference densifies the sparse tensor, unfolds it, and converts back (`x.dense()` -> SQSubM2d -> `from_dense`): on the
    Waymo BEV grid that is a 1504 x 1504 x C unfold per call.  Here the tensor stays sparse: every active site appears at every
    kernel position of that unfold, so the per-column maxima are the per-input-channel maxima of the active rows, and the layer is
    W8A8 SmoothQuant per input channel -- int8 codes x int8 codes -> INT32 through the sparse conv kernel (SQConv3d's path)."""

    def __init__(self, sqsubm2d: SQSubM2d, subm2d: SubMConv2d):
        super().__init__()
        from .quant import SQConv3d
        self.sqsubm2d = sqsubm2d
        self.subm2d = subm2d
        # the reference copies the sparse conv's weight into the dense wrapper in (oc, ic, kh, kw) order (:123)
        self.sqsubm2d.weight.data = self.subm2d.weight.data.permute(0, 3, 1, 2).contiguous()
        self._impl = SQConv3d(spconv3d=subm2d, scaling_factor=sqsubm2d.scaling_factor,
                              w_bits=_bits(sqsubm2d._weight_quantizer), act_bits=_bits(sqsubm2d._input_quantizer))

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        if self.sqsubm2d.scaling_factor is None:
            raise ValueError("Please specify the scaling_factor parameter!")
        self._impl.scaling_factor = self.sqsubm2d.scaling_factor
        return self._impl(x)


def smoothquant_layer(nn_instance, quant_module, scaling_factor, w_bits, act_bits):
    """quant/quantize.py:48-76: build `quant_module` WITHOUT calling its constructor (`__new__`), copy every attribute of the fp32
    layer (tuples collapse to their first element), attach per-output-channel weight / per-tensor input quantisers."""
    if scaling_factor is None:
        raise ValueError("Please specify the scaling_factor parameter!")
    if act_bits is None or w_bits is None:
        raise ValueError("Please specify the num_bits parameter!")
    quant_instance = quant_module.__new__(quant_module)
    for k, val in vars(nn_instance).items():
        if isinstance(val, tuple):
            val = val[0]
        setattr(quant_instance, k, val)
    quant_instance._weight_quantizer = TensorQuantizer(QuantDescriptor(num_bits=w_bits, axis=(0)))
    quant_instance._input_quantizer = TensorQuantizer(QuantDescriptor(num_bits=act_bits))
    quant_instance.scaling_factor = scaling_factor
    return quant_instance


def smoothquant(model, module_dict, curr_path, alpha, w_bits, act_bits, src, tgt, no_list) -> None:
    """quant/quantize.py:79-115: recursive named_children walk; `src` instances whose dotted path is not in no_list become `tgt`."""
    for name, module in model.named_children():
        path = f"{curr_path}.{name}" if curr_path else name
        smoothquant(module, module_dict, path, alpha, w_bits, act_bits, src, tgt, no_list)
        if isinstance(module, src) and path not in no_list:
            model._modules[name] = smoothquant_layer(module, tgt, alpha, w_bits, act_bits)
    return
