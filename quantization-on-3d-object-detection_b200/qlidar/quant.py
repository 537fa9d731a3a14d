"""Drop-in mirror of the reference's 3-D quantisation API -- QConvNd (quant/quant.py:6-58; variants QConv3d/QConv2d
quant/quant_voxelnext.py:75-115,138-169; GQConv3d quant/quant_conv3d.py:70-138), q_conv3d, collect_stats, compute_amax
(quant/quantize.py:13-43,175-207) -- executing REAL integer kernels instead of fake-quant around fp32 spconv.

Mode selection per wrapper:
  act_bits <= 8, cw=False : W8A8-pt  int8 codes x int8 codes -> INT32 (tcgen05 kind::i8), dequant in the epilogue
  act_bits <= 8, cw=True  : W8A8-cw  per-input-channel fake-quant (does not factor out of the sum): fp16 fake-quant
                                      rows x exact int8-code weights (tcgen05 kind::f16, fp32 accumulate)
  act_bits  > 8           : W8A16    fp16 rows (int16 grid vs fp16 grid differ by <= 2^-11 relative) x int8-code weights
  per_row=True            : GQConv3d per-voxel-row amax, same kernel as cw
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .sparse import (SparseConvolution, SparseConvTensor, SparseModule, SubMConv3d, SparseConv3d, SubMConv2d, SparseConv2d,
                     _make_output, _round_up)
from .tensor_quant import QuantDescriptor, TensorQuantizer, MaxCalibrator, HistogramCalibrator, quant_scale, reduce_amax


class QConvNd(SparseModule):
    """Channel-wise Quant Module for SubMConvNd & SparseConvNd (same ctor and attributes as quant/quant.py:6-34)."""

    def __init__(self, module: SparseConvolution, w_bits: int, act_bits: int, cw: bool, per_row: bool = False):
        super().__init__()
        # The reference's fake-quant accepts any bit width; the real-integer kernels hold weight codes in int8 (kind::i8) or as
        # exact fp16 integers (kind::f16, exact only to 2048), so wider weights would be silently corrupted: refuse them.
        if not 2 <= int(w_bits) <= 8:
            raise ValueError(f"QConvNd: weight codes are 8-bit integers on this path (w_bits={w_bits}); the reference's "
                             "configurations are W8A8 and W8A16 (quant/quant_centerpoint.py)")
        if int(act_bits) < 2:
            raise ValueError(f"QConvNd: act_bits={act_bits}")
        self.module = module
        self.w = self.module.weight.data.clone()
        self.w_quant = TensorQuantizer(QuantDescriptor(num_bits=w_bits, axis=(0)))
        self.act_quant = TensorQuantizer(QuantDescriptor(num_bits=act_bits, axis=(1)) if cw else QuantDescriptor(num_bits=act_bits))
        self.cw = bool(cw)
        self.per_row = bool(per_row)
        self._cache = {}

    # ---- weights: per-output-channel codes (QuantDescriptor(axis=(0)) on (oc, ic*K), quant/quant.py:14-18,42-44) ----
    def weight_codes(self):
        """(codes (oc, K, ic) float, amax_w (oc,)) -- codes are integers in [-bound, bound]."""
        wt = self.module.weight.detach()
        oc, ic = wt.shape[0], wt.shape[-1]
        wm = wt.reshape(oc, -1, ic).float()
        amax = self.w_quant.amax.view(-1).to(wm.device) if self.w_quant.amax is not None else wm.abs().amax(dim=(1, 2))
        bound = float(2 ** (self.w_quant.num_bits - 1) - 1)
        codes = torch.round(wm * quant_scale(amax, bound).view(-1, 1, 1)).clamp_(-bound, bound)
        return codes, amax, bound

    def _prepared(self, dev, kind: str):
        wt = self.module.weight
        key = (wt._version, wt.data_ptr(), str(dev), kind, self.w_quant.num_bits, None if self.w_quant.amax is None else self.w_quant.amax.sum().item())
        hit = self._cache.get(kind)
        if hit is not None and hit[0] == key:
            return hit[1]
        codes, amax, bound = self.weight_codes()
        oc, K, ic = codes.shape
        ic_p = _round_up(ic, 16 if kind == "i8" else 8)
        oc_p = _round_up(oc, 16)
        w = torch.zeros((oc_p, K, ic_p), dtype=torch.int8 if kind == "i8" else torch.float16)
        w[:oc, :, :ic] = codes.to(w.dtype).cpu()
        packed = ops.pack_weights(w).to(dev)
        w_scale = torch.zeros(oc_p, dtype=torch.float32, device=dev)
        w_scale[:oc] = (amax.float().cpu() / bound).to(dev)             # de-quantisation scale amax_w[oc] / bound (a true division: on the host)
        shift = torch.zeros(oc_p, dtype=torch.float32, device=dev)
        if self.module.bias is not None:
            shift[:oc] = self.module.bias.detach().float().to(dev)
        val = (packed, ic_p, oc_p, w_scale, shift)
        self._cache[kind] = (key, val)
        return val

    def _act_absmax(self, f: torch.Tensor, n_dev) -> torch.Tensor:
        """Per-channel |x| max vector feeding the quantise kernel: calibrated `_amax` if present, else dynamic."""
        c = f.shape[1]
        a = self.act_quant.amax
        if a is not None:
            a = a.to(f.device).float().reshape(-1)
            return (a.expand(c) if a.numel() == 1 else a).contiguous()
        return ops.absmax_cols(f, n_dev)

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        conv = self.module
        aq = self.act_quant
        if aq._if_calib and isinstance(aq._calibrator, HistogramCalibrator):
            aq._calibrator.collect(x.features)                       # the histogram needs every value, not the reduced maxima
        elif aq._if_calib:                                           # collect_stats: MaxCalibrator running max
            am = ops.absmax_cols(x.features.contiguous(), x._n_dev)
            aq._calibrator.collect(am.view(1, -1) if self.cw else am.max().view(()))  # noqa: the calibrator sees reduced maxima
        if aq._disabled or not aq._if_quant:
            return conv(x)                                           # quantisers off (calibration pass): plain conv
        rb = conv.get_rulebook(x)
        f = x.features
        in_dtype = f.dtype
        f = (f if f.dtype in (torch.float16, torch.float32) else f.float()).contiguous()
        bits = aq.num_bits
        out_dtype = torch.float32 if in_dtype == torch.float32 else torch.float16
        n_dev = x._n_dev
        if bits <= 8 and not self.cw and not self.per_row:
            packed, ic_p, oc_p, w_scale, shift = self._prepared(f.device, "i8")
            q, act_scale = ops.quantize_rows(f, self._act_absmax(f, n_dev), ops.QL_Q_CODES_PER_TENSOR, bits, n_dev)
            if ic_p != conv.in_channels:
                q = torch.nn.functional.pad(q, (0, ic_p - conv.in_channels)).contiguous()
            y = ops.spconv_mma(q, rb.nbr, rb.n_out, rb.n_out_dev, oc_p, packed, w_scale, shift, act_scale=act_scale, out_dtype=out_dtype, kmask=rb.kmask)
        else:
            packed, ic_p, oc_p, w_scale, shift = self._prepared(f.device, "f16")
            if bits > 8:
                fh = f if f.dtype == torch.float16 else f.half()
            elif self.per_row:
                fh, _ = ops.quantize_rows(f, None, ops.QL_Q_FAKE_PER_ROW, bits, n_dev)
            else:
                fh, _ = ops.quantize_rows(f, self._act_absmax(f, n_dev), ops.QL_Q_FAKE_PER_CHANNEL, bits, n_dev)
            if ic_p != conv.in_channels:
                fh = torch.nn.functional.pad(fh, (0, ic_p - conv.in_channels))
            y = ops.spconv_mma(fh.contiguous(), rb.nbr, rb.n_out, rb.n_out_dev, oc_p, packed, w_scale, shift, out_dtype=out_dtype, kmask=rb.kmask)
        if oc_p != conv.out_channels:
            y = y[:, :conv.out_channels].contiguous()
        return _make_output(x, rb, y, conv.ndim)


class QConv3d(QConvNd):
    """quant/quant_voxelnext.py:75-115 (identical to QConvNd for 3-D convs)."""


class QConv2d(QConvNd):
    """quant/quant_voxelnext.py:138-169 for SubMConv2d/SparseConv2d.  The reference forgets to `replace_feature` the
    quantised activations; this mirror implements the evident intent (per-tensor act quant)."""

    def __init__(self, module, w_bits, act_bits):
        super().__init__(module, w_bits, act_bits, cw=False)


class GQConv3d(QConvNd):
    """quant/quant_conv3d.py:70-138: per-voxel-row activation amax (axis=0 on <=64-row groups; grouping does not
    change the values).  The reference writes the fake-quantised rows back in place (a bug); not mirrored.
    Same constructor as the reference -- GQConv3d(spconv3d, act_bits, w_bits, n), called with keywords at
    quant/quant_conv3d.py:290 -- and the same attribute names (.spconv3d, .orig_w)."""

    def __init__(self, spconv3d, act_bits=8, w_bits=8, n=64):
        super().__init__(spconv3d, w_bits, act_bits, cw=False, per_row=True)
        self.n = n

    @property
    def spconv3d(self):
        return self.module

    @property
    def orig_w(self):
        return self.w


def gq_conv3d(model, module_dict, curr_path, w_bits, act_bits, n) -> None:
    """The group-quantisation surgery of quant/quant_conv3d.py:280-292 (its own `q_conv3d`): every SubMConv3d/SparseConv3d
    except backbone_3d.conv_input.0 becomes GQConv3d(spconv3d=module, w_bits=..., act_bits=..., n=n)."""
    for name, module in model.named_children():
        path = f"{curr_path}.{name}" if curr_path else name
        gq_conv3d(module, module_dict, path, w_bits, act_bits, n)
        if isinstance(module, (SubMConv3d, SparseConv3d)) and path != "backbone_3d.conv_input.0":
            model._modules[name] = GQConv3d(spconv3d=module, w_bits=w_bits, act_bits=act_bits, n=n)
    return


class SQConv3d(SparseModule):
    """SmoothQuant for the sparse 3-D convs: W8A8 with the per-input-channel smoothing scale folded into the activation
    quantiser (the gather then reads int8 codes) and into the weights.

    Mirrors the two SQConv3d classes of the reference:
      * SQConv3d(spconv3d)                      quant/collect_act_conv3d.py:68-110, the only functional one: a SCALAR
        s = sqrt(max|x| / max|w|), x /= s, w *= s, per-tensor act / per-oc weight fake-quant.  A scalar cancels in both
        quantisers, so the int8 codes -- and this module's output -- equal QConvNd(8, 8, cw=False) (W8A8-pt).
      * SQConv3d(spconv3d, scaling_factor=a)    quant/quant_conv3d.py:141-236 (non-functional as shipped: dense unfold on
        the CPU with prints): its intent with the quant/smoothquant.py:72-79 formula, per input channel,
        s[ic] = amax_x[ic]^a / (max_{oc,k} |w[oc,k,ic]|)^(1-a), zeros -> 1; x' = x / s, w' = w * s; W8A8-pt on (x', w').
    act_amax: optional calibrated per-channel |x| maxima (static SmoothQuant; weights are then prepared once, on the host).
    Without it the maxima are taken from the incoming features on every call and the smoothed weights are re-quantised into the
    kernel's packed image ON THE DEVICE (ql_sq_prepare_weights) -- no host round trip, so the engine can capture the layer."""

    def __init__(self, spconv3d: SparseConvolution = None, scaling_factor: Optional[float] = None, module: SparseConvolution = None,
                 w_bits: int = 8, act_bits: int = 8, act_amax: Optional[torch.Tensor] = None):
        super().__init__()
        self.spconv3d = spconv3d if spconv3d is not None else module
        if self.spconv3d is None:
            raise ValueError("SQConv3d needs the conv to wrap")
        if act_bits > 8 or w_bits > 8:
            raise ValueError("SQConv3d is the W8A8 integer path")
        self.scaling_factor = scaling_factor
        self.w_quant = TensorQuantizer(QuantDescriptor(num_bits=w_bits, axis=(0)))
        self.act_quant = TensorQuantizer(QuantDescriptor(num_bits=act_bits))
        self.original_weight = self.spconv3d.weight.data.clone()
        self.act_amax = None if act_amax is None else act_amax.detach().float().reshape(-1).clone()
        self._cache = None
        self._dev = None

    def smoothing_scale(self, amax_ic: torch.Tensor) -> Optional[torch.Tensor]:
        """fp32 on the host, the oracle's arithmetic (oracle smoothquant_scale), so that codes are reproducible bit for bit."""
        if self.scaling_factor is None:
            return None
        a = float(self.scaling_factor)
        w = self.spconv3d.weight.detach().float().cpu()
        w_ic = w.abs().amax(dim=tuple(range(w.dim() - 1)))
        s = amax_ic.float().cpu().view(-1) ** a / w_ic ** (1.0 - a)
        return torch.where((s == 0) | ~torch.isfinite(s), torch.ones_like(s), s)

    def _prepare(self, dev, s: Optional[torch.Tensor]):
        conv = self.spconv3d
        wt = conv.weight.detach().float().cpu()
        if s is not None:
            wt = wt * s.view(*([1] * (wt.dim() - 1)), -1)
        oc, ic = wt.shape[0], wt.shape[-1]
        wm = wt.reshape(oc, -1, ic)
        bound = float(2 ** (self.w_quant.num_bits - 1) - 1)
        amax = wm.abs().amax(dim=(1, 2))
        codes = torch.round(wm * quant_scale(amax, bound).view(-1, 1, 1)).clamp_(-bound, bound)
        ic_p, oc_p = _round_up(ic, 16), _round_up(oc, 16)
        w8 = torch.zeros((oc_p, wm.shape[1], ic_p), dtype=torch.int8)
        w8[:oc, :, :ic] = codes.to(torch.int8)
        packed = ops.pack_weights(w8).to(dev)
        w_scale = torch.zeros(oc_p, dtype=torch.float32, device=dev)
        w_scale[:oc] = (amax / bound).to(dev)
        shift = torch.zeros(oc_p, dtype=torch.float32, device=dev)
        if conv.bias is not None:
            shift[:oc] = conv.bias.detach().float().to(dev)
        return packed, ic_p, oc_p, w_scale, shift

    def _device_state(self, dev):
        """fp32 weights (oc_p, K, ic_p), their per-input-channel |w| maxima, bias and the output buffers of the device-side
        preparation (ql_sq_prepare_weights) -- the dynamic per-channel path never leaves the device."""
        conv = self.spconv3d
        key = (conv.weight._version, conv.weight.data_ptr(), str(dev))
        if self._dev is not None and self._dev[0] == key:
            return self._dev[1]
        wt = conv.weight.detach().float()
        oc, ic = wt.shape[0], wt.shape[-1]
        ic_p, oc_p = _round_up(ic, 16), _round_up(oc, 16)
        w = torch.zeros((oc_p, wt.numel() // (oc * ic), ic_p), dtype=torch.float32, device=dev)
        w[:oc, :, :ic] = wt.reshape(oc, -1, ic).to(dev)
        shift = torch.zeros(oc_p, dtype=torch.float32, device=dev)
        if conv.bias is not None:
            shift[:oc] = conv.bias.detach().float().to(dev)
        st = dict(w=w, w_ic=w.abs().amax(dim=(0, 1)).contiguous(), shift=shift, ic_p=ic_p, oc_p=oc_p,
                  smooth=torch.empty(ic_p, dtype=torch.float32, device=dev), scale=torch.empty(oc_p, dtype=torch.float32, device=dev),
                  packed=torch.empty(int(ops.lib().ql_packed_weight_bytes(ic_p, oc_p, w.shape[1], ops.QL_S8)), dtype=torch.uint8, device=dev))
        self._dev = (key, st)
        return st

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        conv = self.spconv3d
        rb = conv.get_rulebook(x)
        f = x.features
        in_dtype = f.dtype
        f = (f if f.dtype in (torch.float16, torch.float32) else f.float()).contiguous()
        n_dev = x._n_dev
        static = self.act_amax is not None
        out_dtype = torch.float32 if in_dtype == torch.float32 else torch.float16
        if not static and self.scaling_factor is not None:
            # dynamic per-channel SmoothQuant, all on the device: abs-max -> smoothing scale + smoothed int8 weight image -> codes -> conv
            st = self._device_state(f.device)
            if st["ic_p"] != conv.in_channels:
                f = torch.nn.functional.pad(f, (0, st["ic_p"] - conv.in_channels)).contiguous()
            amax_ic = ops.absmax_cols(f, n_dev)
            ops.sq_prepare_weights(st["w"], st["w_ic"], amax_ic, float(self.scaling_factor), smooth=st["smooth"], packed=st["packed"], scale=st["scale"])
            q, act_scale = ops.quantize_rows(f, amax_ic, ops.QL_Q_CODES_PER_TENSOR, self.act_quant.num_bits, n_dev, smooth=st["smooth"])
            y = ops.spconv_mma(q, rb.nbr, rb.n_out, rb.n_out_dev, st["oc_p"], st["packed"], st["scale"], st["shift"], act_scale=act_scale,
                               out_dtype=out_dtype, kmask=rb.kmask)
            if st["oc_p"] != conv.out_channels:
                y = y[:, :conv.out_channels].contiguous()
            return _make_output(x, rb, y, conv.ndim)
        amax_ic = self.act_amax.to(f.device) if static else ops.absmax_cols(f, n_dev)
        if self._cache is not None and (static or self.scaling_factor is None):
            s_dev, prep = self._cache
        else:
            s = self.smoothing_scale(amax_ic)                 # None for the scalar variant: the weights are prepared once
            s_dev = None if s is None else s.to(f.device).contiguous()
            prep = self._prepare(f.device, s)
            self._cache = (s_dev, prep)
        packed, ic_p, oc_p, w_scale, shift = prep
        q, act_scale = ops.quantize_rows(f, amax_ic.contiguous(), ops.QL_Q_CODES_PER_TENSOR, self.act_quant.num_bits, n_dev, smooth=s_dev)
        if ic_p != conv.in_channels:
            q = torch.nn.functional.pad(q, (0, ic_p - conv.in_channels)).contiguous()
        y = ops.spconv_mma(q, rb.nbr, rb.n_out, rb.n_out_dev, oc_p, packed, w_scale, shift, act_scale=act_scale, out_dtype=out_dtype, kmask=rb.kmask)
        if oc_p != conv.out_channels:
            y = y[:, :conv.out_channels].contiguous()
        return _make_output(x, rb, y, conv.ndim)


def sq_conv3d(model, module_dict, curr_path, alpha, w_bits, act_bits, src, no_list) -> None:
    """Surgery for the 3-D SmoothQuant wrapper, the walk of quant/quantize.py:46-77 (`smoothquant`) applied to the sparse
    convs the way quant/quant_conv3d.py:286-292 / quant_second.py:84 intend: swap `src` instances not in no_list."""
    for name, module in model.named_children():
        path = f"{curr_path}.{name}" if curr_path else name
        sq_conv3d(module, module_dict, path, alpha, w_bits, act_bits, src, no_list)
        if isinstance(module, src) and path not in no_list:
            model._modules[name] = SQConv3d(spconv3d=module, scaling_factor=alpha, w_bits=w_bits, act_bits=act_bits)
    return


def q_conv3d(model, module_dict, curr_path, w_bits, act_bits, cw, src, no_list) -> None:
    """quant/quantize.py:13-43: recursive named_children walk; swap `src` instances whose dotted path is not in no_list."""
    for name, module in model.named_children():
        path = f"{curr_path}.{name}" if curr_path else name
        q_conv3d(module, module_dict, path, w_bits, act_bits, cw, src, no_list)
        if isinstance(module, src) and path not in no_list:
            model._modules[name] = QConvNd(module=module, w_bits=w_bits, act_bits=act_bits, cw=cw)
    return


def collect_stats(model, data_loader, n_batches=200, to_device=None):
    """quant/quantize.py:175-194 (including its `i > n_batches` break, i.e. n_batches + 2 batches)."""
    model.eval()
    for name, module in model.named_modules():
        if name.endswith("_quantizer") or name.endswith("_quant"):
            module.enable_calib()
            module.disable_quant()
    with torch.no_grad():
        for i, batch_dict in enumerate(data_loader):
            if to_device is not None:
                to_device(batch_dict)
            model(batch_dict)
            if i > n_batches:
                break
    for name, module in model.named_modules():
        if name.endswith("_quantizer") or name.endswith("_quant"):
            module.disable_calib()
            module.enable_quant()
    return


def compute_amax(model, device, **kwargs):
    """quant/quantize.py:198-207."""
    for _, module in model.named_modules():
        if isinstance(module, TensorQuantizer):
            if module._calibrator is not None:
                if isinstance(module._calibrator, MaxCalibrator):
                    module.load_calib_amax(strict=False)
                else:
                    module.load_calib_amax(**kwargs)                # method='entropy' | 'mse' | 'percentile', percentile=...
                if module.amax is not None:
                    module._amax = module._amax.to(device)
    return
