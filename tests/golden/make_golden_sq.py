#!/usr/bin/env python
"""Golden vectors for the dense SmoothQuant wrapper family: the reference's OWN quant/smoothquant.py (SQConv2d, SQConv1d, SQConvT2d,
SQLinear), imported unmodified from /root/reference and constructed the way quant/quantize.py:48-76 does it (`__new__` + attribute
copy from the fp32 torch layer + TensorQuantizer objects), run on CPU.  The file only needs torch plus a TensorQuantizer; the
stand-in from oracle/ext_stubs.py supplies that ([EXT] pytorch_quantization is not installable here) -- so the vectors pin the
wrapper's own arithmetic (unfold, per-column scale, zeros -> 1, which tensor is divided / multiplied, quantiser axes, reshape / fold,
bias) and are only as good as the fake-quant restatement for the [EXT] rounding.  SQSubM2d (quant/SQSubM2d.py) cannot be imported
usefully -- its constructor raises NameError -- and is restated in the oracle instead.

Run in the build container:   python tests/golden/make_golden_sq.py     ->  tests/golden/sq_dense.npz (~100 KB)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ext_stubs

CASES = {
    # name: (fp32 layer ctor, ctor kwargs, input shape)
    "conv2d_3x3_s1": ("Conv2d", dict(in_channels=16, out_channels=32, kernel_size=3, stride=1, padding=1), (2, 16, 12, 10)),
    "conv2d_3x3_s2": ("Conv2d", dict(in_channels=16, out_channels=16, kernel_size=3, stride=2, padding=1), (2, 16, 12, 10)),
    "conv2d_1x1_head": ("Conv2d", dict(in_channels=32, out_channels=3, kernel_size=1), (1, 32, 9, 7)),
    "conv1d_k3": ("Conv1d", dict(in_channels=16, out_channels=24, kernel_size=3, padding=1), (2, 16, 20)),
    # SQConvT2d.forward does `.permute(0, 2, 1).view(-1, ic)` on a non-contiguous tensor (quant/smoothquant.py:231): it raises for every
    # input with more than one pixel, so the only shape the shipped file can run is 1 x 1 (the oracle restates it with reshape)
    "convT2d_k2_s2": ("ConvTranspose2d", dict(in_channels=16, out_channels=8, kernel_size=2, stride=2), (3, 16, 1, 1)),
    "linear": ("Linear", dict(in_features=32, out_features=48), (5, 3, 32)),
}
ALPHA = 0.5


def load_ref():
    ext_stubs.install()
    spec = importlib.util.spec_from_file_location("ref_smoothquant", "/root/reference/quant/smoothquant.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def smoothquant_layer(nn_instance, quant_module, scaling_factor, w_bits, act_bits):
    """quant/quantize.py:48-76, verbatim in behaviour (that file cannot be imported without pcdet; its construction is 12 lines)."""
    q = quant_module.__new__(quant_module)
    for k, val in vars(nn_instance).items():
        if isinstance(val, tuple):
            val = val[0]
        setattr(q, k, val)
    q._weight_quantizer = ext_stubs.TensorQuantizer(ext_stubs.QuantDescriptor(num_bits=w_bits, axis=(0)))
    q._input_quantizer = ext_stubs.TensorQuantizer(ext_stubs.QuantDescriptor(num_bits=act_bits))
    q.scaling_factor = scaling_factor
    return q


def main():
    ref = load_ref()
    tgt = {"Conv2d": ref.SQConv2d, "Conv1d": ref.SQConv1d, "ConvTranspose2d": ref.SQConvT2d, "Linear": ref.SQLinear}
    out = {}
    g = torch.Generator().manual_seed(7)
    for name, (kind, kw, shape) in CASES.items():
        layer = getattr(torch.nn, kind)(**kw)
        with torch.no_grad():
            layer.weight.copy_(torch.randn(layer.weight.shape, generator=g) * 0.2)
            layer.bias.copy_(torch.randn(layer.bias.shape, generator=g) * 0.1)
        x = torch.randn(shape, generator=g)
        x.view(-1)[::37] *= 12.0                                   # outliers: what SmoothQuant is for
        q = smoothquant_layer(layer, tgt[kind], ALPHA, 8, 8)
        with torch.no_grad():
            y = q(x.clone())
        out[name + ":x"] = x.numpy()
        out[name + ":w"] = layer.weight.detach().numpy()
        out[name + ":b"] = layer.bias.detach().numpy()
        out[name + ":y"] = y.numpy()
        print(name, tuple(x.shape), "->", tuple(y.shape), "max|y|", float(y.abs().max()))
    np.savez_compressed(os.path.join(HERE, "sq_dense.npz"), **out)


if __name__ == "__main__":
    main()
