"""Golden vectors for the dense BEV backbone: the reference's OWN pcdet/models/backbones_2d/base_bev_backbone.py
class and its OWN quant/smoothquant.py SQConv2d, imported unmodified from /root/reference and run on CPU -- fp32, and after the
surgery quant/quant_centerpoint.py:96-106 performs (`smoothquant(model, ..., src=(nn.Conv2d), tgt=SQConv2d, no_list)`; the walk and
the __new__ construction of quant/quantize.py:48-115 are restated below because that file imports pcdet, which needs spconv).  The
`_weight_quantizer` / `_input_quantizer` objects are oracle/ext_stubs.py's TensorQuantizer stand-in ([EXT] pytorch_quantization).

Run in the build container:   python tests/golden/make_golden_bev.py     ->  tests/golden/bev_backbone.npz
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

CFG = dict(LAYER_NUMS=[1, 2], LAYER_STRIDES=[1, 2], NUM_FILTERS=[32, 64], UPSAMPLE_STRIDES=[1, 2], NUM_UPSAMPLE_FILTERS=[32, 32])
C_IN, SHAPE, ALPHA = 64, (2, 64, 24, 20), 0.5
NO_LIST = ["blocks.1.4"]                    # one layer kept fp32, like the heads' output layers in quant_centerpoint.py:28-71


class EasyCfg(dict):
    __getattr__ = dict.__getitem__


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def smoothquant_layer(nn_instance, quant_module, scaling_factor, w_bits, act_bits):
    q = quant_module.__new__(quant_module)
    for k, val in vars(nn_instance).items():
        if isinstance(val, tuple):
            val = val[0]
        setattr(q, k, val)
    q._weight_quantizer = ext_stubs.TensorQuantizer(ext_stubs.QuantDescriptor(num_bits=w_bits, axis=(0)))
    q._input_quantizer = ext_stubs.TensorQuantizer(ext_stubs.QuantDescriptor(num_bits=act_bits))
    q.scaling_factor = scaling_factor
    return q


def smoothquant(model, curr_path, alpha, w_bits, act_bits, src, tgt, no_list):
    for name, module in model.named_children():
        path = f"{curr_path}.{name}" if curr_path else name
        smoothquant(module, path, alpha, w_bits, act_bits, src, tgt, no_list)
        if isinstance(module, src) and path not in no_list:
            model._modules[name] = smoothquant_layer(module, tgt, alpha, w_bits, act_bits)


def main():
    ext_stubs.install()
    ref_sq = load("/root/reference/quant/smoothquant.py", "ref_smoothquant")
    ref_bev = load("/root/reference/pcdet/models/backbones_2d/base_bev_backbone.py", "ref_bev")
    torch.manual_seed(11)
    m = ref_bev.BaseBEVBackbone(EasyCfg(CFG), C_IN)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    m.eval()
    x = torch.randn(SHAPE, generator=g).abs()                     # a BEV map is post-ReLU
    x[:, 5] *= 15.0                                               # an outlier channel: what SmoothQuant is for
    x[:, :, ::3, ::2] = 0.0                                       # and it is sparse
    out = {"x": x.numpy()}
    for k, v in m.state_dict().items():
        out["p:" + k] = v.numpy()
    with torch.no_grad():
        out["y_fp32"] = m({"spatial_features": x})["spatial_features_2d"].numpy()
        smoothquant(m, "", ALPHA, 8, 8, (torch.nn.Conv2d), ref_sq.SQConv2d, NO_LIST)
        out["y_sq"] = m({"spatial_features": x})["spatial_features_2d"].numpy()
    n_sq = sum(isinstance(mod, ref_sq.SQConv2d) for mod in m.modules())
    print("SQConv2d layers:", n_sq, "out", out["y_sq"].shape, "max|y|", float(np.abs(out["y_sq"]).max()),
          "sq vs fp32 rel", float(np.abs(out["y_sq"] - out["y_fp32"]).max() / np.abs(out["y_fp32"]).max()))
    np.savez_compressed(os.path.join(HERE, "bev_backbone.npz"), **out)


if __name__ == "__main__":
    main()
