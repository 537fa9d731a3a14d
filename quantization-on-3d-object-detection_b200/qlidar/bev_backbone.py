"""the dense 2-D BEV backbone (pcdet/models/backbones_2d/base_bev_backbone.py:6-113) -- the only tensor-bound
stage of CenterPoint -- as a drop-in module whose Conv2d layers run as INT8 SmoothQuant kernels after the reference's own surgery,

    smoothquant(model, {}, "", alpha, w_bits, act_bits, src=(nn.Conv2d), tgt=SQConv2d, no_list)      (quant/quant_centerpoint.py:96-106)

which replaces every nn.Conv2d (never the ConvTranspose2d de-blocks: the reference's SQConvT2d cannot run on an input with more
than one pixel, quant/smoothquant.py:231) with SQConv2d.  Same constructor, attribute names (`blocks`, `deblocks`,
`num_bev_features`), state-dict keys (`blocks.i.j.*`, `deblocks.i.j.*`) and data_dict contract as the reference class.

forward(): a [ZeroPad2d] -> SQConv2d -> BatchNorm2d(eval) -> ReLU run of a block is ONE quantise pass + ONE tcgen05 kind::i8
launch (the padding joins the unfold, BN and ReLU ride in the conv kernel's epilogue, SQConv2d.forward_fused); anything else -- fp32
Conv2d layers that were left out of the surgery, the ConvTranspose2d de-blocks, training mode -- runs layer by layer as the modules
say."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .backbones import _cfg
from .smoothquant import SQConv2d


def _bn_affine(bn: nn.BatchNorm2d):
    a = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return a, bn.bias.detach().float() - a * bn.running_mean.detach().float()


def run_block(block: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential.forward with the fusable runs collapsed (eval mode only)."""
    mods = list(block)
    i = 0
    while i < len(mods):
        m = mods[i]
        pad = 0
        j = i
        if isinstance(m, nn.ZeroPad2d) and j + 1 < len(mods) and isinstance(mods[j + 1], SQConv2d) and len(set(m.padding)) == 1:
            pad = int(m.padding[0])
            j += 1
        if isinstance(mods[j], SQConv2d) and not block.training:
            conv = mods[j]
            a = b = None
            relu = False
            k = j + 1
            if k < len(mods) and isinstance(mods[k], nn.BatchNorm2d) and mods[k].track_running_stats:
                a, b = _bn_affine(mods[k])
                k += 1
                if k < len(mods) and isinstance(mods[k], nn.ReLU):
                    relu = True
                    k += 1
            x = conv.forward_fused(x, pad, a, b, relu)
            i = k
            continue
        x = m(x)
        i += 1
    return x


class BaseBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = cfg = _cfg(model_cfg)
        layer_nums, layer_strides, num_filters = [], [], []
        if cfg.get('LAYER_NUMS', None) is not None:
            assert len(cfg.LAYER_NUMS) == len(cfg.LAYER_STRIDES) == len(cfg.NUM_FILTERS)
            layer_nums, layer_strides, num_filters = cfg.LAYER_NUMS, cfg.LAYER_STRIDES, cfg.NUM_FILTERS
        upsample_strides, num_upsample_filters = [], []
        if cfg.get('UPSAMPLE_STRIDES', None) is not None:
            assert len(cfg.UPSAMPLE_STRIDES) == len(cfg.NUM_UPSAMPLE_FILTERS)
            upsample_strides, num_upsample_filters = cfg.UPSAMPLE_STRIDES, cfg.NUM_UPSAMPLE_FILTERS

        def bn_relu(c):
            return [nn.BatchNorm2d(c, eps=1e-3, momentum=0.01), nn.ReLU()]

        c_ins = [input_channels, *num_filters[:-1]]
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        for lvl, (n_layers, stride, c_out) in enumerate(zip(layer_nums, layer_strides, num_filters)):
            layers = [nn.ZeroPad2d(1), nn.Conv2d(c_ins[lvl], c_out, kernel_size=3, stride=stride, padding=0, bias=False), *bn_relu(c_out)]
            for _ in range(n_layers):
                layers += [nn.Conv2d(c_out, c_out, kernel_size=3, padding=1, bias=False), *bn_relu(c_out)]
            self.blocks.append(nn.Sequential(*layers))
            if len(upsample_strides) > 0:
                us, c_up = upsample_strides[lvl], num_upsample_filters[lvl]
                if us > 1 or (us == 1 and not cfg.get('USE_CONV_FOR_NO_STRIDE', False)):
                    up = nn.ConvTranspose2d(c_out, c_up, us, stride=us, bias=False)
                else:
                    ds = int(np.round(1 / us))                       # (the reference's np.int is gone from numpy >= 1.24)
                    up = nn.Conv2d(c_out, c_up, ds, stride=ds, bias=False)
                self.deblocks.append(nn.Sequential(up, *bn_relu(c_up)))
        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > len(layer_nums):
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(c_in, c_in, upsample_strides[-1], stride=upsample_strides[-1], bias=False),
                                               *bn_relu(c_in)))
        self.num_bev_features = c_in

    def forward(self, data_dict):
        spatial_features = data_dict['spatial_features']
        ups, x = [], spatial_features
        for i, block in enumerate(self.blocks):
            x = run_block(block, x)
            ups.append(run_block(self.deblocks[i], x) if len(self.deblocks) > 0 else x)   # (the reference's per-stride ret_dict is never returned)
        if len(ups) > 1:
            x = torch.cat(ups, dim=1)
        elif len(ups) == 1:
            x = ups[0]
        if len(self.deblocks) > len(self.blocks):
            x = run_block(self.deblocks[-1], x)
        data_dict['spatial_features_2d'] = x
        return data_dict
