"""Dense BEV backbone (pcdet/models/backbones_2d/base_bev_backbone.py:6-113 + the SmoothQuant surgery of
quant/quant_centerpoint.py:96-106).  tests/golden/bev_backbone.npz holds what the reference's own classes computed
(make_golden_bev.py).  CPU: the oracle's restatement against it.  GPU: qlidar.BaseBEVBackbone after qlidar.smoothquant -- int8
kernels, ZeroPad2d / BatchNorm2d / ReLU folded into the conv launch -- against it, at the north star's feature tolerance."""
import os

import numpy as np
import pytest
import torch

import qlidar_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bev_backbone.npz"))
CFG = dict(LAYER_NUMS=[1, 2], LAYER_STRIDES=[1, 2], NUM_FILTERS=[32, 64], UPSAMPLE_STRIDES=[1, 2], NUM_UPSAMPLE_FILTERS=[32, 32])
NO_LIST = ["blocks.1.4"]
PARAMS = {k[2:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("p:")}


def rel(got, ref):
    return (got.double().cpu() - ref.double()).abs().max().item() / ref.abs().max().item()


def test_oracle_restatement_reproduces_the_reference_classes():
    x = torch.from_numpy(G["x"])
    assert rel(O.base_bev_backbone(x, PARAMS, CFG), torch.from_numpy(G["y_fp32"])) <= 1e-5
    assert rel(O.base_bev_backbone(x, PARAMS, CFG, alpha=0.5, no_list=NO_LIST), torch.from_numpy(G["y_sq"])) <= 1e-5


def test_module_has_the_references_state_dict_and_attributes():
    import qlidar
    m = qlidar.BaseBEVBackbone(dict(CFG), 64)
    assert sorted(m.state_dict().keys()) == sorted(PARAMS.keys())
    assert m.num_bev_features == 64 and len(m.blocks) == 2 and len(m.deblocks) == 2
    m.load_state_dict(PARAMS, strict=True)
    # Waymo CenterPoint's shape (cfgs/waymo_models/centerpoint.yaml BACKBONE_2D) constructs too
    w = qlidar.BaseBEVBackbone(dict(LAYER_NUMS=[5, 5], LAYER_STRIDES=[1, 2], NUM_FILTERS=[128, 256], UPSAMPLE_STRIDES=[1, 2],
                                    NUM_UPSAMPLE_FILTERS=[256, 256]), 256)
    assert w.num_bev_features == 512 and len(w.blocks[0]) == 4 + 3 * 5


@pytest.mark.gpu
def test_int8_bev_backbone_reproduces_the_reference():
    import qlidar
    from qlidar.bev_backbone import run_block
    m = qlidar.BaseBEVBackbone(dict(CFG), 64)
    m.load_state_dict(PARAMS, strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(G["x"]).cuda()
    torch.backends.cudnn.allow_tf32 = False                              # the fp32 comparison below is about the module wiring, not TF32
    with torch.no_grad():
        y32 = m({"spatial_features": x})["spatial_features_2d"]
        assert rel(y32, torch.from_numpy(G["y_fp32"])) <= 1e-4          # un-quantised: plain torch modules (cuDNN fp32)
        qlidar.smoothquant(m, {}, "", 0.5, 8, 8, (torch.nn.Conv2d), qlidar.SQConv2d, NO_LIST)
        kinds = [type(mod).__name__ for mod in m.modules() if isinstance(mod, (torch.nn.Conv2d, qlidar.SQConv2d))]
        assert kinds.count("SQConv2d") == 4 and kinds.count("Conv2d") == 1
        y = m({"spatial_features": x})["spatial_features_2d"]
        ref = torch.from_numpy(G["y_sq"])
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel(y, ref) <= 1e-2, rel(y, ref)
        # the folded launch (pad + conv + BN + ReLU in one kernel) against the same layers run one by one
        z = x
        for mod in m.blocks[0]:
            z = mod(z)
        zf = run_block(m.blocks[0], x)
        assert (z - zf).abs().max().item() <= 1e-5 * z.abs().max().item()
