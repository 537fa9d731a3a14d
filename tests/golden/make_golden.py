#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ by running the REFERENCE'S OWN SOURCES from /root/reference
(pcdet VoxelResBackBone8x / VoxelBackBone8x, MeanVFE, HeightCompression, quant/quant.py::QConvNd, quant/quantize.py::q_conv3d,
collect_stats, compute_amax), unmodified, on CPU.  The two third-party packages they import (spconv 2.x, pytorch_quantization)
are not installed or vendored anywhere; oracle/ext_stubs.py provides stand-ins for them built on the oracle's restatement
of their published algorithms.  So the vectors pin everything the reference itself owns on this path -- network topology,
indice_key sharing, the QConvNd permute / fake-quant / restore sequence, BN / ReLU / residual order, the module surgery
walk and its no_list, MeanVFE, HeightCompression -- and are only as good as the restatement for the [EXT] arithmetic.

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_golden.py
Output: tests/golden/backbone_mini_*.npz  (150-400 KB each).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qlidar_oracle as O
import ext_stubs

# "mini" geometry: KITTI voxel size on a 17.6 m x 20 m x 4 m crop -> grid 352 x 400 x 40, sparse_shape [41, 400, 352],
# encoded grid [2, 50, 44]; keeps every stage shape rule of the full configs with a ~1 MB BEV map.
MINI = dict(pc_range=[0.0, -10.0, -3.0, 17.6, 10.0, 1.0], voxel_size=[0.05, 0.05, 0.1], nfeat=4, max_pts=5, max_voxels=40000)
TAP_ROW_STRIDE = 8           # multi-scale taps are stored every 8th row (indices in full)
ENC_ROW_STRIDE = 4           # encoded features every 4th row; all rows enter the per-channel BEV sums


class Cfg(dict):
    """EasyDict-like: attribute access + .get (the reference reads model_cfg.NUM_BEV_FEATURES and model_cfg.get('USE_BIAS'))."""
    __getattr__ = dict.get


def mini_points(seed=1000, n_az=360):
    pts = O.synth_lidar_frame("kitti", seed, n_az=n_az)
    r = MINI["pc_range"]
    m = (pts[:, 0] >= r[0]) & (pts[:, 0] < r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] < r[4]) & (pts[:, 2] >= r[2]) & (pts[:, 2] < r[5])
    return np.ascontiguousarray(pts[m])


def run_reference(ref, arch, points, mode):
    """One forward of the reference modules.  mode = (label, w_bits, act_bits, cw, no_list, static)."""
    label, w_bits, act_bits, cw, no_list, static = mode
    grid = O.grid_size_xyz(MINI["pc_range"], MINI["voxel_size"])
    voxels, coords3, num = O.voxelize_hard(points, MINI["pc_range"], MINI["voxel_size"], MINI["max_pts"], MINI["max_voxels"])
    coords = np.concatenate([np.zeros((coords3.shape[0], 1), np.int32), coords3], axis=1)       # dataset.py:237-244 batch column
    bb_mod = ref["spconv_backbone"]
    backbone = getattr(bb_mod, arch)(Cfg(), MINI["nfeat"], np.asarray(grid))
    prog = O.backbone_specs(arch, MINI["nfeat"])
    missing = backbone.load_state_dict(O.init_params(prog), strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    backbone.eval()
    vfe = ref["mean_vfe"].MeanVFE(Cfg(), MINI["nfeat"])
    hc = ref["height_compression"].HeightCompression(Cfg(NUM_BEV_FEATURES=256))
    if label != "fp32":
        sp = sys.modules["spconv.pytorch"]
        ref["quantize"].q_conv3d(backbone, {}, "", w_bits, act_bits, cw, (sp.SubMConv3d, sp.SparseConv3d), list(no_list))

    def batch():
        # models/__init__.py:23-36 load_data_to_gpu casts every array to float (coords and num_points included)
        return {"voxels": torch.from_numpy(voxels), "voxel_num_points": torch.from_numpy(num).float(),
                "voxel_coords": torch.from_numpy(coords).float(), "batch_size": 1}

    def forward(bd):
        with torch.no_grad():
            return hc(backbone(vfe(bd)))

    amax = {}
    if static:
        # the reference's own static calibration, quantize.py:175-207: max calibrator over the "data loader", then freeze _amax
        class Pipeline(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.vfe, self.backbone_3d, self.map_to_bev = vfe, backbone, hc

            def forward(self, bd):
                return self.map_to_bev(self.backbone_3d(self.vfe(bd)))

        model = Pipeline()
        ref["quantize"].collect_stats(model, [batch()], n_batches=0)
        ref["quantize"].compute_amax(model, torch.device("cpu"))
        for name, m in backbone.named_modules():
            if name.endswith("act_quant"):
                amax[name] = m.amax.reshape(-1).numpy().copy()
    bd = forward(batch())
    enc = bd["encoded_spconv_tensor"]
    out = {"points": points, "voxel_coords": coords, "voxel_num_points": num, "voxel_features": bd["voxel_features"].numpy(),
           "encoded_features_strided": enc.features.numpy()[::ENC_ROW_STRIDE], "encoded_indices": enc.indices.numpy().astype(np.int32),
           "encoded_shape": np.asarray(enc.spatial_shape, np.int32),
           "spatial_features_shape": np.asarray(bd["spatial_features"].shape, np.int32),
           "spatial_features_sum": np.float64(bd["spatial_features"].double().sum().item()),
           "spatial_features_abs_sum_per_channel": bd["spatial_features"].double().abs().sum(dim=(0, 2, 3)).numpy()}
    if label == "fp32":
        out["spatial_features_f16"] = bd["spatial_features"].numpy().astype(np.float16)
    for k, t in bd["multi_scale_3d_features"].items():
        out[k + "_indices"] = t.indices.numpy().astype(np.int32)
        out[k + "_features_strided"] = t.features.numpy()[::TAP_ROW_STRIDE]
    for k, v in amax.items():
        out["amax:" + k] = v
    quantized = [n for n, m in backbone.named_modules() if type(m).__name__ == "QConvNd"]
    out["quantized_modules"] = np.asarray(quantized)
    return out


MODES = [
    # label, w_bits, act_bits, cw, no_list (named_children dotted paths, quant_centerpoint.py:24-26), static calibration
    ("fp32", 0, 0, False, (), False),
    ("w8a16_cw", 8, 16, True, ("conv_input.0",), False),       # the bench configuration ("progressive", sq=True -> cw=True)
    ("w8a8_cw", 8, 8, True, ("conv_input.0",), False),         # repo default with sq=True
    ("w8a8_pt", 8, 8, False, (), False),                       # quant_centerpoint.py:115 non-SQ mode: every conv, per-tensor act
    ("w8a8_pt_static", 8, 8, False, ("conv_input.0",), True),  # static=True path (collect_stats / compute_amax)
]


def main():
    ref = ext_stubs.load_reference()
    pts = mini_points()
    for arch in ("VoxelResBackBone8x", "VoxelBackBone8x"):
        for mode in MODES:
            if arch == "VoxelBackBone8x" and mode[0] not in ("fp32", "w8a8_cw"):
                continue
            out = run_reference(ref, arch, pts, mode)
            path = os.path.join(HERE, f"backbone_mini_{arch}_{mode[0]}.npz")
            np.savez_compressed(path, **out)
            print(f"{os.path.basename(path)}: {out['voxel_coords'].shape[0]} voxels -> {out['encoded_indices'].shape[0]} encoded sites, "
                  f"{len(out['quantized_modules'])} QConvNd, {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
