#!/usr/bin/env python
"""Full-size fixture for BASELINE configs[1] (CenterPoint Waymo, batch 4, seeds 1000-1003, ~146 k voxels per frame with the
reference's per-frame MAX_NUMBER_OF_VOXELS cap): the oracle run ONCE in the build container (minutes of CPU), reduced to a few
hundred KB.  tests/test_gpu_fullsize.py compares the CUDA engine (BackboneEngine.forward_points on the same synthetic points)
with it on the GPU box, where the oracle itself would take too long.

  w8a16_cw        reference math (QConvNd(8, 16, cw=True), quant/quant.py:36-58 restated): stage counts, encoded indices, every
                  16th encoded feature row, per-channel sums of the encoded features (== the BEV map's per-channel-pair sums)
  w8a8_pt_static  the kernel-numerics mirror (oracle mirror_backbone_w8a8_pt) with per-layer amax taken from an fp32 forward
                  (what collect_stats / compute_amax would freeze, quantize.py:175-207): per-layer checksums of the int8 codes
                  and of the stored fp16 rows, every 16th encoded row -- compared with tolerance ZERO

Run:  python tests/golden/make_golden_fullsize.py   ->  tests/golden/fullsize_waymo_b4.npz
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qlidar_oracle as O

ROW_STRIDE = 16
BATCH = 4


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    c = O.CONFIGS["waymo"]
    pts = O.synth_batch("waymo", BATCH)
    feats, coords, _ = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    sshape = O.sparse_shape_zyx(grid)
    order = np.argsort(O._lin(coords, sshape), kind="stable")               # the engine's stage-1 order
    feats, coords = np.ascontiguousarray(feats[order]), np.ascontiguousarray(coords[order])
    # the mirror leg takes the voxel means in the CUDA voxeliser's (left-to-right) summation order
    feats_seq = np.ascontiguousarray(O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"], sequential=True)[0][order])
    prog = O.backbone_specs("VoxelResBackBone8x", c["nfeat"])
    P = O.init_params(prog)
    out = {"n_points": np.int64(pts.shape[0]), "n_voxels": np.int64(coords.shape[0]), "row_stride": np.int64(ROW_STRIDE)}
    no_list = ("conv_input.0",)

    t0 = time.time()
    ref, taps = O.backbone_forward(prog, P, torch.from_numpy(feats), coords, sshape, BATCH,
                                   O.QuantCfg(mode="ref", w_bits=8, act_bits=16, cw=True, no_list=no_list, fast=True))
    print(f"w8a16_cw reference math: {time.time() - t0:.1f} s, encoded {ref.coords.shape[0]} sites")
    out["stage_counts"] = np.asarray([coords.shape[0]] + [taps[f"x_conv{i}"].coords.shape[0] for i in (2, 3, 4)] + [ref.coords.shape[0]], np.int64)
    out["encoded_indices"] = ref.coords.astype(np.int16)
    out["w8a16_cw:encoded_rows"] = ref.features[::ROW_STRIDE].numpy().astype(np.float16)
    out["w8a16_cw:encoded_channel_sums"] = ref.features.double().sum(dim=0).numpy()
    out["w8a16_cw:encoded_abs_max"] = np.float64(ref.features.abs().max().item())
    for k in ("x_conv1", "x_conv2", "x_conv3", "x_conv4"):
        out[f"w8a16_cw:{k}_channel_sums"] = taps[k].features.double().sum(dim=0).numpy()

    t0 = time.time()
    rec = {}
    O.backbone_forward(prog, P, torch.from_numpy(feats), coords, sshape, BATCH, O.QuantCfg(mode="fp32", fast=True), rec)
    amax = {s.name: float(rec[s.name + ".in"][0].abs().max().item()) for s in O.all_conv_specs(prog) if s.name not in no_list}
    print(f"fp32 forward for the calibration amax: {time.time() - t0:.1f} s")
    del rec
    t0 = time.time()
    mrec, mout, _ = O.mirror_backbone_w8a8_pt(prog, P, feats_seq, coords, sshape, BATCH, no_list=no_list, act_amax=amax, fast=True, keep=False)
    print(f"w8a8_pt static mirror: {time.time() - t0:.1f} s")
    names = [s.name for s in O.all_conv_specs(prog)]
    out["w8a8_pt_static:layers"] = np.asarray(names)
    out["w8a8_pt_static:amax"] = np.asarray([amax.get(n, 0.0) for n in names], np.float64)
    out["w8a8_pt_static:codes_sum"] = np.asarray([mrec[n]["codes_sum"] or 0 for n in names], np.int64)
    out["w8a8_pt_static:codes_abs_sum"] = np.asarray([mrec[n]["codes_abs_sum"] or 0 for n in names], np.int64)
    out["w8a8_pt_static:out_bits_sum"] = np.asarray([mrec[n]["out_bits_sum"] for n in names], np.uint64)
    out["w8a8_pt_static:n_out"] = np.asarray([mrec[n]["n_out"] for n in names], np.int64)
    out["w8a8_pt_static:encoded_rows"] = mout.features[::ROW_STRIDE].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "fullsize_waymo_b4.npz"), **out)
    print("stage counts", out["stage_counts"], "file", os.path.getsize(os.path.join(HERE, "fullsize_waymo_b4.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
