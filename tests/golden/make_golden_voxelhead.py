#!/usr/bin/env python
"""Golden vectors for the VoxelNeXt sparse head's decode: the reference's OWN centernet_utils.decode_bbox_from_voxels_nuscenes
(pcdet/models/model_utils/centernet_utils.py:289-354, with its _topk_1d / gather_feat_idx), imported unmodified from /root/reference
and run on CPU on seeded per-voxel head outputs, called the way VoxelNeXtHead.generate_predicted_boxes calls it
(voxelnext_head.py:433-456: hm.sigmoid(), dim.exp(), rot split, (iou + 1) / 2).  numba is stubbed as in make_golden_centerhead.py.

Run in the build container:   python tests/golden/make_golden_voxelhead.py   ->  tests/golden/voxelhead_decode.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_centerhead import load_ref, PC_RANGE, VOXEL, STRIDE, LIMIT

CASES = {
    # name: (B, C, H, W, voxels per frame, K, with_vel, with_iou, score_thresh)
    "waymo_like_iou": (2, 3, 188, 188, 2500, 100, False, True, 0.1),
    "nusc_like_vel": (3, 2, 64, 64, 700, 60, True, False, 0.1),
    "nothresh": (1, 3, 40, 40, 300, 50, False, False, None),
}


def make_case(seed, B, C, H, W, n_per, with_vel, with_iou):
    g = torch.Generator().manual_seed(seed)
    idx = []
    for b in range(B):
        cells = torch.randperm(H * W, generator=g)[:n_per].sort().values
        idx.append(torch.stack([torch.full_like(cells, b), cells // W, cells % W], 1))
    idx = torch.cat(idx, 0).int()
    N = idx.shape[0]
    d = {"hm": torch.randn(N, C, generator=g) * 1.5 - 2.0, "center": torch.rand(N, 2, generator=g),
         "center_z": torch.randn(N, 1, generator=g) * 1.5 + 0.5, "dim": torch.randn(N, 3, generator=g) * 0.4 + 0.8,
         "rot": torch.randn(N, 2, generator=g)}
    if with_vel:
        d["vel"] = torch.randn(N, 2, generator=g)
    if with_iou:
        d["iou"] = torch.rand(N, 1, generator=g) * 2.4 - 1.2          # beyond [-1, 1]: the clamp is exercised
    return idx, d


def main():
    ref = load_ref()
    out = {}
    for i, (name, (B, C, H, W, n_per, K, wv, wi, st)) in enumerate(CASES.items()):
        idx, pd = make_case(300 + i, B, C, H, W, n_per, wv, wi)
        res = ref.decode_bbox_from_voxels_nuscenes(
            batch_size=B, indices=idx.long(), obj=pd["hm"].sigmoid(), rot_cos=pd["rot"][:, 0].unsqueeze(dim=1),
            rot_sin=pd["rot"][:, 1].unsqueeze(dim=1), center=pd["center"], center_z=pd["center_z"], dim=pd["dim"].exp(), vel=pd.get("vel"),
            iou=(pd["iou"] + 1) * 0.5 if wi else None, point_cloud_range=torch.tensor(PC_RANGE), voxel_size=torch.tensor(VOXEL),
            feature_map_stride=STRIDE, K=K, score_thresh=st, post_center_limit_range=torch.tensor(LIMIT).float())
        out[f"{name}/in/indices"] = idx.numpy()
        for k, v in pd.items():
            out[f"{name}/in/{k}"] = v.numpy()
        out[f"{name}/cfg"] = np.array([B, C, K, int(wv), int(wi), -1.0 if st is None else st], np.float64)
        for b, d in enumerate(res):
            for k in ("pred_boxes", "pred_scores", "pred_labels", "pred_ious"):
                if d[k] is not None:
                    out[f"{name}/out/{b}/{k}"] = d[k].numpy().reshape(d[k].shape[0], -1) if k == "pred_ious" else d[k].numpy()
        print(name, [int(d["pred_scores"].shape[0]) for d in res])
    path = os.path.join(HERE, "voxelhead_decode.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
