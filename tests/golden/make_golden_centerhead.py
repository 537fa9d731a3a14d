#!/usr/bin/env python
"""Golden vectors for the CenterHead decode: the reference's OWN centernet_utils.decode_bbox_from_heatmap
(pcdet/models/model_utils/centernet_utils.py:176-241, with its _topk / _transpose_and_gather_feat), imported unmodified from
/root/reference and run on CPU on seeded head maps, called the way CenterHead.generate_predicted_boxes calls it
(center_head.py:306-327: hm.sigmoid(), dim.exp(), rot split into cos / sin, (iou + 1) / 2).  The module imports numba at the top
(for circle_nms, which the head never reaches: center_head.py:347-348 raises); numba is not installed here, so a stub module
with a pass-through `jit` is registered first.  The NMS half of the path needs the reference's CUDA extension and is pinned on the
GPU box against oracle/_ref/libiou3d_ref.so instead (tests/test_gpu_centerhead.py).

Run in the build container:   python tests/golden/make_golden_centerhead.py   ->  tests/golden/centerhead_decode.npz (~300 KB)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (B, C, H, W, K, with_vel, with_iou, score_thresh)
    "waymo_like": (2, 3, 47, 47, 60, False, False, 0.1),
    "nusc_like_vel": (2, 2, 32, 40, 40, True, False, 0.1),
    "iou_head_nothresh": (1, 3, 24, 24, 30, False, True, None),
}
PC_RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
VOXEL = [0.1, 0.1, 0.15]
STRIDE = 8
LIMIT = [-70.0, -70.0, -2.0, 70.0, 70.0, 4.0]


def load_ref():
    nb = types.ModuleType("numba")
    nb.jit = lambda *a, **k: (lambda f: f)
    sys.modules.setdefault("numba", nb)
    spec = importlib.util.spec_from_file_location("ref_centernet_utils", "/root/reference/pcdet/models/model_utils/centernet_utils.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_maps(seed, B, C, H, W, with_vel, with_iou):
    g = torch.Generator().manual_seed(seed)
    hm = torch.randn(B, C, H, W, generator=g) * 1.5 - 3.0            # mostly below the 0.1 threshold, a few hundred above
    maps = {"hm": hm, "center": torch.rand(B, 2, H, W, generator=g), "center_z": torch.randn(B, 1, H, W, generator=g) * 1.5 + 0.5,
            "dim": torch.randn(B, 3, H, W, generator=g) * 0.4 + 0.8, "rot": torch.randn(B, 2, H, W, generator=g)}
    if with_vel:
        maps["vel"] = torch.randn(B, 2, H, W, generator=g)
    if with_iou:
        maps["iou"] = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    return maps


def main():
    ref = load_ref()
    out = {}
    for i, (name, (B, C, H, W, K, wv, wi, st)) in enumerate(CASES.items()):
        maps = make_maps(100 + i, B, C, H, W, wv, wi)
        pd = maps
        res = ref.decode_bbox_from_heatmap(
            heatmap=pd["hm"].sigmoid(), rot_cos=pd["rot"][:, 0].unsqueeze(dim=1), rot_sin=pd["rot"][:, 1].unsqueeze(dim=1),
            center=pd["center"], center_z=pd["center_z"], dim=pd["dim"].exp(), vel=pd.get("vel"),
            iou=(pd["iou"] + 1) * 0.5 if wi else None, point_cloud_range=PC_RANGE, voxel_size=VOXEL, feature_map_stride=STRIDE, K=K,
            circle_nms=False, score_thresh=st, post_center_limit_range=torch.tensor(LIMIT).float())
        for k, v in maps.items():
            out[f"{name}/in/{k}"] = v.numpy()
        out[f"{name}/cfg"] = np.array([B, C, H, W, K, int(wv), int(wi), -1.0 if st is None else st], np.float64)
        for b, d in enumerate(res):
            for k, v in d.items():
                out[f"{name}/out/{b}/{k}"] = v.numpy()
        print(name, [int(d["pred_scores"].shape[0]) for d in res])
    out["pc_range"] = np.array(PC_RANGE); out["voxel"] = np.array(VOXEL); out["stride"] = np.array(STRIDE); out["limit"] = np.array(LIMIT)
    path = os.path.join(HERE, "centerhead_decode.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
