"""VoxelNeXt sparse head (pcdet/models/dense_heads/voxelnext_head.py + centernet_utils.py:243-354).
  * CPU: the oracle's restatement against what the reference's own decode_bbox_from_voxels_nuscenes computed
    (tests/golden/voxelhead_decode.npz, made by make_golden_voxelhead.py).
  * GPU: ql_voxelhead_decode against the same golden vectors; the whole head -- SeparateHead sparse convs (conv + BN + ReLU in one
    launch) + generate_predicted_boxes, class-agnostic NMS and the IoU-branch's per-class re-scored NMS -- against the oracle.
Tolerances: labels, counts, order exact on the golden cases; fp32 box values 2e-6 relative (expf / atan2f / sigmoid differ by <= 1 ulp
between implementations); head features 1e-2 of max (fp16 activations between the two convs of a branch)."""
import os

import numpy as np
import pytest
import torch

import qlidar_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "voxelhead_decode.npz"))
PC_RANGE, VOXEL, STRIDE, LIMIT = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], [0.1, 0.1, 0.15], 8, [-70.0, -70.0, -2.0, 70.0, 70.0, 4.0]
CASES = ["waymo_like_iou", "nusc_like_vel", "nothresh"]


def _case(name):
    B, C, K, wv, wi, st = Z[f"{name}/cfg"]
    ins = {k: Z[f"{name}/in/{k}"] for k in ("indices", "hm", "center", "center_z", "dim", "rot")}
    ins["vel"] = Z[f"{name}/in/vel"] if wv else None
    ins["iou"] = Z[f"{name}/in/iou"] if wi else None
    ref = []
    for b in range(int(B)):
        d = {k: Z[f"{name}/out/{b}/{k}"] for k in ("pred_boxes", "pred_scores", "pred_labels")}
        if wi:
            d["pred_iou"] = Z[f"{name}/out/{b}/pred_ious"].reshape(-1)
        ref.append(d)
    return int(B), int(C), int(K), (None if st < 0 else float(st)), ins, ref


@pytest.mark.parametrize("name", CASES)
def test_oracle_decode_reproduces_the_reference(name):
    B, C, K, st, ins, ref = _case(name)
    got = O.voxelhead_decode(ins["hm"], ins["center"], ins["center_z"], ins["dim"], ins["rot"], ins["vel"], ins["iou"], ins["indices"], B, K,
                             STRIDE, VOXEL, PC_RANGE, LIMIT, st)
    for g, r in zip(got, ref):
        assert np.array_equal(g["pred_labels"], r["pred_labels"].astype(np.int32))
        np.testing.assert_allclose(g["pred_scores"], r["pred_scores"], rtol=1e-6)
        np.testing.assert_allclose(g["pred_boxes"], r["pred_boxes"], rtol=2e-6, atol=2e-6)
        if "pred_iou" in r:
            np.testing.assert_allclose(g["pred_iou"], r["pred_iou"], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_decode_reproduces_the_reference(name):
    from qlidar import ops
    B, C, K, st, ins, ref = _case(name)
    c = {k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in ins.items()}
    boxes, scores, labels, iou, count = ops.voxelhead_decode(c["hm"], c["center"], c["center_z"], c["dim"], c["rot"], c["vel"], c["iou"],
                                                             c["indices"], None, B, K, STRIDE, VOXEL, PC_RANGE, LIMIT, st)
    count = count.cpu().numpy()
    for b, d in enumerate(ref):
        n = d["pred_scores"].shape[0]
        assert count[b] == n, (b, count[b], n)
        assert np.array_equal(labels[b, :n].cpu().numpy(), d["pred_labels"].astype(np.int32))
        np.testing.assert_allclose(scores[b, :n].cpu().numpy(), d["pred_scores"], rtol=2e-6)
        np.testing.assert_allclose(boxes[b, :n].cpu().numpy(), d["pred_boxes"], rtol=2e-6, atol=2e-6)
        if "pred_iou" in d:
            np.testing.assert_allclose(iou[b, :n].cpu().numpy(), d["pred_iou"], rtol=2e-6)


def _head_cfg(iou_branch, channels):
    hd = {'center': {'out_channels': 2, 'num_conv': 2}, 'center_z': {'out_channels': 1, 'num_conv': 2}, 'dim': {'out_channels': 3, 'num_conv': 2},
          'rot': {'out_channels': 2, 'num_conv': 2}}
    nms = dict(NMS_TYPE='nms_gpu', NMS_THRESH=0.7, NMS_PRE_MAXSIZE=1000, NMS_POST_MAXSIZE=83)
    if iou_branch:                                          # tools/cfgs/waymo_models/voxelnext_ioubranch_large.yaml:20-62
        hd['iou'] = {'out_channels': 1, 'num_conv': 2}
        nms = dict(NMS_TYPE='nms_gpu', NMS_THRESH=[0.8, 0.55, 0.55], NMS_PRE_MAXSIZE=[2048, 1024, 1024], NMS_POST_MAXSIZE=[200, 150, 150])
    return dict(CLASS_AGNOSTIC=False, INPUT_FEATURES=channels, CLASS_NAMES_EACH_HEAD=[['Vehicle', 'Pedestrian', 'Cyclist']],
                SHARED_CONV_CHANNEL=channels, KERNEL_SIZE_HEAD=3, USE_BIAS_BEFORE_NORM=True, NUM_HM_CONV=2, IOU_BRANCH=iou_branch,
                RECTIFIER=[0.68, 0.71, 0.65], SEPARATE_HEAD_CFG=dict(HEAD_ORDER=['center', 'center_z', 'dim', 'rot'], HEAD_DICT=hd),
                TARGET_ASSIGNER_CONFIG=dict(FEATURE_MAP_STRIDE=8),
                POST_PROCESSING=dict(SCORE_THRESH=0.1, POST_CENTER_LIMIT_RANGE=[-75.2, -75.2, -2, 75.2, 75.2, 4], MAX_OBJ_PER_SAMPLE=500, NMS_CONFIG=nms))


def test_head_state_dict_names_are_the_references():
    import qlidar
    h = qlidar.VoxelNeXtHead(_head_cfg(True, 32), 32, 3, ['Vehicle', 'Pedestrian', 'Cyclist'], [1504, 1504, 40], PC_RANGE, VOXEL)
    keys = set(h.state_dict().keys())
    for name in ("center", "center_z", "dim", "rot", "iou", "hm"):
        for k in ("0.0.weight", "0.0.bias", "0.1.weight", "0.1.running_mean", "1.weight", "1.bias"):
            assert f"heads_list.0.{name}.{k}" in keys
    assert tuple(h.heads_list[0].hm[1].weight.shape) == (3, 1, 1, 32) and float(h.heads_list[0].hm[1].bias[0]) == pytest.approx(-2.19)
    with pytest.raises(NotImplementedError):
        h.train()({"encoded_spconv_tensor": None})


@pytest.mark.gpu
@pytest.mark.parametrize("iou_branch", [False, True])
def test_head_forward_matches_the_oracle(iou_branch):
    import qlidar
    rng = np.random.default_rng(21 + int(iou_branch))
    B, H, W, Cc = 2, 188, 188, 64
    yx = [np.unique(np.clip(rng.integers(0, H, size=(1200, 2)) + rng.integers(-1, 2, size=(1200, 2)), 0, H - 1), axis=0) for _ in range(B)]
    # object-like clusters: the NMS must have overlapping boxes to remove
    yx = [np.unique(np.concatenate([c, c[:300] + [0, 1], c[:300] + [1, 0]]), axis=0) for c in yx]
    yx = [c[(c[:, 0] < H) & (c[:, 1] < W)] for c in yx]
    idx = np.concatenate([np.concatenate([np.full((c.shape[0], 1), b), c], 1) for b, c in enumerate(yx)]).astype(np.int32)
    N = idx.shape[0]
    feats = np.maximum(rng.normal(size=(N, Cc)), 0).astype(np.float32)
    torch.manual_seed(5)
    head = qlidar.VoxelNeXtHead(_head_cfg(iou_branch, Cc), Cc, 3, ['Vehicle', 'Pedestrian', 'Cyclist'], [1504, 1504, 40], PC_RANGE, VOXEL)
    with torch.no_grad():
        for m in head.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1); m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
        sh = head.heads_list[0]
        sh.hm[1].weight.mul_(8.0); sh.hm[1].bias.fill_(-2.0)         # scores spread around the 0.1 threshold
        sh.dim[1].bias.copy_(torch.log(torch.tensor([4.5, 2.0, 1.6])))
        sh.center_z[1].bias.fill_(1.0)
    head = head.cuda().eval()
    x = qlidar.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(idx).cuda(), [H, W], B)
    with torch.no_grad():
        out = head({"encoded_spconv_tensor": x, "batch_size": B})
    got_feats = {k: v.float().cpu().numpy() for k, v in head.forward_ret_dict["pred_dicts"][0].items()}

    # ---- oracle: the same branches (fp32 math), then the reference's post-processing restated
    sd = {k: v.detach().float().cpu() for k, v in head.state_dict().items()}
    coords4 = np.concatenate([idx[:, :1], np.zeros((N, 1), np.int32), idx[:, 1:]], 1)
    nbr3 = O.rulebook_subm(coords4, [1, H, W], (1, 3, 3))
    nbr1 = O.rulebook_subm(coords4, [1, H, W], (1, 1, 1))
    ref_feats = {}
    for name in head.heads_list[0].sep_head_dict:
        p = f"heads_list.0.{name}."
        w = sd[p + "0.0.weight"]                                                  # (oc, kh, kw, ic)
        y = O.sparse_conv(torch.from_numpy(feats).half().double(), nbr3, w.half().double().reshape(Cc, 1, 3, 3, Cc)) + sd[p + "0.0.bias"].double()
        a = sd[p + "0.1.weight"].double() / torch.sqrt(sd[p + "0.1.running_var"].double() + 1e-5)
        y = torch.relu((y - sd[p + "0.1.running_mean"].double()) * a + sd[p + "0.1.bias"].double())
        w2 = sd[p + "1.weight"]
        oc = w2.shape[0]
        y2 = O.sparse_conv(y.half().double(), nbr1, w2.half().double().reshape(oc, 1, 1, 1, Cc)) + sd[p + "1.bias"].double()
        ref_feats[name] = y2.float().numpy()
        err = np.abs(got_feats[name] - ref_feats[name]).max() / max(np.abs(ref_feats[name]).max(), 1e-6)
        assert err <= 1e-2, (name, err)

    # post-processing parity is checked on the DEVICE's own head outputs (so that a score next to the threshold cannot flip the lists)
    if iou_branch:
        ref = O.voxelhead_generate_predicted_boxes([got_feats], idx, B, [[0, 1, 2]], 500, 8, VOXEL, PC_RANGE, [-75.2, -75.2, -2, 75.2, 75.2, 4], 0.1,
                                                   [0.8, 0.55, 0.55], [2048, 1024, 1024], [200, 150, 150], iou_branch=True,
                                                   rectifier=[0.68, 0.71, 0.65], num_class=3)
    else:
        ref = O.voxelhead_generate_predicted_boxes([got_feats], idx, B, [[0, 1, 2]], 500, 8, VOXEL, PC_RANGE, [-75.2, -75.2, -2, 75.2, 75.2, 4], 0.1,
                                                   0.7, 1000, 83)
    got = out["final_box_dicts"]
    for b in range(B):
        n = ref[b]["pred_scores"].shape[0]
        assert n > 40 and got[b]["pred_scores"].shape[0] == n, (b, n, got[b]["pred_scores"].shape)
        assert np.array_equal(got[b]["pred_labels"].cpu().numpy(), ref[b]["pred_labels"])
        np.testing.assert_allclose(got[b]["pred_scores"].cpu().numpy(), ref[b]["pred_scores"], rtol=1e-5)
        np.testing.assert_allclose(got[b]["pred_boxes"].cpu().numpy(), ref[b]["pred_boxes"], rtol=2e-6, atol=2e-6)
