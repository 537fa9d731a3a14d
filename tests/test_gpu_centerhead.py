"""CenterHead post-processing on the device (csrc/centerhead.cu) against
  * the golden vectors made by the reference's own centernet_utils.decode_bbox_from_heatmap (tests/golden/make_golden_centerhead.py),
  * the numpy / C oracle (oracle/qlidar_oracle.py centerhead_*, nms_rotated; oracle/qloracle_c.c qlo_rect_iou) at Waymo head size,
  * the REFERENCE's own CUDA kernels (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu compiled in place to oracle/_ref/libiou3d_ref.so):
    pairwise rotated IoU and the NMS bit mask, swept the way iou3d_nms.cpp:137-183 sweeps it.
Tolerances: indices, labels, counts and keep lists exact; fp32 box values 2e-6 relative (expf / atan2f / sigmoid are <= 1 ulp
implementations on both sides, not the same one); IoU 1e-5 absolute against the C oracle, 1e-6 against the reference kernel."""
import ctypes
import os

import numpy as np
import pytest
import torch

import qlidar_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "centerhead_decode.npz")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libiou3d_ref.so")

WAYMO = dict(pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel=[0.1, 0.1, 0.15], stride=8, limit=[-75.2, -75.2, -2, 75.2, 75.2, 4],
             K=500, score_thresh=0.1, nms_thresh=0.7, nms_pre=4096, nms_post=500)


def _maps(seed, B, C, H, W, vel=False, iou=False, hot=600, bias=-6.0):
    """head outputs with `hot` object-like peaks per frame (clustered so that the NMS has work to do)"""
    g = np.random.default_rng(seed)
    hm = (g.standard_normal((B, C, H, W)) * 0.7 + bias).astype(np.float32)
    for b in range(B):
        cy, cx = g.integers(2, H - 2, hot // 3), g.integers(2, W - 2, hot // 3)
        for k in range(hot):
            j = k % (hot // 3)
            y, x = np.clip(cy[j] + g.integers(-1, 2), 0, H - 1), np.clip(cx[j] + g.integers(-1, 2), 0, W - 1)
            hm[b, g.integers(0, C), y, x] = g.uniform(-1.5, 3.0)
    m = {"hm": hm, "center": g.random((B, 2, H, W), dtype=np.float32), "center_z": (g.standard_normal((B, 1, H, W)) * 0.5 + 1).astype(np.float32),
         "dim": (g.standard_normal((B, 3, H, W)) * 0.2 + np.log(np.array([4.5, 2.0, 1.6]))[None, :, None, None]).astype(np.float32),
         "rot": g.standard_normal((B, 2, H, W)).astype(np.float32)}
    if vel:
        m["vel"] = g.standard_normal((B, 2, H, W)).astype(np.float32)
    if iou:
        m["iou"] = (g.random((B, 1, H, W), dtype=np.float32) * 2 - 1).astype(np.float32)
    return m


def _cuda(m):
    return {k: torch.from_numpy(v).cuda() for k, v in m.items()}


def _check_decode(got, ref, rtol=2e-6):
    boxes, scores, labels, iou, count = got
    count = count.cpu().numpy()
    for b, d in enumerate(ref):
        n = d["pred_scores"].shape[0]
        assert count[b] == n, (b, count[b], n)
        assert np.array_equal(labels[b, :n].cpu().numpy(), np.asarray(d["pred_labels"]).astype(np.int32)), b
        np.testing.assert_allclose(scores[b, :n].cpu().numpy(), d["pred_scores"], rtol=rtol, atol=0)
        np.testing.assert_allclose(boxes[b, :n].cpu().numpy(), d["pred_boxes"], rtol=rtol, atol=2e-6)
        if "pred_iou" in d:
            np.testing.assert_allclose(iou[b, :n].cpu().numpy(), d["pred_iou"], rtol=rtol, atol=0)


@pytest.mark.parametrize("case", ["waymo_like", "nusc_like_vel", "iou_head_nothresh"])
def test_decode_reproduces_reference_golden(case):
    from qlidar import ops
    z = np.load(GOLD)
    B, C, H, W, K, wv, wi, st = z[f"{case}/cfg"]
    B, C, H, W, K = int(B), int(C), int(H), int(W), int(K)
    m = {k: torch.from_numpy(z[f"{case}/in/{k}"]).cuda() for k in ("hm", "center", "center_z", "dim", "rot")}
    vel = torch.from_numpy(z[f"{case}/in/vel"]).cuda() if wv else None
    iou = torch.from_numpy(z[f"{case}/in/iou"]).cuda() if wi else None
    got = ops.centerhead_decode(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"], vel, iou, K, float(z["stride"]), z["voxel"], z["pc_range"],
                                z["limit"], None if st < 0 else float(st))
    ref = []
    for b in range(B):
        d = {k: z[f"{case}/out/{b}/{k}"] for k in ("pred_boxes", "pred_scores", "pred_labels")}
        if wi:
            d["pred_iou"] = z[f"{case}/out/{b}/pred_iou"]
        ref.append(d)
    _check_decode(got, ref)


@pytest.mark.parametrize("score_thresh", [0.1, None])
def test_decode_waymo_head_size_matches_oracle(score_thresh):
    """4 frames x 3 classes x 468 x 468 (2.5 x the Waymo head's 188 x 188 per side; feature stride 3.2 keeps the cells inside the range),
    K = 500; score_thresh None sends all 657 072 cells of a frame through the radix select"""
    from qlidar import ops
    B, C, H, W = 4, 3, 468, 468
    m = _maps(7, B, C, H, W, hot=900)
    WAYMO = dict(globals()["WAYMO"], stride=3.2)
    ref = O.centerhead_decode(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"], None, None, WAYMO["K"], WAYMO["stride"], WAYMO["voxel"],
                              WAYMO["pc_range"], WAYMO["limit"], score_thresh, class_map=[2, 0, 1])
    c = _cuda(m)
    cmap = torch.tensor([2, 0, 1], dtype=torch.int32, device="cuda")
    got = ops.centerhead_decode(c["hm"], c["center"], c["center_z"], c["dim"], c["rot"], None, None, WAYMO["K"], WAYMO["stride"], WAYMO["voxel"],
                                WAYMO["pc_range"], WAYMO["limit"], score_thresh, class_map=cmap)
    assert min(d["pred_scores"].shape[0] for d in ref) > 300
    _check_decode(got, ref)


def test_decode_empty_and_saturated_frames():
    from qlidar import ops
    B, C, H, W = 2, 3, 40, 40
    m = _maps(3, B, C, H, W, hot=30)
    m["hm"][0] = -20.0                                        # frame 0: nothing above the threshold
    c = _cuda(m)
    got = ops.centerhead_decode(c["hm"], c["center"], c["center_z"], c["dim"], c["rot"], None, None, 50, 8, WAYMO["voxel"], WAYMO["pc_range"],
                                WAYMO["limit"], 0.1)
    ref = O.centerhead_decode(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"], None, None, 50, 8, WAYMO["voxel"], WAYMO["pc_range"],
                              WAYMO["limit"], 0.1)
    assert ref[0]["pred_scores"].shape[0] == 0
    _check_decode(got, ref)


# ---------------------------------------------------------------------------------------------- NMS
def _boxes(seed, n, spread=40.0):
    g = np.random.default_rng(seed)
    ctr = g.uniform(-spread, spread, (max(n // 4, 1), 2))
    b = np.zeros((n, 7), np.float32)
    j = g.integers(0, ctr.shape[0], n)
    b[:, :2] = ctr[j] + g.normal(0, 0.6, (n, 2))
    b[:, 2] = g.normal(0, 0.5, n)
    b[:, 3:6] = np.abs(g.normal([4.5, 2.0, 1.6], [0.5, 0.2, 0.2], (n, 3)))
    b[:, 6] = g.uniform(-np.pi, np.pi, n)
    s = np.sort(g.random(n).astype(np.float32))[::-1].copy()
    return b, s


def _ref_lib():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libiou3d_ref.so not built (make -C oracle ref in the build container)")
    lib = ctypes.CDLL(REF_SO)
    vp = ctypes.c_void_p
    lib._Z11nmsLauncherPKfPyif.argtypes = [vp, vp, ctypes.c_int, ctypes.c_float]
    lib._Z19boxesioubevLauncheriPKfiS0_Pf.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, vp]
    return lib


def _ref_sweep(mask: np.ndarray, n: int) -> np.ndarray:
    """the host loop of iou3d_nms.cpp:160-176 over the reference kernel's bit mask"""
    cb = mask.shape[1]
    remv = np.zeros(cb, np.uint64)
    keep = []
    for i in range(n):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            keep.append(i)
            remv[i // 64:] |= mask[i, i // 64:]
    return np.asarray(keep, np.int64)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 500, 1000])
def test_nms_matches_reference_kernel_and_oracle(n):
    from qlidar import ops
    b, s = _boxes(n, n)
    tb = torch.from_numpy(b).cuda()[None].contiguous()
    out = ops.nms_rotated(tb, torch.from_numpy(s).cuda()[None].contiguous(), None, None, 0.7, 4096, n, return_iou=True)
    torch.cuda.synchronize()
    nk = int(out["keep_count"][0])
    keep = out["keep"][0, :nk].cpu().numpy()
    iou = out["iou"][0].cpu().numpy()
    # (a) the C oracle
    m = O.rect_iou_matrix(b)
    assert np.abs(np.triu(iou, 1) - np.triu(m, 1)).max() <= 1e-5
    ok = O.nms_rotated(b, s, 0.7)
    assert np.array_equal(keep, ok), (nk, len(ok))
    # (b) the reference's own kernels
    lib = _ref_lib()
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device="cuda")
    ref_iou = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    flat = tb[0].contiguous()
    torch.cuda.synchronize()
    lib._Z11nmsLauncherPKfPyif(flat.data_ptr(), mask.data_ptr(), n, ctypes.c_float(0.7))
    lib._Z19boxesioubevLauncheriPKfiS0_Pf(n, flat.data_ptr(), n, flat.data_ptr(), ref_iou.data_ptr())
    torch.cuda.synchronize()
    assert np.abs(np.triu(iou, 1) - np.triu(ref_iou.cpu().numpy(), 1)).max() <= 1e-6
    rk = _ref_sweep(mask.cpu().numpy().view(np.uint64), n)
    assert np.array_equal(keep, rk), (nk, len(rk))
    assert 0 < nk <= n and (n < 100 or nk < n)             # the clustered boxes do suppress each other


def test_nms_batched_counts_pre_and_post_max():
    from qlidar import ops
    B, cap = 3, 300
    bs, ss, counts = [], [], [300, 0, 130]
    for f in range(B):
        b, s = _boxes(50 + f, cap)
        bs.append(b); ss.append(s)
    tb = torch.from_numpy(np.stack(bs)).cuda()
    ts = torch.from_numpy(np.stack(ss)).cuda()
    tl = torch.arange(cap, dtype=torch.int32, device="cuda")[None].repeat(B, 1).contiguous()
    tc = torch.tensor(counts, dtype=torch.int32, device="cuda")
    out = ops.nms_rotated(tb, ts, tl, tc, 0.5, 200, 40, label_offset=1)
    kc = out["keep_count"].cpu().numpy()
    for f in range(B):
        n = min(counts[f], 200)
        ok = O.nms_rotated(bs[f][:n], ss[f][:n], 0.5, 200, 40)
        assert kc[f] == len(ok)
        assert np.array_equal(out["keep"][f, :kc[f]].cpu().numpy(), ok)
        assert np.array_equal(out["boxes"][f, :kc[f]].cpu().numpy(), bs[f][ok])
        assert np.array_equal(out["scores"][f, :kc[f]].cpu().numpy(), ss[f][ok])
        assert np.array_equal(out["labels"][f, :kc[f]].cpu().numpy(), ok.astype(np.int32) + 1)
        assert (out["keep"][f, kc[f]:] == -1).all()


def test_generate_predicted_boxes_matches_oracle_waymo_config():
    """the reference-facing call: CenterHead.generate_predicted_boxes(batch_size, pred_dicts) (center_head.py:297-365), Waymo CenterPoint
    POST_PROCESSING (tools/cfgs/waymo_models/centerpoint.yaml:61-69), one head with 3 classes + a second head to cover the concat"""
    import qlidar
    B, H, W = 2, 188, 188
    names = ["Vehicle", "Pedestrian", "Cyclist", "Sign"]
    heads = [["Vehicle", "Pedestrian", "Cyclist"], ["Sign"]]
    post = {"SCORE_THRESH": 0.1, "POST_CENTER_LIMIT_RANGE": WAYMO["limit"], "MAX_OBJ_PER_SAMPLE": 500,
            "NMS_CONFIG": {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.7, "NMS_PRE_MAXSIZE": 4096, "NMS_POST_MAXSIZE": 500}}
    pp = qlidar.CenterHeadPostProcessor(names, heads, WAYMO["pc_range"], WAYMO["voxel"], 8, post)
    maps = [_maps(11, B, 3, H, W, hot=700), _maps(12, B, 1, H, W, hot=200)]
    got = pp.generate_predicted_boxes(B, [_cuda(m) for m in maps])
    ref = O.centerhead_generate_predicted_boxes(maps, [[0, 1, 2], [3]], 500, 8, WAYMO["voxel"], WAYMO["pc_range"], WAYMO["limit"], 0.1, 0.7, 4096, 500)
    for b in range(B):
        n = ref[b]["pred_scores"].shape[0]
        assert got[b]["pred_scores"].shape[0] == n and n > 100
        assert np.array_equal(got[b]["pred_labels"].cpu().numpy(), ref[b]["pred_labels"])
        np.testing.assert_allclose(got[b]["pred_scores"].cpu().numpy(), ref[b]["pred_scores"], rtol=2e-6)
        np.testing.assert_allclose(got[b]["pred_boxes"].cpu().numpy(), ref[b]["pred_boxes"], rtol=2e-6, atol=2e-6)
        assert set(np.unique(ref[b]["pred_labels"])) <= {1, 2, 3, 4} and 4 in ref[b]["pred_labels"]
