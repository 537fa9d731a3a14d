"""CPU: pin the oracle.  The reference ships no golden vectors for this path, so the oracle is
pinned by independent restatements: dense torch conv3d, float64 integer arithmetic, and structural properties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import qlidar_oracle as O
from helpers import random_coords

CASES = [(3, 1, 1, True), (3, 2, 1, False), (3, 2, (0, 1, 1), False), ((3, 1, 1), (2, 1, 1), 0, False), (5, 2, 2, False),
         ((1, 3, 3), 1, (0, 1, 1), False), ((1, 3, 3), 1, (0, 1, 1), True)]


@pytest.mark.parametrize("k,s,p,subm", CASES)
def test_sparse_conv_equals_dense_conv3d(k, s, p, subm):
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    B, D, H, W = 2, 7, 9, 8
    coords = random_coords(rng, B, D, H, W, 0.15)
    cin, cout = 5, 6
    x = torch.randn(len(coords), cin, dtype=torch.float64)
    dense = O.to_dense(x, coords, [D, H, W], B)
    k3 = O._triple(k)
    w = torch.randn((cout,) + k3 + (cin,), dtype=torch.float64)
    ref = F.conv3d(dense, w.permute(0, 4, 1, 2, 3), stride=O._triple(s), padding=O._triple(p))
    if subm:
        nbr, oc, osh = O.rulebook_subm(coords, [D, H, W], k), coords, [D, H, W]
    else:
        oc, osh, nbr = O.rulebook_strided(coords, [D, H, W], k, s, p)
        occ = torch.zeros((B, 1, D, H, W), dtype=torch.float64)
        occ[coords[:, 0], 0, coords[:, 1], coords[:, 2], coords[:, 3]] = 1
        act = F.conv3d(occ, torch.ones((1, 1) + k3, dtype=torch.float64), stride=O._triple(s), padding=O._triple(p)) > 0
        assert int(act.sum()) == len(oc)                      # out-set = dilate o subsample
        assert list(act.shape[2:]) == list(osh)
        assert len(np.unique(O._lin(oc, osh))) == len(oc)
    y = O.sparse_conv(x, nbr, w)
    refs = ref.permute(0, 2, 3, 4, 1)[oc[:, 0], oc[:, 1], oc[:, 2], oc[:, 3]]
    assert (y - refs).abs().max().item() < 1e-10


def test_rulebook_permutation_invariance():
    rng = np.random.default_rng(1)
    coords = random_coords(rng, 1, 6, 12, 12, 0.2)
    perm = rng.permutation(len(coords))
    a = O.pairs_in_coord_space(O.rulebook_subm(coords, [6, 12, 12], 3), coords, coords)
    b = O.pairs_in_coord_space(O.rulebook_subm(coords[perm], [6, 12, 12], 3), coords[perm], coords[perm])
    assert np.array_equal(a, b)
    oc1, _, n1 = O.rulebook_strided(coords, [6, 12, 12], 3, 2, 1)
    oc2, _, n2 = O.rulebook_strided(coords[perm], [6, 12, 12], 3, 2, 1)
    assert np.array_equal(O.pairs_in_coord_space(n1, coords, oc1), O.pairs_in_coord_space(n2, coords[perm], oc2))


def test_strided_sorted_key_order_matches_naive_loop():
    rng = np.random.default_rng(2)
    coords = random_coords(rng, 2, 5, 9, 9, 0.2)
    oc, osh, _ = O.rulebook_strided(coords, [5, 9, 9], 3, 2, 1)
    seen, order = {}, []
    for c in coords:                                          # pure-Python restatement of the numbering rule
        for kz in range(3):
            for ky in range(3):
                for kx in range(3):
                    n = (c[1] + 1 - kz, c[2] + 1 - ky, c[3] + 1 - kx)
                    if any(v % 2 for v in n) or any(v < 0 for v in n):
                        continue
                    o = (int(c[0]), n[0] // 2, n[1] // 2, n[2] // 2)
                    if o[1] >= osh[0] or o[2] >= osh[1] or o[3] >= osh[2]:
                        continue
                    if o not in seen:
                        seen[o] = len(order)
                        order.append(o)
    # the numbering rule: ascending linear key ((b*Do+z)*Ho+y)*Wo+x == lexicographic (b, z, y, x)
    assert np.array_equal(oc, np.asarray(sorted(order), dtype=np.int32))
    assert len(order) == len(set(order))


def test_backbone_shapes_match_reference_comments():
    # spconv_backbone.py:206,213,220,229 and SURVEY.md 8: KITTI [41,1600,1408]->[21,800,704]->[11,400,352]->[5,200,176]->[2,200,176]
    s = [41, 1600, 1408]
    s = O.conv_out_shape(s, 3, 2, 1); assert s == [21, 800, 704]
    s = O.conv_out_shape(s, 3, 2, 1); assert s == [11, 400, 352]
    s = O.conv_out_shape(s, 3, 2, (0, 1, 1)); assert s == [5, 200, 176]
    s = O.conv_out_shape(s, (3, 1, 1), (2, 1, 1), 0); assert s == [2, 200, 176]
    assert O.sparse_shape_zyx(O.grid_size_xyz(**{k: O.CONFIGS["waymo"][k] for k in ("pc_range", "voxel_size")})) == [41, 1504, 1504]


def test_fake_quant_semantics():
    # TensorQuantizer defaults: 8 bit, narrow range, symmetric, round half to even, amax<=2^-24 -> 0
    t = torch.tensor([[0.5, -1.0, 1.0, 0.0039370079 * 0.5]])
    fq = O.fake_quant(t, 8)
    assert fq[0, 1].item() == -1.0 and fq[0, 2].item() == 1.0
    q = O.quantize_codes(torch.tensor([0.5, 1.5, 2.5, -0.5, 127.0, -300.0]), torch.tensor(127.0), 8)
    assert q.tolist() == [0, 2, 2, 0, 127, -127]              # half-to-even and clamp to +-127
    assert O.fake_quant(torch.zeros(4, 3), 8).abs().sum().item() == 0
    x = torch.randn(50, 6)
    per_c = O.fake_quant(x, 8, axis=1)
    for c in range(6):
        assert torch.equal(per_c[:, c], O.fake_quant(x[:, c], 8))
    assert O.quant_bound(16) == 32767.0


def test_weight_matrix_roundtrip_and_per_oc_quant():
    w = torch.randn(8, 3, 3, 3, 4)
    assert torch.equal(O.weight_from_matrix(O.weight_matrix(w), w), w)
    q, amax = O.quantize_weight_per_oc(w)
    assert q.abs().amax(dim=(1, 2, 3, 4)).tolist() == [127] * 8
    assert torch.allclose(amax, w.abs().amax(dim=(1, 2, 3, 4)))


def test_w8a8_pt_int_path_is_the_reference_math():
    """Per-tensor activation scale factors out of the sum: int32 accumulate * scales == QConvNd fake-quant conv."""
    rng = np.random.default_rng(3)
    coords = random_coords(rng, 1, 6, 14, 14, 0.25)
    nbr = O.rulebook_subm(coords, [6, 14, 14], 3)
    x = torch.randn(len(coords), 16)
    w = torch.randn(16, 3, 3, 3, 16) * 0.1
    b = torch.randn(16) * 0.01
    acc, y, amax_x, amax_w = O.qconv_w8a8_pt(x, nbr, w, b)
    ref = O.qconv_reference_math(x, nbr, w, b, 8, 8, cw=False)
    assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # exactness of the integer path against float64
    qw, _ = O.quantize_weight_per_oc(w)
    qx = O.quantize_codes(x, amax_x, 8)
    ref64 = O.sparse_conv(qx.double(), nbr, qw.double())
    assert torch.equal(acc.double(), ref64)


def test_voxelize_hard_caps_and_order():
    pts = np.array([[0.01, 0.01, 0.01, 1], [5.0, 5.0, 0.5, 2], [0.02, 0.02, 0.02, 3], [0.03, 0.03, 0.03, 4],
                    [100.0, 0, 0, 5], [9.9, 0.0, 0.0, 6]], dtype=np.float32)
    v, c, n = O.voxelize_hard(pts, [0, 0, 0, 10, 10, 1], [1.0, 1.0, 1.0], max_pts=2, max_voxels=2)
    assert c.tolist() == [[0, 0, 0], [0, 5, 5]] and n.tolist() == [2, 1]
    assert v[0, :, 3].tolist() == [1.0, 3.0]                  # first two points in point order; 4th dropped by the cap
    m = O.mean_vfe(v, n)
    assert np.allclose(m[0], [0.015, 0.015, 0.015, 2.0])


def test_dense_and_height_compression_layout():
    rng = np.random.default_rng(4)
    coords = random_coords(rng, 2, 2, 5, 6, 0.4)
    f = torch.randn(len(coords), 3)
    hc = O.height_compression(f, coords, [2, 5, 6], 2)
    assert hc.shape == (2, 6, 5, 6)
    i = 7
    b, d, y, x = coords[i]
    assert torch.equal(hc[b, torch.arange(3) * 2 + d, y, x], f[i])    # channel index = c*D + d


def test_backbone_oracle_runs_small():
    prog = O.backbone_specs("VoxelResBackBone8x", 4)
    P = O.init_params(prog)
    assert len(O.all_conv_specs(prog)) == 21
    rng = np.random.default_rng(5)
    coords = random_coords(rng, 1, 41, 64, 64, 0.01)
    f = torch.randn(len(coords), 4)
    out, taps = O.backbone_forward(prog, P, f, coords, [41, 64, 64], 1)
    assert out.features.shape[1] == 128 and out.spatial_shape == [2, 8, 8]
    assert taps["x_conv4"].spatial_shape == [5, 8, 8]
    rec = {}
    out_q, _ = O.backbone_forward(prog, P, f, coords, [41, 64, 64], 1, O.QuantCfg(mode="w8a8_pt", no_list=("conv_input.0",)), rec)
    rel = (out_q.features - out.features).abs().max() / out.features.abs().max()
    assert rel < 0.2 and "conv1.0.conv1.acc" in rec


def test_im2col_conv_equals_per_offset_conv():
    rng = np.random.default_rng(6)
    coords = random_coords(rng, 2, 6, 14, 14, 0.2)
    oc, osh, nbr = O.rulebook_strided(coords, [6, 14, 14], 3, 2, 1)
    x = torch.randn(len(coords), 16, dtype=torch.float64)
    w = torch.randn(32, 3, 3, 3, 16, dtype=torch.float64)
    b = torch.randn(32, dtype=torch.float64)
    assert (O.sparse_conv(x, nbr, w, b) - O.sparse_conv_im2col(x, nbr, w, b, chunk=100)).abs().max().item() < 1e-10


def test_mirror_layers_agree_with_reference_math_teacher_forced():
    """The mirror is not a second opinion about the network: fed layer by layer with ITS OWN inputs, the reference's fake-quant
    math (quant/quant.py:36-58 restated, qconv_reference_math) + BatchNorm (+residual) + ReLU reproduces each mirror layer to 1e-3
    of the layer's max (fp16 storage of the output is 2^-11 relative; the rest is fp32 summation order)."""
    c = O.CONFIGS["kitti"]
    pts = O.synth_batch("kitti", 1, n_az=260)
    f_np, coords, _ = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    order = np.argsort(O._lin(coords, O.sparse_shape_zyx(grid)), kind="stable")
    feats, coords = torch.from_numpy(f_np[order]).contiguous(), np.ascontiguousarray(coords[order])
    prog = O.backbone_specs("VoxelResBackBone8x", 4)
    P = O.init_params(prog)
    rec, out, taps = O.mirror_backbone_w8a8_pt(prog, P, feats, coords, O.sparse_shape_zyx(grid), 1)
    x = O.SpT(None, coords.astype(np.int32), O.sparse_shape_zyx(grid), 1)
    prev_h = None
    checked = 0
    for op in prog:
        if op["op"] == "tap":
            continue
        specs = [(op["conv"], op["bn"], None)] if op["op"] == "conv_bn_relu" else [(op["conv1"], op["bn1"], None), (op["conv2"], op["bn2"], "res")]
        block_in = prev_h
        for spec, bnname, res in specs:
            r = rec[spec.name]
            if spec.name != "conv_input.0":
                xin = O.SpT(torch.from_numpy(prev_h.astype(np.float32)), x.coords, x.spatial_shape, 1, x.rulebooks)
                # same amax as the mirror (the producing layer's fp32 maximum; the fp16-stored maximum differs by <= 2^-11 relative,
                # enough to move codes that sit on a rounding boundary by one step = 0.8 % of amax)
                y = O.run_conv(xin, spec, P, O.QuantCfg(mode="ref", w_bits=8, act_bits=8, cw=False,
                                                         act_amax={spec.name: torch.tensor(float(r["amax_in"]))}))
                y = O.bn_relu(y, bnname, P, 1e-3, True, None if res is None else torch.from_numpy(block_in.astype(np.float32)))
                ref = y.features.numpy()
                err = np.abs(r["out"].astype(np.float32) - ref).max() / max(np.abs(ref).max(), 1e-12)
                assert err <= 1e-3, (spec.name, err)
                checked += 1
                x = O.SpT(None, y.coords, y.spatial_shape, 1, y.rulebooks)
            else:
                xin = O.SpT(feats, x.coords, x.spatial_shape, 1, x.rulebooks)
                y = O.bn_relu(O.run_conv(xin, spec, P, O.QuantCfg()), bnname, P, 1e-3)
                err = np.abs(r["out"].astype(np.float32) - y.features.numpy()).max() / np.abs(y.features.numpy()).max()
                assert err <= 1e-3, err
                x = O.SpT(None, y.coords, y.spatial_shape, 1, y.rulebooks)
            prev_h = r["out"]
    assert checked == 20


def test_sq_dense_wrappers_match_the_references_own_file():
    """oracle/sq_* restate quant/smoothquant.py; tests/golden/sq_dense.npz holds what that file itself computed (imported unmodified,
    built the way quantize.py:48-76 builds it).  1e-5 of max|y|: same operations, same order."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sq_dense.npz"))
    T = lambda k: torch.from_numpy(g[k])
    cases = {
        "conv2d_3x3_s1": lambda n: O.sq_conv2d(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5, 1, 1),
        "conv2d_3x3_s2": lambda n: O.sq_conv2d(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5, 2, 1),
        "conv2d_1x1_head": lambda n: O.sq_conv2d(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5, 1, 0),
        "conv1d_k3": lambda n: O.sq_conv1d(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5, 1, 1),
        "convT2d_k2_s2": lambda n: O.sq_convT2d(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5, 2),
        "linear": lambda n: O.sq_linear(T(n + ":x"), T(n + ":w"), T(n + ":b"), 0.5),
    }
    for name, fn in cases.items():
        y, ref = fn(name), T(name + ":y")
        assert y.shape == ref.shape, name
        assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), name


def test_sq_surgery_builds_wrappers_like_the_reference():
    """smoothquant() / smoothquant_layer() (quant/quantize.py:48-115): `__new__` + attribute copy from the fp32 layer (tuples
    collapse to their first element), quantisers attached, dotted-path no_list honoured; forward without a scaling_factor raises."""
    import qlidar
    net = torch.nn.Sequential()
    net.add_module("blocks", torch.nn.Sequential(torch.nn.Conv2d(16, 32, 3, stride=2, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(32, 32, 3, padding=1)))
    net.add_module("head", torch.nn.Conv2d(32, 3, 1))
    w0 = net.blocks[0].weight
    qlidar.smoothquant(net, {}, "", 0.5, 8, 8, torch.nn.Conv2d, qlidar.SQConv2d, ["head"])
    assert isinstance(net.blocks[0], qlidar.SQConv2d) and isinstance(net.blocks[2], qlidar.SQConv2d) and isinstance(net.head, torch.nn.Conv2d)
    q = net.blocks[0]
    assert q.weight is w0 and q.kernel_size == 3 and q.stride == 2 and q.padding == 1 and q.scaling_factor == 0.5
    assert q._weight_quantizer.num_bits == 8 and q._weight_quantizer.axis in ((0,), 0) and q._input_quantizer.axis is None
    q.scaling_factor = None
    with pytest.raises(ValueError):
        q(torch.zeros(1, 16, 4, 4))
    with pytest.raises(ValueError):
        qlidar.smoothquant_layer(torch.nn.Conv2d(4, 4, 1), qlidar.SQConv2d, None, 8, 8)
    # SQSubM2d: the constructor the reference intended (its own raises NameError), ValueError without a scaling factor
    sq = qlidar.SQSubM2d(16, 16, 3, 1, 1, device="cpu", scaling_factor=None)
    with pytest.raises(ValueError):
        sq(torch.zeros(1, 16, 4, 4))


def test_sqsubm2d_forward_matches_the_restated_file():
    """qlidar.SQSubM2d.forward(dense) -> (weight, x): the literal unfold -> scale -> quantise -> fold sequence of quant/SQSubM2d.py:22-91
    (whose class cannot be constructed as shipped) against the oracle's restatement; weight in the sparse conv's (oc, kh, kw, ic) layout."""
    import qlidar
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 8, 9, 7), generator=g)
    x[:, 3] *= 9
    sq = qlidar.SQSubM2d(8, 12, 3, 1, 1, input_quantizer=qlidar.TensorQuantizer(qlidar.QuantDescriptor(num_bits=8)),
                         weight_quantizer=qlidar.TensorQuantizer(qlidar.QuantDescriptor(num_bits=8, axis=(0))), device="cpu", scaling_factor=0.5)
    with torch.no_grad():
        sq.weight.copy_(torch.randn(sq.weight.shape, generator=g) * 0.3)
        w, xo = sq(x)
    w_ref, x_ref = O.sq_subm2d(x, sq.weight.detach(), 0.5)
    assert tuple(w.shape) == (12, 3, 3, 8) and tuple(xo.shape) == (2, 9, 7, 8)
    assert (w - w_ref).abs().max().item() <= 1e-6 * w_ref.abs().max().item()
    assert (xo - x_ref).abs().max().item() <= 1e-5 * x_ref.abs().max().item()


def test_sparse_conv_int_f64_is_exact():
    rng = np.random.default_rng(12)
    coords = O.synth_surface_sheet(40, seed=4, depth=10)
    nbr = O.rulebook_subm(coords, [10, 40, 40], 3)
    qx = torch.from_numpy(rng.integers(-127, 128, size=(coords.shape[0], 64)).astype(np.int8))
    qw = torch.from_numpy(rng.integers(-127, 128, size=(32, 3, 3, 3, 64)).astype(np.int8))
    assert torch.equal(O.sparse_conv_int_f64(qx, nbr, qw), O.sparse_conv_int(qx, nbr, qw))


# ---------------------------------------------------------------------------------------------- CenterHead post-processing
def test_centerhead_decode_oracle_reproduces_reference_golden():
    """the numpy restatement against the reference's own decode_bbox_from_heatmap (tests/golden/make_golden_centerhead.py)"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "centerhead_decode.npz"))
    for case in ("waymo_like", "nusc_like_vel", "iou_head_nothresh"):
        B, C, H, W, K, wv, wi, st = z[f"{case}/cfg"]
        m = {k: z[f"{case}/in/{k}"] for k in ("hm", "center", "center_z", "dim", "rot")}
        dec = O.centerhead_decode(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"], z[f"{case}/in/vel"] if wv else None,
                                  z[f"{case}/in/iou"] if wi else None, int(K), float(z["stride"]), z["voxel"], z["pc_range"], z["limit"],
                                  None if st < 0 else float(st))
        for b, d in enumerate(dec):
            assert np.array_equal(d["pred_labels"], z[f"{case}/out/{b}/pred_labels"].astype(np.int32)), (case, b)
            np.testing.assert_allclose(d["pred_scores"], z[f"{case}/out/{b}/pred_scores"], rtol=1e-6)
            np.testing.assert_allclose(d["pred_boxes"], z[f"{case}/out/{b}/pred_boxes"], rtol=2e-6, atol=2e-6)
            if wi:
                np.testing.assert_allclose(d["pred_iou"], z[f"{case}/out/{b}/pred_iou"], rtol=1e-6)


def test_rotated_iou_and_nms_oracle_known_answers():
    b = np.array([[0, 0, 0, 4, 2, 1, 0], [0, 0, 0, 4, 2, 1, 0], [2, 0, 0, 4, 2, 1, 0], [10, 10, 0, 1, 1, 1, 0.3], [0, 0, 0, 2, 4, 1, np.pi / 2],
                  [0, 0, 0, 2, 2, 1, np.pi / 4]], np.float32)
    m = O.rect_iou_matrix(b)
    assert abs(m[0, 1] - 1.0) < 1e-6 and abs(m[0, 2] - 1.0 / 3.0) < 1e-6 and m[0, 3] == 0.0
    assert abs(m[0, 4] - 1.0) < 1e-5                          # the same rectangle, described with swapped sides and a quarter turn
    # a 2 x 2 square turned by 45 degrees inside a 4 x 2 box: the octagon-ish overlap = 4 - 2 * (sqrt(2) - 1)^2 = 3.6569; union 8 + 4 - that
    assert abs(m[0, 5] - 3.65685 / (8 + 4 - 3.65685)) < 1e-4
    s = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4], np.float32)
    assert list(O.nms_rotated(b, s, 0.5)) == [0, 2, 3, 5]
    assert list(O.nms_rotated(b, s, 0.3)) == [0, 3]
    assert list(O.nms_rotated(b, s[::-1].copy(), 0.5)) == [5, 4, 3, 2]      # score order decides who survives
    assert list(O.nms_rotated(b, s, 0.5, pre_max=2)) == [0] and list(O.nms_rotated(b, s, 0.5, post_max=2)) == [0, 2]


# ---------------------------------------------------------------------------------------------- histogram calibration
def test_histogram_calibrator_matches_the_numpy_restatement_and_closed_forms():
    """qlidar.HistogramCalibrator (torch) against the oracle's independent numpy restatement on a multi-batch activation stream whose
    range grows, for the three compute_amax methods; plus closed-form cases: a uniform distribution's 50th percentile, and an outlier
    that max-calibration would follow while entropy / percentile clip it."""
    import qlidar
    g = torch.Generator().manual_seed(3)
    batches = [torch.relu(torch.randn(20000, generator=g)) * s for s in (1.0, 1.5, 0.8)]
    batches[1][7] = 40.0                                                        # one outlier, far above the bulk
    cal = qlidar.HistogramCalibrator(8, None, False)
    for b in batches:
        cal.collect(b)
    hist, edges = O.hist_collect([b.numpy() for b in batches])
    assert cal._calib_hist.numel() == len(hist)
    assert np.array_equal(cal._calib_hist.numpy().astype(np.int64), hist.astype(np.int64))
    np.testing.assert_allclose(cal._calib_bin_edges.numpy(), edges, rtol=1e-6)
    a_pct = float(cal.compute_amax("percentile", percentile=99.9))
    a_mse = float(cal.compute_amax("mse", stride=16))
    a_ent = float(cal.compute_amax("entropy", stride=8))
    assert abs(a_pct - float(O.hist_amax_percentile(hist, edges, 99.9))) <= 1e-6 * a_pct
    assert abs(a_mse - float(O.hist_amax_mse(hist, edges, 8, 16))) <= 1e-6 * a_mse
    assert abs(a_ent - float(O.hist_amax_entropy(hist, edges, 8, 8))) <= 1e-6 * a_ent
    for a in (a_pct, a_ent):
        assert 2.0 < a < 10.0                                                   # the bulk (sigma <= 1.5), not the outlier at 40
    assert 10.0 < a_mse < 40.0                   # squared error weighs the one clipped outlier (34^2) above the bulk's finer steps
    # closed form: uniform on [0, 1): the median edge
    u = qlidar.HistogramCalibrator(8, None, False, num_bins=1000)
    u.collect(torch.arange(100000, dtype=torch.float32) / 100000)
    assert abs(float(u.compute_amax("percentile", percentile=50.0)) - 0.5) <= 2e-3
    # through the quantiser, the way quant/quantize.py:175-207 drives it
    q = qlidar.TensorQuantizer(qlidar.QuantDescriptor(num_bits=8, calib_method="histogram"))
    q.enable_calib(); q.disable_quant()
    for b in batches:
        q(b)
    q.disable_calib(); q.enable_quant()
    q.load_calib_amax("percentile", percentile=99.9)
    assert abs(float(q.amax) - a_pct) <= 1e-6 * a_pct
    y = q(batches[0])
    assert float(y.abs().max()) <= a_pct * (1 + 1e-6) and torch.unique(y).numel() <= 255
