"""Golden vectors produced by the REFERENCE'S OWN Python sources (tests/golden/make_golden.py imports pcdet's
VoxelResBackBone8x / VoxelBackBone8x / MeanVFE / HeightCompression and quant/quant.py::QConvNd, quant/quantize.py::q_conv3d,
collect_stats, compute_amax from /root/reference, unmodified, on CPU; spconv and pytorch_quantization -- absent everywhere --
are replaced by oracle/ext_stubs.py).  They pin what the reference itself owns on this path: network topology, indice_key
sharing, the QConvNd permute / fake-quant / restore sequence, BN / ReLU / residual order, the surgery walk with its no_list,
static calibration, MeanVFE and HeightCompression.

CPU half (`-m "not gpu"`): the oracle's stand-alone restatement (backbone_specs / backbone_forward / QuantCfg) reproduces them.
GPU half (`-m gpu`): the product's drop-in module API, called exactly as the reference drivers call theirs, reproduces them
through libqlidar_b200.so.  Neither half reads /root/reference.

Tolerances: indices bit-exact; oracle features <= 1e-5 * max|golden| (same fp32 math, different association only);
product features <= 1e-2 * max|golden| for 16-bit / unquantised activations (north star), and the documented end-to-end
bound for stacked 8-bit activation quantisation (test_gpu_backbone.py: max 1e-1, mean 5e-3 of max|golden|)."""
import glob
import os

import numpy as np
import pytest
import torch

import qlidar_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MINI = dict(pc_range=[0.0, -10.0, -3.0, 17.6, 10.0, 1.0], voxel_size=[0.05, 0.05, 0.1], nfeat=4, max_pts=5, max_voxels=40000)
TAP_ROW_STRIDE, ENC_ROW_STRIDE = 8, 4

# label -> (w_bits, act_bits, cw, no_list, static)   (make_golden.py MODES)
MODES = {
    "fp32": (0, 0, False, (), False),
    "w8a16_cw": (8, 16, True, ("conv_input.0",), False),
    "w8a8_cw": (8, 8, True, ("conv_input.0",), False),
    "w8a8_pt": (8, 8, False, (), False),
    "w8a8_pt_static": (8, 8, False, ("conv_input.0",), True),
}
CASES = sorted((os.path.basename(p)[len("backbone_mini_"):-len(".npz")] for p in glob.glob(os.path.join(GOLD, "backbone_mini_*.npz"))))


def split_case(case):
    arch, label = case.split("_", 1)
    return arch, label


def load(case):
    return np.load(os.path.join(GOLD, f"backbone_mini_{case}.npz"))


def rel(got, ref):
    got = torch.as_tensor(np.asarray(got)).double() if not torch.is_tensor(got) else got.double().cpu()
    ref = torch.as_tensor(np.asarray(ref)).double()
    e = (got - ref).abs()
    m = max(ref.abs().max().item(), 1e-12)
    return e.max().item() / m, e.mean().item() / m


def test_golden_files_present():
    assert len(CASES) == 7, CASES


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_sources(case):
    arch, label = split_case(case)
    g = load(case)
    w_bits, act_bits, cw, no_list, static = MODES[label]
    pts = g["points"]
    voxels, coords3, num = O.voxelize_hard(pts, MINI["pc_range"], MINI["voxel_size"], MINI["max_pts"], MINI["max_voxels"])
    coords = np.concatenate([np.zeros((coords3.shape[0], 1), np.int32), coords3], axis=1)
    assert np.array_equal(coords, g["voxel_coords"]) and np.array_equal(num, g["voxel_num_points"])
    feats = O.mean_vfe(voxels, num)
    assert np.array_equal(feats, g["voxel_features"])                      # mean_vfe.py:25-29, same fp32 ops
    grid = O.grid_size_xyz(MINI["pc_range"], MINI["voxel_size"])
    prog = O.backbone_specs(arch, MINI["nfeat"])
    P = O.init_params(prog)
    qc = O.QuantCfg()
    if label != "fp32":
        amax = None
        if static:
            amax = {k[len("amax:"):-len(".act_quant")]: torch.from_numpy(g[k]) for k in g.files if k.startswith("amax:")}
            assert len(amax) == 20
        qc = O.QuantCfg(mode="ref", w_bits=w_bits, act_bits=act_bits, cw=cw, no_list=no_list, act_amax=amax)
        n_q = len(O.all_conv_specs(prog)) - len(no_list)
        assert len(g["quantized_modules"]) == n_q                           # q_conv3d walk + no_list (quantize.py:35-41)
    enc, taps = O.backbone_forward(prog, P, torch.from_numpy(feats), coords, O.sparse_shape_zyx(grid), 1, qc)
    assert np.array_equal(enc.coords, g["encoded_indices"])
    assert list(enc.spatial_shape) == list(g["encoded_shape"])
    assert rel(enc.features[::ENC_ROW_STRIDE], g["encoded_features_strided"])[0] <= 1e-5
    for k, t in taps.items():
        assert np.array_equal(t.coords, g[k + "_indices"])
        assert rel(t.features[::TAP_ROW_STRIDE], g[k + "_features_strided"])[0] <= 1e-5, k
    bev = O.height_compression(enc.features, enc.coords, enc.spatial_shape, 1)
    assert list(bev.shape) == list(g["spatial_features_shape"])
    assert rel(bev.double().abs().sum(dim=(0, 2, 3)), g["spatial_features_abs_sum_per_channel"])[0] <= 1e-5
    if label == "fp32":
        assert rel(bev, g["spatial_features_f16"].astype(np.float32))[0] <= 1e-3     # fp16 storage of the golden map


def test_static_calibration_amax_is_the_dynamic_amax_of_the_calibration_batch():
    """collect_stats / compute_amax (quantize.py:175-207) with a one-batch loader: the frozen per-tensor _amax of every
    act_quant equals max|x| of the un-quantised forward's input to that conv (calibration disables quantisation)."""
    g = load("VoxelResBackBone8x_w8a8_pt_static")
    prog = O.backbone_specs("VoxelResBackBone8x", 4)
    P = O.init_params(prog)
    rec = {}
    grid = O.grid_size_xyz(MINI["pc_range"], MINI["voxel_size"])
    O.backbone_forward(prog, P, torch.from_numpy(g["voxel_features"]), g["voxel_coords"], O.sparse_shape_zyx(grid), 1, O.QuantCfg(), rec)
    # the input of block conv1 / of a strided conv is the previous op's output; conv2's input is relu(bn1(conv1)) which the
    # record does not hold, so check the 11 convs whose input is a recorded ".post" tensor
    order = [s.name for s in O.all_conv_specs(prog)]
    checked = 0
    for i, name in enumerate(order):
        if i == 0 or name.endswith(".conv2"):
            continue
        prev = order[i - 1]
        if prev + ".post" not in rec:
            continue
        want = rec[prev + ".post"].abs().max().item()
        got = float(g[f"amax:{name}.act_quant"][0])
        assert abs(got - want) <= 1e-6 * max(want, 1.0), (name, got, want)
        checked += 1
    assert checked >= 10


# ------------------------------------------------------------------------------------------------------------ GPU half
def _reference_style_forward(arch, label, g):
    """The calls a reference driver makes (quant_centerpoint.py:80-116, quantize.py:175-207), on the product's classes."""
    import qlidar
    w_bits, act_bits, cw, no_list, static = MODES[label]
    grid = O.grid_size_xyz(MINI["pc_range"], MINI["voxel_size"])
    prog = O.backbone_specs(arch, MINI["nfeat"])
    bb = getattr(qlidar, arch)(qlidar.Cfg(), MINI["nfeat"], np.asarray(grid))
    missing, unexpected = bb.load_state_dict(O.init_params(prog), strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    bb = bb.cuda().eval()
    vfe = qlidar.MeanVFE(qlidar.Cfg(), MINI["nfeat"])
    hc = qlidar.HeightCompression(qlidar.Cfg(NUM_BEV_FEATURES=256))
    if label != "fp32":
        qlidar.q_conv3d(bb, {}, "", w_bits, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), list(no_list))
    voxels, coords3, num = O.voxelize_hard(g["points"], MINI["pc_range"], MINI["voxel_size"], MINI["max_pts"], MINI["max_voxels"])

    def batch():
        return {"voxels": torch.from_numpy(voxels).cuda(), "voxel_num_points": torch.from_numpy(num).float().cuda(),
                "voxel_coords": torch.from_numpy(g["voxel_coords"]).float().cuda(), "batch_size": 1}

    class Pipeline(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vfe, self.backbone_3d, self.map_to_bev = vfe, bb, hc

        def forward(self, bd):
            return self.map_to_bev(self.backbone_3d(self.vfe(bd)))

    model = Pipeline()
    if static:
        qlidar.collect_stats(model, [batch()], n_batches=0)
        qlidar.compute_amax(model, torch.device("cuda"))
    with torch.no_grad():
        return model(batch()), bb


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_product_modules_reproduce_reference_sources(case):
    import qlidar
    arch, label = split_case(case)
    g = load(case)
    bd, bb = _reference_style_forward(arch, label, g)
    a8 = MODES[label][1] == 8
    tol_max, tol_mean = (1e-1, 5e-3) if a8 else (1e-2, 1e-2)
    quantized = sorted(n for n, m in bb.named_modules() if isinstance(m, qlidar.QConvNd))
    if label != "fp32":
        assert quantized == sorted(g["quantized_modules"].tolist())
    assert rel(bd["voxel_features"], g["voxel_features"])[0] <= 1e-6
    enc = bd["encoded_spconv_tensor"]
    assert np.array_equal(enc.indices.cpu().numpy(), g["encoded_indices"])
    assert [int(v) for v in enc.spatial_shape] == list(g["encoded_shape"])
    mx, mn = rel(enc.features[::ENC_ROW_STRIDE], g["encoded_features_strided"])
    assert mx <= tol_max and mn <= tol_mean, (mx, mn)
    for k, t in bd["multi_scale_3d_features"].items():
        # the plugin call replays the engine, which returns every stage in ascending-key order (the reference keeps the voxeliser's
        # order for x_conv1): same set of sites, rows aligned by coordinate
        idx_e, idx_g = t.indices.cpu().numpy().astype(np.int64), g[k + "_indices"].astype(np.int64)
        key = lambda c: ((c[:, 0] * 4096 + c[:, 1]) * 4096 + c[:, 2]) * 4096 + c[:, 3]
        order_e = np.argsort(key(idx_e), kind="stable")
        assert np.array_equal(np.sort(key(idx_e)), np.sort(key(idx_g))), k
        row_of = order_e[np.searchsorted(key(idx_e)[order_e], key(idx_g))]          # engine row of every reference row
        assert np.array_equal(idx_e[row_of], idx_g), k
        mx, mn = rel(t.features[torch.from_numpy(row_of).to(t.features.device)][::TAP_ROW_STRIDE], g[k + "_features_strided"])
        assert mx <= tol_max and mn <= tol_mean, (k, mx, mn)
    sf = bd["spatial_features"]
    assert list(sf.shape) == list(g["spatial_features_shape"])
    mx, mn = rel(sf.double().abs().sum(dim=(0, 2, 3)), g["spatial_features_abs_sum_per_channel"])
    assert mx <= (5e-2 if a8 else 1e-2), mx
    if label == "fp32":
        assert rel(sf, g["spatial_features_f16"].astype(np.float32))[0] <= 1e-2
    if MODES[label][4]:
        for n, m in bb.named_modules():
            if n.endswith("act_quant"):
                want = float(g["amax:" + n][0])
                got = float(m.amax.reshape(-1)[0])
                assert abs(got - want) <= 2e-3 * want, (n, got, want)          # fp16 activations between layers
