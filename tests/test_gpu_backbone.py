"""GPU parity of the drop-in module API and of the graph-captured engine against the CPU oracle's restatement of
VoxelResBackBone8x / VoxelBackBone8x / QConvNd (fp32 fake-quant "reference math").

Tolerance (north star): de-quantised features max-abs error <= 1e-2 * max|ref| per tensor; coordinates bit-exact."""
import numpy as np
import pytest
import torch

import qlidar_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-2
# Whole-network runs with 8-bit ACTIVATION quantisation are chaotic at the code level: an fp16-vs-fp32 difference of
# 1e-4 in a layer input flips a few round-half-even decisions, each flip moves that activation by amax/127 (0.8 %),
# and 20 stacked layers amplify it.  Layer-level parity (teacher forced, same inputs) is held to 2e-3 and the INT32
# accumulators to bit-exactness elsewhere in this file; every one of the 21 layers is
# additionally checked teacher-forced (test_every_backbone_layer_teacher_forced); end to end we bound max-abs by 1e-1 and
# mean-abs by 5e-3 of max|ref|.
TOL_A8_MAX, TOL_A8_MEAN = 1e-1, 5e-3


def check_feats(got, ref, a8):
    e = (got.double().cpu() - ref.double()).abs()
    m = max(ref.abs().max().item(), 1e-12)
    if a8:
        assert e.max().item() / m <= TOL_A8_MAX and e.mean().item() / m <= TOL_A8_MEAN, (e.max().item() / m, e.mean().item() / m)
    else:
        assert e.max().item() / m <= TOL, e.max().item() / m


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    return (got.double().cpu() - ref.double()).abs().max().item() / max(ref.abs().max().item(), 1e-12)


def make_frame(cfg="kitti", batch=1, **kw):
    c = O.CONFIGS[cfg]
    kw = kw or (dict(n_az=300) if cfg == "kitti" else dict(n_beams=24, n_az=500))
    pts = O.synth_batch(cfg, batch, **kw)
    feats, coords, _ = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    return pts, torch.from_numpy(feats), coords, grid, c


def build(arch, nfeat, grid, seed=4, **cfg):
    import qlidar
    prog = O.backbone_specs(arch, nfeat, cfg.get("CHANNELS"), cfg.get("SPCONV_KERNEL_SIZES"), cfg.get("OUT_CHANNEL"))
    P = O.init_params(prog, seed)
    bb = getattr(qlidar, arch)(cfg, nfeat, np.asarray(grid))
    missing, unexpected = bb.load_state_dict(P, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    return prog, P, bb.cuda().eval()


def batch_dict(feats, coords, batch):
    # load_data_to_gpu casts every array, coords included, to float32 (pcdet/models/__init__.py:36)
    return {"voxel_features": feats.cuda(), "voxel_coords": torch.from_numpy(coords).float().cuda(), "batch_size": batch}


@pytest.mark.parametrize("use_engine", [True, False])
@pytest.mark.parametrize("arch", ["VoxelResBackBone8x", "VoxelBackBone8x"])
def test_module_path_unquantized(arch, use_engine):
    """The reference's plugin call, backbone(batch_dict): by default it replays the compiled engine (the fast path), with
    use_engine = False it walks the module tree op by op.  Same contract either way; the engine returns x_conv1 in ascending-key
    order (coordinates travel with the rows), so the taps are compared aligned by key."""
    _, feats, coords, grid, c = make_frame("kitti")
    prog, P, bb = build(arch, 4, grid)
    bb.use_engine = use_engine
    ref, taps = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 1)
    with torch.no_grad():
        for _ in range(2):                                            # the second call replays the captured graph
            out = bb(batch_dict(feats, coords, 1))
    assert (getattr(bb, "_engine_state", None) is not None) == use_engine
    enc = out["encoded_spconv_tensor"]
    assert np.array_equal(enc.indices.cpu().numpy(), ref.coords)
    assert enc.spatial_shape == ref.spatial_shape
    assert enc.features.dtype == torch.float32                        # fp32 in -> fp32 out, like the reference
    assert rel_err(enc.features, ref.features) <= TOL
    for k, t in taps.items():
        got = out["multi_scale_3d_features"][k]
        o = np.argsort(O._lin(t.coords, t.spatial_shape), kind="stable") if use_engine else np.arange(t.coords.shape[0])
        assert np.array_equal(got.indices.cpu().numpy(), t.coords[o])
        assert rel_err(got.features, t.features[torch.from_numpy(o)]) <= TOL
    assert out["encoded_spconv_tensor_stride"] == 8


@pytest.mark.parametrize("w_bits,act_bits,cw", [(8, 8, False), (8, 8, True), (8, 16, False), (8, 16, True), (4, 8, False)])
def test_qconvnd_single_layer_matches_reference_math(w_bits, act_bits, cw):
    """QConvNd(module, w_bits, act_bits, cw) on one SubMConv3d/SparseConv3d vs quant/quant.py:36-58 restated in fp32."""
    import qlidar
    rng = np.random.default_rng(7)
    coords = O.synth_surface_sheet(50, seed=11, depth=12)
    x = torch.from_numpy(rng.normal(size=(coords.shape[0], 32)).astype(np.float32))
    x[:, 3] *= 8
    for subm in (True, False):
        conv = (qlidar.SubMConv3d(32, 64, 3, padding=1, bias=True, indice_key="k") if subm
                else qlidar.SparseConv3d(32, 64, 3, stride=2, padding=1, bias=True, indice_key="k")).cuda()
        q = qlidar.QConvNd(conv, w_bits, act_bits, cw)
        w, b = conv.weight.detach().cpu(), conv.bias.detach().cpu()
        if subm:
            nbr, oc = O.rulebook_subm(coords, [12, 50, 50], 3), coords
        else:
            oc, _, nbr = O.rulebook_strided(coords, [12, 50, 50], 3, 2, 1)
        ref = O.qconv_reference_math(x, nbr, w, b, w_bits, act_bits, cw)
        st = qlidar.SparseConvTensor(x.cuda(), torch.from_numpy(coords).cuda(), [12, 50, 50], 1)
        w_before = conv.weight.detach().clone()
        with torch.no_grad():
            y = q(st)
        assert torch.equal(conv.weight.detach(), w_before)            # the wrapper must not alter module.weight
        assert np.array_equal(y.indices.cpu().numpy(), oc)
        assert rel_err(y.features, ref) <= 2e-3, (subm, rel_err(y.features, ref))


def test_gqconv3d_per_row():
    import qlidar
    rng = np.random.default_rng(8)
    coords = O.synth_surface_sheet(40, seed=12, depth=12)
    x = torch.from_numpy(rng.normal(size=(coords.shape[0], 16)).astype(np.float32))
    conv = qlidar.SubMConv3d(16, 16, 3, padding=1, bias=False, indice_key="k").cuda()
    nbr = O.rulebook_subm(coords, [12, 40, 40], 3)
    w = conv.weight.detach().cpu()
    wq = O.weight_from_matrix(O.fake_quant(O.weight_matrix(w), 8, axis=0), w)
    ref = O.sparse_conv(O.fake_quant(x, 8, axis=0), nbr, wq)          # quant_conv3d.py:112-131: amax per voxel row
    with torch.no_grad():
        y = qlidar.GQConv3d(conv)(qlidar.SparseConvTensor(x.cuda(), torch.from_numpy(coords).cuda(), [12, 40, 40], 1))
    assert rel_err(y.features, ref) <= 2e-3


@pytest.mark.parametrize("use_engine", [True, False])
@pytest.mark.parametrize("w_bits,act_bits,cw,mode", [(8, 8, False, "ref"), (8, 8, True, "ref"), (8, 16, True, "ref")])
def test_q_conv3d_surgery_whole_backbone(w_bits, act_bits, cw, mode, use_engine):
    """quant_centerpoint.quant(): q_conv3d over the backbone, conv_input.0 in the no_list when sq (=cw) is on.  (8, 8, False) wraps
    conv_input.0 too: 8-bit activations on raw point features are served by the eager tree, the plugin call falls back to it."""
    import qlidar
    _, feats, coords, grid, c = make_frame("kitti")
    prog, P, bb = build("VoxelResBackBone8x", 4, grid)
    bb.use_engine = use_engine
    no_list = ["conv_input.0"] if cw else []
    qlidar.q_conv3d(bb, {}, "", w_bits, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), no_list)
    n_wrapped = sum(isinstance(m, qlidar.QConvNd) for m in bb.modules())
    assert n_wrapped == (20 if cw else 21)
    ref, _ = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 1,
                                O.QuantCfg(mode=mode, w_bits=w_bits, act_bits=act_bits, cw=cw, no_list=tuple(no_list)))
    with torch.no_grad():
        out = bb(batch_dict(feats, coords, 1))
    enc = out["encoded_spconv_tensor"]
    assert np.array_equal(enc.indices.cpu().numpy(), ref.coords)
    check_feats(enc.features, ref.features, act_bits <= 8)


@pytest.mark.parametrize("mode", ["eager", "engine", "engine_fused_bev"])
def test_height_compression_module(mode):
    """HeightCompression after the backbone plugin call: dense() of the encoded tensor, or -- attach() -- the map the engine wrote
    inside its graph."""
    import qlidar
    _, feats, coords, grid, c = make_frame("kitti")
    prog, P, bb = build("VoxelResBackBone8x", 4, grid)
    bb.use_engine = mode != "eager"
    hc = qlidar.HeightCompression(qlidar.Cfg(NUM_BEV_FEATURES=256))
    if mode == "engine_fused_bev":
        hc.attach(bb)
    ref, _ = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 1)
    with torch.no_grad():
        bd = bb(batch_dict(feats, coords, 1))
        bd = hc(bd)
    sf = bd["spatial_features"]
    assert sf.dtype == torch.float32
    ref_bev = O.height_compression(ref.features, ref.coords, ref.spatial_shape, 1)
    assert tuple(sf.shape) == tuple(ref_bev.shape) == (1, 256, 200, 176)
    assert rel_err(sf, ref_bev) <= TOL
    assert torch.equal(sf.cpu() != 0, ref_bev != 0) or ((sf.cpu() != 0) ^ (ref_bev != 0)).float().mean().item() < 1e-3


@pytest.mark.parametrize("quant", [None, (8, 16, True), (8, 8, False), (8, 8, True)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_from_points_matches_oracle(quant, use_graph):
    """points -> voxelize+meanVFE -> backbone -> BEV in one graph replay vs the oracle end to end (Waymo-shaped, batch 2)."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    qc = O.QuantCfg()
    if quant is not None:
        w_bits, act_bits, cw = quant
        no_list = ["conv_input.0"]
        qlidar.q_conv3d(bb, {}, "", w_bits, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), no_list)
        qc = O.QuantCfg(mode="ref", w_bits=w_bits, act_bits=act_bits, cw=cw, no_list=tuple(no_list))
    rec = {}
    ref, taps = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 2, qc, rec)
    eng = qlidar.BackboneEngine(bb, 2, coords.shape[0] + 1000, max_points=pts.shape[0] + 500, pc_range=c["pc_range"],
                                voxel_size=c["voxel_size"], max_pts_per_voxel=c["max_pts"], use_graph=use_graph, stage_cap_ratio=4.0)
    for _ in range(2):                                                # second call replays the captured graph
        out = eng.forward_points(torch.from_numpy(pts))
    torch.cuda.synchronize()
    counts = eng.counts()
    assert not eng.overflowed()
    assert counts[0] == coords.shape[0] and counts[-1] == ref.coords.shape[0]
    n = counts[-1]
    assert np.array_equal(out["encoded_coords"][:n].cpu().numpy(), ref.coords)
    a8 = quant is not None and quant[1] <= 8
    check_feats(out["encoded_features"][:n], ref.features, a8)
    for name, (f, st) in out["taps"].items():
        m = taps[name].coords.shape[0]
        # the engine returns every stage in ascending-key order (stage 1 is renumbered after the voxeliser); the oracle's
        # stage 1 is in first-touch order like the reference's: compare the rows as sets, aligned by key
        o = np.argsort(O._lin(taps[name].coords, taps[name].spatial_shape), kind="stable")
        assert np.array_equal(st.coords[:m].cpu().numpy(), taps[name].coords[o])
        check_feats(f[:m], taps[name].features[torch.from_numpy(o)], a8)
    ref_bev = O.height_compression(ref.features, ref.coords, ref.spatial_shape, 2)
    check_feats(out["spatial_features"], ref_bev, a8)


def test_engine_int32_accumulators_match_oracle_on_first_quantized_layer():
    """W8A8-pt: the i8 path's quantised codes and INT32 accumulators are bit-exact when fed identical fp16 features."""
    import qlidar
    from qlidar import ops
    rng = np.random.default_rng(9)
    coords = O.synth_surface_sheet(60, seed=13, depth=12)
    x = torch.from_numpy(rng.normal(size=(coords.shape[0], 64)).astype(np.float32)).half()
    w = torch.from_numpy(rng.normal(size=(64, 3, 3, 3, 64)).astype(np.float32)) * 0.05
    nbr = O.rulebook_subm(coords, [12, 60, 60], 3)
    acc_ref, y_ref, amax_x, amax_w = O.qconv_w8a8_pt(x.float(), nbr, w, None)
    conv = qlidar.SubMConv3d(64, 64, 3, padding=1, bias=False, indice_key="k").cuda()
    conv.weight.data.copy_(w)
    q = qlidar.QConvNd(conv, 8, 8, False)
    packed, ic_p, oc_p, w_scale, shift = q._prepared(torch.device("cuda"), "i8")
    st = qlidar.SparseConvTensor(x.cuda(), torch.from_numpy(coords).cuda(), [12, 60, 60], 1)
    rb = conv.get_rulebook(st)
    codes, act_scale = ops.quantize_rows(x.cuda(), ops.absmax_cols(x.cuda()), ops.QL_Q_CODES_PER_TENSOR)
    acc = torch.zeros((coords.shape[0], 64), dtype=torch.int32, device="cuda")
    ops.spconv_mma(codes, rb.nbr, rb.n_out, None, 64, packed, w_scale, shift, out=acc, kmask=rb.kmask)
    assert torch.equal(acc.cpu(), acc_ref)
    with torch.no_grad():
        y = q(st)
    assert rel_err(y.features, y_ref) <= 2e-3


@pytest.mark.parametrize("w_bits,act_bits,cw", [(8, 8, False), (8, 8, True), (8, 16, True)])
def test_every_backbone_layer_teacher_forced(w_bits, act_bits, cw):
    """Each of the 21 VoxelResBackBone8x convs, fed the ORACLE's input for that layer, must reproduce the oracle's
    QConvNd output (quant/quant.py:36-58 math) to 2e-3 of max|ref| with bit-exact output coordinates."""
    import qlidar
    _, feats, coords, grid, c = make_frame("kitti")
    prog = O.backbone_specs("VoxelResBackBone8x", 4)
    P = O.init_params(prog)
    rec = {}
    O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 1,
                       O.QuantCfg(mode="ref", w_bits=w_bits, act_bits=act_bits, cw=cw), rec)
    worst = 0.0
    for spec in O.all_conv_specs(prog):
        x, xc, shape, oc = rec[spec.name + ".in"]
        cls = qlidar.SubMConv3d if spec.subm else qlidar.SparseConv3d
        conv = cls(spec.cin, spec.cout, spec.ksize, stride=spec.stride, padding=spec.pad, bias=spec.bias, indice_key="k").cuda()
        conv.weight.data.copy_(P[spec.name + ".weight"])
        if spec.bias:
            conv.bias.data.copy_(P[spec.name + ".bias"])
        st = qlidar.SparseConvTensor(x.cuda().contiguous(), torch.from_numpy(xc).cuda(), shape, 1)
        with torch.no_grad():
            y = qlidar.QConvNd(conv, w_bits, act_bits, cw)(st)
        assert np.array_equal(y.indices.cpu().numpy(), oc), spec.name
        e = rel_err(y.features, rec[spec.name])
        worst = max(worst, e)
        assert e <= 2e-3, (spec.name, e)


# ------------------------------------------------------------------------------------------------ SmoothQuant (W8A8-sq)
@pytest.mark.parametrize("alpha", [None, 0.5, 0.8])
def test_sqconv3d_matches_oracle(alpha):
    """SQConv3d: alpha=None is the scalar SmoothQuant of quant/collect_act_conv3d.py:68-110 (== W8A8-pt, the scalar cancels);
    alpha=a the per-input-channel SmoothQuant of SURVEY.md 8a-Q (intent of quant/quant_conv3d.py:141-236 with the
    quant/smoothquant.py:72-79 formula).  Activations carry x20 outliers in two channels."""
    import qlidar
    torch.manual_seed(21)                                              # the convs' random weights
    rng = np.random.default_rng(21)
    coords = O.synth_surface_sheet(50, seed=15, depth=12)
    x = torch.from_numpy(rng.normal(size=(coords.shape[0], 64)).astype(np.float32))
    x[::100, 5] *= 20
    x[::100, 40] *= 20
    x = x.half().float()                                               # the features the kernel sees
    for subm in (True, False):
        conv = (qlidar.SubMConv3d(64, 64, 3, padding=1, bias=True, indice_key="k") if subm
                else qlidar.SparseConv3d(64, 32, 3, stride=2, padding=1, bias=False, indice_key="k")).cuda()
        w = conv.weight.detach().cpu()
        b = None if conv.bias is None else conv.bias.detach().cpu()
        if subm:
            nbr, oc = O.rulebook_subm(coords, [12, 50, 50], 3), coords
        else:
            oc, _, nbr = O.rulebook_strided(coords, [12, 50, 50], 3, 2, 1)
        if alpha is None:
            _, ref, _, _ = O.qconv_w8a8_pt(x, nbr, w, b)
        else:
            _, ref, _, _ = O.qconv_w8a8_sq(x, nbr, w, b, alpha)
        q = qlidar.SQConv3d(conv) if alpha is None else qlidar.SQConv3d(conv, scaling_factor=alpha)
        st = qlidar.SparseConvTensor(x.cuda().half(), torch.from_numpy(coords).cuda(), [12, 50, 50], 1)
        with torch.no_grad():
            y = q(st)
        assert np.array_equal(y.indices.cpu().numpy(), oc)
        # The dynamic path derives the smoothing scale ON THE DEVICE (ql_sq_prepare_weights: powf), the oracle on the host (pow):
        # the two can differ in the last bit of s[ic], which moves an int8 code that sits on a round-half-even boundary by one step
        # -- ~1.4e-3 of max|ref| per flipped code.  Otherwise the codes and INT32 accumulators are identical and only the fp16 store
        # differs (< 5e-4).  3e-3 leaves room for two such codes; an order of magnitude inside the north star's 1e-2.
        assert rel_err(y.features, ref) <= 3e-3, (alpha, subm, rel_err(y.features, ref))
        if alpha is not None:
            # static SmoothQuant: calibrated per-channel maxima, weights prepared once
            amax = x.abs().amax(dim=0)
            qs = qlidar.SQConv3d(conv, scaling_factor=alpha, act_amax=amax)
            with torch.no_grad():
                y1, y2 = qs(st), qs(st)
            assert torch.equal(y1.features, y2.features)
            assert rel_err(y1.features, ref) <= 3e-3                      # host-prepared (static) vs the oracle: same bound
            assert rel_err(y1.features, y.features.float().cpu()) <= 3e-3  # ... and vs the device-prepared (dynamic) path


def test_smoothquant_helps_with_channel_outliers():
    """The point of SmoothQuant: with per-channel outliers the W8A8-sq error against the un-quantised conv is lower than W8A8-pt's."""
    import qlidar
    rng = np.random.default_rng(22)
    coords = O.synth_surface_sheet(50, seed=16, depth=12)
    x = torch.from_numpy(rng.normal(size=(coords.shape[0], 64)).astype(np.float32))
    x[:, 5] *= 30
    x[:, 40] *= 30
    conv = qlidar.SubMConv3d(64, 64, 3, padding=1, bias=False, indice_key="k").cuda()
    ref = O.sparse_conv(x, O.rulebook_subm(coords, [12, 50, 50], 3), conv.weight.detach().cpu())
    st = qlidar.SparseConvTensor(x.cuda().half(), torch.from_numpy(coords).cuda(), [12, 50, 50], 1)
    with torch.no_grad():
        e_pt = rel_err(qlidar.QConvNd(conv, 8, 8, False)(st).features, ref)
        e_sq = rel_err(qlidar.SQConv3d(conv, scaling_factor=0.5)(st).features, ref)
    assert e_sq < 0.6 * e_pt, (e_sq, e_pt)


def test_second_backbone_w8a8_smoothquant_surgery():
    """BASELINE config 3: SECOND's sparse middle extractor (VoxelBackBone8x, kitti_models/second.yaml:13-14) W8A8 with
    SmoothQuant on KITTI-shaped frames; sq_conv3d swaps every sparse conv but the stem (quant_second.py no_list)."""
    import qlidar
    _, feats, coords, grid, c = make_frame("kitti", batch=2, n_az=260)
    prog, P, bb = build("VoxelBackBone8x", 4, grid)
    no_list = ["conv_input.0"]
    qlidar.sq_conv3d(bb, {}, "", 0.5, 8, 8, (qlidar.SubMConv3d, qlidar.SparseConv3d), no_list)
    assert sum(isinstance(m, qlidar.SQConv3d) for m in bb.modules()) == len(O.all_conv_specs(prog)) - 1
    ref, taps = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 2,
                                   O.QuantCfg(mode="w8a8_sq", alpha=0.5, no_list=tuple(no_list)))
    with torch.no_grad():
        for _ in range(2):
            out = bb(batch_dict(feats, coords, 2))
    assert bb._engine_state is not None and bb._engine_state["eng"].layers_sq == len(O.all_conv_specs(prog)) - 1    # the plugin call runs the engine
    enc = out["encoded_spconv_tensor"]
    assert np.array_equal(enc.indices.cpu().numpy(), ref.coords)
    check_feats(enc.features, ref.features, True)
    for k, t in out["multi_scale_3d_features"].items():
        o = np.argsort(O._lin(taps[k].coords, taps[k].spatial_shape), kind="stable")
        assert np.array_equal(t.indices.cpu().numpy(), taps[k].coords[o])
    # the eager tree (device-side weight preparation per call, no host round trip) agrees with the engine
    bb.use_engine = False
    with torch.no_grad():
        out2 = bb(batch_dict(feats, coords, 2))
    check_feats(out2["encoded_spconv_tensor"].features, enc.features.cpu(), True)


# ------------------------------------------------------------------------------------------------ VoxelNeXt (config 4)
def _voxelnext_case(cfg, dataset, pc_range, batch, quant, **synth_kw):
    import qlidar
    c = O.CONFIGS[dataset]
    pts = O.synth_batch(dataset, batch, **synth_kw)
    if pc_range is None:
        pc_range = c["pc_range"]
    else:
        m = (pts[:, 1] >= pc_range[0]) & (pts[:, 1] < pc_range[3]) & (pts[:, 2] >= pc_range[1]) & (pts[:, 2] < pc_range[4])
        pts = pts[m]
    feats, coords, _ = O.voxelize_mean_batch(pts, pc_range, c["voxel_size"], c["max_pts"], c["max_voxels"])
    grid = O.grid_size_xyz(pc_range, c["voxel_size"])
    nfeat = c["nfeat"]
    prog = O.backbone_specs("VoxelResBackBone8xVoxelNeXt", nfeat, cfg["CHANNELS"], cfg["SPCONV_KERNEL_SIZES"], cfg["OUT_CHANNEL"])
    P = O.init_params(prog)
    bb = qlidar.VoxelResBackBone8xVoxelNeXt(cfg, nfeat, np.asarray(grid))
    sd = {k: (v.reshape(v.shape[0], *v.shape[2:]) if k in ("conv_out.0.weight", "shared_conv.0.weight") else v) for k, v in P.items()}
    missing, unexpected = bb.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    bb = bb.cuda().eval()
    qc = O.QuantCfg()
    if quant is not None:
        w_bits, act_bits, cw = quant
        qlidar.q_conv3d(bb, {}, "", w_bits, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), ["conv_input.0"])
        no_list = ("conv_input.0", "conv_out.0", "shared_conv.0")                      # the 2-D tail is not a 3-D conv: not swapped
        qc = O.QuantCfg(mode="ref", w_bits=w_bits, act_bits=act_bits, cw=cw, no_list=no_list)
        assert sum(isinstance(mm, qlidar.QConvNd) for mm in bb.modules()) == 5 * 5 + 4
    ref, taps = O.backbone_forward(prog, P, torch.from_numpy(feats), coords, O.sparse_shape_zyx(grid), batch, qc)
    return bb, feats, coords, ref, taps


def _check_voxelnext(bb, feats, coords, batch, ref, taps, a8, use_engine):
    bb.use_engine = use_engine
    with torch.no_grad():
        for _ in range(2 if use_engine else 1):                                         # the second call replays the captured graph
            out = bb(batch_dict(torch.from_numpy(feats), coords, batch))
    assert (getattr(bb, "_engine_state", None) is not None) == use_engine
    enc = out["encoded_spconv_tensor"]
    got_idx = enc.indices.cpu().numpy()
    assert np.array_equal(got_idx, ref.coords[:, [0, 2, 3]])                            # 2-D indices [b, y, x], same (sorted) order
    assert list(enc.spatial_shape) == list(ref.spatial_shape[1:])
    check_feats(enc.features, ref.features, a8)
    assert out["encoded_spconv_tensor_stride"] == 8
    for k in ("x_conv1", "x_conv2", "x_conv3"):
        t = taps[k]
        o = np.argsort(O._lin(t.coords, t.spatial_shape), kind="stable") if use_engine else np.arange(t.coords.shape[0])
        assert np.array_equal(out["multi_scale_3d_features"][k].indices.cpu().numpy(), t.coords[o])
        check_feats(out["multi_scale_3d_features"][k].features, t.features[torch.from_numpy(o)], a8)


@pytest.mark.parametrize("use_engine", [True, False])
@pytest.mark.parametrize("quant", [None, (8, 8, False)])
def test_voxelnext_backbone_module_path(quant, use_engine):
    """VoxelResBackBone8xVoxelNeXt (spconv_backbone_voxelnext.py:70-225) with the Waymo-large kernel sizes [5,5,3,3]: six 3-D
    stages, stage-5/6 indices scaled onto the stage-4 grid, 2-D merge of duplicate (b,y,x) rows, SparseConv2d + SubMConv2d tail --
    through the plugin call, engine (graph replay) and eager tree."""
    cfg = dict(SPCONV_KERNEL_SIZES=[5, 5, 3, 3], CHANNELS=[16, 32, 64, 128, 128], OUT_CHANNEL=128)
    pc_range = [0.0, -12.8, -3.0, 25.6, 12.8, 1.0]                     # 512 x 512 x 40 crop: keeps all six stage shapes non-trivial
    bb, feats, coords, ref, taps = _voxelnext_case(cfg, "kitti", pc_range, 2, quant, n_az=500)
    _check_voxelnext(bb, feats, coords, 2, ref, taps, quant is not None, use_engine)


@pytest.mark.parametrize("quant", [(8, 16, True), (8, 8, False)])
def test_voxelnext_waymo_large_through_the_engine(quant):
    """BASELINE config 4 at its real shape: waymo_models/voxelnext_ioubranch_large.yaml:13-16 -- CHANNELS [32, 64, 128, 256, 256],
    SPCONV_KERNEL_SIZES [5, 5, 3, 3], OUT_CHANNEL 256 -- on the full Waymo grid [41, 1504, 1504] (a thinned frame so that the CPU
    oracle finishes in seconds): 5^3 strided rulebooks (K = 125), 512-byte rows (C = 256), the three-stage 2-D merge and the 2-D
    tail, all inside one graph replay behind the plugin call."""
    cfg = dict(SPCONV_KERNEL_SIZES=[5, 5, 3, 3], CHANNELS=[32, 64, 128, 256, 256], OUT_CHANNEL=256)
    bb, feats, coords, ref, taps = _voxelnext_case(cfg, "waymo", None, 1, quant, n_beams=20, n_az=500)
    assert feats.shape[0] > 5000 and ref.coords.shape[0] > 500
    _check_voxelnext(bb, feats, coords, 1, ref, taps, quant[1] <= 8, True)


def test_engine_static_calibration_fuses_requantisation():
    """static=True flow of the reference drivers (collect_stats -> compute_amax, quant/quantize.py:175-207) through the engine:
    with frozen per-tensor amax every W8A8 layer but the first receives its int8 codes from the previous layer's epilogue."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    no_list = ["conv_input.0"]
    qlidar.q_conv3d(bb, {}, "", 8, 8, False, (qlidar.SubMConv3d, qlidar.SparseConv3d), no_list)

    class Pipe(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone_3d = bb

        def forward(self, bd):
            return self.backbone_3d(bd)

    qlidar.collect_stats(Pipe(), [batch_dict(feats, coords, 2)], n_batches=0)
    qlidar.compute_amax(bb, torch.device("cuda"))
    amax = {n[:-len(".act_quant")]: m.amax.detach().float().cpu().reshape(-1) for n, m in bb.named_modules() if n.endswith("act_quant")}
    assert len(amax) == 20 and all(v.numel() == 1 and v.item() > 0 for v in amax.values())
    ref, taps = O.backbone_forward(prog, P, feats, coords, O.sparse_shape_zyx(grid), 2,
                                   O.QuantCfg(mode="ref", w_bits=8, act_bits=8, cw=False, no_list=tuple(no_list), act_amax=amax))
    eng = qlidar.BackboneEngine(bb, 2, coords.shape[0] + 1000, max_points=pts.shape[0] + 500, pc_range=c["pc_range"],
                                voxel_size=c["voxel_size"], max_pts_per_voxel=c["max_pts"], use_graph=True, stage_cap_ratio=4.0)
    assert sum(L.fused_q for L in eng.layers) == 19                        # all but the first quantised layer (its input is the stem)
    for _ in range(2):
        out = eng.forward_points(torch.from_numpy(pts))
    torch.cuda.synchronize()
    counts = eng.counts()
    assert not eng.overflowed() and counts[-1] == ref.coords.shape[0]
    n = counts[-1]
    assert np.array_equal(out["encoded_coords"][:n].cpu().numpy(), ref.coords)
    check_feats(out["encoded_features"][:n], ref.features, True)
    # and the eager module path with the same frozen amax agrees with the engine to the same bound
    with torch.no_grad():
        mod = bb(batch_dict(feats, coords, 2))["encoded_spconv_tensor"]
    check_feats(out["encoded_features"][:n], mod.features.float().cpu(), True)


def test_engine_grouped_rulebooks_give_identical_outputs():
    """group_rows=True (submanifold rulebooks of the ranked stages binned by line key) changes the tiling only: every
    output of the engine is bit-identical to the ungrouped run."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    qlidar.q_conv3d(bb, {}, "", 8, 16, True, (qlidar.SubMConv3d, qlidar.SparseConv3d), ["conv_input.0"])
    outs = []
    for grouped in (False, True):
        eng = qlidar.BackboneEngine(bb, 2, coords.shape[0] + 1000, max_points=pts.shape[0] + 500, pc_range=c["pc_range"],
                                    voxel_size=c["voxel_size"], max_pts_per_voxel=c["max_pts"], stage_cap_ratio=4.0, group_rows=grouped)
        for _ in range(2):
            out = eng.forward_points(torch.from_numpy(pts))
        torch.cuda.synchronize()
        assert any(p is not None for p in eng.row_perms.values()) == grouped
        outs.append({"enc": out["encoded_features"].clone(), "bev": out["spatial_features"].clone(),
                     **{k: f.clone() for k, (f, _) in out["taps"].items()}})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_engine_empty_ragged_and_single_voxel_inputs():
    """Edge cases of the batch contract: no in-range point at all (every stage empty, BEV all zero), a batch whose second
    frame is empty (ragged), and a single voxel -- each against the oracle, through the same captured graph."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    eng = qlidar.BackboneEngine(bb, 2, coords.shape[0] + 1000, max_points=pts.shape[0] + 500, pc_range=c["pc_range"],
                                voxel_size=c["voxel_size"], max_pts_per_voxel=c["max_pts"], stage_cap_ratio=4.0)
    shape = O.sparse_shape_zyx(grid)

    def run(p):
        out = eng.forward_points(torch.from_numpy(p))
        torch.cuda.synchronize()
        return out, eng.counts()

    # (1) nothing in range
    far = pts.copy()
    far[:, 1:4] += 1e4
    out, counts = run(far)
    assert counts == [0] * len(counts) and not eng.overflowed()
    assert out["spatial_features"].abs().sum().item() == 0
    # (2) ragged batch: frame 1 has no points; (3) one voxel (two points in the same cell of frame 0)
    one = pts[pts[:, 0] == 0][:1].copy()
    one = np.concatenate([one, one + np.array([0, 1e-3, 1e-3, 0, 0.1, 0.1], np.float32)])
    for p in (pts[pts[:, 0] == 0], one):
        f, co, _ = O.voxelize_mean_batch(p, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
        ref, _ = O.backbone_forward(prog, P, torch.from_numpy(f), co, shape, 2, O.QuantCfg(), {})
        out, counts = run(p)
        assert counts[0] == co.shape[0] and counts[-1] == ref.coords.shape[0]
        n = counts[-1]
        assert np.array_equal(out["encoded_coords"][:n].cpu().numpy(), ref.coords)
        check_feats(out["encoded_features"][:n], ref.features, False)
        ref_bev = O.height_compression(ref.features, ref.coords, ref.spatial_shape, 2)
        check_feats(out["spatial_features"], ref_bev, False)
        assert out["spatial_features"][1].abs().sum().item() == 0          # the empty frame's BEV map


def test_engine_sorted_voxelizer_steps_aside_when_a_frame_exceeds_its_voxel_cap():
    """The key-sorted front end cannot drop a frame's surplus voxels in the reference's first-touch order: it reports the frame counts,
    frame_cap_exceeded() says so, and after use_hash_voxelizer() the engine gives exactly what an engine built on the hash voxeliser
    gives (which test_voxelize_per_frame_voxel_cap holds to the oracle).  Under the cap both front ends agree row for row."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    per_frame = np.bincount(coords[:, 0], minlength=2)
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    mk = lambda cap_pf, sv: qlidar.BackboneEngine(bb, 2, 2 * max(cap_pf, 1), max_points=pts.shape[0] + 500, pc_range=c["pc_range"],
                                                   voxel_size=c["voxel_size"], max_pts_per_voxel=c["max_pts"], stage_cap_ratio=4.0,
                                                   max_voxels_per_frame=cap_pf, sorted_voxelizer=sv)

    def run(eng):
        for _ in range(2):
            out = eng.forward_points(torch.from_numpy(pts))
        torch.cuda.synchronize()
        n = eng.counts()
        return n, out["encoded_coords"][:n[-1]].cpu().numpy(), out["encoded_features"][:n[-1]].cpu().numpy()

    roomy = int(per_frame.max()) + 10
    es, eh = mk(roomy, True), mk(roomy, False)
    assert es.sorted_voxelizer and not eh.sorted_voxelizer
    ns, cs, fs = run(es)
    nh, ch, fh = run(eh)
    assert not es.frame_cap_exceeded() and ns == nh and np.array_equal(cs, ch) and np.array_equal(fs, fh)
    assert es.frame_counts.cpu().tolist() == per_frame.tolist()
    tight = int(per_frame.min()) - 100                                  # both frames over their cap
    es, eh = mk(tight, True), mk(tight, False)
    run(es)
    assert es.frame_cap_exceeded()
    es.use_hash_voxelizer()
    ns, cs, fs = run(es)
    nh, ch, fh = run(eh)
    assert ns[0] == 2 * tight and ns == nh and np.array_equal(cs, ch) and np.array_equal(fs, fh)
    assert mk(tight, None).sorted_voxelizer is False and mk(0, None).sorted_voxelizer is True


@pytest.mark.parametrize("arch", ["VoxelResBackBone8x", "VoxelResBackBone8xVoxelNeXt"])
def test_plugin_call_with_deferred_counts_gives_the_same_tensors(arch):
    """backbone.engine_lazy_counts: after one synchronous call has sized the engine, the plugin call queues the graph replay and returns
    without reading the row counts back; the published tensors resolve their shapes on first access and equal the synchronous call's.
    From raw points (VoxelizeMeanVFE attached) and from a voxelised batch_dict; an overflowing batch raises when its shape is read."""
    import qlidar
    pts, feats, coords, grid, c = make_frame("waymo", batch=2)
    nfeat = 5
    if arch == "VoxelResBackBone8xVoxelNeXt":
        bb = qlidar.VoxelResBackBone8xVoxelNeXt(qlidar.Cfg(SPCONV_KERNEL_SIZES=[3, 3, 3, 3], OUT_CHANNEL=128, CHANNELS=[16, 32, 64, 128, 128]),
                                                nfeat, np.asarray(grid)).cuda().eval()
    else:
        _, _, bb = build(arch, nfeat, grid)
    vfe = qlidar.VoxelizeMeanVFE({}, nfeat, c["voxel_size"], c["pc_range"], c["max_pts"], coords.shape[0] + 100).attach(bb)
    tp = torch.from_numpy(pts).cuda()

    def call():
        with torch.no_grad():
            return bb(vfe({"points": tp, "batch_size": 2}))

    sync = call()
    e0 = sync["encoded_spconv_tensor"]
    ref_f, ref_i = e0.features.clone(), e0.indices.clone()
    ref_taps = {k: (t.features.clone(), t.indices.clone()) for k, t in sync["multi_scale_3d_features"].items()}
    bb.engine_lazy_counts = True
    for _ in range(3):
        lazy = call()
    e1 = lazy["encoded_spconv_tensor"]
    assert type(e1).__name__ == "LazySparseConvTensor" and e1._lazy_f is None          # nothing read back yet
    assert e1.spatial_shape == e0.spatial_shape and e1.batch_size == 2
    assert e1.num_rows() == ref_f.shape[0]
    assert torch.equal(e1.indices, ref_i) and torch.equal(e1.features, ref_f) and e1.features.dtype == ref_f.dtype
    for k, t in lazy["multi_scale_3d_features"].items():
        assert torch.equal(t.features, ref_taps[k][0]) and torch.equal(t.indices, ref_taps[k][1]), k
    n0 = int(lazy["voxel_count"][0].item())
    assert n0 == coords.shape[0] and lazy["voxel_coords"].shape[0] >= n0
    # a batch twice as dense as anything the engine was sized for: the deferred call cannot re-run itself, it says so
    dense = torch.cat([tp, tp + torch.tensor([0, 0.05, 0.05, 0.07, 0, 0], device="cuda")[: tp.shape[1]]])[: int(tp.shape[0] * 1.2)].contiguous()
    eng = bb._engine_state["eng"]
    if dense.shape[0] <= eng.max_points:
        with torch.no_grad():
            over = bb(vfe({"points": dense, "batch_size": 2}))
        try:
            over["encoded_spconv_tensor"].num_rows()
            overflowed = False
        except qlidar.QlidarError:
            overflowed = True
        c_all = eng.counts_all.cpu()
        assert overflowed == bool((c_all[2:2 * len(eng.stages)].view(-1, 2)[:, 1] > c_all[2:2 * len(eng.stages)].view(-1, 2)[:, 0]).any()
                                  or eng.frame_cap_exceeded(c_all))
