"""Full-size parity (BASELINE configs[1]: Waymo-shaped batch of 4 synthetic ~146 k-voxel frames) through size-independent
properties -- the CPU oracle needs ~14 s per frame here, so at this size the CUDA path is checked against invariants of the
reference algorithm instead of the oracle's outputs:
  * voxel set: every in-range point's cell is a voxel, voxel count == number of distinct cells (torch.unique on the GPU);
  * strided outputs: ascending unique keys, and out-set == dilate-subsample of the in-set (checked by membership);
  * submanifold rulebook: centre tap is the identity and the rulebook is symmetric (j = nbr[k][i]  <=>  i = nbr[K-1-k][j]);
  * ranked (bitmap) rulebooks == hash rulebooks bit for bit;
  * INT8 conv: checksum of checksums -- sum over output rows of the INT32 accumulators equals, per output channel,
    sum_k (sum of the gathered int8 rows at offset k) . W_k, evaluated exactly in int64 with torch ops;
  * f16 conv: linearity, conv(a*x1 + b*x2) == a*conv(x1) + b*conv(x2) within fp16 rounding;
  * BEV: the dense map's per-channel sums equal the sparse features' per-channel sums.
All through the C ABI (qlidar.ops)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame():
    from qlidar import ops, synth
    c = synth.CONFIGS["waymo"]
    pts = torch.from_numpy(synth.synth_batch("waymo", 4)).cuda()
    grid_xyz = synth.grid_size_xyz(c["pc_range"], c["voxel_size"])
    cap = 700000
    feats, coords, npts, n_dev, table = ops.voxelize_mean(pts, c["pc_range"], c["voxel_size"], grid_xyz, 4, c["max_pts"], cap)
    n = int(n_dev[0].item())
    assert int(n_dev[1].item()) == n and 500000 < n < cap
    sshape = synth.sparse_shape_zyx(grid_xyz)
    return dict(ops=ops, cfg=c, pts=pts, feats=feats, coords=coords, n=n, n_dev=n_dev, table=table, grid=(4, *sshape), grid_xyz=grid_xyz)


def _keys(coords, grid):
    B, D, H, W = grid
    c = coords.long()
    return ((c[:, 0] * D + c[:, 1]) * H + c[:, 2]) * W + c[:, 3]


def test_voxel_set_is_the_set_of_occupied_cells(frame):
    c, pts, grid = frame["cfg"], frame["pts"], frame["grid"]
    r = torch.tensor(c["pc_range"], device="cuda")
    vs = torch.tensor(c["voxel_size"], device="cuda")
    cell = torch.floor((pts[:, 1:4] - r[:3]) / vs)                       # fp32 floor((p - min) / vs), dynamic_mean_vfe.py:53
    g = torch.tensor([float(v) for v in frame["grid_xyz"]], device="cuda")
    ok = ((cell >= 0) & (cell < g)).all(dim=1)
    cz = torch.stack([pts[ok, 0].long(), cell[ok, 2].long(), cell[ok, 1].long(), cell[ok, 0].long()], dim=1)
    want = torch.unique(_keys(cz, grid))
    got = _keys(frame["coords"][:frame["n"]], grid)
    assert got.numel() == want.numel() and torch.equal(torch.sort(got)[0], want)


def test_subm_rulebook_identity_and_symmetry(frame):
    ops, n = frame["ops"], frame["n"]
    nbr, kmask = ops.rulebook_subm(frame["coords"], frame["n_dev"], frame["grid"], 3, frame["table"], with_mask=True)
    tiles = (n + 127) // 128
    flat = ops.expand_rulebook(nbr[:tiles], kmask[:tiles]).permute(1, 0, 2).reshape(27, -1)[:, :n]          # (K, n)
    rows = torch.arange(n, device="cuda", dtype=torch.int32)
    assert torch.equal(flat[13], rows)                                  # centre tap
    for k in (0, 4, 9, 12):
        j = flat[k]
        m = j >= 0
        assert torch.equal(flat[26 - k][j[m].long()], rows[m])          # symmetric pairs
    assert int((flat >= 0).sum().item()) > 3 * n


def test_strided_outputs_sorted_unique_and_ranked_paths_match_hash(frame):
    ops = frame["ops"]
    grid = frame["grid"]
    cap = 800000
    oc, n_out, table, nbr, ogrid, kmask = ops.rulebook_strided(frame["coords"], frame["n_dev"], grid, 3, 2, 1, cap)
    m = int(n_out[0].item())
    assert int(n_out[1].item()) == m
    keys = _keys(oc[:m], ogrid)
    assert bool((keys[1:] > keys[:-1]).all())                           # ascending, unique
    # out-set == { (c + pad - k) / stride exact, in range }: every pair's output coordinate is consistent with its input's
    tiles = (m + 127) // 128
    flat = ops.expand_rulebook(nbr[:tiles], kmask[:tiles]).permute(1, 0, 2).reshape(27, -1)[:, :m]
    assert bool(((flat >= 0).sum(dim=0) > 0).all())                     # every output site has at least one input
    ins = frame["coords"][:frame["n"]].long()
    for k in (0, 13, 26):
        kz, ky, kx = k // 9, (k // 3) % 3, k % 3
        o = torch.nonzero(flat[k] >= 0).squeeze(1)
        i = flat[k][o].long()
        oc_k = oc[o].long()
        assert torch.equal(ins[i, 0], oc_k[:, 0])
        assert torch.equal(ins[i, 1], oc_k[:, 1] * 2 - 1 + kz) and torch.equal(ins[i, 2], oc_k[:, 2] * 2 - 1 + ky) and torch.equal(ins[i, 3], oc_k[:, 3] * 2 - 1 + kx)
    # every input reaches at least one output (k=3, s=2, p=1 covers the grid)
    used = torch.zeros(frame["n"], dtype=torch.bool, device="cuda")
    used[flat[flat >= 0].long()] = True
    assert bool(used.all())
    # rank-index paths on this key-sorted stage == hash paths, bit for bit
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes(grid, 3, 2, 1), dtype=torch.uint8, device="cuda")
    oc2 = torch.zeros_like(oc); n2 = torch.zeros_like(n_out); nbr2 = torch.zeros_like(nbr)
    km2 = ops.rulebook_strided(frame["coords"], frame["n_dev"], grid, 3, 2, 1, cap, out=(oc2, n2, None, nbr2), workspace=ws)[-1]
    ex = ops.expand_rulebook
    assert torch.equal(oc2[:m], oc[:m]) and torch.equal(ex(nbr2[:tiles], km2[:tiles]), ex(nbr[:tiles], kmask[:tiles]))
    index = ops.rulebook_strided_index(grid, 3, 2, 1, ws)
    a, ka = ops.rulebook_subm_ranked(oc, n_out, ogrid, 3, index)
    b, kb = ops.rulebook_subm(oc, n_out, ogrid, 3, table, with_mask=True)
    assert torch.equal(ex(a[:tiles], ka[:tiles]), ex(b[:tiles], kb[:tiles])) and torch.equal(ka[:tiles], kb[:tiles])
    oc3_h, n3_h, _, nbr3_h, og3, km3_h = ops.rulebook_strided(oc, n_out, ogrid, 3, 2, 1, 400000)
    oc3_r, n3_r, _, nbr3_r, _, km3_r = ops.rulebook_strided(oc, n_out, ogrid, 3, 2, 1, 400000, in_index=index)
    m3 = int(n3_h[0].item())
    t3 = (m3 + 127) // 128
    assert n3_r.tolist() == n3_h.tolist() and torch.equal(oc3_r[:m3], oc3_h[:m3])
    assert torch.equal(ex(nbr3_r[:t3], km3_r[:t3]), ex(nbr3_h[:t3], km3_h[:t3])) and torch.equal(km3_r[:t3], km3_h[:t3])


@pytest.mark.parametrize("C", [16, 64, 128])
def test_int8_accumulator_checksum_of_checksums(frame, C):
    ops, n = frame["ops"], frame["n"]
    nbr, kmask = ops.rulebook_subm(frame["coords"], frame["n_dev"], frame["grid"], 3, frame["table"], with_mask=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(-127, 128, (frame["coords"].shape[0], C), generator=g, device="cuda", dtype=torch.int8)
    w = torch.randint(-127, 128, (C, 27, C), generator=g, device="cuda", dtype=torch.int8)
    packed = ops.pack_weights(w.cpu()).cuda()
    acc = torch.zeros((frame["coords"].shape[0], C), dtype=torch.int32, device="cuda")
    one = torch.ones(C, device="cuda")
    ops.spconv_mma(x, nbr, frame["coords"].shape[0], frame["n_dev"], C, packed, one, torch.zeros(C, device="cuda"), out=acc, kmask=kmask)
    got = acc[:n].long().sum(dim=0)                                     # per output channel, exact
    tiles = (n + 127) // 128
    flat = ops.expand_rulebook(nbr[:tiles], kmask[:tiles]).permute(1, 0, 2).reshape(27, -1)[:, :n]
    want = torch.zeros(C, dtype=torch.int64, device="cuda")
    for k in range(27):
        idx = flat[k][flat[k] >= 0].long()
        s_k = x[idx].long().sum(dim=0)                                  # sum of the gathered rows at offset k, (C_in,)
        want += (w[:, k, :].long() * s_k.view(1, -1)).sum(dim=1)
    assert torch.equal(got, want)
    # and a sampled set of rows against a direct int64 evaluation
    rows = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(1)).tolist()
    for r in rows[:16]:
        ref = torch.zeros(C, dtype=torch.int64, device="cuda")
        for k in range(27):
            j = int(flat[k, r].item())
            if j >= 0:
                ref += (w[:, k, :].long() * x[j].long().view(1, -1)).sum(dim=1)
        assert torch.equal(acc[r].long(), ref)


def test_f16_conv_linearity(frame):
    ops, n = frame["ops"], frame["n"]
    C = 32
    nbr, kmask = ops.rulebook_subm(frame["coords"], frame["n_dev"], frame["grid"], 3, frame["table"], with_mask=True)
    g = torch.Generator(device="cuda").manual_seed(6)
    N = frame["coords"].shape[0]
    # small integers: every product and partial sum is exact in fp16 inputs / fp32 accumulation, so linearity holds exactly
    x1 = torch.randint(-8, 9, (N, C), generator=g, device="cuda").half()
    x2 = torch.randint(-8, 9, (N, C), generator=g, device="cuda").half()
    w = torch.randint(-4, 5, (C, 27, C), generator=g, device="cuda").half()
    packed = ops.pack_weights(w.cpu()).cuda()
    one, zero = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")

    def conv(x):
        return ops.spconv_mma(x.contiguous(), nbr, N, frame["n_dev"], C, packed, one, zero, out_dtype=torch.float32, kmask=kmask)[:n]

    y1, y2, y12 = conv(x1), conv(x2), conv(2 * x1 - 3 * x2)
    assert torch.equal(y12, 2 * y1 - 3 * y2)
    assert y1.abs().max().item() > 0


def test_bev_channel_sums(frame):
    ops = frame["ops"]
    grid = frame["grid"]
    oc, n_out, table, nbr, ogrid, _ = ops.rulebook_strided(frame["coords"], frame["n_dev"], grid, 3, 2, 1, 800000)
    # collapse to a coarse [2, H, W]-like stage the way conv_out does: use the stage itself with D planes
    m = int(n_out[0].item())
    C = 16
    f = torch.randint(-50, 51, (oc.shape[0], C), device="cuda").half()
    out = ops.bev_densify(f, table, ogrid, out_dtype=torch.float32)
    B, D, H, W = ogrid
    assert out.shape == (B, C * D, H, W)
    per_c = out.view(B, C, D, H, W).double().sum(dim=(0, 2, 3, 4))
    assert torch.equal(per_c, f[:m].double().sum(dim=0))
    assert int((out != 0).sum().item()) == int((f[:m] != 0).sum().item())


# ------------------------------------------------------------------------------------------------ full-size oracle fixture
# BASELINE configs[1] at its real size against the ORACLE'S OWN OUTPUTS (tests/golden/fullsize_waymo_b4.npz, produced once in the
# build container by tests/golden/make_golden_fullsize.py: minutes of CPU, too long for the GPU box).
@pytest.fixture(scope="module")
def full():
    import os
    import qlidar_oracle as O
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_waymo_b4.npz")
    g = np.load(p)
    pts = O.synth_batch("waymo", 4)                                       # seeds 1000-1003, the bench's frames
    assert pts.shape[0] == int(g["n_points"])
    return g, pts, O


def _full_engine(O, pts, act_bits, cw, amax=None):
    import qlidar
    from test_gpu_backbone import build
    c = O.CONFIGS["waymo"]
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    prog, P, bb = build("VoxelResBackBone8x", 5, grid)
    qlidar.q_conv3d(bb, {}, "", 8, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), ["conv_input.0"])
    if amax is not None:
        for n, m in bb.named_modules():
            if n.endswith("act_quant"):
                m.amax = torch.tensor(float(amax[n[:-len(".act_quant")]]), dtype=torch.float32, device="cuda")
    cap = 4 * c["max_voxels"]
    eng = qlidar.BackboneEngine(bb, 4, cap, max_points=pts.shape[0], pc_range=c["pc_range"], voxel_size=c["voxel_size"],
                                max_pts_per_voxel=c["max_pts"], use_graph=True, max_voxels_per_frame=c["max_voxels"],
                                stage_caps=[cap, int(1.25 * cap), int(0.75 * cap), int(0.5 * cap), int(0.5 * cap)])
    for _ in range(2):
        out = eng.forward_points(torch.from_numpy(pts))
    torch.cuda.synchronize()
    assert not eng.overflowed()
    return eng, out


def test_fullsize_w8a16_cw_matches_the_oracles_outputs(full):
    """The headline configuration (W8A16, cw) on the bench's four frames: stage counts and encoded indices equal the oracle's, every
    16th encoded row within 1e-2 of max|ref| (the north star's feature tolerance), per-channel sums of every stage within 1e-3."""
    g, pts, O = full
    eng, out = _full_engine(O, pts, 16, True)
    counts = eng.counts()
    assert counts == [int(v) for v in g["stage_counts"]]
    n = counts[-1]
    assert np.array_equal(out["encoded_coords"][:n].cpu().numpy(), g["encoded_indices"].astype(np.int32))
    stride = int(g["row_stride"])
    ref_rows = torch.from_numpy(g["w8a16_cw:encoded_rows"].astype(np.float32))
    got_rows = out["encoded_features"][:n][::stride].float().cpu()
    m = float(g["w8a16_cw:encoded_abs_max"])
    assert (got_rows - ref_rows).abs().max().item() <= 1e-2 * m
    sums = out["encoded_features"][:n].double().sum(dim=0).cpu().numpy()
    ref_sums = g["w8a16_cw:encoded_channel_sums"]
    assert np.abs(sums - ref_sums).max() <= 1e-3 * np.abs(ref_sums).max()
    for name, (f, st) in out["taps"].items():
        k = counts[eng.stages.index(st)]
        s = f[:k].double().sum(dim=0).cpu().numpy()
        r = g[f"w8a16_cw:{name}_channel_sums"]
        assert np.abs(s - r).max() <= 1e-3 * np.abs(r).max(), name
    # the BEV hand-off carries exactly those rows: channel c*D + d sums back to channel c
    bev = out["spatial_features"].double().sum(dim=(0, 2, 3)).view(-1, 2).sum(dim=1).cpu().numpy()
    assert np.abs(bev - ref_sums).max() <= 1e-3 * np.abs(ref_sums).max()


def test_fullsize_w8a8_pt_static_is_bit_exact_against_the_mirror(full):
    """W8A8 per-tensor with frozen amax on the same four frames, from raw points: the int8 codes of all 20 quantised layers and the fp16
    rows of all 21 layers have the oracle mirror's checksums, and every 16th encoded row is identical -- tolerance zero at 582 k voxels."""
    g, pts, O = full
    names = [str(s) for s in g["w8a8_pt_static:layers"]]
    amax = {n: a for n, a in zip(names, g["w8a8_pt_static:amax"])}
    eng, out = _full_engine(O, pts, 8, False, amax)
    counts = eng.counts()
    assert counts == [int(v) for v in g["stage_counts"]]
    assert [L.name for L in eng.layers] == names
    for i, L in enumerate(eng.layers):
        n_in, n_out = counts[L.stage_in], counts[L.stage_out]
        assert n_out == int(g["w8a8_pt_static:n_out"][i])
        if L.kind == "i8":
            codes = L.q_buf[:n_in].to(torch.int64)
            assert int(codes.sum().item()) == int(g["w8a8_pt_static:codes_sum"][i]), L.name
            assert int(codes.abs().sum().item()) == int(g["w8a8_pt_static:codes_abs_sum"][i]), L.name
        bits = int(L.out[:n_out].view(torch.int16).to(torch.int64).bitwise_and(0xFFFF).sum().item())
        assert bits == int(g["w8a8_pt_static:out_bits_sum"][i]), L.name
    n = counts[-1]
    got = out["encoded_features"][:n][::int(g["row_stride"])].cpu().numpy()
    assert np.array_equal(got.view(np.uint16), g["w8a8_pt_static:encoded_rows"].view(np.uint16))
