"""GPU parity tests: every kernel is called through the C ABI (qlidar.ops -> libqlidar_b200.so) and compared with the
CPU oracle on the same seeded inputs.  Integer / index results are bit-exact; floating point tolerances are stated."""
import numpy as np
import pytest
import torch

import qlidar_oracle as O
from helpers import nbr_to_tiles, tiles_to_nbr, random_coords, dense_nbr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from qlidar import ops as _ops
    return _ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


# ------------------------------------------------------------------------------------------------ rulebooks
@pytest.mark.parametrize("ksize", [3, (3, 1, 1), (1, 3, 3), 5])
def test_rulebook_subm_bit_exact(ops, ksize):
    rng = np.random.default_rng(1)
    B, D, H, W = 2, 9, 40, 36
    coords = random_coords(rng, B, D, H, W, 0.08)
    ref = O.rulebook_subm(coords, [D, H, W], ksize)
    c = dev(coords)
    table = ops.hash_build(c, None, (B, D, H, W))
    nbr, kmask = ops.rulebook_subm(c, None, (B, D, H, W), ksize, table, with_mask=True)
    got = dense_nbr(ops, nbr, kmask, coords.shape[0])
    assert np.array_equal(got, ref)
    assert np.array_equal(kmask.cpu().numpy().view(np.uint32), O.tile_kmask(ref))
    # rows of the last tile beyond N are -1
    tail = ops.expand_rulebook(nbr, kmask).cpu().numpy().transpose(1, 0, 2).reshape(ref.shape[0], -1)[:, coords.shape[0]:]
    assert (tail == -1).all()
    # the same build without a mask is the dense layout
    nbr_d = ops.rulebook_subm(c, None, (B, D, H, W), ksize, table)
    assert np.array_equal(tiles_to_nbr(nbr_d.cpu().numpy(), coords.shape[0]), ref)


def test_rulebook_subm_device_count(ops):
    """n_dev < n_cap: only the first n rows take part; later tiles are untouched."""
    rng = np.random.default_rng(2)
    B, D, H, W = 1, 5, 30, 30
    coords = random_coords(rng, B, D, H, W, 0.2)
    n = coords.shape[0] - 57
    ref = O.rulebook_subm(coords[:n], [D, H, W], 3)
    c = dev(coords)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    table = ops.hash_build(c, n_dev, (B, D, H, W))
    nbr = ops.rulebook_subm(c, n_dev, (B, D, H, W), 3, table)
    assert np.array_equal(tiles_to_nbr(nbr.cpu().numpy(), n), ref)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0), (5, 2, 2), ((1, 3, 3), 1, (0, 1, 1))])
def test_rulebook_strided_bit_exact(ops, k, s, p):
    rng = np.random.default_rng(3)
    B, D, H, W = 2, 11, 33, 38
    coords = random_coords(rng, B, D, H, W, 0.05)
    oc_ref, osh, nbr_ref = O.rulebook_strided(coords, [D, H, W], k, s, p)
    c = dev(coords)
    cap = oc_ref.shape[0] + 300
    out_coords, n_out, out_table, nbr, ogrid, kmask = ops.rulebook_strided(c, None, (B, D, H, W), k, s, p, cap)
    n = int(n_out[0].item())
    assert int(n_out[1].item()) == n
    assert n == oc_ref.shape[0]
    assert list(ogrid[1:]) == list(osh)
    assert np.array_equal(out_coords[:n].cpu().numpy(), oc_ref)          # same order: ascending linear key
    assert np.array_equal(dense_nbr(ops, nbr, kmask, n), nbr_ref)
    tiles = (n + 127) // 128
    assert np.array_equal(kmask[:tiles].cpu().numpy().view(np.uint32), O.tile_kmask(nbr_ref))
    # order-free form required by the parity gate: (k, in_coord, out_coord) sorted sets
    a = O.pairs_in_coord_space(dense_nbr(ops, nbr, kmask, n), coords, out_coords[:n].cpu().numpy())
    b = O.pairs_in_coord_space(nbr_ref, coords, oc_ref)
    assert np.array_equal(a, b)
    # the output table maps out coords -> rows: a submanifold rulebook on the outputs must hit every centre
    nbr2 = ops.rulebook_subm(out_coords, n_out, ogrid, (1, 1, 1), out_table)
    assert np.array_equal(tiles_to_nbr(nbr2.cpu().numpy(), n)[0], np.arange(n))


@pytest.mark.parametrize("k,s,p,ksub", [(3, 2, 1, 3), (3, 2, (0, 1, 1), 3), ((3, 1, 1), (2, 1, 1), 0, (1, 3, 3)), (5, 2, 2, 5), (3, 2, 1, (3, 1, 1))])
def test_rank_index_subm_rulebook_and_bev_bit_exact(ops, k, s, p, ksub):
    """A strided build WITHOUT a hash table (out_table NULL) leaves the stage's rank index (bitmap + prefix) in its
    workspace; the submanifold rulebook and the BEV hand-off taken through it equal the hash-based ones and the oracle."""
    rng = np.random.default_rng(31)
    B, D, H, W = 2, 11, 35, 70
    coords = random_coords(rng, B, D, H, W, 0.07)
    oc_ref, osh, nbr_ref = O.rulebook_strided(coords, [D, H, W], k, s, p)
    n = oc_ref.shape[0]
    cap = n + 77
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes((B, D, H, W), k, s, p), dtype=torch.uint8, device="cuda")
    out_coords = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    n_out = torch.zeros(2, dtype=torch.int32, device="cuda")
    K = int(np.prod(O._triple(k)))
    nbr = torch.zeros(((cap + 127) // 128, K, 128), dtype=torch.int32, device="cuda")
    _, _, tbl, _, ogrid, kmask = ops.rulebook_strided(dev(coords), None, (B, D, H, W), k, s, p, cap, out=(out_coords, n_out, None, nbr), workspace=ws)
    assert tbl is None and n_out.tolist() == [n, n]
    assert np.array_equal(out_coords[:n].cpu().numpy(), oc_ref)
    assert np.array_equal(dense_nbr(ops, nbr, kmask, n), nbr_ref)
    index = ops.rulebook_strided_index((B, D, H, W), k, s, p, ws)
    sub_ref = O.rulebook_subm(oc_ref, osh, ksub)
    nbr2, kmask2 = ops.rulebook_subm_ranked(out_coords, n_out, ogrid, ksub, index)
    assert np.array_equal(dense_nbr(ops, nbr2, kmask2, n), sub_ref)
    tiles = (n + 127) // 128
    assert np.array_equal(kmask2[:tiles].cpu().numpy().view(np.uint32), O.tile_kmask(sub_ref))
    # identical to the hash path
    table = ops.hash_build(out_coords, n_out, ogrid)
    nbr3, kmask3 = ops.rulebook_subm(out_coords, n_out, ogrid, ksub, table, with_mask=True)
    assert torch.equal(ops.expand_rulebook(nbr2[:tiles], kmask2[:tiles]), ops.expand_rulebook(nbr3[:tiles], kmask3[:tiles])) and torch.equal(kmask2[:tiles], kmask3[:tiles])
    # BEV hand-off through the rank index
    C = 64
    f = torch.from_numpy(rng.normal(size=(cap, C)).astype(np.float32)).half()
    ref = O.height_compression(f[:n].float(), oc_ref, osh, B)
    out = ops.bev_densify_ranked(dev(f), index, n_out, ogrid)
    assert torch.equal(out.cpu().float(), ref)
    # a second strided conv ON this key-sorted stage, pairs taken through its rank index: identical to the scatter-form build
    for k2, s2, p2 in [(3, 2, 1), ((3, 1, 1), (2, 1, 1), 0)]:
        oc2_ref, osh2, nbr2_ref = O.rulebook_strided(oc_ref, osh, k2, s2, p2)
        m = oc2_ref.shape[0]
        oc2, n2, t2, nb2, og2, km2 = ops.rulebook_strided(out_coords, n_out, ogrid, k2, s2, p2, m + 50, in_index=index)
        assert t2 is None and n2.tolist() == [m, m]
        assert np.array_equal(oc2[:m].cpu().numpy(), oc2_ref) and list(og2[1:]) == list(osh2)
        assert np.array_equal(dense_nbr(ops, nb2, km2, m), nbr2_ref)
        assert np.array_equal(km2[:(m + 127) // 128].cpu().numpy().view(np.uint32), O.tile_kmask(nbr2_ref))
    # a row cap below the number of sites: the dropped (largest-key) sites are absent everywhere
    n_small = torch.tensor([n - 40, n], dtype=torch.int32, device="cuda")
    nbr4, km4 = ops.rulebook_subm_ranked(out_coords, n_small, ogrid, ksub, index)
    ref4 = O.rulebook_subm(oc_ref[:n - 40], osh, ksub)
    assert np.array_equal(dense_nbr(ops, nbr4, km4, n - 40), ref4)


def test_rulebook_strided_overflow_is_safe(ops):
    rng = np.random.default_rng(4)
    B, D, H, W = 1, 8, 20, 20
    coords = random_coords(rng, B, D, H, W, 0.1)
    oc_ref, _, _ = O.rulebook_strided(coords, [D, H, W], 3, 2, 1)
    c = dev(coords)
    cap = oc_ref.shape[0] // 2
    out_coords, n_out, out_table, nbr, ogrid, kmask = ops.rulebook_strided(c, None, (B, D, H, W), 3, 2, 1, cap)
    assert n_out.tolist() == [cap, oc_ref.shape[0]]                     # (kept, found): overflow is reported
    assert np.array_equal(out_coords.cpu().numpy(), oc_ref[:cap])
    assert ops.expand_rulebook(nbr[:(cap + 127) // 128], kmask[:(cap + 127) // 128]).max().item() < coords.shape[0]


# ------------------------------------------------------------------------------------------------ voxelization
@pytest.mark.parametrize("cfg,batch", [("kitti", 1), ("waymo", 2)])
def test_voxelize_sorted_is_the_hash_voxelizer_in_key_order(ops, cfg, batch):
    """ql_voxelize_sorted_*: the same voxels, point counts and (bit for bit) the same means as the reference-ordered voxeliser, rows in
    ascending linear key; the rank index it leaves serves the stage-1 rulebook; frame_counts = voxels per frame."""
    c = O.CONFIGS[cfg]
    kw = dict(n_az=500) if cfg == "kitti" else dict(n_beams=32, n_az=700)
    pts = O.synth_batch(cfg, batch, **kw)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    sshape = O.sparse_shape_zyx(grid)
    f_ref, c_ref, n_ref = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], 10 ** 7, sequential=True)
    order = np.argsort(O._lin(c_ref, sshape), kind="stable")
    cap = c_ref.shape[0] + 100
    g4 = (batch, *sshape)
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes(g4, 1, 1, 0), dtype=torch.uint8, device="cuda")
    fc = torch.zeros(batch, dtype=torch.int32, device="cuda")
    feats, coords, npts, n_dev = ops.voxelize_sorted(dev(pts), c["pc_range"], c["voxel_size"], grid, batch, c["max_pts"], cap, ws, frame_counts=fc)
    n = int(n_dev[0].item())
    assert n == c_ref.shape[0] and int(n_dev[1].item()) == n
    assert np.array_equal(coords[:n].cpu().numpy(), c_ref[order])
    assert np.array_equal(npts[:n].cpu().numpy(), n_ref[order])
    assert np.array_equal(feats[:n].cpu().numpy().view(np.uint32), f_ref[order].view(np.uint32))      # left-to-right sums: bit-exact
    assert np.array_equal(fc.cpu().numpy(), np.bincount(c_ref[:, 0], minlength=batch))
    # the same features as the hash voxeliser, row for row after sorting
    hf, hc, hn, hnd, _ = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, batch, c["max_pts"], cap)
    ho = np.argsort(O._lin(hc[:n].cpu().numpy(), sshape), kind="stable")
    assert np.array_equal(hf[:n].cpu().numpy()[ho].view(np.uint32), feats[:n].cpu().numpy().view(np.uint32))
    # rank index -> identity rulebook
    rank = ops.rulebook_strided_index(g4, 1, 1, 0, ws)
    nbr, km = ops.rulebook_subm_ranked(coords, n_dev, g4, (1, 1, 1), rank)
    assert np.array_equal(dense_nbr(ops, nbr, km, n)[0], np.arange(n))
    # batch capacity smaller than the voxel count: the first `cap` voxels in key order are kept, found is reported
    small = n // 2
    f2, c2, n2, nd2 = ops.voxelize_sorted(dev(pts), c["pc_range"], c["voxel_size"], grid, batch, c["max_pts"], small, ws)
    assert nd2.cpu().tolist() == [small, n]
    assert np.array_equal(c2[:small].cpu().numpy(), c_ref[order][:small])
    assert np.array_equal(f2[:small].cpu().numpy().view(np.uint32), f_ref[order][:small].view(np.uint32))


@pytest.mark.parametrize("cfg,batch", [("kitti", 1), ("waymo", 2)])
def test_voxelize_hard_matches_cpu_voxelizer(ops, cfg, batch):
    c = O.CONFIGS[cfg]
    kw = dict(n_az=500) if cfg == "kitti" else dict(n_beams=32, n_az=700)
    pts = O.synth_batch(cfg, batch, **kw)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    f_ref, c_ref, n_ref = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], 10 ** 7)
    cap = c_ref.shape[0] + 100
    feats, coords, npts, n_dev, table = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, batch, c["max_pts"], cap)
    n = int(n_dev[0].item())
    assert n == c_ref.shape[0]
    assert np.array_equal(coords[:n].cpu().numpy(), c_ref)               # identical first-touch order
    assert np.array_equal(npts[:n].cpu().numpy(), n_ref)
    # mean of <= 5 fp32 values: summation order may differ by 1 ulp
    np.testing.assert_allclose(feats[:n].cpu().numpy(), f_ref, rtol=1e-6, atol=1e-6)
    # the table hands coords -> row to the stage-1 rulebook
    sparse_shape = O.sparse_shape_zyx(grid)
    nbr = ops.rulebook_subm(coords, n_dev, (batch, *sparse_shape), (1, 1, 1), table)
    assert np.array_equal(tiles_to_nbr(nbr.cpu().numpy(), n)[0], np.arange(n))


def test_voxelize_padded_feature_rows(ops):
    """out_feats wider than n_feat: same means in the first columns, zeros in the pad (the engine uses a stride of 8)."""
    c = O.CONFIGS["waymo"]
    pts = O.synth_batch("waymo", 1, n_beams=16, n_az=400)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    f0, c0, n0, nd0, _ = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, 1, c["max_pts"], 20000)
    n = int(nd0[0].item())
    out = (torch.full((20000, 8), 7.0, device="cuda"), torch.zeros((20000, 4), dtype=torch.int32, device="cuda"),
           torch.zeros(20000, dtype=torch.int32, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda"),
           torch.empty(ops.hash_capacity(pts.shape[0]), dtype=torch.int64, device="cuda"))
    f1, c1, n1, nd1, _ = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, 1, c["max_pts"], 20000, out=out)
    assert int(nd1[0].item()) == n and torch.equal(c1[:n], c0[:n])
    assert torch.equal(f1[:n, :5], f0[:n]) and (f1[:n, 5:] == 0).all()


def test_voxelize_per_frame_voxel_cap(ops):
    """MAX_NUMBER_OF_VOXELS is a per-FRAME cap in the reference (each frame goes through the CPU voxeliser on its own,
    data_processor.py:151-153, then collate_batch concatenates): frames over the cap keep their first voxels in first-touch
    order, frames under it are untouched, ids stay dense."""
    c = O.CONFIGS["waymo"]
    pts = O.synth_batch("waymo", 3, n_beams=20, n_az=400)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    per_frame = [O.voxelize_hard(pts[pts[:, 0] == b][:, 1:], c["pc_range"], c["voxel_size"], c["max_pts"], 10 ** 7)[1].shape[0] for b in range(3)]
    cap = sorted(per_frame)[1] - 7                                      # two frames over the cap, one under it
    assert sum(v > cap for v in per_frame) == 2
    feats_ref, coords_ref, _ = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], c["max_pts"], cap)
    f, co, npts, nd, table = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, 3, c["max_pts"], 3 * cap + 100,
                                               max_voxels_per_frame=cap)
    n = int(nd[0].item())
    assert n == coords_ref.shape[0] == sum(min(v, cap) for v in per_frame) and int(nd[1].item()) == sum(per_frame)
    assert np.array_equal(co[:n].cpu().numpy(), coords_ref)
    assert torch.equal(f[:n].cpu(), torch.from_numpy(feats_ref))
    # the table knows exactly the kept voxels: a (1,1,1) submanifold rulebook hits every centre, and a dropped cell misses
    nbr = ops.rulebook_subm(co, nd, (3, *O.sparse_shape_zyx(grid)), (1, 1, 1), table)
    assert np.array_equal(tiles_to_nbr(nbr.cpu().numpy(), n)[0], np.arange(n))


def test_voxelize_voxel_cap(ops):
    c = O.CONFIGS["kitti"]
    pts = O.synth_batch("kitti", 1, n_az=400)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    f_ref, c_ref, n_ref = O.voxelize_mean_batch(pts, c["pc_range"], c["voxel_size"], 3, 1000)
    feats, coords, npts, n_dev, table = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, 1, 3, 1000)
    assert int(n_dev[0].item()) == 1000 == c_ref.shape[0] and int(n_dev[1].item()) > 1000
    assert np.array_equal(coords.cpu().numpy(), c_ref)
    assert np.array_equal(npts.cpu().numpy(), n_ref)
    np.testing.assert_allclose(feats.cpu().numpy(), f_ref, rtol=1e-6, atol=1e-6)


def test_voxelize_dynamic_matches_dynamic_mean_vfe(ops):
    c = O.CONFIGS["waymo"]
    pts = O.synth_batch("waymo", 2, n_beams=24, n_az=600)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    f_ref, c_ref, n_ref = O.voxelize_dynamic_mean(pts, c["pc_range"], c["voxel_size"])
    feats, coords, npts, n_dev, _ = ops.voxelize_mean(dev(pts), c["pc_range"], c["voxel_size"], grid, 2, 0, c_ref.shape[0] + 10)
    n = int(n_dev[0].item())
    assert n == c_ref.shape[0]
    got_c = coords[:n].cpu().numpy().astype(np.int64)
    order = np.lexsort((got_c[:, 1], got_c[:, 2], got_c[:, 3], got_c[:, 0]))   # DynamicMeanVFE key: b, x, y, z
    assert np.array_equal(got_c[order], c_ref)                          # coordinate SETS bit-exact
    assert np.array_equal(npts[:n].cpu().numpy()[order], n_ref)
    # unbounded number of fp32 atomic adds per voxel: tolerance 1e-5 relative to the coordinate magnitude (~75 m)
    np.testing.assert_allclose(feats[:n].cpu().numpy()[order], f_ref, rtol=1e-5, atol=1e-4)


def test_mean_vfe(ops):
    rng = np.random.default_rng(5)
    V, T, F = 1000, 5, 4
    num = rng.integers(0, T + 1, size=V).astype(np.int32)
    vox = rng.normal(size=(V, T, F)).astype(np.float32)
    vox[np.arange(T)[None, :] >= num[:, None]] = 0
    ref = O.mean_vfe(vox, num)
    got = ops.mean_vfe(dev(vox), dev(num.astype(np.float32)))
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-6, atol=1e-6)
    got = ops.mean_vfe(dev(vox), dev(num))
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------ sparse conv
def _conv_case(rng, n_target, cin, cout, ksize=3, subm=True, stride=1, pad=1):
    S = int(np.sqrt(n_target))
    coords = O.synth_surface_sheet(S, seed=int(rng.integers(1 << 30)), depth=12)
    shape = [12, S, S]
    if subm:
        nbr = O.rulebook_subm(coords, shape, ksize)
    else:
        _, _, nbr = O.rulebook_strided(coords, shape, ksize, stride, pad)
    return coords, nbr


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (64, 64), (128, 128), (16, 32), (64, 128), (128, 256), (256, 256)])
def test_spconv_i8_accumulators_bit_exact(ops, cin, cout):
    rng = np.random.default_rng(10 + cin + cout)
    coords, nbr = _conv_case(rng, 3000, cin, cout)
    N = coords.shape[0]
    qx = torch.from_numpy(rng.integers(-127, 128, size=(N, cin)).astype(np.int8))
    qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, 3, 3, 3, cin)).astype(np.int8))
    ref = O.sparse_conv_int(qx, nbr, qw)
    wp = ops.pack_weights(qw.reshape(cout, 27, cin)).cuda()
    one = torch.ones(cout, device="cuda")
    zero = torch.zeros(cout, device="cuda")
    out = torch.full((N, cout), -12345, dtype=torch.int32, device="cuda")
    ops.spconv_mma(dev(qx), dev(nbr_to_tiles(nbr)), N, None, cout, wp, one, zero, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref), f"mismatches: {(out.cpu() != ref).sum().item()} of {ref.numel()}"


@pytest.mark.parametrize("cin,cout,ksize,subm,stride,pad", [
    (16, 16, 3, True, 1, 1), (32, 64, 3, False, 2, 1), (64, 64, 3, True, 1, 1), (128, 128, (3, 1, 1), False, (2, 1, 1), 0),
    (32, 64, 5, False, 2, 2), (128, 128, 3, True, 1, 1), (64, 64, (1, 3, 3), True, 1, (0, 1, 1))])
def test_spconv_f16_matches_fp64_oracle(ops, cin, cout, ksize, subm, stride, pad):
    rng = np.random.default_rng(20 + cin)
    coords, nbr = _conv_case(rng, 2500, cin, cout, ksize, subm, stride, pad)
    n_in, n_out = coords.shape[0], nbr.shape[1]
    K = nbr.shape[0]
    x = torch.from_numpy(rng.normal(size=(n_in, cin)).astype(np.float32)).half()
    qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, K, cin)).astype(np.int8))
    k3 = O._triple(ksize)
    ref = O.sparse_conv(x.double(), nbr, qw.double().reshape((cout,) + k3 + (cin,)))
    wp = ops.pack_weights(qw.half()).cuda()
    one = torch.ones(cout, device="cuda")
    zero = torch.zeros(cout, device="cuda")
    out = ops.spconv_mma(dev(x), dev(nbr_to_tiles(nbr)), n_out, None, cout, wp, one, zero, out_dtype=torch.float32)
    # fp16 operands are exact, products exact in fp32, only the fp32 accumulation order differs: 1e-4 of max|ref|
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item(), err


def _ranked_stage(ops, coords, grid):
    """Key-sorted stage + rank index from a 1x1x1 stride-1 'strided' build over the given sites."""
    B, D, H, W = grid
    n = coords.shape[0]
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes(grid, 1, 1, 0), dtype=torch.uint8, device="cuda")
    oc, n_out, _, _, ogrid, _ = ops.rulebook_strided(dev(coords), None, grid, 1, 1, 0, n, out=None, workspace=ws)
    assert n_out.tolist() == [n, n] and tuple(ogrid) == tuple(grid)
    return oc, n_out, ops.rulebook_strided_index(grid, 1, 1, 0, ws)


@pytest.mark.parametrize("ksub", [3, (3, 1, 1), (1, 3, 3), 5])
def test_rulebook_subm_grouped_is_a_row_permutation_of_the_plain_rulebook(ops, ksub):
    """Grouped rulebook (rows binned by line key): row_perm is a permutation of the live rows (-1 padding), slot s holds
    exactly the plain rulebook's column of row row_perm[s] (so the pairs, compared in row space, are bit-exact against
    the oracle), the per-tile masks match the slots, and on a surface-like input the tiles need fewer live offsets."""
    S = 180
    coords = O.synth_surface_sheet(S, seed=11, depth=12)
    grid = (1, 12, S, S)
    oc, n_out, index = _ranked_stage(ops, coords, grid)
    n = coords.shape[0]
    oc_np = oc[:n].cpu().numpy()
    ref = O.rulebook_subm(oc_np, list(grid[1:]), ksub)                       # (K, n), rows in key order
    nbr_p, kmask_p = ops.rulebook_subm_ranked(oc, n_out, grid, ksub, index)
    assert np.array_equal(dense_nbr(ops, nbr_p, kmask_p, n), ref)
    nbr_g, kmask_g, perm = ops.rulebook_subm_ranked_grouped(oc, n_out, grid, ksub, index)
    tiles = (n + 127) // 128
    perm_np = perm[:tiles * 128].cpu().numpy()
    assert np.array_equal(np.sort(perm_np[:n]), np.arange(n)) and (perm_np[n:] == -1).all()
    g = dense_nbr(ops, nbr_g, kmask_g, tiles * 128)                          # (K, slots)
    assert np.array_equal(g[:, :n], ref[:, perm_np[:n]]) and (g[:, n:] == -1).all()
    km = kmask_g[:tiles].cpu().numpy().view(np.uint32)
    assert np.array_equal(km, O.tile_kmask(g))
    live = lambda m: sum(bin(int(v)).count("1") for v in m.ravel())
    if np.prod(O._triple(ksub)) == 27:
        assert live(km) < 0.8 * live(kmask_p[:tiles].cpu().numpy().view(np.uint32)), (live(km), live(kmask_p[:tiles].cpu().numpy().view(np.uint32)))
    # a device-side row count below the capacity: only the live rows are placed
    n_small = torch.tensor([n - 300, n], dtype=torch.int32, device="cuda")
    _, _, perm2 = ops.rulebook_subm_ranked_grouped(oc, n_small, grid, ksub, index)
    p2 = perm2.cpu().numpy()
    t2 = (n - 300 + 127) // 128
    assert np.array_equal(np.sort(p2[:n - 300]), np.arange(n - 300)) and (p2[n - 300:t2 * 128] == -1).all()


@pytest.mark.parametrize("cin,cout,int8", [(16, 16, True), (16, 16, False), (32, 32, False), (64, 64, True), (128, 128, False)])
def test_spconv_through_grouped_rulebook_is_bit_identical(ops, cin, cout, int8):
    """ql_spconv_mma_rows through a grouped rulebook == ql_spconv_mma through the plain one, bit for bit (outputs, fused
    int8 re-quantisation, residual, abs-max), and the int32 accumulators equal the oracle's."""
    rng = np.random.default_rng(5 + cin)
    S = 200
    coords = O.synth_surface_sheet(S, seed=3, depth=10)
    grid = (1, 10, S, S)
    oc, n_out, index = _ranked_stage(ops, coords, grid)
    n = coords.shape[0]
    nbr_p, kmask_p = ops.rulebook_subm_ranked(oc, n_out, grid, 3, index)
    nbr_g, kmask_g, perm = ops.rulebook_subm_ranked_grouped(oc, n_out, grid, 3, index)
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)).cuda() * 1e-3
    shift = torch.from_numpy(rng.normal(size=cout).astype(np.float32)).cuda()
    qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, 27, cin)).astype(np.int8))
    if int8:
        x = torch.from_numpy(rng.integers(-127, 128, size=(n, cin)).astype(np.int8)).cuda()
        w = ops.pack_weights(qw).cuda()
        acc = torch.zeros((n, cout), dtype=torch.int32, device="cuda")
        ops.spconv_mma(x, nbr_g, n, n_out, cout, w, torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda"), out=acc, kmask=kmask_g, row_perm=perm)
        ref = O.sparse_conv_int(x.cpu(), dense_nbr(ops, nbr_p, kmask_p, n), qw.reshape(cout, 3, 3, 3, cin))
        assert torch.equal(acc.cpu(), ref)
    else:
        x = torch.from_numpy(rng.normal(size=(n, cin)).astype(np.float32)).half().cuda()
        w = ops.pack_weights(qw.half()).cuda()
    res = torch.from_numpy(rng.normal(size=(n, cout)).astype(np.float32)).half().cuda()
    qs = torch.full((cout,), 20.0, device="cuda")
    outs = []
    for nbr, km, pm in ((nbr_p, kmask_p, None), (nbr_g, kmask_g, perm)):
        out = torch.zeros((n, cout), dtype=torch.float16, device="cuda")
        out_q = torch.zeros((n, cout), dtype=torch.int8, device="cuda")
        am = torch.zeros(cout, device="cuda")
        ops.spconv_mma(x, nbr, n, n_out, cout, w, scale, shift, residual=res, relu=True, out=out, out_q=out_q, out_qscale=qs, absmax=am,
                       kmask=km, row_perm=pm)
        outs.append((out, out_q, am))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert outs[0][0].abs().sum().item() > 0


@pytest.mark.parametrize("cin,cout,int8", [(16, 16, True), (32, 32, False), (64, 64, True), (128, 128, False)])
def test_spconv_many_tiles_per_cta_with_offset_mask(ops, cin, cout, int8):
    """> 3 tiles per persistent CTA (148 SMs) so the A/B ring and the accumulator double buffer wrap many times, through
    the GPU-built rulebook and its per-tile offset mask; int8: accumulators bit-exact, f16: 1e-4 of max|ref|."""
    rng = np.random.default_rng(70 + cin)
    S = 250
    coords = O.synth_surface_sheet(S, seed=7, depth=12)                 # 62 500 sites -> 489 tiles
    coords = coords[np.argsort(O._lin(coords, [12, S, S]), kind="stable")]
    N = coords.shape[0]
    nbr_ref = O.rulebook_subm(coords, [12, S, S], 3)
    c = dev(coords)
    table = ops.hash_build(c, None, (1, 12, S, S))
    nbr, kmask = ops.rulebook_subm(c, None, (1, 12, S, S), 3, table, with_mask=True)
    assert np.array_equal(dense_nbr(ops, nbr, kmask, N), nbr_ref)
    km = kmask.cpu().numpy().view(np.uint32)
    assert np.array_equal(km, O.tile_kmask(nbr_ref))
    assert np.mean([bin(int(v)).count("1") for v in km[:, 0]]) < 24      # the mask really skips slabs on this input
    one = torch.ones(cout, device="cuda")
    zero = torch.zeros(cout, device="cuda")
    if int8:
        qx = torch.from_numpy(rng.integers(-127, 128, size=(N, cin)).astype(np.int8))
        qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, 3, 3, 3, cin)).astype(np.int8))
        ref = O.sparse_conv_int(qx, nbr_ref, qw)
        out = torch.full((N, cout), -12345, dtype=torch.int32, device="cuda")
        ops.spconv_mma(dev(qx), nbr, N, None, cout, ops.pack_weights(qw.reshape(cout, 27, cin)).cuda(), one, zero, out=out, kmask=kmask)
        assert torch.equal(out.cpu(), ref), f"mismatches: {(out.cpu() != ref).sum().item()} of {ref.numel()}"
        out2 = torch.full((N, cout), -12345, dtype=torch.int32, device="cuda")
        ops.spconv_mma(dev(qx), ops.expand_rulebook(nbr, kmask), N, None, cout, ops.pack_weights(qw.reshape(cout, 27, cin)).cuda(), one, zero, out=out2)
        assert torch.equal(out2.cpu(), ref)                              # dense rulebook, no mask: same accumulators
    else:
        x = torch.from_numpy(rng.normal(size=(N, cin)).astype(np.float32)).half()
        qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, 27, cin)).astype(np.int8))
        ref = O.sparse_conv_auto(x.float(), nbr_ref, qw.float().reshape(cout, 3, 3, 3, cin)).double()
        out = ops.spconv_mma(dev(x), nbr, N, None, cout, ops.pack_weights(qw.half()).cuda(), one, zero, out_dtype=torch.float32, kmask=kmask)
        err = (out.cpu().double() - ref).abs().max().item()
        assert err <= 1e-4 * ref.abs().max().item(), err


def test_spconv_epilogue_fusion(ops):
    """dequant scale * BN scale/shift + residual + ReLU, fp16 out, int8 re-quantised out, per-channel absmax, device count."""
    rng = np.random.default_rng(30)
    cin = cout = 64
    coords, nbr = _conv_case(rng, 4000, cin, cout)
    N = coords.shape[0]
    n_live = N - 100
    x = torch.from_numpy(rng.normal(size=(N, cin)).astype(np.float32)).half()
    qw = torch.from_numpy(rng.integers(-127, 128, size=(cout, 27, cin)).astype(np.int8))
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)) * 1e-3
    shift = torch.from_numpy(rng.normal(size=cout).astype(np.float32))
    res = torch.from_numpy(rng.normal(size=(N, cout)).astype(np.float32)).half()
    act_scale = torch.tensor([0.37], dtype=torch.float32)
    qscale = torch.from_numpy(rng.uniform(20, 40, cout).astype(np.float32))
    nbr_live = nbr.copy()
    ref = O.sparse_conv(x.double(), nbr_live, qw.double().reshape(cout, 3, 3, 3, cin))
    y = torch.relu(ref * (scale.double() * 0.37) + shift.double() + res.double())[:n_live]
    n_dev = torch.tensor([n_live], dtype=torch.int32, device="cuda")
    out = torch.zeros((N, cout), dtype=torch.float16, device="cuda")
    out_q = torch.zeros((N, cout), dtype=torch.int8, device="cuda")
    absmax = torch.zeros(cout, dtype=torch.float32, device="cuda")
    ops.spconv_mma(dev(x), dev(nbr_to_tiles(nbr)), N, n_dev, cout, ops.pack_weights(qw.half()).cuda(), dev(scale), dev(shift),
                   act_scale=dev(act_scale), residual=dev(res), relu=True, out=out, out_q=out_q, out_qscale=dev(qscale), absmax=absmax)
    got = out.cpu().double()
    tol = 2e-3 * y.abs().max().item()                                   # fp16 output rounding (2^-11 relative)
    assert (got[:n_live] - y).abs().max().item() <= tol
    assert (got[n_live:] == 0).all()                                    # rows past the device count are not written
    np.testing.assert_allclose(absmax.cpu().numpy(), y.abs().amax(dim=0).numpy(), rtol=2e-3)
    q_ref = torch.clamp(torch.round(y * qscale.double()), -127, 127)
    dq = (out_q.cpu().double()[:n_live] - q_ref).abs()
    assert dq.max().item() <= 1 and (dq > 0).double().mean().item() < 0.02   # only rounding-boundary flips


@pytest.mark.parametrize("cin,cout,int8", [(16, 16, False), (32, 32, False), (16, 32, False), (32, 64, False), (16, 16, True),
                                            (16, 32, True), (32, 32, True), (64, 32, True)])
def test_spconv_register_gather_kernel_equals_the_tcgen05_kernel(ops, monkeypatch, cin, cout, int8):
    """csrc/spconv_warp.cu (mma.sync, one warp per 16 rows; opt-in through QL_SPCONV_WARP, profiles/r02_warp_kernel_ab.md) against
    k_spconv_ts on the same compact rulebook, full epilogue: INT32 accumulators / int8 paths bit-identical, fp16 within the fp32
    summation-order bound."""
    rng = np.random.default_rng(70 + cin + cout)
    S = 60
    coords_np = O.synth_surface_sheet(S, seed=71, depth=12)
    N = coords_np.shape[0]
    coords = torch.from_numpy(coords_np).cuda()
    grid = (1, 12, S, S)
    table = ops.hash_build(coords, None, grid)
    nbr, kmask = ops.rulebook_subm(coords, None, grid, 3, table, with_mask=True)
    n_live = N - 37
    n_dev = torch.tensor([n_live], dtype=torch.int32, device="cuda")
    if int8:
        x = torch.from_numpy(rng.integers(-127, 128, size=(N, cin)).astype(np.int8)).cuda()
        w = torch.from_numpy(rng.integers(-127, 128, size=(cout, 27, cin)).astype(np.int8))
    else:
        x = torch.from_numpy(rng.normal(size=(N, cin)).astype(np.float32)).half().cuda()
        w = torch.from_numpy(rng.integers(-127, 128, size=(cout, 27, cin)).astype(np.float32)).half()
    wp = ops.pack_weights(w).cuda()
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)).cuda() * 1e-3
    shift = torch.from_numpy(rng.normal(size=cout).astype(np.float32)).cuda()
    res = torch.from_numpy(rng.normal(size=(N, cout)).astype(np.float32)).half().cuda()
    qscale = torch.from_numpy(rng.uniform(20, 40, cout).astype(np.float32)).cuda()
    act = torch.tensor([0.37], dtype=torch.float32, device="cuda")

    def run(mode, raw):
        monkeypatch.setenv("QL_SPCONV_WARP", mode)
        if raw:
            out = torch.full((N, cout), -777, dtype=torch.int32, device="cuda")
            ops.spconv_mma(x, nbr, N, n_dev, cout, wp, torch.ones_like(scale), torch.zeros_like(shift), out=out, kmask=kmask)
            return (out,)
        out = torch.zeros((N, cout), dtype=torch.float16, device="cuda")
        out_q = torch.zeros((N, cout), dtype=torch.int8, device="cuda")
        absmax = torch.zeros(cout, dtype=torch.float32, device="cuda")
        ops.spconv_mma(x, nbr, N, n_dev, cout, wp, scale, shift, act_scale=act, residual=res, relu=True, out=out, out_q=out_q,
                       out_qscale=qscale, absmax=absmax, kmask=kmask)
        return out, out_q, absmax

    if int8:
        a, = run("0", True)
        b, = run("2", True)
        assert torch.equal(a, b)
        for ta, tb in zip(run("0", False), run("2", False)):
            assert torch.equal(ta, tb)
    else:
        a, aq, am = run("0", False)
        b, bq, bm = run("2", False)
        tol = 1e-3 * a.float().abs().max().item()
        assert (a.float() - b.float()).abs().max().item() <= tol
        assert (a[n_live:] == 0).all() and (b[n_live:] == 0).all()
        assert (aq.int() - bq.int()).abs().max().item() <= 1
        np.testing.assert_allclose(am.cpu().numpy(), bm.cpu().numpy(), rtol=1e-3)


def test_stem_conv(ops):
    rng = np.random.default_rng(40)
    coords, nbr = _conv_case(rng, 3000, 5, 16)
    N = coords.shape[0]
    x = torch.from_numpy(rng.normal(size=(N, 5)).astype(np.float32) * 10)
    w = torch.from_numpy(rng.normal(size=(16, 3, 3, 3, 5)).astype(np.float32))
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, 16).astype(np.float32))
    shift = torch.from_numpy(rng.normal(size=16).astype(np.float32))
    ref = torch.relu(O.sparse_conv(x.double(), nbr, w.double()) * scale.double() + shift.double())
    w_kio = w.reshape(16, 27, 5).permute(1, 2, 0).contiguous()
    absmax = torch.zeros(16, device="cuda")
    out = ops.stem_conv(dev(x), dev(nbr_to_tiles(nbr)), N, None, dev(w_kio), dev(scale), dev(shift), relu=True,
                        out_dtype=torch.float32, absmax=absmax)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    np.testing.assert_allclose(absmax.cpu().numpy(), ref.abs().amax(dim=0).numpy(), rtol=1e-5)
    # rows padded to 8 floats (the engine's voxel-feature layout): the 256-bit-load form gives the same bits
    x8 = torch.zeros((N, 8))
    x8[:, :5] = x
    out8 = ops.stem_conv(dev(x8), dev(nbr_to_tiles(nbr)), N, None, dev(w_kio), dev(scale), dev(shift), relu=True, out_dtype=torch.float32)
    assert torch.equal(out8, out)
    # the compact (masked) rulebook of the same sites gives the same bits
    c = dev(coords)
    S = int(np.sqrt(3000))
    table = ops.hash_build(c, None, (1, 12, S, S))
    nbr_c, km_c = ops.rulebook_subm(c, None, (1, 12, S, S), 3, table, with_mask=True)
    assert np.array_equal(dense_nbr(ops, nbr_c, km_c, N), nbr)
    outc8 = ops.stem_conv(dev(x8), nbr_c, N, None, dev(w_kio), dev(scale), dev(shift), relu=True, out_dtype=torch.float32, kmask=km_c)
    assert torch.equal(outc8, out)
    # 4 raw features (KITTI) and a width the specialised kernels do not cover
    for cin in (4, 7):
        xc = torch.from_numpy(rng.normal(size=(N, cin)).astype(np.float32))
        wc = torch.from_numpy(rng.normal(size=(16, 3, 3, 3, cin)).astype(np.float32))
        refc = torch.relu(O.sparse_conv(xc.double(), nbr, wc.double()) * scale.double() + shift.double())
        outc = ops.stem_conv(dev(xc), dev(nbr_to_tiles(nbr)), N, None, dev(wc.reshape(16, 27, cin).permute(1, 2, 0).contiguous()),
                             dev(scale), dev(shift), relu=True, out_dtype=torch.float32)
        assert (outc.cpu().double() - refc).abs().max().item() <= 1e-5 * refc.abs().max().item()


# ------------------------------------------------------------------------------------------------ quantizer
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_quantize_rows_codes_bit_exact(ops, dtype):
    rng = np.random.default_rng(50)
    x = torch.from_numpy(rng.normal(size=(5000, 64)).astype(np.float32))
    x[::97, 3] *= 20
    x = x.to(dtype)
    xf = x.float()
    amax = O.dynamic_amax(xf)
    q_ref = O.quantize_codes(xf, amax, 8)
    absmax = ops.absmax_cols(dev(x))
    np.testing.assert_array_equal(absmax.cpu().numpy(), xf.abs().amax(dim=0).numpy())
    q, act_scale = ops.quantize_rows(dev(x), absmax, ops.QL_Q_CODES_PER_TENSOR)
    assert torch.equal(q.cpu().int(), q_ref)
    assert abs(act_scale.item() - amax.item() / 127.0) <= 1e-7 * amax.item()


def test_quantize_rows_smooth_and_fake(ops):
    rng = np.random.default_rng(51)
    x = torch.from_numpy(rng.normal(size=(3000, 32)).astype(np.float32))
    s = torch.from_numpy(rng.uniform(0.5, 2.0, 32).astype(np.float32))
    absmax = ops.absmax_cols(dev(x))
    q, act_scale = ops.quantize_rows(dev(x), absmax, ops.QL_Q_CODES_PER_TENSOR, smooth=dev(s))
    xs = x / s
    assert torch.equal(q.cpu().int(), O.quantize_codes(xs, O.dynamic_amax(xs), 8))
    fq, _ = ops.quantize_rows(dev(x), absmax, ops.QL_Q_FAKE_PER_CHANNEL)
    ref = O.fake_quant(x, 8, axis=1)
    assert (fq.cpu().float() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()   # fp16 storage of the fake-quant value
    fr, _ = ops.quantize_rows(dev(x), None, ops.QL_Q_FAKE_PER_ROW)
    ref = O.fake_quant(x, 8, axis=0)
    assert (fr.cpu().float() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


# ------------------------------------------------------------------------------------------------ BEV
@pytest.mark.parametrize("D,H,W,C", [(2, 47, 45, 128), (2, 188, 188, 128), (5, 20, 24, 64)])
def test_bev_densify(ops, D, H, W, C):
    rng = np.random.default_rng(60)
    B = 2
    coords = random_coords(rng, B, D, H, W, 0.15)
    f = torch.from_numpy(rng.normal(size=(coords.shape[0], C)).astype(np.float32)).half()
    ref = O.height_compression(f.float(), coords, [D, H, W], B)
    c = dev(coords)
    table = ops.hash_build(c, None, (B, D, H, W))
    out = ops.bev_densify(dev(f), table, (B, D, H, W))
    assert out.shape == (B, C * D, H, W)
    assert torch.equal(out.cpu().float(), ref)
    out32 = ops.bev_densify(dev(f.float()), table, (B, D, H, W), out_dtype=torch.float32)
    assert torch.equal(out32.cpu(), ref)


# ------------------------------------------------------------------------------------------------ VoxelNeXt 2-D merge
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_bev_merge2d_matches_torch_unique_index_add(ops, dtype):
    """ql_bev_merge2d == VoxelResBackBone8xVoxelNeXt.bev_out (spconv_backbone_voxelnext.py:149-164) restated by the oracle
    (drop z, unique (b,y,x) ascending, duplicates summed): coordinates bit-exact, sums within fp32 re-association."""
    rng = np.random.default_rng(70)
    B, D, H, W, C = 2, 6, 47, 52, 64
    coords = random_coords(rng, B, D, H, W, 0.2)                      # many (b,y,x) duplicates across z
    f = torch.from_numpy(rng.normal(size=(coords.shape[0], C)).astype(np.float32)).to(dtype)
    ref_f, ref_c = O.bev_merge2d(f.float(), coords)
    out_f, out_c, n_out = ops.bev_merge2d(dev(f), dev(coords), None, (B, H, W))
    n = int(n_out[0].item())
    assert n == ref_c.shape[0] and int(n_out[1].item()) == n
    assert np.array_equal(out_c[:n].cpu().numpy(), ref_c)
    tol = 1e-6 if dtype == torch.float32 else 2e-3
    assert (out_f[:n].float().cpu() - ref_f).abs().max().item() <= tol * ref_f.abs().max().item()
    # device-side row count and an output cap below the number of sites
    n_dev = torch.tensor([coords.shape[0] - 500], dtype=torch.int32, device="cuda")
    ref_f2, ref_c2 = O.bev_merge2d(f[:coords.shape[0] - 500].float(), coords[:coords.shape[0] - 500])
    out_f2, out_c2, n_out2 = ops.bev_merge2d(dev(f), dev(coords), n_dev, (B, H, W), n_out_cap=ref_c2.shape[0] - 10)
    assert n_out2.tolist() == [ref_c2.shape[0] - 10, ref_c2.shape[0]]
    assert np.array_equal(out_c2.cpu().numpy(), ref_c2[:-10])
    assert (out_f2.float().cpu() - ref_f2[:-10]).abs().max().item() <= tol * ref_f2.abs().max().item()


def test_zero_rows_are_safe_everywhere(ops):
    """Device-side row count 0 (an empty batch): every op returns without touching its outputs' live rows or faulting."""
    B, D, H, W = 1, 5, 16, 16
    cap = 256
    coords = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    n0 = torch.zeros(2, dtype=torch.int32, device="cuda")
    table = ops.hash_build(coords, n0, (B, D, H, W))
    nbr, kmask = ops.rulebook_subm(coords, n0, (B, D, H, W), 3, table, with_mask=True)
    oc, n_out, _, nbr_s, ogrid, _ = ops.rulebook_strided(coords, n0, (B, D, H, W), 3, 2, 1, cap)
    assert n_out.tolist() == [0, 0]
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes((B, D, H, W), 3, 2, 1), dtype=torch.uint8, device="cuda")
    oc2, n_out2, _, _, ogrid2, _ = ops.rulebook_strided(coords, n0, (B, D, H, W), 3, 2, 1, cap, workspace=ws)
    index = ops.rulebook_strided_index((B, D, H, W), 3, 2, 1, ws)
    nbr_g, km_g, perm = ops.rulebook_subm_ranked_grouped(oc2, n_out2, ogrid2, 3, index)
    x = torch.ones((cap, 16), dtype=torch.float16, device="cuda")
    w = ops.pack_weights(torch.ones(16, 27, 16).half()).cuda()
    out = torch.full((cap, 16), 7.0, dtype=torch.float16, device="cuda")
    ops.spconv_mma(x, nbr, cap, n0, 16, w, torch.ones(16, device="cuda"), torch.zeros(16, device="cuda"), out=out, kmask=kmask)
    ops.spconv_mma(x, nbr_g, cap, n_out2, 16, w, torch.ones(16, device="cuda"), torch.zeros(16, device="cuda"), out=out, kmask=km_g, row_perm=perm)
    torch.cuda.synchronize()
    assert (out == 7.0).all()
    bev = ops.bev_densify_ranked(x, index, n_out2, ogrid2)
    assert bev.abs().sum().item() == 0
    pts = torch.full((10, 6), 1e9, dtype=torch.float32, device="cuda")
    pts[:, 0] = 0
    f, c, npts, nd, _ = ops.voxelize_mean(pts, [0, -40, -3, 70.4, 40, 1], [0.05, 0.05, 0.1], [1408, 1600, 40], 1, 5, 100)
    assert nd.tolist() == [0, 0]
    f, c, npts, nd, _ = ops.voxelize_mean(pts[:0], [0, -40, -3, 70.4, 40, 1], [0.05, 0.05, 0.1], [1408, 1600, 40], 1, 5, 100)
    assert nd.tolist() == [0, 0]


def test_renumber_by_key_sorts_and_leaves_a_rank_index(ops):
    """ql_renumber_by_key: distinct sites in random order -> ascending linear key (bit-exact vs numpy), src_row is the sorting
    permutation, payload rows move with it, and the rank index it leaves serves the ranked submanifold rulebook (== oracle)."""
    rng = np.random.default_rng(77)
    B, D, H, W = 2, 9, 150, 170
    coords = random_coords(rng, B, D, H, W, 0.02)                      # random order
    n = coords.shape[0]
    cap = n + 333
    grid = (B, D, H, W)
    feats = torch.from_numpy(rng.normal(size=(cap, 8)).astype(np.float32)).cuda()
    c_in = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    c_in[:n] = dev(coords)
    n_dev = torch.tensor([n, n], dtype=torch.int32, device="cuda")
    ws = torch.zeros(ops.rulebook_strided_workspace_bytes(grid, 1, 1, 0), dtype=torch.uint8, device="cuda")
    oc, n_out, src, rows = ops.renumber_by_key(c_in, n_dev, grid, ws, rows_in=feats)
    order = np.argsort(O._lin(coords, [D, H, W]), kind="stable")
    assert n_out.tolist() == [n, n]
    assert np.array_equal(oc[:n].cpu().numpy(), coords[order])
    assert np.array_equal(src[:n].cpu().numpy(), order.astype(np.int32))
    assert torch.equal(rows[:n].cpu(), feats[:n].cpu()[torch.from_numpy(order)])
    index = ops.rulebook_strided_index(grid, 1, 1, 0, ws)
    nbr, kmask = ops.rulebook_subm_ranked(oc, n_out, grid, 3, index)
    ref = O.rulebook_subm(coords[order], [D, H, W], 3)
    assert np.array_equal(dense_nbr(ops, nbr, kmask, n), ref)
    # coordinates only (no payload), and an empty list
    oc2, n2, src2, r2 = ops.renumber_by_key(c_in, n_dev, grid, ws)
    assert r2 is None and torch.equal(oc2[:n], oc[:n]) and torch.equal(src2[:n], src[:n])
    zero = torch.zeros(2, dtype=torch.int32, device="cuda")
    _, n3, _, _ = ops.renumber_by_key(c_in, zero, grid, ws)
    assert n3.tolist() == [0, 0]
    # permute_rows with the same table
    moved = ops.permute_rows(feats, src, n_out)
    assert torch.equal(moved[:n], rows[:n])


def test_voxelize_phases_equal_the_single_call(ops):
    """ql_voxelize_coords followed by ql_voxelize_features (same buffers) == ql_voxelize_mean; coordinates and the voxel count
    are already final after the first phase."""
    c = O.CONFIGS["waymo"]
    pts = O.synth_batch("waymo", 2, n_beams=24, n_az=400)
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    p = dev(pts)
    cap = 60000
    a = ops.voxelize_mean(p, c["pc_range"], c["voxel_size"], grid, 2, c["max_pts"], cap)
    ws = torch.zeros(int(ops.lib().ql_voxelize_workspace_bytes(pts.shape[0], cap, 5, c["max_pts"])), dtype=torch.uint8, device="cuda")
    out = (torch.full((cap, 5), -7.0, device="cuda"), torch.zeros((cap, 4), dtype=torch.int32, device="cuda"),
           torch.zeros(cap, dtype=torch.int32, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda"),
           torch.empty(ops.hash_capacity(pts.shape[0]), dtype=torch.int64, device="cuda"))
    ops.voxelize_mean(p, c["pc_range"], c["voxel_size"], grid, 2, c["max_pts"], cap, out=out, workspace=ws, phase="coords")
    n = int(a[3][0].item())
    assert out[3].tolist() == a[3].tolist() and torch.equal(out[1][:n], a[1][:n])
    assert (out[0][:n] == -7.0).all()                                   # features untouched so far
    ops.voxelize_mean(p, c["pc_range"], c["voxel_size"], grid, 2, c["max_pts"], cap, out=out, workspace=ws, phase="features")
    assert torch.equal(out[0][:n], a[0][:n]) and torch.equal(out[2][:n], a[2][:n]) and torch.equal(out[1][:n], a[1][:n])


# ------------------------------------------------------------------------------------------------ SmoothQuant weight preparation
@pytest.mark.parametrize("cin,cout,K", [(16, 16, 27), (32, 64, 27), (64, 64, 27), (128, 128, 3), (256, 256, 9), (48, 32, 5)])
def test_sq_prepare_weights_layout_and_codes(ops, cin, cout, K):
    """ql_sq_prepare_weights writes the conv kernel's packed image: with act_absmax == the weights' own per-channel maxima and
    alpha = 0.5 the smoothing scale is exactly 1 (a^0.5 / a^0.5), so the image must equal, byte for byte, ql_pack_weights_host of
    the plain per-output-channel codes, and the scale amax_w / 127 * bn; with real activation maxima the codes equal the host
    formula's up to the device's powf (<= 1 code step on a handful of weights)."""
    rng = np.random.default_rng(cin + cout)
    w = torch.from_numpy(rng.normal(size=(cout, K, cin)).astype(np.float32))
    bn = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    w_ic = w.abs().amax(dim=(0, 1))
    smooth, packed, scale = ops.sq_prepare_weights(dev(w), dev(w_ic), dev(w_ic), 0.5, bn_scale=dev(bn))
    assert torch.equal(smooth.cpu(), torch.ones(cin))
    amax = w.abs().amax(dim=(1, 2))
    codes = torch.round(w * (torch.full_like(amax, 127.0) / amax).view(-1, 1, 1)).clamp_(-127, 127).to(torch.int8)
    assert torch.equal(packed.cpu(), ops.pack_weights(codes))
    assert torch.equal(scale.cpu(), (amax / 127.0) * bn)
    # real statistics: s = amax_x^a / amax_w^(1-a)
    ax = torch.from_numpy(rng.uniform(0.1, 30.0, cin).astype(np.float32))
    ax[3] = 0.0                                                               # a dead channel: s -> 1
    smooth, packed, scale = ops.sq_prepare_weights(dev(w), dev(w_ic), dev(ax), 0.8)
    s_ref = O.smoothquant_scale(ax, w.reshape(cout, K, 1, 1, cin), 0.8)
    assert torch.allclose(smooth.cpu(), s_ref, rtol=2e-6) and smooth[3].item() == 1.0
    ws = w * smooth.cpu().view(1, 1, -1)
    amax2 = ws.abs().amax(dim=(1, 2))
    codes2 = torch.round(ws * (torch.full_like(amax2, 127.0) / amax2).view(-1, 1, 1)).clamp_(-127, 127).to(torch.int8)
    assert torch.equal(packed.cpu(), ops.pack_weights(codes2))                # same smoothing vector -> same image
    assert torch.equal(scale.cpu(), amax2 / 127.0)


def test_bev_merge2d_multi_equals_concat_then_merge(ops):
    """ql_bev_merge2d_multi (VoxelNeXt: stages 4 / 5 / 6 merged in one pass, coarser stages scaled onto the target grid) == the
    oracle's concat -> drop z -> unique -> index_add_ (spconv_backbone_voxelnext.py:149-164,194-199); device-side counts, 4-column
    output coordinates, and the bitmap + prefix it leaves serve as the merged stage's rank index."""
    rng = np.random.default_rng(91)
    B, H, W, C = 2, 48, 40, 32
    segs_np, scales = [], [1, 2, 4]
    for sc in scales:
        cs = random_coords(rng, B, 3, H // sc, W // sc, 0.15)
        f = rng.normal(size=(cs.shape[0] + 17, C)).astype(np.float32)
        segs_np.append((f, cs))
    cat_f = torch.cat([torch.from_numpy(f[:c.shape[0]]).half().float() for f, c in segs_np])
    cat_c = np.concatenate([np.concatenate([c[:, :2], c[:, 2:] * sc], axis=1) for (f, c), sc in zip(segs_np, scales)])
    ref_f, ref_c = O.bev_merge2d(cat_f, cat_c)
    segs = []
    for (f, c), sc in zip(segs_np, scales):
        cc = torch.zeros((f.shape[0], 4), dtype=torch.int32, device="cuda")
        cc[:c.shape[0]] = dev(c)
        segs.append((dev(torch.from_numpy(f).half()), cc, torch.tensor([c.shape[0], c.shape[0]], dtype=torch.int32, device="cuda"), sc))
    cap = ref_c.shape[0] + 9
    out_f = ops.zero_led_rows(cap, C)
    out_c = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    n_out = torch.zeros(2, dtype=torch.int32, device="cuda")
    ws = torch.zeros(int(ops.lib().ql_bev_merge2d_workspace_bytes(B, H, W, cap, C, ops.QL_F16)), dtype=torch.uint8, device="cuda")
    ops.bev_merge2d_multi(segs, (B, H, W), cap, out_f, out_c, n_out, ws)
    n = ref_c.shape[0]
    assert n_out.tolist() == [n, n]
    got_c = out_c[:n].cpu().numpy()
    assert np.array_equal(got_c[:, [0, 2, 3]], ref_c) and (got_c[:, 1] == 0).all()
    assert (out_f[:n].float().cpu() - ref_f).abs().max().item() <= 2e-3 * ref_f.abs().max().item()
    # rank index left in the workspace: a submanifold 2-D rulebook through it == the oracle's
    n_words = (B * H * W + 31) // 32
    index = ops.RankIndex(ws, ws.data_ptr(), ws.data_ptr() + ((n_words * 4 + 255) & ~255), n_words)
    nbr, km = ops.rulebook_subm_ranked(out_c, n_out, (B, 1, H, W), (1, 3, 3), index)
    assert np.array_equal(dense_nbr(ops, nbr, km, n), O.rulebook_subm(got_c, [1, H, W], (1, 3, 3)))
