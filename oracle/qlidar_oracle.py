"""CPU ORACLE (test infrastructure, NOT product code) for the Q-LiDAR quantized sparse-3D-conv path.

This file is a CPU restatement (numpy for the integer/index work, torch-CPU for the matmuls) of the
algorithms on the reference's hot path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
(`quantization-on-3d-object-detection_b200/qlidar`) never does.

PARITY STATUS -- pinned for the code the reference owns, UNPINNED for the third-party arithmetic under it.
The reference (BiboyQG/Quantization-on-3D-Object-Detection == OpenPCDet v0.6.0 + quant/) ships no tests, golden
vectors or fixtures for this path, and the arithmetic lives in two un-vendored, un-pinned
third-party packages that are not installed in this image (and cannot be built here: spconv needs pccm/cumm codegen):
  * spconv 2.x (traveller59/spconv; `pip install spconv-cu116`, unversioned: docker/cu116.Dockerfile:66;
    observed 2.x with ConvAlgo.MaskImplicitGemm: tools/demo.ipynb:255)
  * pytorch_quantization (NVIDIA TensorRT repo; imported at quant/quant.py:1-2, listed nowhere)
What IS pinned: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the reference's own
sources from /root/reference unmodified (pcdet VoxelResBackBone8x / VoxelBackBone8x / MeanVFE / HeightCompression,
quant/quant.py::QConvNd, quant/quantize.py::q_conv3d / collect_stats / compute_amax) and runs them on CPU with
oracle/ext_stubs.py standing in for the two absent packages.  tests/test_golden.py holds this file's stand-alone
restatement (backbone_specs / backbone_forward / QuantCfg) to those vectors at 1e-5 and indices bit-exact -- topology,
indice_key sharing, the QConvNd permute / fake-quant / restore sequence, BN / ReLU / residual order, no_list handling
and static calibration are therefore the reference's.  What is NOT pinned: the two packages' *published* algorithms
(rulebook, gather-GEMM-scatter, TensorQuantizer arithmetic) are restated here, anchored on the reference's call sites
(cited per function as `file:line` relative to /root/reference), and earn trust only through independent cross-checks
in tests/test_oracle.py: sparse conv vs dense torch.nn.functional.conv3d for every (k, stride, pad) used, int32
path vs float64, torch.round half-to-even, and property tests (permutation invariance, subm preserves the active
set, strided out-set = dilate o subsample).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Configs (constants only; tools/cfgs/dataset_configs/{kitti,waymo}_dataset.yaml,
# tools/cfgs/nuscenes_models/cbgs_voxel0075_res3d_centerpoint.yaml -- SURVEY.md §8 table)
# ----------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (point_cloud_range xyzxyz, voxel_size xyz, num point features, max pts/voxel, max voxels @test)
    "kitti": dict(pc_range=[0.0, -40.0, -3.0, 70.4, 40.0, 1.0], voxel_size=[0.05, 0.05, 0.1], nfeat=4,
                  max_pts=5, max_voxels=40000),        # kitti_dataset.yaml:4,50,65-70
    "waymo": dict(pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15], nfeat=5,
                  max_pts=5, max_voxels=150000),       # waymo_dataset.yaml:5,63,79-84
    "nuscenes": dict(pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], voxel_size=[0.075, 0.075, 0.2], nfeat=5,
                     max_pts=10, max_voxels=160000),   # cbgs_voxel0075_res3d_centerpoint.yaml:6,55-60
}


def grid_size_xyz(pc_range, voxel_size) -> np.ndarray:
    """pcdet/datasets/processor/data_processor.py:135-137: round((max-min)/voxel) as int64, xyz order."""
    r = np.asarray(pc_range, dtype=np.float64)
    g = (r[3:6] - r[0:3]) / np.asarray(voxel_size, dtype=np.float64)
    return np.round(g).astype(np.int64)


def sparse_shape_zyx(grid_xyz) -> List[int]:
    """pcdet/models/backbones_3d/spconv_backbone.py:191: sparse_shape = grid_size[::-1] + [1, 0, 0]."""
    g = [int(v) for v in grid_xyz]
    return [g[2] + 1, g[1], g[0]]


# ----------------------------------------------------------------------------------------------
# a1 / a4 / a4'  Voxelization + mean VFE
# ----------------------------------------------------------------------------------------------
def point_to_cell(points_xyz: np.ndarray, pc_range, voxel_size, grid_xyz):
    """fp32 floor((p - min) / vs) per axis, in-grid mask with exclusive upper bound.

    Follows pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:53-54 (torch fp32 floor + `< grid_size`)
    and [EXT] spconv Point2VoxelCPU3d (same fp32 expression, out-of-range points skipped).
    """
    p = np.asarray(points_xyz, dtype=np.float32)
    mn = np.asarray(pc_range[:3], dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    c = np.floor((p - mn) / vs)                       # fp32 arithmetic (numpy keeps float32)
    g = np.asarray(grid_xyz, dtype=np.float32)
    mask = np.all((c >= 0) & (c < g), axis=1)
    return c.astype(np.int32), mask


def voxelize_hard(points: np.ndarray, pc_range, voxel_size, max_pts: int, max_voxels: int):
    """Hard voxelization of ONE frame, first-touch order, caps applied in point order.

    Restates [EXT] spconv.utils.Point2VoxelCPU3d.point_to_voxel as called from
    pcdet/datasets/processor/data_processor.py:45-61,151-153: points outside the grid are skipped; a new
    voxel is opened for the first point that lands in an empty cell unless `max_voxels` voxels already
    exist (then that point is dropped, later points may still join existing voxels); a voxel keeps its
    first `max_pts` points.  Returns (voxels (V,T,F) zero padded, coords (V,3) zyx int32, num_points (V,) int32).
    """
    points = np.ascontiguousarray(points, dtype=np.float32)
    grid = grid_size_xyz(pc_range, voxel_size)
    cell, mask = point_to_cell(points[:, :3], pc_range, voxel_size, grid)
    idx = np.nonzero(mask)[0]
    cell = cell[idx].astype(np.int64)
    key = (cell[:, 2] * grid[1] + cell[:, 1]) * grid[0] + cell[:, 0]      # z,y,x linearisation
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # voxels in first-touch order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inv]                                   # voxel id per (in-range) point
    keep_v = vid < max_voxels
    V = int(min(order.size, max_voxels))
    # position of each point inside its voxel, in point order
    srt = np.argsort(vid, kind="stable")
    vs_sorted = vid[srt]
    start = np.searchsorted(vs_sorted, np.arange(order.size))
    pos = np.empty(idx.size, dtype=np.int64)
    pos[srt] = np.arange(idx.size) - start[vs_sorted]
    keep = keep_v & (pos < max_pts)
    F_ = points.shape[1]
    voxels = np.zeros((V, max_pts, F_), dtype=np.float32)
    voxels[vid[keep], pos[keep]] = points[idx[keep]]
    num = np.zeros(V, dtype=np.int32)
    np.add.at(num, vid[keep], 1)
    kz = uniq[order][:V]
    coords = np.stack([kz // (grid[1] * grid[0]), (kz // grid[0]) % grid[1], kz % grid[0]], axis=1).astype(np.int32)
    return voxels, coords, num


def mean_vfe(voxels: np.ndarray, num_points: np.ndarray, sequential: bool = False) -> np.ndarray:
    """pcdet/models/backbones_3d/vfe/mean_vfe.py:25-29: sum over T / clamp_min(num_points, 1).
    The reference's summation order is whatever torch.sum(dim=1) does on its device (a cascade on the CPU: it differs from a
    left-to-right sum in the last bit of ~5 % of the means).  sequential=True adds the T slots left to right in fp32 -- the order
    of the CUDA voxeliser (csrc/voxelize.cu k_vox_finalize) -- for the kernel-numerics mirror, whose int8 codes must match bit for bit."""
    if sequential:
        s = np.zeros((voxels.shape[0], voxels.shape[2]), dtype=np.float32)
        for t in range(voxels.shape[1]):
            s = s + voxels[:, t]
        n = np.maximum(num_points.astype(np.float32), np.float32(1.0)).reshape(-1, 1)
        return (s / n).astype(np.float32)
    s = torch.from_numpy(voxels).sum(dim=1)
    n = torch.clamp_min(torch.from_numpy(num_points).view(-1, 1).float(), 1.0)
    return (s / n).numpy()


def collate_voxels(per_frame):
    """pcdet/datasets/dataset.py:232-244: concat voxels/num_points, prepend batch index to coords."""
    feats, coords, nums = [], [], []
    for b, (v, c, n) in enumerate(per_frame):
        feats.append(v)
        nums.append(n)
        coords.append(np.concatenate([np.full((c.shape[0], 1), b, dtype=np.int32), c], axis=1))
    return np.concatenate(feats), np.concatenate(coords), np.concatenate(nums)


def voxelize_mean_batch(points_b: np.ndarray, pc_range, voxel_size, max_pts: int, max_voxels: int, sequential: bool = False):
    """Fused a1+a2+a4 on a collated `points (sum P, 1+F)` array (batch index in column 0, frames
    contiguous): per-frame hard voxelization, mean VFE, `[b,z,y,x]` coords.  `max_voxels` is per frame.
    This is the semantics of the GPU op `ql_voxelize_mean` (first-touch order over the whole array)."""
    out = []
    bidx = points_b[:, 0].astype(np.int64)
    B = int(bidx.max()) + 1 if points_b.shape[0] else 0
    for b in range(B):
        out.append(voxelize_hard(points_b[bidx == b, 1:], pc_range, voxel_size, max_pts, max_voxels))
    voxels, coords, nums = collate_voxels(out)
    return mean_vfe(voxels, nums, sequential), coords, nums


def voxelize_dynamic_mean(points_b: np.ndarray, pc_range, voxel_size):
    """pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:53-72: no caps, mean over all points of a voxel,
    output sorted by key b*XYZ + x*YZ + y*Z + z, coords as [b,z,y,x]."""
    grid = grid_size_xyz(pc_range, voxel_size)
    cell, mask = point_to_cell(points_b[:, 1:4], pc_range, voxel_size, grid)
    pts = points_b[mask]
    cell = cell[mask].astype(np.int64)
    sxyz, syz, sz = grid[0] * grid[1] * grid[2], grid[1] * grid[2], grid[2]
    key = pts[:, 0].astype(np.int64) * sxyz + cell[:, 0] * syz + cell[:, 1] * sz + cell[:, 2]
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    sums = np.zeros((uniq.size, pts.shape[1] - 1), dtype=np.float64)
    np.add.at(sums, inv, pts[:, 1:].astype(np.float64))
    mean = (sums / cnt[:, None]).astype(np.float32)
    coords = np.stack([uniq // sxyz, uniq % sz, (uniq % syz) // sz, (uniq % sxyz) // syz], axis=1).astype(np.int32)
    return mean, coords, cnt.astype(np.int32)


# ----------------------------------------------------------------------------------------------
# a7  Rulebook (indice pairs)
# ----------------------------------------------------------------------------------------------
def _triple(v) -> Tuple[int, int, int]:
    if isinstance(v, (int, np.integer)):
        return (int(v),) * 3
    v = tuple(int(x) for x in v)
    assert len(v) == 3
    return v


def conv_out_shape(in_shape, ksize, stride, pad) -> List[int]:
    """[EXT] spconv output-shape rule per dim: (in + 2*pad - k)//stride + 1 (dilation 1); consistent with
    the shape comments at spconv_backbone.py:206,213,220,229."""
    k, s, p = _triple(ksize), _triple(stride), _triple(pad)
    return [(int(in_shape[d]) + 2 * p[d] - k[d]) // s[d] + 1 for d in range(3)]


def _lin(coords: np.ndarray, shape) -> np.ndarray:
    c = coords.astype(np.int64)
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def _lookup(sorted_keys, sorted_idx, q):
    pos = np.searchsorted(sorted_keys, q)
    pos_c = np.minimum(pos, sorted_keys.size - 1)
    hit = sorted_keys[pos_c] == q
    return np.where(hit, sorted_idx[pos_c], -1)


def kernel_offsets(ksize) -> np.ndarray:
    """Offsets enumerated as k = (kz*KH + ky)*KW + kx, the order of the weight tensor
    (C_out, kd, kh, kw, C_in) (spconv-2 layout, quant/quant.py:37-39, detector3d_template.py:346)."""
    k = _triple(ksize)
    kz, ky, kx = np.meshgrid(np.arange(k[0]), np.arange(k[1]), np.arange(k[2]), indexing="ij")
    return np.stack([kz.ravel(), ky.ravel(), kx.ravel()], axis=1).astype(np.int64)


def rulebook_subm(coords: np.ndarray, spatial_shape, ksize) -> np.ndarray:
    """Submanifold rulebook ([EXT] spconv SubMConv: stride 1, pad = k//2, outputs == inputs, same order).
    Returns nbr (K, N) int32: nbr[k, o] = input row at coord(o) + offset_k - k//2, or -1.
    Cross-correlation orientation (== nn.Conv3d with W.permute(0,4,1,2,3)); pinned by the dense cross-check."""
    N = coords.shape[0]
    k = _triple(ksize)
    offs = kernel_offsets(k)
    nbr = np.full((offs.shape[0], N), -1, dtype=np.int32)
    if N == 0:
        return nbr
    keys = _lin(coords, spatial_shape)
    order = np.argsort(keys, kind="stable")
    sk, si = keys[order], order.astype(np.int64)
    c = coords.astype(np.int64)
    shp = np.asarray(spatial_shape, dtype=np.int64)
    for ki, (dz, dy, dx) in enumerate(offs):
        q = c.copy()
        q[:, 1] += dz - k[0] // 2
        q[:, 2] += dy - k[1] // 2
        q[:, 3] += dx - k[2] // 2
        ok = np.all((q[:, 1:] >= 0) & (q[:, 1:] < shp), axis=1)
        res = _lookup(sk, si, _lin(q, spatial_shape))
        nbr[ki] = np.where(ok, res, -1)
    return nbr


def rulebook_strided(coords: np.ndarray, spatial_shape, ksize, stride, pad):
    """Regular (strided) sparse conv rulebook ([EXT] spconv SparseConv3d): an output site is active iff its
    window contains >= 1 active input.  Output ORDER is implementation-defined in spconv; this framework
    fixes it to ascending linear key ((b*Do + z)*Ho + y)*Wo + x (== np.unique of the candidate keys).
    Returns (out_coords (M,4) int32, out_shape, nbr (K, M) int32) with nbr[k,o] = row of input at
    o*stride - pad + offset_k."""
    k, s, p = _triple(ksize), _triple(stride), _triple(pad)
    out_shape = conv_out_shape(spatial_shape, k, s, p)
    offs = kernel_offsets(k)
    K = offs.shape[0]
    N = coords.shape[0]
    c = coords.astype(np.int64)
    osh = np.asarray(out_shape, dtype=np.int64)
    cand_keys = []
    for ki, off in enumerate(offs):
        num = c[:, 1:] + np.asarray(p) - off
        o = num // np.asarray(s)
        ok = np.all((num % np.asarray(s) == 0) & (o >= 0) & (o < osh), axis=1)
        oc = np.concatenate([c[ok, :1], o[ok]], axis=1)
        cand_keys.append(_lin(oc, out_shape))
    cand_keys = np.concatenate(cand_keys) if N else np.zeros(0, np.int64)
    okeys = np.unique(cand_keys)
    M = okeys.size
    W_, H_, D_ = out_shape[2], out_shape[1], out_shape[0]
    out_coords = np.stack([okeys // (D_ * H_ * W_), (okeys // (H_ * W_)) % D_, (okeys // W_) % H_, okeys % W_],
                          axis=1).astype(np.int32)
    nbr = np.full((K, M), -1, dtype=np.int32)
    if N and M:
        ikeys = _lin(coords, spatial_shape)
        iorder = np.argsort(ikeys, kind="stable")
        sk, si = ikeys[iorder], iorder.astype(np.int64)
        ish = np.asarray(spatial_shape, dtype=np.int64)
        oc64 = out_coords.astype(np.int64)
        for ki, off in enumerate(offs):
            q = oc64.copy()
            q[:, 1:] = oc64[:, 1:] * np.asarray(s) - np.asarray(p) + off
            ok = np.all((q[:, 1:] >= 0) & (q[:, 1:] < ish), axis=1)
            res = _lookup(sk, si, _lin(np.where(ok[:, None], q, 0), spatial_shape))
            nbr[ki] = np.where(ok, res, -1)
    return out_coords, out_shape, nbr


def tile_kmask(nbr: np.ndarray, tile_m: int = 128) -> np.ndarray:
    """Per-tile offset mask of a (K, N) rulebook: uint32 [tiles, ceil(K/32)], bit k set iff any row of the tile has a
    neighbour through offset k (what the rulebook kernels emit for the conv kernel's slab skipping)."""
    K, N = nbr.shape
    tiles = (N + tile_m - 1) // tile_m
    out = np.zeros((tiles, (K + 31) // 32), dtype=np.uint32)
    for t in range(tiles):
        any_k = (nbr[:, t * tile_m:(t + 1) * tile_m] >= 0).any(axis=1)
        for k in np.nonzero(any_k)[0]:
            out[t, k >> 5] |= np.uint32(1 << (k & 31))
    return out


def pairs_in_coord_space(nbr: np.ndarray, in_coords: np.ndarray, out_coords: np.ndarray) -> np.ndarray:
    """(k, in_coord, out_coord) rows sorted lexicographically -- the order-free form in which indice pairs
    are compared (SURVEY.md §8d parity gates)."""
    k, o = np.nonzero(nbr >= 0)
    i = nbr[k, o]
    rows = np.concatenate([k[:, None].astype(np.int64), in_coords[i].astype(np.int64), out_coords[o].astype(np.int64)], axis=1)
    if rows.shape[0] == 0:
        return rows
    return rows[np.lexsort(rows.T[::-1])]


# ----------------------------------------------------------------------------------------------
# a9 / §8a-Q  TensorQuantizer ([EXT] pytorch_quantization) semantics
# ----------------------------------------------------------------------------------------------
def quant_bound(bits: int) -> float:
    return float(2 ** (bits - 1) - 1)                 # narrow_range symmetric: 127 / 32767


def dynamic_amax(t: torch.Tensor, axis=None) -> torch.Tensor:
    """max|t| over every dim NOT in `axis`, keepdim ([EXT] TensorQuantizer dynamic amax; descriptors built at
    quant/quant.py:14-32: weights axis=(0) on (oc, ic*K); activations axis=(1) if cw else None)."""
    a = t.detach().abs()
    if axis is None:
        return a.max() if a.numel() else a.new_zeros(())
    axes = (axis,) if isinstance(axis, int) else tuple(axis)
    red = [d for d in range(t.dim()) if d not in axes]
    return a.amax(dim=red, keepdim=True) if red else a


def quant_scale(amax: torch.Tensor, bits: int) -> torch.Tensor:
    """scale = bound / amax, with amax <= 2^-24 -> scale 0 ([EXT] fake_tensor_quant epsilon rule)."""
    bound = quant_bound(bits)
    amax = amax.to(torch.float32)
    tiny = amax <= (1.0 / (1 << 24))
    # ONE correctly rounded division per element, like the library (`max_bound / amax`, both tensors).  NOT `bound / tensor`:
    # Python evaluates that as tensor.reciprocal() * bound -- two roundings, an ulp off in a quarter of the cases.
    return torch.where(tiny, torch.zeros_like(amax), torch.full_like(amax, bound) / torch.where(tiny, torch.ones_like(amax), amax))


def quantize_codes(t: torch.Tensor, amax: torch.Tensor, bits: int) -> torch.Tensor:
    """q = clamp(round_half_even(t * scale), -bound, +bound) as int32 (torch.round == rint)."""
    bound = quant_bound(bits)
    q = torch.round(t.to(torch.float32) * quant_scale(amax, bits)).clamp_(-bound, bound)
    return q.to(torch.int32)


def fake_quant(t: torch.Tensor, bits: int, axis=None, amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """t_fq = q / scale (scale==0 -> 0): quantize -> round -> clamp -> DEquantize, still fp32.  This is what
    every reference wrapper feeds to the unmodified fp32 spconv layer (quant/quant.py:44,51)."""
    if amax is None:
        amax = dynamic_amax(t, axis)
    sc = quant_scale(amax, bits)
    q = quantize_codes(t, amax, bits).to(torch.float32)
    return torch.where(sc == 0, torch.zeros_like(q), q / torch.where(sc == 0, torch.ones_like(sc), sc))


def weight_matrix(weight: torch.Tensor) -> torch.Tensor:
    """(oc, kd, kh, kw, ic) -> (oc, ic*K) exactly as quant/quant.py:37-43 (permute(0,4,1,2,3).view(oc,-1))."""
    oc = weight.shape[0]
    dim = weight.dim()
    return weight.permute([0, dim - 1] + list(range(1, dim - 1))).contiguous().view(oc, -1)


def weight_from_matrix(wm: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Inverse of weight_matrix (quant/quant.py:45-48)."""
    oc, ic = like.shape[0], like.shape[-1]
    kdim = tuple(like.shape[1:-1])
    dim = like.dim()
    return wm.view((oc, ic) + kdim).permute([0] + list(range(2, dim)) + [1]).contiguous()


def quantize_weight_per_oc(weight: torch.Tensor, bits: int = 8):
    """Per-output-channel weight codes + amax (QuantDescriptor(num_bits=w_bits, axis=(0)), quant/quant.py:14-18)."""
    wm = weight_matrix(weight)
    amax = dynamic_amax(wm, axis=0)                   # (oc,1)
    q = quantize_codes(wm, amax, bits)
    return weight_from_matrix(q, weight), amax.view(-1)


# ----------------------------------------------------------------------------------------------
# a8  Sparse convolution (spconv "Native" algorithm: per-offset gather -> GEMM -> scatter-add)
# ----------------------------------------------------------------------------------------------
def sparse_conv(features: torch.Tensor, nbr: np.ndarray, weight: torch.Tensor,
                bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[o,:] = sum_k W[:,k,:] @ x[nbr[k,o],:] (+bias); weight (oc, kd,kh,kw, ic).  fp32 (or fp64)."""
    oc, ic = weight.shape[0], weight.shape[-1]
    K = nbr.shape[0]
    M = nbr.shape[1]
    w = weight.reshape(oc, K, ic)
    out = torch.zeros((M, oc), dtype=features.dtype)
    nbr_t = torch.from_numpy(nbr.astype(np.int64))
    for k in range(K):
        o = torch.nonzero(nbr_t[k] >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        g = features.index_select(0, nbr_t[k][o])
        out.index_add_(0, o, g @ w[:, k, :].to(features.dtype).t())
    if bias is not None:
        out += bias.to(out.dtype)
    return out


def sparse_conv_im2col(features: torch.Tensor, nbr: np.ndarray, weight: torch.Tensor,
                       bias: Optional[torch.Tensor] = None, chunk: int = 32768) -> torch.Tensor:
    """Same result as sparse_conv (up to fp32 summation order) computed as gather -> [M, K*C_in] @ [K*C_in, C_out]; faster
    on CPU for small channel counts.  Only used to time the CPU baseline (bench.py); tests check it against sparse_conv."""
    oc, ic = weight.shape[0], weight.shape[-1]
    K, M = nbr.shape
    wm = weight.reshape(oc, K * ic).to(features.dtype).t().contiguous()
    fz = torch.cat([features, features.new_zeros((1, ic))])
    idx = torch.from_numpy(np.ascontiguousarray(nbr.T).astype(np.int64))
    idx = torch.where(idx < 0, torch.full_like(idx, features.shape[0]), idx)
    out = torch.empty((M, oc), dtype=features.dtype)
    for s in range(0, M, chunk):
        out[s:s + chunk] = fz[idx[s:s + chunk].reshape(-1)].view(-1, K * ic) @ wm
    if bias is not None:
        out += bias.to(out.dtype)
    return out


def sparse_conv_auto(features, nbr, weight, bias=None):
    """Fastest CPU variant per layer shape (measured on 8 cores): im2col GEMM up to 64 channels, per-offset above."""
    return (sparse_conv_im2col if weight.shape[-1] <= 64 else sparse_conv)(features, nbr, weight, bias)


def sparse_conv_int(qx: torch.Tensor, nbr: np.ndarray, qw: torch.Tensor) -> torch.Tensor:
    """INT8 x INT8 -> INT32 accumulators, exact (int64 matmul on CPU, result fits int32)."""
    oc, ic = qw.shape[0], qw.shape[-1]
    K, M = nbr.shape
    w = qw.reshape(oc, K, ic).to(torch.int64)
    out = torch.zeros((M, oc), dtype=torch.int64)
    nbr_t = torch.from_numpy(nbr.astype(np.int64))
    x64 = qx.to(torch.int64)
    for k in range(K):
        o = torch.nonzero(nbr_t[k] >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        g = x64.index_select(0, nbr_t[k][o])
        out.index_add_(0, o, g @ w[:, k, :].t())
    assert out.abs().max() < 2 ** 31 if out.numel() else True
    return out.to(torch.int32)


def sparse_conv_int_f64(qx: torch.Tensor, nbr: np.ndarray, qw: torch.Tensor) -> torch.Tensor:
    """sparse_conv_int through float64 BLAS: int8 x int8 products and their sums (|acc| < 2^31) are exact in float64, and dgemm is
    ~100x faster than torch's int64 matmul on the CPU -- what makes a full-size (146 k voxels x 4 frames) mirror run affordable.
    tests/test_oracle.py holds it to sparse_conv_int bit for bit."""
    oc, ic = qw.shape[0], qw.shape[-1]
    K, M = nbr.shape
    w = qw.reshape(oc, K, ic).to(torch.float64)
    out = torch.zeros((M, oc), dtype=torch.float64)
    nbr_t = torch.from_numpy(nbr.astype(np.int64))
    x64 = qx.to(torch.float64)
    for k in range(K):
        o = torch.nonzero(nbr_t[k] >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        out.index_add_(0, o, x64.index_select(0, nbr_t[k][o]) @ w[:, k, :].t())
    assert out.abs().max() < 2 ** 31 if out.numel() else True
    return out.to(torch.int32)


def bn_fold(gamma, beta, mean, var, eps):
    """Eval-mode BatchNorm1d as y = a*x + b (spconv_backbone.py:189 eps=1e-3; read eps from the module)."""
    a = gamma / torch.sqrt(var + eps)
    return a, beta - a * mean


# ----------------------------------------------------------------------------------------------
# Quantized conv modes
# ----------------------------------------------------------------------------------------------
def qconv_reference_math(features, nbr, weight, bias, w_bits, act_bits, cw, act_amax=None, conv_fn=None):
    """QConvNd.forward exactly (quant/quant.py:36-58): fake-quant(w) per oc, fake-quant(x) per tensor
    (cw=False) or per input channel (cw=True, axis=1), fp32 sparse conv, +bias."""
    wq = weight_from_matrix(fake_quant(weight_matrix(weight), w_bits, axis=0), weight)
    xq = fake_quant(features, act_bits, axis=1 if cw else None, amax=act_amax)
    return (conv_fn or sparse_conv)(xq, nbr, wq, bias)


def qconv_w8a8_pt(features, nbr, weight, bias, act_amax=None):
    """W8A8 per-tensor: int8 codes, INT32 accumulate, dequant with (amax_x/127)*(amax_w[oc]/127).
    Returns (acc int32, y fp32, amax_x, amax_w)."""
    qw, amax_w = quantize_weight_per_oc(weight, 8)
    amax_x = dynamic_amax(features) if act_amax is None else act_amax
    qx = quantize_codes(features, amax_x, 8)
    acc = sparse_conv_int(qx, nbr, qw)
    y = acc.to(torch.float32) * ((amax_x.to(torch.float32) / 127.0) * (amax_w / 127.0)).view(1, -1)
    if bias is not None:
        y = y + bias
    return acc, y, amax_x, amax_w


def smoothquant_scale(act_amax_ic: torch.Tensor, weight: torch.Tensor, alpha: float) -> torch.Tensor:
    """s[ic] = amax_x[ic]^alpha / amax_w[ic]^(1-alpha), zeros -> 1 (quant/smoothquant.py:69-76 formula,
    per input channel: in a dense unfold every voxel appears at every kernel position)."""
    w_ic = weight.abs().amax(dim=tuple(range(weight.dim() - 1)))           # max over oc,k
    s = act_amax_ic.view(-1) ** alpha / w_ic ** (1.0 - alpha)
    s = torch.where((s == 0) | ~torch.isfinite(s), torch.ones_like(s), s)
    return s


def qconv_w8a8_sq(features, nbr, weight, bias, alpha=0.5):
    """SmoothQuant: x' = x/s, w' = w*s, then W8A8 per-tensor on (x', w')."""
    s = smoothquant_scale(dynamic_amax(features, axis=1), weight, alpha)
    return qconv_w8a8_pt(features / s.view(1, -1), nbr, weight * s.view(*([1] * (weight.dim() - 1)), -1), bias)


# ----------------------------------------------------------------------------------------------
# a12  dense() / HeightCompression
# ----------------------------------------------------------------------------------------------
def to_dense(features: torch.Tensor, coords: np.ndarray, spatial_shape, batch_size: int) -> torch.Tensor:
    """[EXT] SparseConvTensor.dense(): scatter rows into zeros (B, D,H,W, C) -> (B, C, D,H,W)."""
    D, H, W = [int(v) for v in spatial_shape]
    C = features.shape[1]
    out = torch.zeros((batch_size, D, H, W, C), dtype=features.dtype)
    c = torch.from_numpy(coords.astype(np.int64))
    out[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = features
    return out.permute(0, 4, 1, 2, 3).contiguous()


def height_compression(features, coords, spatial_shape, batch_size):
    """pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-24: dense().view(N, C*D, H, W)."""
    d = to_dense(features, coords, spatial_shape, batch_size)
    N, C, D, H, W = d.shape
    return d.view(N, C * D, H, W)


def bev_merge2d(features: torch.Tensor, coords: np.ndarray):
    """VoxelResBackBone8xVoxelNeXt.bev_out (spconv_backbone_voxelnext.py:149-164): drop z, torch.unique(dim=0)
    (lexicographically sorted [b,y,x]) and index_add_ duplicates."""
    ind = torch.from_numpy(coords[:, [0, 2, 3]].astype(np.int64))
    uniq, inv = torch.unique(ind, dim=0, return_inverse=True)
    out = features.new_zeros((uniq.shape[0], features.shape[1]))
    out.index_add_(0, inv, features)
    return out, uniq.numpy().astype(np.int32)


# ----------------------------------------------------------------------------------------------
# Sparse tensor + network restatement (a5, a6, a13)
# ----------------------------------------------------------------------------------------------
@dataclass
class SpT:
    features: torch.Tensor
    coords: np.ndarray                      # (N,4) int32 [b,z,y,x]
    spatial_shape: List[int]
    batch_size: int
    rulebooks: Dict[str, tuple] = field(default_factory=dict)

    def replace(self, f):
        return SpT(f, self.coords, self.spatial_shape, self.batch_size, self.rulebooks)


@dataclass
class ConvSpec:
    name: str
    cin: int
    cout: int
    ksize: Tuple[int, int, int]
    stride: Tuple[int, int, int]
    pad: Tuple[int, int, int]
    subm: bool
    key: str
    bias: bool
    bn_eps: float = 1e-3


@dataclass
class QuantCfg:
    """mode: 'fp32' | 'ref' (reference fake-quant math, QConvNd) | 'w8a8_pt' | 'w8a8_sq' ;
    w_bits/act_bits/cw as in quant/quant.py:8; no_list = dotted conv names left un-quantized
    (quant/quant_centerpoint.py:24-26)."""
    mode: str = "fp32"
    w_bits: int = 8
    act_bits: int = 8
    cw: bool = False
    alpha: float = 0.5
    no_list: Tuple[str, ...] = ()
    fast: bool = False            # CPU-baseline timing: pick the faster of the two equivalent conv formulations
    act_amax: Optional[Dict[str, torch.Tensor]] = None   # static calibration: conv name -> calibrated activation amax
                                                         # (quant/quantize.py:175-207); None/missing = dynamic


def run_conv(x: SpT, spec: ConvSpec, params: Dict[str, torch.Tensor], q: QuantCfg, record=None) -> SpT:
    """One (optionally quantized) sparse conv incl. rulebook caching per indice_key ([EXT] spconv indice_dict)."""
    if spec.key in x.rulebooks:
        out_coords, out_shape, nbr = x.rulebooks[spec.key]
    else:
        if spec.subm:
            out_coords, out_shape, nbr = x.coords, x.spatial_shape, rulebook_subm(x.coords, x.spatial_shape, spec.ksize)
        else:
            out_coords, out_shape, nbr = rulebook_strided(x.coords, x.spatial_shape, spec.ksize, spec.stride, spec.pad)
        x.rulebooks[spec.key] = (out_coords, out_shape, nbr)
    w = params[spec.name + ".weight"]
    b = params.get(spec.name + ".bias") if spec.bias else None
    mode = "fp32" if spec.name in q.no_list else q.mode
    conv_fn = sparse_conv_auto if q.fast else sparse_conv
    if mode == "fp32":
        y = conv_fn(x.features, nbr, w, b)
    elif mode == "ref":
        am = q.act_amax.get(spec.name) if q.act_amax else None
        y = qconv_reference_math(x.features, nbr, w, b, q.w_bits, q.act_bits, q.cw, act_amax=am, conv_fn=conv_fn)
    elif mode == "w8a8_pt":
        acc, y, _, _ = qconv_w8a8_pt(x.features, nbr, w, b)
        if record is not None:
            record[spec.name + ".acc"] = acc
    elif mode == "w8a8_sq":
        acc, y, _, _ = qconv_w8a8_sq(x.features, nbr, w, b, q.alpha)
    else:
        raise ValueError(mode)
    out = SpT(y, out_coords, list(out_shape), x.batch_size, x.rulebooks if spec.subm else {})
    if record is not None:
        record[spec.name] = y
        record[spec.name + ".in"] = (x.features, x.coords, list(x.spatial_shape), out_coords)
    return out


def bn_relu(x: SpT, name: str, params, eps: float, relu=True, residual: Optional[torch.Tensor] = None) -> SpT:
    """nn.BatchNorm1d in eval mode applied to .features by SparseSequential / replace_feature (spconv_backbone.py:16-17,
    56-65): torch's own batch_norm on the running statistics, so the fp32 rounding is the reference's, not a folded a*x+b."""
    y = F.batch_norm(x.features, params[name + ".running_mean"], params[name + ".running_var"], params[name + ".weight"],
                     params[name + ".bias"], False, 0.0, eps)
    if residual is not None:
        y = y + residual
    if relu:
        y = torch.relu(y)
    return x.replace(y)


def backbone_specs(arch: str, input_channels: int, channels=None, kernel_sizes=None, out_channel=None) -> List[dict]:
    """Layer list of VoxelBackBone8x (spconv_backbone.py:78-118), VoxelResBackBone8x (:193-234) and
    VoxelResBackBone8xVoxelNeXt (spconv_backbone_voxelnext.py:81-138) as a flat program of ops."""
    prog: List[dict] = []

    def conv(name, cin, cout, k, s, p, subm, key, bias):
        return ConvSpec(name, cin, cout, _triple(k), _triple(s), _triple(p), subm, key, bias)

    def post_act(prefix, cin, cout, k, s, p, subm, key):
        prog.append(dict(op="conv_bn_relu", conv=conv(prefix + ".0", cin, cout, k, s, p, subm, key, False), bn=prefix + ".1"))

    def basic(prefix, c, key, bias=True):
        prog.append(dict(op="basic_block", conv1=conv(prefix + ".conv1", c, c, 3, 1, 1, True, key, bias), bn1=prefix + ".bn1",
                         conv2=conv(prefix + ".conv2", c, c, 3, 1, 1, True, key, bias), bn2=prefix + ".bn2"))

    if arch == "VoxelBackBone8x":
        post_act("conv_input", input_channels, 16, 3, 1, 1, True, "subm1")
        post_act("conv1.0", 16, 16, 3, 1, 1, True, "subm1")
        prog.append(dict(op="tap", name="x_conv1"))
        for si, (ci, co, pad) in enumerate([(16, 32, 1), (32, 64, 1), (64, 64, (0, 1, 1))], start=2):
            post_act(f"conv{si}.0", ci, co, 3, 2, pad, False, f"spconv{si}")
            post_act(f"conv{si}.1", co, co, 3, 1, 1, True, f"subm{si}")
            post_act(f"conv{si}.2", co, co, 3, 1, 1, True, f"subm{si}")
            prog.append(dict(op="tap", name=f"x_conv{si}"))
        post_act("conv_out", 64, 128, (3, 1, 1), (2, 1, 1), 0, False, "spconv_down2")
    elif arch == "VoxelResBackBone8x":
        post_act("conv_input", input_channels, 16, 3, 1, 1, True, "subm1")
        basic("conv1.0", 16, "res1")
        basic("conv1.1", 16, "res1")
        prog.append(dict(op="tap", name="x_conv1"))
        for si, (ci, co, pad) in enumerate([(16, 32, 1), (32, 64, 1), (64, 128, (0, 1, 1))], start=2):
            post_act(f"conv{si}.0", ci, co, 3, 2, pad, False, f"spconv{si}")
            basic(f"conv{si}.1", co, f"res{si}")
            basic(f"conv{si}.2", co, f"res{si}")
            prog.append(dict(op="tap", name=f"x_conv{si}"))
        post_act("conv_out", 128, 128, (3, 1, 1), (2, 1, 1), 0, False, "spconv_down2")
    elif arch == "VoxelResBackBone8xVoxelNeXt":
        ch = list(channels or [16, 32, 64, 128, 128])
        ks = list(kernel_sizes or [3, 3, 3, 3])
        oc = int(out_channel or 128)
        post_act("conv_input", input_channels, ch[0], 3, 1, 1, True, "subm1")
        basic("conv1.0", ch[0], "res1")
        basic("conv1.1", ch[0], "res1")
        prog.append(dict(op="tap", name="x_conv1"))
        stage = [(ch[0], ch[1], ks[0]), (ch[1], ch[2], ks[1]), (ch[2], ch[3], ks[2]), (ch[3], ch[4], ks[3]), (ch[4], ch[4], ks[3])]
        for si, (ci, co, k) in enumerate(stage, start=2):
            post_act(f"conv{si}.0", ci, co, k, 2, k // 2, False, f"spconv{si}")
            basic(f"conv{si}.1", co, f"res{si}")
            basic(f"conv{si}.2", co, f"res{si}")
            prog.append(dict(op="tap", name=f"x_conv{si}"))
        prog.append(dict(op="voxelnext_bev"))
        prog.append(dict(op="conv_bn_relu", conv=conv("conv_out.0", ch[3], oc, (1, 3, 3), 1, (0, 1, 1), False, "spconv_down2", False), bn="conv_out.1"))
        prog.append(dict(op="conv_bn_relu", conv=conv("shared_conv.0", oc, oc, (1, 3, 3), 1, (0, 1, 1), True, "subm_shared", True),
                         bn="shared_conv.1", bn_eps=1e-5))
    else:
        raise ValueError(arch)
    return prog


def all_conv_specs(prog) -> List[ConvSpec]:
    out = []
    for op in prog:
        if op["op"] == "conv_bn_relu":
            out.append(op["conv"])
        elif op["op"] == "basic_block":
            out += [op["conv1"], op["conv2"]]
    return out


def init_params(prog, seed: int = 4) -> Dict[str, torch.Tensor]:
    """Random-init weights per SURVEY.md §8d: W ~ N(0, sqrt(2/(K*C_in))) in (oc,kd,kh,kw,ic); bias N(0,.01);
    BN gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5).  torch.manual_seed(4) is the reference's
    seed (quant/quant_centerpoint.py:174)."""
    g = np.random.default_rng(seed)                    # numpy PCG64: the same parameters on every machine / torch version
    P: Dict[str, torch.Tensor] = {}

    def randn(*shape):
        return torch.from_numpy(g.standard_normal(shape, dtype=np.float32))

    def rand(*shape):
        return torch.from_numpy(g.random(shape, dtype=np.float32))

    def bn(name, c):
        P[name + ".weight"] = rand(c) + 0.5
        P[name + ".bias"] = randn(c) * 0.1
        P[name + ".running_mean"] = randn(c) * 0.1
        P[name + ".running_var"] = rand(c) + 0.5

    for op in prog:
        convs = []
        if op["op"] == "conv_bn_relu":
            convs = [(op["conv"], op["bn"])]
        elif op["op"] == "basic_block":
            convs = [(op["conv1"], op["bn1"]), (op["conv2"], op["bn2"])]
        for spec, bnname in convs:
            K = spec.ksize[0] * spec.ksize[1] * spec.ksize[2]
            std = math.sqrt(2.0 / (K * spec.cin))
            P[spec.name + ".weight"] = randn(spec.cout, *spec.ksize, spec.cin) * std
            if spec.bias:
                P[spec.name + ".bias"] = randn(spec.cout) * 0.01
            bn(bnname, spec.cout)
    return P


def backbone_forward(prog, params, features: torch.Tensor, coords: np.ndarray, sparse_shape, batch_size: int,
                     q: QuantCfg = QuantCfg(), record: Optional[dict] = None):
    """VoxelResBackBone8x.forward (spconv_backbone.py:243-295) / VoxelBackBone8x (:129-181) /
    VoxelResBackBone8xVoxelNeXt.forward (spconv_backbone_voxelnext.py:166-225) on the CPU.
    Returns (encoded SpT, dict of taps)."""
    x = SpT(features, coords.astype(np.int32), list(sparse_shape), batch_size)
    taps: Dict[str, SpT] = {}
    for op in prog:
        kind = op["op"]
        if kind == "conv_bn_relu":
            x = run_conv(x, op["conv"], params, q, record)
            x = bn_relu(x, op["bn"], params, op.get("bn_eps", 1e-3))
        elif kind == "basic_block":
            # SparseBasicBlock.forward (spconv_backbone.py:51-67); identity is the un-quantized block input
            identity = x.features
            out = run_conv(x, op["conv1"], params, q, record)
            out = bn_relu(out, op["bn1"], params, 1e-3)
            out = run_conv(out, op["conv2"], params, q, record)
            x = bn_relu(out, op["bn2"], params, 1e-3, relu=True, residual=identity)
        elif kind == "tap":
            taps[op["name"]] = x
        elif kind == "voxelnext_bev":
            # spconv_backbone_voxelnext.py:194-199: scale stage-5/6 indices onto the stage-4 grid, concat, merge in 2-D
            x4, x5, x6 = taps["x_conv4"], taps["x_conv5"], taps["x_conv6"]
            c5 = x5.coords.copy(); c5[:, 1:] *= 2
            c6 = x6.coords.copy(); c6[:, 1:] *= 4
            feats = torch.cat([x4.features, x5.features, x6.features])
            cc = np.concatenate([x4.coords, c5, c6])
            f2, c2 = bev_merge2d(feats, cc)
            c2 = np.stack([c2[:, 0], np.zeros_like(c2[:, 0]), c2[:, 1], c2[:, 2]], axis=1).astype(np.int32)
            x = SpT(f2, c2, [1] + list(x4.spatial_shape[1:]), batch_size)
        else:
            raise ValueError(kind)
        if record is not None and kind in ("conv_bn_relu", "basic_block"):
            nm = op["conv"].name if kind == "conv_bn_relu" else op["conv2"].name
            record[nm + ".post"] = x.features
    return x, taps


# ----------------------------------------------------------------------------------------------
# Synthetic frame generators ("G1 lidar-like", "G2 surface sheet")
# ----------------------------------------------------------------------------------------------
def synth_lidar_frame(cfg: str, seed: int, n_az: Optional[int] = None, n_beams: Optional[int] = None) -> np.ndarray:
    """G1: sensor at origin above a ground plane, `n_beams` elevation rings x `n_az` azimuth steps; each ray hits
    the ground or one of a set of random vertical cylinders; range noise N(0,0.02). Returns (P, F) float32."""
    rng = np.random.default_rng(seed)
    if cfg == "kitti":
        elev = np.deg2rad(np.linspace(-24.8, 2.0, n_beams or 64)); n_az = n_az or 1400
        sensor_h, n_cyl, ext, max_r, nf = 1.73, 60, 60.0, 80.0, 4
    else:
        elev = np.deg2rad(np.linspace(-17.6, 2.4, n_beams or 192)); n_az = n_az or 2650
        sensor_h, n_cyl, ext, max_r, nf = 2.0, 120, 75.0, 75.0, 5
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    cyl_c = rng.uniform(-ext, ext, size=(n_cyl, 2))
    cyl_r = rng.uniform(0.5, 2.5, size=n_cyl)
    cyl_h = rng.uniform(1.5, 3.5, size=n_cyl)
    A, E = np.meshgrid(az, elev, indexing="ij")
    dx, dy, dz = np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dz < -1e-6, -sensor_h / dz, np.inf)
    t_best = np.minimum(t_ground, max_r * 1.5)
    dxy2 = dx * dx + dy * dy
    for c, r, h in zip(cyl_c, cyl_r, cyl_h):
        # ray-circle intersection in the xy plane
        b = dx * c[0] + dy * c[1]
        cc = c[0] * c[0] + c[1] * c[1] - r * r
        disc = b * b - dxy2 * cc
        with np.errstate(invalid="ignore"):
            t = (b - np.sqrt(np.where(disc >= 0, disc, np.nan))) / dxy2
        z = t * dz                                    # relative to the sensor
        ok = (disc >= 0) & (t > 0.5) & (z >= -sensor_h) & (z <= -sensor_h + h)
        t_best = np.where(ok & (t < t_best), t, t_best)
    t_best = t_best + rng.normal(0.0, 0.02, size=t_best.shape)
    valid = np.isfinite(t_best) & (t_best < max_r) & (t_best > 0.5)
    x, y, z = (t_best * dx)[valid], (t_best * dy)[valid], (t_best * dz)[valid]
    if cfg != "kitti":
        z = z + sensor_h                              # vehicle frame: ground at z = 0 (KITTI stays in the sensor frame)
    if cfg == "kitti":
        keep = (x > 0) & (np.abs(np.arctan2(y, x)) < np.pi / 4)
        x, y, z = x[keep], y[keep], z[keep]
    feats = [x, y, z] + [rng.uniform(0, 1, size=x.shape) for _ in range(nf - 3)]
    pts = np.stack(feats, axis=1).astype(np.float32)
    if cfg != "kitti":                                # Waymo/nuScenes shuffle points at test time (waymo_dataset.yaml:72-76)
        pts = pts[rng.permutation(pts.shape[0])]
    return pts


def synth_batch(cfg: str, batch: int, first_seed: int = 1000, **kw) -> np.ndarray:
    """Collated `points (sum P, 1+F)` with the batch index in column 0 (dataset.py collate of 'points')."""
    fr = [synth_lidar_frame(cfg, first_seed + i, **kw) for i in range(batch)]
    return np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f], axis=1) for i, f in enumerate(fr)])


def synth_surface_sheet(S: int, seed: int = 2000, depth: int = 40) -> np.ndarray:
    """G2: S x S (x,y) patch with z0(x,y) a clipped 2-D random walk in [0,depth): N = S^2 voxels. Returns (N,4) [b,z,y,x]."""
    rng = np.random.default_rng(seed)
    steps_y = rng.integers(-1, 2, size=(S, 1)).cumsum(axis=0)
    steps_x = rng.integers(-1, 2, size=(S, S)).cumsum(axis=1)
    z = np.clip(depth // 2 + steps_y + steps_x, 0, depth - 1)
    yy, xx = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    return np.stack([np.zeros(S * S, np.int64), z.ravel(), yy.ravel(), xx.ravel()], axis=1).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# Kernel-numerics mirror (W8A8 per-tensor): the SAME network as backbone_forward(mode="w8a8_pt"), with the inter-layer
# arithmetic restated the way the device performs it, so that int8 CODES and INT32 accumulators can be compared with the
# CUDA path bit for bit through all 21 layers (a plain fp32 restatement diverges chaotically: one fp16-vs-fp32 ulp flips a
# round-half-even decision, and every flipped code is 0.8 % of amax).  What is mirrored -- and nothing else differs from
# qconv_w8a8_pt + bn_relu:
#   * activations are STORED as fp16 between layers (round-to-nearest-even of the fp32 epilogue value);
#   * de-quantisation + BatchNorm1d (eval, spconv_backbone.py:16-17,56-65) are one folded fp32 FMA per element,
#     y = fmaf(float(acc), (amax_w[oc]/127 * a[oc]) * (amax_x/127), bias[oc]*a[oc] + b[oc]), a = gamma/sqrt(var+eps), b = beta - a*mean,
#     then + fp16 residual (SparseBasicBlock, :64), then ReLU;
#   * dynamic amax (quant/quant.py:28-32, axis=None) is the max of the PRODUCING layer's fp32 values before the fp16 store;
#     static amax (quantize.py:175-207) is the calibrated scalar, and a layer fed by another conv takes its codes from that
#     conv's fp32 epilogue values, rint(y * (127/amax)), instead of re-reading the fp16 rows;
#   * the un-quantised stem (no_list, quant_centerpoint.py:24-26) accumulates in fp32 FMAs over (k, ic) ascending.
# The exact-FMA pieces are C (oracle/qloracle_c.c).
# ----------------------------------------------------------------------------------------------
_QLO = None


def _qlo():
    global _QLO
    if _QLO is None:
        import ctypes
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, "libqloracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-s", "-C", here], check=True)
        lib = ctypes.CDLL(so)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        lib.qlo_stem_conv.argtypes = [vp, i64, i32, vp, i32, i64, vp, i32, vp, vp, i32, vp]
        lib.qlo_stem_conv.restype = None
        lib.qlo_epilogue.argtypes = [vp, i64, i32, vp, vp, vp, i32, vp]
        lib.qlo_epilogue.restype = None
        lib.qlo_rect_iou_matrix.argtypes = [vp, i64, i64, vp]
        lib.qlo_rect_iou_matrix.restype = None
        _QLO = lib
    return _QLO


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float32)


def mirror_codes(x32: np.ndarray, amax: np.float32, bits: int = 8) -> np.ndarray:
    """int8 codes the way the device computes them: scale = fl(bound / amax) (0 if amax <= 2^-24), q = clamp(rint(fl(x * scale)))."""
    bound = np.float32(quant_bound(bits))
    amax = np.float32(amax)
    scale = np.float32(0.0) if amax <= np.float32(1.0 / (1 << 24)) else np.float32(bound / amax)
    q = np.rint(x32.astype(np.float32) * scale)
    return np.clip(q, -bound, bound).astype(np.int8)


def mirror_stem(x: np.ndarray, nbr: np.ndarray, weight: torch.Tensor, scale, shift, relu=True) -> np.ndarray:
    oc, ic = weight.shape[0], weight.shape[-1]
    K, n = nbr.shape
    w_kio = _f32c(weight.reshape(oc, K, ic).permute(1, 2, 0))
    x = _f32c(x)
    nb = np.ascontiguousarray(nbr, dtype=np.int32)
    sc, sh = _f32c(scale), _f32c(shift)
    y = np.empty((n, oc), dtype=np.float32)
    _qlo().qlo_stem_conv(x.ctypes.data, x.shape[1], ic, nb.ctypes.data, K, n, w_kio.ctypes.data, oc, sc.ctypes.data, sh.ctypes.data,
                         1 if relu else 0, y.ctypes.data)
    return y


def mirror_epilogue(acc: np.ndarray, s, shift, residual_h: Optional[np.ndarray], relu=True) -> np.ndarray:
    acc = np.ascontiguousarray(acc, dtype=np.int32)
    n, c = acc.shape
    s, sh = _f32c(s), _f32c(shift)
    res = None if residual_h is None else np.ascontiguousarray(residual_h.astype(np.float32))
    y = np.empty((n, c), dtype=np.float32)
    _qlo().qlo_epilogue(acc.ctypes.data, n, c, s.ctypes.data, sh.ctypes.data, None if res is None else res.ctypes.data, 1 if relu else 0,
                        y.ctypes.data)
    return y


def mirror_backbone_w8a8_pt(prog, params, features, coords: np.ndarray, sparse_shape, batch_size: int,
                            no_list=("conv_input.0",), act_amax: Optional[Dict[str, float]] = None, bits: int = 8, fast: bool = False,
                            keep: bool = True):
    """Returns (record, encoded SpT with fp16-valued features).  record[name] = dict(codes int8 (N_in, C_in) or None for the
    stem, acc int32 or None, y32 fp32 epilogue values, out fp16, out_coords, amax_in).  act_amax: calibrated per-layer scalar
    amax (static); None = dynamic."""
    x = SpT(None, coords.astype(np.int32), list(sparse_shape), batch_size)
    x_h = None                                   # stored activations (fp16) of the current tensor
    y32_prev = None                              # fp32 epilogue values they were rounded from
    prev_was_conv = False                        # the previous layer is a tensor-core conv (its epilogue can emit codes)
    rec: Dict[str, dict] = {}
    feats32 = _f32c(features)
    bound = np.float32(quant_bound(bits))

    def one_conv(x: SpT, spec: ConvSpec, bnname: str, eps: float, residual_h):
        nonlocal x_h, y32_prev, prev_was_conv
        if spec.key in x.rulebooks:
            out_coords, out_shape, nbr = x.rulebooks[spec.key]
        else:
            if spec.subm:
                out_coords, out_shape, nbr = x.coords, x.spatial_shape, rulebook_subm(x.coords, x.spatial_shape, spec.ksize)
            else:
                out_coords, out_shape, nbr = rulebook_strided(x.coords, x.spatial_shape, spec.ksize, spec.stride, spec.pad)
            x.rulebooks[spec.key] = (out_coords, out_shape, nbr)
        w = params[spec.name + ".weight"].float()
        a, b = bn_fold(params[bnname + ".weight"].float(), params[bnname + ".bias"].float(), params[bnname + ".running_mean"].float(),
                       params[bnname + ".running_var"].float(), eps)
        bias = params[spec.name + ".bias"].float() if spec.bias else torch.zeros(spec.cout)
        shift = bias * a + b
        if spec.name in no_list:
            if x_h is not None:
                raise ValueError("only the stem may be left un-quantised in the mirror")
            y32 = mirror_stem(feats32, nbr, w, a, shift)
            r = dict(codes=None, acc=None, amax_in=None)
            prev_conv_now = False
        else:
            static = act_amax is not None and spec.name in act_amax
            if static:
                m = np.float32(act_amax[spec.name])
            else:
                m = np.float32(np.abs(y32_prev).max()) if y32_prev.size else np.float32(0)
            if static and prev_was_conv:
                qscale = np.float32(0.0) if m <= np.float32(1.0 / (1 << 24)) else np.float32(bound / m)
                codes = np.clip(np.rint(y32_prev * qscale), -bound, bound).astype(np.int8)     # the producing epilogue's out_q
            else:
                codes = mirror_codes(x_h.astype(np.float32), m, bits)
            act_scale = np.float32(m / bound)
            qw, amax_w = quantize_weight_per_oc(w, 8)
            acc = (sparse_conv_int_f64 if fast else sparse_conv_int)(torch.from_numpy(codes), nbr, qw).numpy()
            s = _f32c((amax_w / quant_bound(8)) * a) * act_scale
            y32 = mirror_epilogue(acc, s, shift, residual_h)
            r = dict(codes=codes, acc=acc, amax_in=m)
            prev_conv_now = True
        out_h = y32.astype(np.float16)
        r.update(y32=y32, out=out_h, out_coords=out_coords)
        if not keep:                              # full-size runs: keep checksums only
            r = dict(amax_in=r["amax_in"], n_out=out_h.shape[0], codes_sum=None if r["codes"] is None else int(r["codes"].astype(np.int64).sum()),
                     codes_abs_sum=None if r["codes"] is None else int(np.abs(r["codes"].astype(np.int64)).sum()),
                     out_bits_sum=int(out_h.view(np.uint16).astype(np.uint64).sum()))
        rec[spec.name] = r
        x_h, y32_prev, prev_was_conv = out_h, y32, prev_conv_now
        return SpT(None, out_coords, list(out_shape), x.batch_size, x.rulebooks if spec.subm else {})

    taps = {}
    for op in prog:
        kind = op["op"]
        if kind == "conv_bn_relu":
            x = one_conv(x, op["conv"], op["bn"], op.get("bn_eps", 1e-3), None)
        elif kind == "basic_block":
            identity_h = x_h
            x = one_conv(x, op["conv1"], op["bn1"], 1e-3, None)
            x = one_conv(x, op["conv2"], op["bn2"], 1e-3, identity_h)
        elif kind == "tap":
            taps[op["name"]] = (x.coords, x_h)
        else:
            raise ValueError(f"mirror: unsupported op {kind}")
    out = SpT(torch.from_numpy(x_h.astype(np.float32)), x.coords, x.spatial_shape, batch_size)
    return rec, out, taps


# ----------------------------------------------------------------------------------------------
# a15 / boundary: the dense SmoothQuant wrapper family (quant/smoothquant.py) and SQSubM2d (quant/SQSubM2d.py), restated.
# Pinned by tests/golden/sq_dense.npz, which the reference's own smoothquant.py produced (tests/golden/make_golden_sq.py).
# ----------------------------------------------------------------------------------------------
def sq_dense_matmul(cols: torch.Tensor, w2d: torch.Tensor, alpha: float, w_bits: int = 8, act_bits: int = 8) -> torch.Tensor:
    """The core every wrapper shares (quant/smoothquant.py:69-85): cols [M, K] unfolded activations, w2d [N, K];
    scale = max|cols|^a / max|w|^(1-a) per column, zeros -> 1; w * scale, cols / scale; weight fake-quant per row (axis 0), input
    per tensor; cols @ w.T."""
    w_scale = w2d.abs().max(dim=0)[0]
    act_scale = cols.abs().max(dim=0)[0]
    scale = act_scale ** alpha / w_scale ** (1 - alpha)
    scale[scale == 0] = 1
    w = fake_quant(w2d * scale, w_bits, axis=0)
    x = fake_quant(cols / scale, act_bits, axis=None)
    return x @ w.t()


def sq_conv2d(x, weight, bias, alpha, stride=1, padding=0, dilation=1):
    """quant/smoothquant.py:38-99.  x (B, C, H, W), weight (oc, ic, k, k)."""
    oc, ic, k, _ = weight.shape
    B, _, H, W = x.shape
    cols = F.unfold(x, kernel_size=k, dilation=dilation, padding=padding, stride=stride).permute(0, 2, 1).reshape(-1, ic * k * k)
    y = sq_dense_matmul(cols, weight.reshape(oc, -1), alpha)
    ho = (H + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    wo = (W + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    y = y.view(B, -1, oc).permute(0, 2, 1).reshape(B, oc, ho, wo)
    return y if bias is None else y + bias.view(1, oc, 1, 1)



def base_bev_backbone(x, params, cfg, alpha=None, no_list=()):
    """pcdet/models/backbones_2d/base_bev_backbone.py:26-113 (eval mode) with the surgery of quant/quant_centerpoint.py:96-106
    applied when `alpha` is given: every nn.Conv2d whose dotted path is not in no_list runs as quant/smoothquant.py SQConv2d
    (sq_conv2d above); ConvTranspose2d layers stay fp32.  params: the module's state dict (torch tensors), cfg: dict with
    LAYER_NUMS / LAYER_STRIDES / NUM_FILTERS / UPSAMPLE_STRIDES / NUM_UPSAMPLE_FILTERS."""
    def bn_relu(y, pre):
        a = params[pre + ".weight"] / torch.sqrt(params[pre + ".running_var"] + 1e-3)
        b = params[pre + ".bias"] - a * params[pre + ".running_mean"]
        return torch.relu(y * a.view(1, -1, 1, 1) + b.view(1, -1, 1, 1))

    def conv(y, path, stride, pad):
        w = params[path + ".weight"]
        if alpha is not None and path not in no_list:
            return sq_conv2d(y, w, None, alpha, stride, pad)
        return F.conv2d(y, w, None, stride, pad)

    ups = []
    n_up = len(cfg.get("UPSAMPLE_STRIDES") or [])
    for lvl, (n_layers, stride) in enumerate(zip(cfg["LAYER_NUMS"], cfg["LAYER_STRIDES"])):
        pre = "blocks.%d." % lvl
        x = bn_relu(conv(F.pad(x, (1, 1, 1, 1)), pre + "1", stride, 0), pre + "2")         # ZeroPad2d(1) + conv(padding=0)
        for k in range(n_layers):
            x = bn_relu(conv(x, pre + str(4 + 3 * k), 1, 1), pre + str(5 + 3 * k))
        if n_up:
            us = cfg["UPSAMPLE_STRIDES"][lvl]
            dpre = "deblocks.%d." % lvl
            if us >= 1:
                u = F.conv_transpose2d(x, params[dpre + "0.weight"], None, stride=us)
            else:
                ds = int(round(1 / us))
                u = conv(x, dpre + "0", ds, 0)
            ups.append(bn_relu(u, dpre + "1"))
        else:
            ups.append(x)
    return torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]


def sq_conv1d(x, weight, bias, alpha, stride=1, padding=0, dilation=1):
    """quant/smoothquant.py:133-176.  x (B, C, L), weight (oc, ic, k)."""
    oc, ic, k = weight.shape
    B, _, L = x.shape
    cols = F.unfold(x.unsqueeze(2), kernel_size=(1, k), dilation=(1, dilation), padding=(0, padding), stride=(1, stride))
    cols = cols.permute(0, 2, 1).reshape(-1, ic * k)
    y = sq_dense_matmul(cols, weight.reshape(oc, -1), alpha)
    lo = (L + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    y = y.view(B, lo, oc).permute(0, 2, 1)
    return y if bias is None else y + bias.view(1, oc, 1)


def sq_convT2d(x, weight, bias, alpha, stride=1, padding=0, output_padding=0, dilation=1):
    """quant/smoothquant.py:214-270 (with `.reshape` where the file's `.view` raises on inputs of more than one pixel).
    x (B, ic, H, W), weight (ic, oc, k, k)."""
    ic, oc, k, _ = weight.shape
    B, _, ih, iw = x.shape
    w = weight.reshape(ic, -1).t()
    rows = x.reshape(B, ic, ih * iw).permute(0, 2, 1).reshape(-1, ic)
    y = sq_dense_matmul(rows, w, alpha).view(B, ih * iw, -1).permute(0, 2, 1)
    ho = (ih - 1) * stride - 2 * padding + dilation * (k - 1) + output_padding + 1
    wo = (iw - 1) * stride - 2 * padding + dilation * (k - 1) + output_padding + 1
    y = F.fold(y, output_size=(ho, wo), kernel_size=k, dilation=dilation, padding=padding, stride=stride)
    return y if bias is None else y + bias.view(1, oc, 1, 1)


def sq_linear(x, weight, bias, alpha):
    """quant/smoothquant.py:299-322.  x (S, B, in)."""
    S, B, fin = x.shape
    y = sq_dense_matmul(x.reshape(-1, fin), weight, alpha)
    if bias is not None:
        y = y + bias.view(1, -1)
    return y.view(-1, B, weight.shape[0])


def sq_subm2d(x, weight, alpha, kernel_size=3, stride=1, padding=1, dilation=1):
    """quant/SQSubM2d.py:22-91 (the class itself cannot be constructed: NameError in __init__): unfold -> per-column scale ->
    fake-quant -> fold (which SUMS overlapping patches) -> (weight (oc, k, k, ic), x (B, H, W, C))."""
    oc, ic, k, _ = weight.shape
    B, _, H, W = x.shape
    ks = ic * k * k
    cols = torch.transpose(F.unfold(x, kernel_size=k, padding=padding, stride=stride), 1, 2).reshape(-1, ks)
    w_flat = weight.reshape(oc, ks).clone()
    scale = cols.abs().max(dim=0)[0] ** alpha / w_flat.abs().max(dim=0)[0] ** (1 - alpha)
    scale[scale == 0] = 1
    cols = fake_quant(cols / scale, 8, axis=None)
    w_flat = fake_quant(w_flat * scale, 8, axis=0)
    xo = F.fold(cols.reshape(B, -1, ks).transpose(1, 2), (H, W), kernel_size=k, dilation=dilation, padding=padding, stride=stride)
    return w_flat.view(oc, ic, k, k).permute(0, 2, 3, 1).contiguous(), xo.permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------------------------
# CenterHead post-processing: heat-map top-K, decode, range/score mask, rotated NMS.
# numpy fp32, one rounding per reference torch op; the rotated IoU is C (oracle/qloracle_c.c, qlo_rect_iou).
# ----------------------------------------------------------------------------------------------
def sigmoid32(x: np.ndarray) -> np.ndarray:
    """torch.sigmoid, fp32: 1 / (1 + exp(-x)) (center_head.py:307)."""
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def centerhead_topk(scores: np.ndarray, K: int):
    """centernet_utils._topk (centernet_utils.py:155-173): per-class top-K over the H*W cells, then top-K over the C*K survivors.
    Ties (torch.topk leaves their order unspecified) go to the lower index at both levels.  Returns per frame
    (score, cell index, class, y, x), each [B, K]."""
    B, C, H, W = scores.shape
    flat = scores.reshape(B, C, H * W)
    K1 = min(K, H * W)
    o1 = np.argsort(-flat, axis=2, kind="stable")[:, :, :K1]                                   # (B, C, K1) cell indices
    s1 = np.take_along_axis(flat, o1, axis=2)
    s1f, o1f = s1.reshape(B, C * K1), o1.reshape(B, C * K1)
    o2 = np.argsort(-s1f, axis=1, kind="stable")[:, :min(K, C * K1)]
    score = np.take_along_axis(s1f, o2, axis=1)
    cls = (o2 // K1).astype(np.int32)
    ind = np.take_along_axis(o1f, o2, axis=1)
    ys = (ind // W).astype(np.float32)
    xs = (ind % W).astype(np.int32).astype(np.float32)
    return score, ind, cls, ys, xs


def centerhead_decode(hm, center, center_z, dim, rot, vel, iou, K, feature_map_stride, voxel_size, point_cloud_range,
                      post_center_limit_range, score_thresh, class_map=None):
    """CenterHead.generate_predicted_boxes up to the NMS (center_head.py:297-327) -> decode_bbox_from_heatmap
    (centernet_utils.py:176-241).  Inputs are the RAW head outputs (hm logits, dim log-sizes, rot = (cos, sin), iou raw), numpy
    NCHW fp32.  Returns a list (one per frame) of dicts pred_boxes (n, 7|9), pred_scores, pred_labels (class_map applied, 0-based),
    [pred_iou]."""
    f32 = np.float32
    hm = sigmoid32(hm)
    dim = np.exp(np.asarray(dim, f32), dtype=f32)
    B = hm.shape[0]
    score, ind, cls, ys, xs = centerhead_topk(hm, K)
    Kk = score.shape[1]

    def gather(m):                                                       # _transpose_and_gather_feat (centernet_utils.py:148-152)
        m = np.asarray(m, f32)
        c = m.shape[1]
        return np.take_along_axis(m.reshape(B, c, -1), np.broadcast_to(ind[:, None, :], (B, c, Kk)), axis=2).transpose(0, 2, 1)

    ctr, cz, dm, rt = gather(center), gather(center_z), gather(dim), gather(rot)
    angle = np.arctan2(rt[:, :, 1:2], rt[:, :, 0:1]).astype(f32)
    xs = (xs[:, :, None] + ctr[:, :, 0:1]).astype(f32)
    ys = (ys[:, :, None] + ctr[:, :, 1:2]).astype(f32)
    xs = ((xs * f32(feature_map_stride)).astype(f32) * f32(voxel_size[0])).astype(f32) + f32(point_cloud_range[0])
    ys = ((ys * f32(feature_map_stride)).astype(f32) * f32(voxel_size[1])).astype(f32) + f32(point_cloud_range[1])
    parts = [xs.astype(f32), ys.astype(f32), cz, dm, angle]
    if vel is not None:
        parts.append(gather(vel))
    boxes = np.concatenate(parts, axis=-1).astype(f32)
    piou = None
    if iou is not None:
        piou = ((gather(iou)[:, :, 0] + f32(1.0)) * f32(0.5)).astype(f32)
    lim = np.asarray(post_center_limit_range, f32)
    mask = (boxes[..., :3] >= lim[:3]).all(2) & (boxes[..., :3] <= lim[3:]).all(2)
    if score_thresh is not None:
        mask &= score > f32(score_thresh)
    out = []
    for b in range(B):
        m = mask[b]
        lab = cls[b, m]
        if class_map is not None:
            lab = np.asarray(class_map)[lab]
        d = {"pred_boxes": boxes[b, m], "pred_scores": score[b, m], "pred_labels": lab.astype(np.int32)}
        if piou is not None:
            d["pred_iou"] = piou[b, m]
        out.append(d)
    return out


def rect_iou_matrix(boxes: np.ndarray) -> np.ndarray:
    """pairwise rotated BEV IoU, upper triangle (iou_bev, iou3d_nms_kernel.cu:227-234)."""
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    n = b.shape[0]
    out = np.zeros((n, n), np.float32)
    if n:
        _qlo().qlo_rect_iou_matrix(b.ctypes.data, n, b.shape[1], out.ctypes.data)
    return out


def nms_rotated(boxes: np.ndarray, scores: np.ndarray, thresh: float, pre_max=None, post_max=None, iou=None) -> np.ndarray:
    """iou3d_nms_utils.nms_gpu (iou3d_nms_utils.py:120-135) + the sweep of iou3d_nms.cpp:137-183 + class_agnostic_nms's
    [:NMS_POST_MAXSIZE] (model_nms_utils.py:19): sort by score (descending, stable), keep the first pre_max, walk them in order, a
    box is kept unless an earlier KEPT box overlaps it with IoU > thresh.  Returns indices into `boxes`.  `iou` (optional): a
    precomputed pairwise matrix of the SORTED, truncated boxes (e.g. the reference kernel's, for bit-identical decisions)."""
    order = np.argsort(-np.asarray(scores, np.float32), kind="stable")
    if pre_max is not None:
        order = order[:pre_max]
    b = np.asarray(boxes, np.float32)[order][:, :7]
    m = rect_iou_matrix(b) if iou is None else iou
    n = b.shape[0]
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        removed[i + 1:] |= m[i, i + 1:] > np.float32(thresh)
    keep = np.asarray(keep, np.int64)
    if post_max is not None:
        keep = keep[:post_max]
    return order[keep]


def centerhead_generate_predicted_boxes(pred_dicts, class_id_mapping_each_head, K, feature_map_stride, voxel_size, point_cloud_range,
                                        post_center_limit_range, score_thresh, nms_thresh, nms_pre, nms_post, use_vel=False):
    """CenterHead.generate_predicted_boxes with NMS_TYPE nms_gpu (center_head.py:297-365).  pred_dicts: one dict of numpy NCHW maps per
    head.  Returns a list per frame of pred_boxes / pred_scores / pred_labels (1-based)."""
    B = pred_dicts[0]["hm"].shape[0]
    ret = [{"pred_boxes": [], "pred_scores": [], "pred_labels": []} for _ in range(B)]
    for h, pd in enumerate(pred_dicts):
        dec = centerhead_decode(pd["hm"], pd["center"], pd["center_z"], pd["dim"], pd["rot"], pd.get("vel") if use_vel else None, None, K,
                                feature_map_stride, voxel_size, point_cloud_range, post_center_limit_range, score_thresh,
                                class_map=class_id_mapping_each_head[h])
        for b, d in enumerate(dec):
            sel = nms_rotated(d["pred_boxes"], d["pred_scores"], nms_thresh, nms_pre, nms_post) if len(d["pred_scores"]) else np.zeros(0, np.int64)
            ret[b]["pred_boxes"].append(d["pred_boxes"][sel])
            ret[b]["pred_scores"].append(d["pred_scores"][sel])
            ret[b]["pred_labels"].append(d["pred_labels"][sel])
    for b in range(B):
        ret[b] = {"pred_boxes": np.concatenate(ret[b]["pred_boxes"], 0), "pred_scores": np.concatenate(ret[b]["pred_scores"], 0),
                  "pred_labels": np.concatenate(ret[b]["pred_labels"], 0) + 1}
    return ret


# ----------------------------------------------------------------------------------------------
# VoxelNeXt sparse head: per-voxel head outputs -> boxes.
# ----------------------------------------------------------------------------------------------
def voxelhead_decode(hm, center, center_z, dim, rot, vel, iou, indices, batch_size, K, feature_map_stride, voxel_size,
                     point_cloud_range, post_center_limit_range, score_thresh, class_map=None):
    """VoxelNeXtHead.generate_predicted_boxes up to the NMS (voxelnext_head.py:418-456) -> decode_bbox_from_voxels_nuscenes
    (centernet_utils.py:289-354) with _topk_1d's `nuscenes=True` branch (:243-276: per class top-K over the frame's voxels, then top-K
    of the C*K survivors; ties to the lower index) and gather_feat_idx (:278-287).  Inputs are the RAW per-voxel head outputs, numpy
    fp32: hm [N, C] logits, center [N, 2], center_z [N, 1], dim [N, 3] log-sizes, rot [N, 2] = (cos, sin), vel [N, 2] / None, iou
    [N, 1] / None (raw: (iou + 1) / 2 clamped to [0, 1] here), indices [N, 3] = (batch, y, x).  Frames with fewer than K voxels take
    min(K, .) at both levels (the reference's torch.stack would raise on ragged frames; its `topk_ind // K` class rule is kept for full
    frames and restated as `// K1` otherwise).  Returns a list per frame of dicts pred_boxes / pred_scores / pred_labels [/ pred_iou]."""
    f32 = np.float32
    hm = sigmoid32(hm)
    dim = np.exp(np.asarray(dim, f32), dtype=f32)
    indices = np.asarray(indices)
    lim = np.asarray(post_center_limit_range, f32)
    out = []
    for b in range(batch_size):
        rows = np.nonzero(indices[:, 0] == b)[0]
        sc = hm[rows].T                                                                 # (C, Nb)
        C, Nb = sc.shape
        K1 = min(K, Nb)
        o1 = np.argsort(-sc, axis=1, kind="stable")[:, :K1]
        s1 = np.take_along_axis(sc, o1, axis=1).reshape(-1)
        o2 = np.argsort(-s1, kind="stable")[:min(K, s1.shape[0])]
        score = s1[o2]
        cls = (o2 // max(K1, 1)).astype(np.int32)
        r = rows[o1.reshape(-1)[o2]]
        ctr, cz, dm, rt = np.asarray(center, f32)[r], np.asarray(center_z, f32)[r], dim[r], np.asarray(rot, f32)[r]
        angle = np.arctan2(rt[:, 1:2], rt[:, 0:1]).astype(f32)
        xs = (indices[r, 2:3].astype(f32) + ctr[:, 0:1]).astype(f32)
        ys = (indices[r, 1:2].astype(f32) + ctr[:, 1:2]).astype(f32)
        xs = ((xs * f32(feature_map_stride)).astype(f32) * f32(voxel_size[0])).astype(f32) + f32(point_cloud_range[0])
        ys = ((ys * f32(feature_map_stride)).astype(f32) * f32(voxel_size[1])).astype(f32) + f32(point_cloud_range[1])
        parts = [xs.astype(f32), ys.astype(f32), cz, dm, angle]
        if vel is not None:
            parts.append(np.asarray(vel, f32)[r])
        boxes = np.concatenate(parts, axis=-1).astype(f32)
        mask = (boxes[:, :3] >= lim[:3]).all(1) & (boxes[:, :3] <= lim[3:]).all(1)
        if score_thresh is not None:
            mask &= score > f32(score_thresh)
        lab = cls[mask]
        if class_map is not None:
            lab = np.asarray(class_map)[lab]
        d = {"pred_boxes": boxes[mask], "pred_scores": score[mask], "pred_labels": lab.astype(np.int32)}
        if iou is not None:
            pi = ((np.asarray(iou, f32)[r, 0] + f32(1.0)) * f32(0.5)).astype(f32)
            d["pred_iou"] = np.clip(pi, f32(0.0), f32(1.0))[mask]
        out.append(d)
    return out


def voxelhead_generate_predicted_boxes(pred_dicts, indices, batch_size, class_id_mapping_each_head, K, feature_map_stride, voxel_size,
                                       point_cloud_range, post_center_limit_range, score_thresh, nms_thresh, nms_pre, nms_post,
                                       iou_branch=False, rectifier=None, num_class=None, use_vel=False):
    """VoxelNeXtHead.generate_predicted_boxes (voxelnext_head.py:418-488).  Without the IoU branch: class-agnostic NMS per head
    (scalar nms_thresh / pre / post).  With it: the heads' boxes are concatenated per frame and rotate_class_specific_nms_iou (:308-331)
    runs one NMS per class on the scores score^(1-r) * iou^r with that class's threshold / pre / post sizes (lists)."""
    ret = [{"pred_boxes": [], "pred_scores": [], "pred_labels": [], "pred_iou": []} for _ in range(batch_size)]
    for h, pd in enumerate(pred_dicts):
        dec = voxelhead_decode(pd["hm"], pd["center"], pd["center_z"], pd["dim"], pd["rot"], pd.get("vel") if use_vel else None,
                               pd.get("iou") if iou_branch else None, indices, batch_size, K, feature_map_stride, voxel_size,
                               point_cloud_range, post_center_limit_range, score_thresh, class_map=class_id_mapping_each_head[h])
        for b, d in enumerate(dec):
            if not iou_branch:
                sel = nms_rotated(d["pred_boxes"], d["pred_scores"], nms_thresh, nms_pre, nms_post) if len(d["pred_scores"]) else np.zeros(0, np.int64)
            else:
                sel = np.arange(len(d["pred_scores"]))
                ret[b]["pred_iou"].append(d["pred_iou"])
            ret[b]["pred_boxes"].append(d["pred_boxes"][sel])
            ret[b]["pred_scores"].append(d["pred_scores"][sel])
            ret[b]["pred_labels"].append(d["pred_labels"][sel])
    out = []
    for b in range(batch_size):
        boxes, scores, labels = (np.concatenate(ret[b][k], 0) for k in ("pred_boxes", "pred_scores", "pred_labels"))
        if iou_branch:
            ious = np.concatenate(ret[b]["pred_iou"], 0)
            bl, sl, ll = [], [], []
            for c in range(num_class):
                m = labels == c
                r = np.float32(rectifier[c])
                sc = (np.power(scores[m], np.float32(1.0) - r, dtype=np.float32) * np.power(ious[m], r, dtype=np.float32)).astype(np.float32)
                sel = nms_rotated(boxes[m], sc, nms_thresh[c], nms_pre[c], nms_post[c]) if m.any() else np.zeros(0, np.int64)
                bl.append(boxes[m][sel]); sl.append(sc[sel]); ll.append(labels[m][sel])
            boxes, scores, labels = np.concatenate(bl, 0), np.concatenate(sl, 0), np.concatenate(ll, 0)
        out.append({"pred_boxes": boxes, "pred_scores": scores, "pred_labels": labels + 1})
    return out


# ----------------------------------------------------------------------------------------------
# Histogram calibration ([EXT] pytorch_quantization calib.HistogramCalibrator, published algorithm; parity UNPINNED: the package is
# neither in this image nor vendored by the reference -- call sites quant/quantize.py:138-145,198-207, count_time_n_memory.py:304-365).
# numpy, float64, written independently of qlidar/tensor_quant.py.
# ----------------------------------------------------------------------------------------------
def hist_collect(batches, num_bins=2048):
    """|x| histogram with equal bins over [0, max of the FIRST batch]; later batches extend the range with bins of the same width."""
    hist = edges = None
    for x in batches:
        a = np.abs(np.asarray(x, dtype=np.float32)).reshape(-1)
        if a.size == 0:
            continue
        if hist is None:
            top = float(a.max()) if a.max() > 0 else 1.0
            edges = np.linspace(0.0, top, num_bins + 1, dtype=np.float32)
            hist, _ = np.histogram(a, bins=num_bins, range=(0.0, top))
            hist = hist.astype(np.float64)
        else:
            width = edges[1] - edges[0]
            if a.max() > edges[-1]:
                n = int(np.ceil(np.float32(a.max()) / width))
                edges = (np.arange(n + 1, dtype=np.float32) * width).astype(np.float32)
            h, _ = np.histogram(a, bins=len(edges) - 1, range=(0.0, float(edges[-1])))
            h = h.astype(np.float64)
            h[:len(hist)] += hist
            hist = h
    return hist, edges


def hist_amax_percentile(hist, edges, percentile=99.99):
    cdf = np.cumsum(hist / hist.sum())
    return np.float32(edges[min(int(np.searchsorted(cdf, percentile / 100.0)), len(edges) - 1)])


def hist_amax_mse(hist, edges, num_bits=8, stride=1, start_bin=128):
    centers = ((edges[1:].astype(np.float64) + edges[:-1]) / 2).astype(np.float32)
    bound = np.float32(2 ** (num_bits - 1) - 1)
    best = None
    for i in range(start_bin, len(hist) + 1, stride):
        amax = centers[i - 1]
        scale = bound / amax
        q = np.clip(np.rint(centers * scale), -bound, bound) / scale
        err = float(np.mean(((q - centers).astype(np.float32) ** 2) * hist.astype(np.float32)))
        if best is None or err < best[0]:
            best = (err, amax)
    return np.float32(best[1])


def hist_amax_entropy(hist, edges, num_bits=8, stride=1, start_bin=128):
    bins = hist.astype(np.float64).copy()
    bins[0] = bins[1]
    levels = 1 << (num_bits - 1)
    best = None
    for i in range(max(start_bin, levels), len(bins) + 1, stride):
        ref = bins[:i].copy()
        ref[i - 1] += bins[i:].sum()
        cand = np.zeros(i)
        lvl = (np.arange(i) * levels) // i
        for L in range(levels):
            sel = lvl == L
            nz = sel & (bins[:i] != 0)
            if nz.any():
                cand[nz] = bins[:i][sel].sum() / nz.sum()
        p, q = ref / ref.sum(), cand / cand.sum()
        m = p > 0
        if (q[m] == 0).any():
            continue
        kl = float(np.sum(p[m] * np.log(p[m] / q[m])))
        if best is None or kl < best[0]:
            best = (kl, i)
    return np.float32(edges[best[1] if best else len(bins)])
