"""Tensor-level wrappers over the C ABI (include/qlidar.h).  PyTorch is used only for device memory and the
current stream; every computation below is a hand-written sm_100a kernel in libqlidar_b200.so.  There is no
CPU path: CPU tensors are rejected."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (QL_F16, QL_F32, QL_S8, QL_S32, QL_Q_CODES_PER_TENSOR, QL_Q_FAKE_PER_CHANNEL, QL_Q_FAKE_PER_TENSOR,
                   QL_Q_FAKE_PER_ROW, TILE_M, QlidarError, check, lib)

_DT = {torch.float16: QL_F16, torch.float32: QL_F32, torch.int8: QL_S8, torch.int32: QL_S32}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise QlidarError("qlidar ops need contiguous CUDA tensors (there is no CPU fallback)")


def _i32x3(v):
    if isinstance(v, int):
        v = (v, v, v)
    a = (C.c_int32 * 3)(*[int(x) for x in v])
    return a


def triple(v) -> Tuple[int, int, int]:
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(x) for x in v)
    if len(v) != 3:
        raise ValueError("expected an int or a 3-tuple")
    return v


def conv_out_shape(in_shape, ksize, stride, pad):
    k, s, p = triple(ksize), triple(stride), triple(pad)
    return [(int(in_shape[d]) + 2 * p[d] - k[d]) // s[d] + 1 for d in range(3)]


def num_tiles(n_cap: int) -> int:
    return (int(n_cap) + TILE_M - 1) // TILE_M


def hash_capacity(n: int) -> int:
    return int(lib().ql_hash_capacity(int(n)))


def hash_build(coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid: Sequence[int], table: Optional[torch.Tensor] = None):
    """grid = (B, D, H, W).  Returns the uint64-slot table as an int64 tensor."""
    _need_cuda(coords, n_dev, table)
    n_cap = coords.shape[0]
    if table is None:
        table = torch.empty(hash_capacity(n_cap), dtype=torch.int64, device=coords.device)
    B, D, H, W = [int(v) for v in grid]
    check(lib().ql_hash_build(_ptr(coords), n_cap, _ptr(n_dev), B, D, H, W, _ptr(table), table.numel(), _stream()), "ql_hash_build")
    return table


def voxelize_mean(points: torch.Tensor, pc_range, voxel_size, grid_xyz, batch_size: int, max_pts: int, max_voxels: int,
                  has_batch_col: bool = True, n_feat: Optional[int] = None, out=None, workspace=None,
                  max_voxels_per_frame: int = 0, phase: str = "mean"):
    """Fused hard voxelization + mean VFE (+ DynamicMeanVFE semantics when max_pts == 0).
    Returns (feats [max_voxels,F] f32, coords [max_voxels,4] i32, npts [max_voxels] i32, n_dev [2] i32 = (kept, found), table).
    A caller-provided feats buffer may be wider than F (row stride = its second dimension, pad columns written as zeros).
    max_voxels_per_frame > 0: the reference's per-frame MAX_NUMBER_OF_VOXELS (points frame-contiguous, frames ascending).
    phase: "mean" (everything), or "coords" then "features" with the same buffers (coords / n_dev are final after "coords")."""
    _need_cuda(points)
    if points.dtype != torch.float32 or points.dim() != 2:
        raise QlidarError("points must be a float32 (P, stride) tensor")
    P, stride = points.shape
    F = int(n_feat) if n_feat is not None else stride - (1 if has_batch_col else 0)
    dev = points.device
    if out is None:
        feats = torch.empty((max_voxels, F), dtype=torch.float32, device=dev)
        coords = torch.empty((max_voxels, 4), dtype=torch.int32, device=dev)
        npts = torch.empty((max_voxels,), dtype=torch.int32, device=dev)
        n_dev = torch.zeros((2,), dtype=torch.int32, device=dev)
        table = torch.empty(hash_capacity(max(P, 1)), dtype=torch.int64, device=dev)
    else:
        feats, coords, npts, n_dev, table = out
    ws_bytes = int(lib().ql_voxelize_workspace_bytes(P, max_voxels, F, max_pts))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rmin = (C.c_float * 3)(*[float(v) for v in pc_range[:3]])
    vs = (C.c_float * 3)(*[float(v) for v in voxel_size])
    g = (C.c_int32 * 3)(*[int(v) for v in grid_xyz])
    fn = {"mean": lib().ql_voxelize_mean, "coords": lib().ql_voxelize_coords, "features": lib().ql_voxelize_features}[phase]
    check(fn(_ptr(points), P, stride, 1 if has_batch_col else 0, F, rmin, vs, g, int(batch_size), int(max_pts),
                                 int(max_voxels), int(max_voxels_per_frame), _ptr(feats), int(feats.shape[1]), _ptr(coords), _ptr(npts), _ptr(n_dev), _ptr(table), table.numel(),
                                 _ptr(workspace), workspace.numel(), _stream()), "ql_voxelize_" + phase)
    return feats, coords, npts, n_dev, table


def voxelize_sorted(points: torch.Tensor, pc_range, voxel_size, grid_xyz, batch_size: int, max_pts: int, max_voxels: int,
                    rank_workspace: torch.Tensor, has_batch_col: bool = True, n_feat: Optional[int] = None, out=None, workspace=None,
                    frame_counts: Optional[torch.Tensor] = None, phase: str = "both"):
    """Key-sorted voxelisation straight from the points (include/qlidar.h: ql_voxelize_sorted_*): voxel ids are the ranks of the
    occupied cells in ascending linear key; rank_workspace (rulebook_strided_workspace_bytes((B, gz + 1, gy, gx), 1, 1, 0)) receives the
    stage's rank index.  phase "coords" / "features" (same buffers, in that order) or "both".
    Returns (feats [max_voxels, F], coords [max_voxels, 4] i32, npts, n_dev [2] = (kept, found))."""
    _need_cuda(points, rank_workspace, frame_counts)
    if points.dtype != torch.float32 or points.dim() != 2:
        raise QlidarError("points must be a float32 (P, stride) tensor")
    P, stride = points.shape
    F = int(n_feat) if n_feat is not None else stride - (1 if has_batch_col else 0)
    dev = points.device
    if out is None:
        feats = torch.empty((max_voxels, F), dtype=torch.float32, device=dev)
        coords = torch.empty((max_voxels, 4), dtype=torch.int32, device=dev)
        npts = torch.empty((max_voxels,), dtype=torch.int32, device=dev)
        n_dev = torch.zeros((2,), dtype=torch.int32, device=dev)
    else:
        feats, coords, npts, n_dev = out
    ws_bytes = int(lib().ql_voxelize_sorted_workspace_bytes(P, max_voxels, max_pts))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rmin = (C.c_float * 3)(*[float(v) for v in pc_range[:3]])
    vs = (C.c_float * 3)(*[float(v) for v in voxel_size])
    g = (C.c_int32 * 3)(*[int(v) for v in grid_xyz])
    hb = 1 if has_batch_col else 0
    if phase in ("both", "coords"):
        check(lib().ql_voxelize_sorted_coords(_ptr(points), P, stride, hb, F, rmin, vs, g, int(batch_size), int(max_voxels), _ptr(coords), _ptr(n_dev),
                                              _ptr(frame_counts), _ptr(rank_workspace), rank_workspace.numel(), _ptr(workspace), workspace.numel(),
                                              _stream()), "ql_voxelize_sorted_coords")
    if phase in ("both", "features"):
        check(lib().ql_voxelize_sorted_features(_ptr(points), P, stride, hb, F, rmin, vs, g, int(batch_size), int(max_pts), int(max_voxels),
                                                _ptr(n_dev), _ptr(feats), int(feats.shape[1]), _ptr(npts), _ptr(workspace), workspace.numel(),
                                                _stream()), "ql_voxelize_sorted_features")
    return feats, coords, npts, n_dev


def mean_vfe(voxels: torch.Tensor, num_points: torch.Tensor) -> torch.Tensor:
    _need_cuda(voxels, num_points)
    V, T, F = voxels.shape
    if voxels.dtype != torch.float32 or num_points.dtype not in (torch.float32, torch.int32):
        raise QlidarError("mean_vfe expects float32 voxels and float32/int32 num_points")
    out = torch.empty((V, F), dtype=torch.float32, device=voxels.device)
    check(lib().ql_mean_vfe(_ptr(voxels), _ptr(num_points), 1 if num_points.dtype == torch.float32 else 0, V, T, F, _ptr(out), _stream()),
          "ql_mean_vfe")
    return out


def mask_words(kvol: int) -> int:
    return (int(kvol) + 31) // 32


def expand_rulebook(nbr: torch.Tensor, kmask: Optional[torch.Tensor]) -> torch.Tensor:
    """Dense [tiles, K, 128] view (-1 = no neighbour) of a COMPACT rulebook (include/qlidar.h: with a mask only the live
    slabs of a tile are stored, first in its block).  For tests and accounting -- the kernels consume the compact form."""
    if kmask is None:
        return nbr
    tiles, K, _ = nbr.shape
    km = kmask[:tiles]
    sh = torch.arange(32, device=nbr.device, dtype=torch.int32)
    bits = ((km.unsqueeze(-1) >> sh) & 1).reshape(tiles, -1)[:, :K].bool()
    pos = bits.long().cumsum(1) - 1
    dense = torch.full_like(nbr, -1)
    t_idx, k_idx = bits.nonzero(as_tuple=True)
    dense[t_idx, k_idx] = nbr[t_idx, pos[t_idx, k_idx]]
    return dense


def rulebook_subm(coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid, ksize, table: torch.Tensor,
                  nbr: Optional[torch.Tensor] = None, kmask: Optional[torch.Tensor] = None, with_mask: bool = False):
    """nbr int32 [tiles, K, 128]; with_mask (or a kmask buffer) also returns the per-tile offset mask int32 [tiles, ceil(K/32)]
    and makes the rulebook COMPACT (expand_rulebook gives the dense view)."""
    _need_cuda(coords, n_dev, table, nbr, kmask)
    k = triple(ksize)
    K = k[0] * k[1] * k[2]
    n_cap = coords.shape[0]
    if nbr is None:
        nbr = torch.empty((num_tiles(n_cap), K, TILE_M), dtype=torch.int32, device=coords.device)
    if kmask is None and with_mask:
        kmask = torch.zeros((num_tiles(n_cap), mask_words(K)), dtype=torch.int32, device=coords.device)
    B, D, H, W = [int(v) for v in grid]
    check(lib().ql_rulebook_subm(_ptr(coords), n_cap, _ptr(n_dev), B, D, H, W, _i32x3(k), _ptr(table), table.numel(), _ptr(nbr),
                                 _ptr(kmask), _stream()), "ql_rulebook_subm")
    return (nbr, kmask) if (with_mask or kmask is not None) else nbr


class RankIndex:
    """(bitmap, word_prefix) device pointers inside a strided-rulebook workspace: the rank index of the stage that build
    produced (include/qlidar.h).  Holds the workspace tensor so the memory outlives the pointers."""
    def __init__(self, workspace: torch.Tensor, bitmap_ptr: int, prefix_ptr: int, n_words: int):
        self.workspace, self.bitmap_ptr, self.prefix_ptr, self.n_words = workspace, bitmap_ptr, prefix_ptr, n_words


def rulebook_strided_index(in_grid, ksize, stride, pad, workspace: torch.Tensor) -> RankIndex:
    _need_cuda(workspace)
    B, D, H, W = [int(v) for v in in_grid]
    bm, pf, nw = C.c_void_p(), C.c_void_p(), C.c_int64()
    check(lib().ql_rulebook_strided_index(B, D, H, W, _i32x3(triple(ksize)), _i32x3(triple(stride)), _i32x3(triple(pad)),
                                          _ptr(workspace), C.byref(bm), C.byref(pf), C.byref(nw)), "ql_rulebook_strided_index")
    return RankIndex(workspace, bm.value, pf.value, nw.value)


def renumber_by_key(coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid, workspace: torch.Tensor, out_coords: Optional[torch.Tensor] = None,
                    n_out_dev: Optional[torch.Tensor] = None, src_row: Optional[torch.Tensor] = None, rows_in: Optional[torch.Tensor] = None,
                    rows_out: Optional[torch.Tensor] = None):
    """Sort a list of distinct sites by linear key; the workspace (rulebook_strided_workspace_bytes(grid, 1, 1, 0)) keeps the
    stage's rank index (rulebook_strided_index(grid, 1, 1, 0, workspace)).  Returns (out_coords, n_out_dev, src_row, rows_out)."""
    _need_cuda(coords, n_dev, workspace, out_coords, n_out_dev, src_row, rows_in, rows_out)
    n_cap = coords.shape[0]
    dev = coords.device
    if out_coords is None:
        out_coords = torch.empty_like(coords)
    if n_out_dev is None:
        n_out_dev = torch.zeros((2,), dtype=torch.int32, device=dev)
    if src_row is None:
        src_row = torch.empty((n_cap,), dtype=torch.int32, device=dev)
    if rows_in is not None and rows_out is None:
        rows_out = torch.empty_like(rows_in)
    B, D, H, W = [int(v) for v in grid]
    row_bytes = 0 if rows_in is None else rows_in.shape[1] * rows_in.element_size()
    check(lib().ql_renumber_by_key(_ptr(coords), n_cap, _ptr(n_dev), B, D, H, W, _ptr(out_coords), _ptr(n_out_dev), _ptr(src_row),
                                   _ptr(rows_in), _ptr(rows_out), row_bytes, _ptr(workspace), workspace.numel(), _stream()), "ql_renumber_by_key")
    return out_coords, n_out_dev, src_row, rows_out


def rulebook_subm_ranked(coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid, ksize, index: RankIndex,
                         nbr: Optional[torch.Tensor] = None, kmask: Optional[torch.Tensor] = None):
    """Submanifold rulebook of a key-sorted stage through its rank index (no hash).  Returns (nbr, kmask)."""
    _need_cuda(coords, n_dev, nbr, kmask)
    k = triple(ksize)
    K = k[0] * k[1] * k[2]
    n_cap = coords.shape[0]
    if nbr is None:
        nbr = torch.empty((num_tiles(n_cap), K, TILE_M), dtype=torch.int32, device=coords.device)
    if kmask is None:
        kmask = torch.zeros((num_tiles(n_cap), mask_words(K)), dtype=torch.int32, device=coords.device)
    B, D, H, W = [int(v) for v in grid]
    check(lib().ql_rulebook_subm_ranked(_ptr(coords), n_cap, _ptr(n_dev), B, D, H, W, _i32x3(k), C.c_void_p(index.bitmap_ptr),
                                        C.c_void_p(index.prefix_ptr), _ptr(nbr), _ptr(kmask), _stream()), "ql_rulebook_subm_ranked")
    return nbr, kmask


def rulebook_subm_ranked_grouped(coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid, ksize, index: RankIndex,
                                 nbr: Optional[torch.Tensor] = None, kmask: Optional[torch.Tensor] = None,
                                 row_perm: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """Grouped submanifold rulebook (rows binned by their 9-bit line key so a tile's rows share live offsets).
    Returns (nbr, kmask, row_perm): nbr / kmask are indexed by tile SLOT, row_perm [tiles * 128] maps slot -> output row."""
    _need_cuda(coords, n_dev, nbr, kmask, row_perm, workspace)
    k = triple(ksize)
    K = k[0] * k[1] * k[2]
    n_cap = coords.shape[0]
    dev = coords.device
    if nbr is None:
        nbr = torch.empty((num_tiles(n_cap), K, TILE_M), dtype=torch.int32, device=dev)
    if kmask is None:
        kmask = torch.zeros((num_tiles(n_cap), mask_words(K)), dtype=torch.int32, device=dev)
    if row_perm is None:
        row_perm = torch.empty((num_tiles(n_cap) * TILE_M,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib().ql_rulebook_group_workspace_bytes(n_cap))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    B, D, H, W = [int(v) for v in grid]
    check(lib().ql_rulebook_subm_ranked_grouped(_ptr(coords), n_cap, _ptr(n_dev), B, D, H, W, _i32x3(k), C.c_void_p(index.bitmap_ptr),
                                                C.c_void_p(index.prefix_ptr), _ptr(nbr), _ptr(kmask), _ptr(row_perm), _ptr(workspace),
                                                workspace.numel(), _stream()), "ql_rulebook_subm_ranked_grouped")
    return nbr, kmask, row_perm


def rulebook_strided_workspace_bytes(grid, ksize, stride, pad) -> int:
    B, D, H, W = [int(v) for v in grid]
    n = int(lib().ql_rulebook_strided_workspace_bytes(B, D, H, W, _i32x3(triple(ksize)), _i32x3(triple(stride)), _i32x3(triple(pad))))
    if n == 0:
        raise QlidarError("invalid strided-conv geometry")
    return n


def rulebook_strided(coords: torch.Tensor, n_in_dev: Optional[torch.Tensor], grid, ksize, stride, pad,
                     n_out_cap: int, out=None, workspace=None, kmask: Optional[torch.Tensor] = None,
                     in_index: Optional["RankIndex"] = None, want_kmask: bool = True):
    """Returns (out_coords [n_out_cap,4] sorted by linear key, n_out_dev [2] = (kept, found), out_table, nbr [tiles,K,128],
    out_grid (B,D,H,W), kmask [tiles, ceil(K/32)]).  in_index: the rank index of a key-sorted INPUT stage -> the pairs come
    from the output side through it (ql_rulebook_strided_ranked; no hash table is produced, out_table must be None)."""
    _need_cuda(coords, n_in_dev, kmask)
    k, s, p = triple(ksize), triple(stride), triple(pad)
    K = k[0] * k[1] * k[2]
    n_in_cap = coords.shape[0]
    dev = coords.device
    B, D, H, W = [int(v) for v in grid]
    od, oh, ow = conv_out_shape((D, H, W), k, s, p)
    if out is None:
        out_coords = torch.empty((n_out_cap, 4), dtype=torch.int32, device=dev)
        n_out_dev = torch.zeros((2,), dtype=torch.int32, device=dev)
        out_table = None if in_index is not None else torch.empty(hash_capacity(n_out_cap), dtype=torch.int64, device=dev)
        nbr = torch.empty((num_tiles(n_out_cap), K, TILE_M), dtype=torch.int32, device=dev)
    else:
        out_coords, n_out_dev, out_table, nbr = out                   # out_table may be None: rank-index consumers only
    if kmask is None and want_kmask:
        kmask = torch.zeros((num_tiles(n_out_cap), mask_words(K)), dtype=torch.int32, device=dev)
    ws_bytes = rulebook_strided_workspace_bytes(grid, k, s, p)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if in_index is not None:
        if out_table is not None:
            raise QlidarError("rulebook_strided(in_index=...) produces no hash table: pass out_table=None")
        check(lib().ql_rulebook_strided_ranked(_ptr(coords), n_in_cap, _ptr(n_in_dev), B, D, H, W, _i32x3(k), _i32x3(s), _i32x3(p),
                                               C.c_void_p(in_index.bitmap_ptr), C.c_void_p(in_index.prefix_ptr), _ptr(out_coords),
                                               int(n_out_cap), _ptr(n_out_dev), _ptr(nbr), _ptr(kmask), _ptr(workspace),
                                               workspace.numel(), _stream()), "ql_rulebook_strided_ranked")
        return out_coords, n_out_dev, None, nbr, (B, od, oh, ow), kmask
    check(lib().ql_rulebook_strided(_ptr(coords), n_in_cap, _ptr(n_in_dev), B, D, H, W, _i32x3(k), _i32x3(s), _i32x3(p),
                                    _ptr(out_coords), int(n_out_cap), _ptr(n_out_dev), _ptr(out_table), 0 if out_table is None else out_table.numel(), _ptr(nbr),
                                    _ptr(kmask), _ptr(workspace), workspace.numel(), _stream()), "ql_rulebook_strided")
    return out_coords, n_out_dev, out_table, nbr, (B, od, oh, ow), kmask


def bev_densify_ranked(feats: torch.Tensor, index: RankIndex, n_dev: Optional[torch.Tensor], grid, out: Optional[torch.Tensor] = None,
                       out_dtype: torch.dtype = torch.float16, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bev_densify for a key-sorted stage: cell -> row from the stage's rank index."""
    _need_cuda(feats, n_dev, out)
    B, D, H, W = [int(v) for v in grid]
    c = feats.shape[1]
    if out is None:
        out = torch.empty((B, c * D, H, W), dtype=out_dtype, device=feats.device)
    ws_bytes = int(lib().ql_bev_densify_workspace_bytes(B, D, H, W))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
    check(lib().ql_bev_densify_ranked(_ptr(feats), _DT[feats.dtype], c, C.c_void_p(index.bitmap_ptr), C.c_void_p(index.prefix_ptr),
                                      feats.shape[0], _ptr(n_dev), B, D, H, W, _ptr(out), _DT[out.dtype], _ptr(workspace),
                                      workspace.numel(), _stream()), "ql_bev_densify_ranked")
    return out


def bev_merge2d(feats: torch.Tensor, coords: torch.Tensor, n_dev: Optional[torch.Tensor], grid_bhw, n_out_cap: Optional[int] = None,
                out_dtype: Optional[torch.dtype] = None):
    """VoxelNeXt bev_out: drop z, unique (b, y, x) rows in ascending order, duplicates summed.
    Returns (out_feats [n_out_cap, c], out_coords [n_out_cap, 3] int32, n_out_dev [2] = (kept, found))."""
    _need_cuda(feats, coords, n_dev)
    if coords.dtype != torch.int32 or coords.shape[1] != 4:
        raise QlidarError("bev_merge2d expects int32 [n, 4] coords [b, z, y, x]")
    n, c = feats.shape
    B, H, W = [int(v) for v in grid_bhw]
    cap = int(n_out_cap) if n_out_cap is not None else max(int(n), 1)
    out_dtype = out_dtype or feats.dtype
    dev = feats.device
    out_feats = torch.empty((cap, c), dtype=out_dtype, device=dev)
    out_coords = torch.zeros((cap, 3), dtype=torch.int32, device=dev)
    n_out = torch.zeros((2,), dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib().ql_bev_merge2d_workspace_bytes(B, H, W, cap, c, _DT[out_dtype])), dtype=torch.uint8, device=dev)
    check(lib().ql_bev_merge2d(_ptr(feats), _DT[feats.dtype], c, _ptr(coords), n, _ptr(n_dev), B, H, W, _ptr(out_feats), _DT[out_dtype],
                               _ptr(out_coords), cap, _ptr(n_out), _ptr(ws), ws.numel(), _stream()), "ql_bev_merge2d")
    return out_feats, out_coords, n_out


def bev_merge2d_multi(segments, grid_bhw, n_out_cap: int, out_feats: torch.Tensor, out_coords: torch.Tensor, n_out_dev: torch.Tensor,
                      workspace: torch.Tensor):
    """VoxelNeXt's stage-4/5/6 merge in one pass.  segments: list of (feats [cap, c], coords [cap, 4] int32, n_dev, coord_scale);
    out_coords [n_out_cap, 3 or 4] int32 (4: [b, 0, y, x]); everything preallocated (graph-capturable)."""
    n = len(segments)
    feats = [s[0] for s in segments]
    _need_cuda(*feats, *[s[1] for s in segments], *[s[2] for s in segments], out_feats, out_coords, n_out_dev, workspace)
    c = feats[0].shape[1]
    B, H, W = [int(v) for v in grid_bhw]
    fp = (C.c_void_p * n)(*[f.data_ptr() for f in feats])
    cp = (C.c_void_p * n)(*[s[1].data_ptr() for s in segments])
    caps = (C.c_int64 * n)(*[int(s[0].shape[0]) for s in segments])
    nd = (C.c_void_p * n)(*[None if s[2] is None else s[2].data_ptr() for s in segments])
    sc = (C.c_int32 * n)(*[int(s[3]) for s in segments])
    check(lib().ql_bev_merge2d_multi(n, fp, _DT[feats[0].dtype], c, cp, caps, nd, sc, B, H, W, _ptr(out_feats), _DT[out_feats.dtype],
                                     _ptr(out_coords), int(out_coords.shape[1]), int(n_out_cap), _ptr(n_out_dev), _ptr(workspace),
                                     workspace.numel(), _stream()), "ql_bev_merge2d_multi")
    return out_feats, out_coords, n_out_dev


def zero_led_rows(n: int, c: int, dtype=torch.float16, device="cuda", fill: bool = True) -> torch.Tensor:
    """[n, c] zero-initialised rows with one extra all-zero row IN FRONT of row 0 (same allocation): the layout the conv kernel
    gathers from -- rulebook index -1 reads that row (include/qlidar.h, zero-row contract).  The returned view is tagged;
    spconv_mma takes tagged tensors as they are and copies anything else into such a buffer.  fill=False zeroes only the leading
    row (for a caller that writes every other row itself)."""
    if fill:
        buf = torch.zeros((int(n) + 1, int(c)), dtype=dtype, device=device)
    else:
        buf = torch.empty((int(n) + 1, int(c)), dtype=dtype, device=device)
        buf[0].zero_()
    v = buf[1:]
    v._ql_zero_led = buf                      # keeps the allocation (and the zero row) alive with the view
    return v


def _zero_led(feats: torch.Tensor) -> torch.Tensor:
    if getattr(feats, "_ql_zero_led", None) is not None and feats.is_contiguous():
        return feats
    v = zero_led_rows(feats.shape[0], feats.shape[1], feats.dtype, feats.device, fill=False)
    v.copy_(feats)
    return v


def pack_weights(w: torch.Tensor) -> torch.Tensor:
    """w: CPU tensor (c_out, K, c_in) int8 codes or float16 -> CPU uint8 tensor with the shared-memory image."""
    if w.is_cuda:
        w = w.cpu()
    w = w.contiguous()
    if w.dtype not in (torch.int8, torch.float16) or w.dim() != 3:
        raise QlidarError("pack_weights expects (c_out, K, c_in) int8 or float16")
    c_out, K, c_in = w.shape
    dt = _DT[w.dtype]
    nbytes = int(lib().ql_packed_weight_bytes(c_in, c_out, K, dt))
    if nbytes == 0:
        raise QlidarError("unsupported weight shape")
    out = torch.empty(nbytes, dtype=torch.uint8)
    check(lib().ql_pack_weights_host(C.c_void_p(w.data_ptr()), dt, c_in, c_out, K, C.c_void_p(out.data_ptr())), "ql_pack_weights_host")
    return out


def sq_prepare_weights(w_okc: torch.Tensor, w_ic_absmax: torch.Tensor, act_absmax: torch.Tensor, alpha: float,
                       bn_scale: Optional[torch.Tensor] = None, smooth: Optional[torch.Tensor] = None,
                       packed: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None):
    """SmoothQuant on the device: w_okc (c_out, K, c_in) fp32 -> (smooth [c_in], packed int8 weight image, scale [c_out])."""
    _need_cuda(w_okc, w_ic_absmax, act_absmax, bn_scale, smooth, packed, scale)
    if w_okc.dtype != torch.float32 or w_okc.dim() != 3:
        raise QlidarError("sq_prepare_weights expects (c_out, K, c_in) float32 weights")
    c_out, K, c_in = w_okc.shape
    dev = w_okc.device
    if smooth is None:
        smooth = torch.empty(c_in, dtype=torch.float32, device=dev)
    if packed is None:
        packed = torch.empty(int(lib().ql_packed_weight_bytes(c_in, c_out, K, QL_S8)), dtype=torch.uint8, device=dev)
    if scale is None:
        scale = torch.empty(c_out, dtype=torch.float32, device=dev)
    check(lib().ql_sq_prepare_weights(_ptr(w_okc), _ptr(w_ic_absmax), _ptr(act_absmax), C.c_float(float(alpha)), c_in, c_out, K,
                                      _ptr(bn_scale), _ptr(smooth), _ptr(packed), _ptr(scale), _stream()), "ql_sq_prepare_weights")
    return smooth, packed, scale


def weights_streamed(c_in: int, c_out: int, kvol: int, dtype: torch.dtype = torch.float16) -> bool:
    """True when the conv kernel streams this layer's weights per tile (they do not fit in shared memory)."""
    return bool(lib().ql_spconv_weights_streamed(int(c_in), int(c_out), int(kvol), _DT[dtype]))


def spconv_mma(feats: torch.Tensor, nbr: torch.Tensor, n_out_cap: int, n_out_dev: Optional[torch.Tensor], c_out: int,
               w_packed: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, *, act_scale: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None, relu: bool = False, out: Optional[torch.Tensor] = None,
               out_dtype: torch.dtype = torch.float16, out_q: Optional[torch.Tensor] = None,
               out_qscale: Optional[torch.Tensor] = None, absmax: Optional[torch.Tensor] = None,
               kmask: Optional[torch.Tensor] = None, row_perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """kmask: the rulebook's per-tile offset mask [tiles, ceil(K/32)] int32 (None = visit every offset).
    row_perm: slot -> output row table of a grouped rulebook (rulebook_subm_ranked_grouped)."""
    _need_cuda(feats, nbr, n_out_dev, w_packed, scale, shift, act_scale, residual, out, out_q, out_qscale, absmax, kmask, row_perm)
    if row_perm is not None and (row_perm.dtype != torch.int32 or row_perm.numel() < num_tiles(n_out_cap) * TILE_M):
        raise QlidarError("row_perm must be int32 [tiles * 128]")
    if feats.dtype not in (torch.float16, torch.int8):
        raise QlidarError("spconv_mma gathers float16 rows or int8 codes")
    if residual is not None and residual.dtype != torch.float16:
        raise QlidarError("residual must be float16")
    c_in = feats.shape[1]
    K = nbr.shape[1]
    feats = _zero_led(feats)                      # zero-row contract of ql_spconv_mma (a copy unless allocated by zero_led_rows)
    if out is None:
        out = torch.empty((n_out_cap, c_out), dtype=out_dtype, device=feats.device)
    raw = out.dtype == torch.int32
    if not raw and out.dtype not in (torch.float16, torch.float32):
        raise QlidarError("out must be float16/float32 (or int32 for the raw accumulators)")
    if kmask is not None and (kmask.dtype != torch.int32 or kmask.shape[0] < num_tiles(n_out_cap) or kmask.shape[1] != mask_words(K)):
        raise QlidarError("kmask must be int32 [tiles, ceil(K/32)]")
    check(lib().ql_spconv_mma_rows(_ptr(feats), _DT[feats.dtype], _ptr(nbr), _ptr(kmask), _ptr(row_perm), int(n_out_cap), _ptr(n_out_dev),
                                   c_in, int(c_out), K, _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(act_scale), _ptr(residual),
                                   1 if relu else 0, _ptr(out), _DT[out.dtype], _ptr(out_q), _ptr(out_qscale), _ptr(absmax), _stream()),
          "ql_spconv_mma_rows")
    return out


def stem_conv(feats: torch.Tensor, nbr: torch.Tensor, n_out_cap: int, n_out_dev: Optional[torch.Tensor], w_kio: torch.Tensor,
              scale: torch.Tensor, shift: torch.Tensor, relu: bool = True, out: Optional[torch.Tensor] = None,
              out_dtype: torch.dtype = torch.float16, absmax: Optional[torch.Tensor] = None,
              kmask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """w_kio: (K, c_in, c_out) fp32; feats (N, stride >= c_in) fp32 (stride 8 = the engine's padded rows, one 256-bit load each).
    kmask: the rulebook's per-tile offset mask (the rulebook is then compact), None = dense rulebook."""
    _need_cuda(feats, nbr, n_out_dev, w_kio, scale, shift, out, absmax, kmask)
    if feats.dtype != torch.float32 or w_kio.dtype != torch.float32:
        raise QlidarError("stem_conv is the fp32 path")
    K, c_in, c_out = w_kio.shape
    if feats.shape[1] < c_in:
        raise QlidarError("stem_conv: feature rows narrower than c_in")
    if out is None:
        out = torch.empty((n_out_cap, c_out), dtype=out_dtype, device=feats.device)
    if kmask is not None and (kmask.dtype != torch.int32 or kmask.shape[0] < num_tiles(n_out_cap) or kmask.shape[1] != mask_words(K)):
        raise QlidarError("kmask must be int32 [tiles, ceil(K/32)]")
    check(lib().ql_stem_conv(_ptr(feats), int(feats.shape[1]), c_in, _ptr(nbr), _ptr(kmask), int(n_out_cap), _ptr(n_out_dev), c_out, K, _ptr(w_kio), _ptr(scale), _ptr(shift),
                             1 if relu else 0, _ptr(out), _DT[out.dtype], _ptr(absmax), _stream()), "ql_stem_conv")
    return out


def permute_rows(x: torch.Tensor, src_row: torch.Tensor, n_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = x[src_row[r]] for r < n (src_row int32, e.g. the [tiles, 1, 128] rulebook of a 1x1x1 renumbering build)."""
    _need_cuda(x, src_row, n_dev, out)
    if out is None:
        out = torch.empty_like(x)
    row_bytes = x.shape[1] * x.element_size()
    n_cap = min(x.shape[0], out.shape[0], src_row.numel())
    check(lib().ql_permute_rows(_ptr(x), _ptr(out), row_bytes, _ptr(src_row), n_cap, _ptr(n_dev), _stream()), "ql_permute_rows")
    return out


def absmax_cols(x: torch.Tensor, n_dev: Optional[torch.Tensor] = None, absmax: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, n_dev, absmax)
    n, c = x.shape
    if absmax is None:
        absmax = torch.zeros((c,), dtype=torch.float32, device=x.device)
    check(lib().ql_absmax_cols(_ptr(x), _DT[x.dtype], n, _ptr(n_dev), c, _ptr(absmax), _stream()), "ql_absmax_cols")
    return absmax


def quantize_rows(x: torch.Tensor, absmax: Optional[torch.Tensor], mode: int, bits: int = 8, n_dev: Optional[torch.Tensor] = None,
                  smooth: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, act_scale: Optional[torch.Tensor] = None):
    """Returns (out, act_scale): int8 codes + device scalar de-quantisation scale (CODES_PER_TENSOR) or fp16 fake-quant rows."""
    _need_cuda(x, absmax, n_dev, smooth, out, act_scale)
    n, c = x.shape
    if out is None:
        out = torch.empty((n, c), dtype=torch.int8 if mode == QL_Q_CODES_PER_TENSOR else torch.float16, device=x.device)
    if act_scale is None and mode == QL_Q_CODES_PER_TENSOR:
        act_scale = torch.zeros((1,), dtype=torch.float32, device=x.device)
    check(lib().ql_quantize_rows(_ptr(x), _DT[x.dtype], n, _ptr(n_dev), c, _ptr(absmax), _ptr(smooth), int(bits), int(mode), _ptr(out),
                                 _ptr(act_scale), _stream()), "ql_quantize_rows")
    return out, act_scale


def bev_densify(feats: torch.Tensor, table: torch.Tensor, grid, out: Optional[torch.Tensor] = None,
                out_dtype: torch.dtype = torch.float16, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """grid = (B, D, H, W) -> out (B, C*D, H, W) with channel index c*D + d (== dense().view(N, C*D, H, W))."""
    _need_cuda(feats, table, out)
    B, D, H, W = [int(v) for v in grid]
    c = feats.shape[1]
    if out is None:
        out = torch.empty((B, c * D, H, W), dtype=out_dtype, device=feats.device)
    ws_bytes = int(lib().ql_bev_densify_workspace_bytes(B, D, H, W))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
    check(lib().ql_bev_densify(_ptr(feats), _DT[feats.dtype], c, _ptr(table), table.numel(), B, D, H, W, _ptr(out), _DT[out.dtype],
                               _ptr(workspace), workspace.numel(), _stream()), "ql_bev_densify")
    return out


# ------------------------------------------------------------------ CenterHead post-processing (csrc/centerhead.cu)
def _host_f32(vals, n):
    a = (C.c_float * n)(*[float(v) for v in list(vals)[:n]])
    return a


def centerhead_decode(hm: torch.Tensor, center: torch.Tensor, center_z: torch.Tensor, dim: torch.Tensor, rot: torch.Tensor,
                      vel: Optional[torch.Tensor], iou: Optional[torch.Tensor], K: int, feature_map_stride, voxel_size, point_cloud_range,
                      post_center_limit_range, score_thresh: Optional[float], class_map: Optional[torch.Tensor] = None,
                      workspace: Optional[torch.Tensor] = None):
    """One head of CenterHead.generate_predicted_boxes up to (not including) NMS (center_head.py:297-327 ->
    centernet_utils.decode_bbox_from_heatmap, centernet_utils.py:176-241).  `hm` are LOGITS, `dim` log-sizes, `rot` = (cos, sin), `iou`
    the raw head output.  Returns (boxes [B,K,7|9], scores [B,K], labels [B,K] int32 (class_map applied, 0-based), iou [B,K] or None,
    count [B] int32): per frame the first count[b] rows are the masked candidates in descending score order."""
    maps = [hm, center, center_z, dim, rot, vel, iou]
    _need_cuda(*maps, class_map)
    for t in maps:
        if t is not None and t.dtype != torch.float32:
            raise QlidarError("centerhead_decode expects fp32 head maps")
    B, Cn, H, W = [int(v) for v in hm.shape]
    dev = hm.device
    bd = 9 if vel is not None else 7
    boxes = torch.zeros((B, K, bd), dtype=torch.float32, device=dev)
    scores = torch.zeros((B, K), dtype=torch.float32, device=dev)
    labels = torch.zeros((B, K), dtype=torch.int32, device=dev)
    out_iou = torch.zeros((B, K), dtype=torch.float32, device=dev) if iou is not None else None
    count = torch.zeros((B,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib().ql_centerhead_decode_workspace_bytes(B, Cn, H, W))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib().ql_centerhead_decode(_ptr(hm), _ptr(center), _ptr(center_z), _ptr(dim), _ptr(rot), _ptr(vel), _ptr(iou), B, Cn, H, W, int(K),
                                     float(feature_map_stride), _host_f32(voxel_size, 2), _host_f32(point_cloud_range, 2),
                                     _host_f32(post_center_limit_range, 6), -1.0 if score_thresh is None else float(score_thresh),
                                     _ptr(class_map), _ptr(boxes), _ptr(scores), _ptr(labels), _ptr(out_iou), _ptr(count), _ptr(workspace),
                                     workspace.numel(), _stream()), "ql_centerhead_decode")
    return boxes, scores, labels, out_iou, count


def voxelhead_decode(hm: torch.Tensor, center: torch.Tensor, center_z: torch.Tensor, dim: torch.Tensor, rot: torch.Tensor,
                     vel: Optional[torch.Tensor], iou: Optional[torch.Tensor], indices_byx: torch.Tensor, n_dev: Optional[torch.Tensor],
                     batch_size: int, K: int, feature_map_stride, voxel_size, point_cloud_range, post_center_limit_range,
                     score_thresh: Optional[float], class_map: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """One head of VoxelNeXtHead.generate_predicted_boxes up to (not including) NMS (voxelnext_head.py:418-456 ->
    centernet_utils.decode_bbox_from_voxels_nuscenes, centernet_utils.py:289-354).  Per-voxel row-major fp32 arrays: `hm` [N, C] LOGITS,
    `dim` [N, 3] log-sizes, `rot` [N, 2] = (cos, sin), `iou` [N, 1] the raw head output; indices_byx int32 [N, 3] = the sparse tensor's
    (batch, y, x).  Returns (boxes [B,K,7|9], scores, labels int32 (class_map applied), iou [B,K] or None, count [B])."""
    arrs = [hm, center, center_z, dim, rot, vel, iou]
    _need_cuda(*arrs, indices_byx, n_dev, class_map)
    for t in arrs:
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise QlidarError("voxelhead_decode expects contiguous fp32 per-voxel arrays")
    if indices_byx.dtype != torch.int32 or indices_byx.dim() != 2 or indices_byx.shape[1] != 3 or not indices_byx.is_contiguous():
        raise QlidarError("indices must be int32 [N, 3] (batch, y, x)")
    N, Cn = int(hm.shape[0]), int(hm.shape[1])
    B = int(batch_size)
    dev = hm.device
    bd = 9 if vel is not None else 7
    boxes = torch.zeros((B, K, bd), dtype=torch.float32, device=dev)
    scores = torch.zeros((B, K), dtype=torch.float32, device=dev)
    labels = torch.zeros((B, K), dtype=torch.int32, device=dev)
    out_iou = torch.zeros((B, K), dtype=torch.float32, device=dev) if iou is not None else None
    count = torch.zeros((B,), dtype=torch.int32, device=dev)
    if N == 0:                                    # an empty sparse tensor: no candidates in any frame
        return boxes, scores, labels, out_iou, count
    ws_bytes = int(lib().ql_voxelhead_decode_workspace_bytes(B, Cn, N))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    check(lib().ql_voxelhead_decode(_ptr(hm), _ptr(center), _ptr(center_z), _ptr(dim), _ptr(rot), _ptr(vel), _ptr(iou), _ptr(indices_byx), N,
                                    _ptr(n_dev), B, Cn, int(K), float(feature_map_stride), _host_f32(voxel_size, 2),
                                    _host_f32(point_cloud_range, 2), _host_f32(post_center_limit_range, 6),
                                    -1.0 if score_thresh is None else float(score_thresh), _ptr(class_map), _ptr(boxes), _ptr(scores),
                                    _ptr(labels), _ptr(out_iou), _ptr(count), _ptr(workspace), workspace.numel(), _stream()),
          "ql_voxelhead_decode")
    return boxes, scores, labels, out_iou, count


def voxelhead_class_split(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, ious: torch.Tensor, count: torch.Tensor,
                          rectifier: torch.Tensor):
    """The first half of rotate_class_specific_nms_iou (voxelnext_head.py:308-331): per class c the frame's boxes with label c,
    re-scored score^(1-r[c]) * iou^r[c] and sorted by the new score.  Returns (boxes [C,B,K,bd], scores [C,B,K], labels [C,B,K],
    counts [C,B])."""
    _need_cuda(boxes, scores, labels, ious, count, rectifier)
    B, K, bd = [int(v) for v in boxes.shape]
    Cn = int(rectifier.numel())
    dev = boxes.device
    ob = torch.zeros((Cn, B, K, bd), dtype=torch.float32, device=dev)
    os_ = torch.zeros((Cn, B, K), dtype=torch.float32, device=dev)
    ol = torch.zeros((Cn, B, K), dtype=torch.int32, device=dev)
    oc = torch.zeros((Cn, B), dtype=torch.int32, device=dev)
    check(lib().ql_voxelhead_class_split(_ptr(boxes.contiguous()), bd, _ptr(scores.contiguous()), _ptr(labels.contiguous()), _ptr(ious.contiguous()),
                                         _ptr(count), B, K, Cn, _ptr(rectifier.float().contiguous()), _ptr(ob), _ptr(os_), _ptr(ol), _ptr(oc),
                                         _stream()), "ql_voxelhead_class_split")
    return ob, os_, ol, oc


def nms_rotated(boxes: torch.Tensor, scores: Optional[torch.Tensor], labels: Optional[torch.Tensor], counts: Optional[torch.Tensor],
                thresh: float, pre_max: int, post_max: int, label_offset: int = 0, box_dim: Optional[int] = None, return_iou: bool = False,
                workspace: Optional[torch.Tensor] = None):
    """Greedy rotated-BEV-IoU NMS per frame with the sweep on the device (iou3d_nms_utils.nms_gpu, iou3d_nms_utils.py:120-135; nms_kernel,
    iou3d_nms_kernel.cu:295-339; the host sweep of iou3d_nms.cpp:137-183).  boxes [B, n_cap, >=7] fp32, each frame's first counts[b]
    rows in DESCENDING score order.  Returns a dict: keep [B, post_max] int32 (-1 padded), keep_count [B], boxes [B, post_max, box_dim],
    scores / labels [B, post_max] when given, iou [B, n_cap, n_cap] (upper triangle) when return_iou."""
    _need_cuda(boxes, scores, labels, counts)
    if boxes.dtype != torch.float32 or boxes.dim() != 3:
        raise QlidarError("nms_rotated expects fp32 [B, n_cap, box_stride] boxes")
    B, n_cap, stride = [int(v) for v in boxes.shape]
    bd = int(box_dim) if box_dim is not None else stride
    dev = boxes.device
    keep = torch.empty((B, post_max), dtype=torch.int32, device=dev)
    keep_count = torch.zeros((B,), dtype=torch.int32, device=dev)
    out_boxes = torch.zeros((B, post_max, bd), dtype=torch.float32, device=dev)
    out_scores = torch.zeros((B, post_max), dtype=torch.float32, device=dev) if scores is not None else None
    out_labels = torch.zeros((B, post_max), dtype=torch.int32, device=dev) if labels is not None else None
    iou = torch.zeros((B, n_cap, n_cap), dtype=torch.float32, device=dev) if return_iou else None
    ws_bytes = int(lib().ql_nms_rotated_workspace_bytes(B, n_cap))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib().ql_nms_rotated(_ptr(boxes), stride, bd, _ptr(scores), _ptr(labels), _ptr(counts), B, n_cap, float(thresh), int(pre_max),
                               int(post_max), int(label_offset), _ptr(keep), _ptr(keep_count), _ptr(out_boxes), _ptr(out_scores),
                               _ptr(out_labels), _ptr(iou), _ptr(workspace), workspace.numel(), _stream()), "ql_nms_rotated")
    return {"keep": keep, "keep_count": keep_count, "boxes": out_boxes, "scores": out_scores, "labels": out_labels, "iou": iou}
