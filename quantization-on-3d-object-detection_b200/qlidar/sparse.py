"""Drop-in mirror of the spconv 2.x Python surface that pcdet and the quant/ wrappers use
(SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, SparseConv3d, SubMConv2d, SparseConv2d), backed by
the sm_100a kernels.  Reference call sites: pcdet/utils/spconv_utils.py:3-38,
pcdet/models/backbones_3d/spconv_backbone.py:12-17,39-46,78-118,256-261, height_compression.py:21.

Weight layout is spconv-2: (C_out, *kernel, C_in) (quant/quant.py:37-39, detector3d_template.py:346)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from ._lib import QlidarError


class IndiceData:
    """One cached rulebook ([EXT] spconv indice_dict entry): the tile-major neighbour table plus the output geometry."""

    def __init__(self, nbr, n_out, n_out_dev, out_indices, out_table, out_grid, ksize, stride, pad, subm, in_indices, kmask=None):
        self.nbr = nbr
        self.kmask = kmask                # per-tile offset mask [tiles, ceil(K/32)] (empty slabs are skipped by the conv)
        self.n_out = n_out                # capacity (== exact row count in the module path)
        self.n_out_dev = n_out_dev        # device-side count or None
        self.out_indices = out_indices
        self.out_table = out_table
        self.out_grid = out_grid          # (B, D, H, W)
        self.ksize, self.stride, self.pad, self.subm = ksize, stride, pad, subm
        self.in_indices = in_indices


class SparseConvTensor:
    """features (N, C) + indices (N, 1+ndim) int32 [b, z, y, x] (or [b, y, x] for 2-D)."""

    def __init__(self, features: torch.Tensor, indices: torch.Tensor, spatial_shape: Sequence[int], batch_size: int,
                 grid=None, voxel_num=None, indice_dict: Optional[dict] = None, benchmark: bool = False, _check: bool = True):
        if indices.dtype != torch.int32:
            raise QlidarError("indices must be int32 (the reference passes voxel_coords.int(), spconv_backbone.py:258)")
        if _check and features.shape[0] != indices.shape[0]:
            raise QlidarError("features and indices disagree on the number of active sites")
        self._features = features
        self._features_as = None          # (dtype, cast copy): engine outputs are stored fp16 and surface in the caller's dtype on demand
        self._surface_dtype = None
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict: Dict[str, IndiceData] = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self._table: Optional[torch.Tensor] = None
        self._table_src = None
        self._n_dev: Optional[torch.Tensor] = None

    # ---- spconv API ----
    @property
    def features(self) -> torch.Tensor:
        """The (N, C) feature rows.  A tensor published by the engine keeps its fp16 rows and converts to the caller's dtype
        (the reference's fp32) the first time someone reads them -- CenterPoint never reads the multi-scale taps."""
        if self._surface_dtype is None or self._features.dtype == self._surface_dtype:
            return self._features
        if self._features_as is None:
            self._features_as = self._features.to(self._surface_dtype)
        return self._features_as

    @features.setter
    def features(self, value: torch.Tensor):
        self._features, self._features_as, self._surface_dtype = value, None, None

    @property
    def spatial_size(self):
        n = 1
        for s in self.spatial_shape:
            n *= s
        return n

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def replace_feature(self, feature: torch.Tensor) -> "SparseConvTensor":
        """New tensor sharing indices, the rulebook cache and the hash table (pcdet/utils/spconv_utils.py:32-38).
        Like spconv's, it does not insist that the row counts agree: VoxelNeXt replaces the features with a concatenation
        first and assigns the matching indices afterwards (spconv_backbone_voxelnext.py:196-197)."""
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                             self.indice_dict, self.benchmark, _check=False)
        t._table, t._table_src, t._n_dev = self._table, self._table_src, self._n_dev
        return t

    def dense(self, channels_first: bool = True) -> torch.Tensor:
        """(B, C, *spatial) -- [EXT] SparseConvTensor.dense(), used by HeightCompression (height_compression.py:21)."""
        g = self._grid4()
        f = self.features
        if f.dtype not in (torch.float16, torch.float32):
            f = f.float()
        out = ops.bev_densify(f.contiguous(), self.table(), g, out_dtype=f.dtype)       # (B, C*D, H, W), channel = c*D + d
        B, D, H, W = g
        C = f.shape[1]
        out = out.view(B, C, D, H, W) if len(self.spatial_shape) == 3 else out.view(B, C, H, W)
        if not channels_first:
            out = out.permute(0, *range(2, out.dim()), 1).contiguous()
        return out

    @classmethod
    def from_dense(cls, x: torch.Tensor) -> "SparseConvTensor":
        """x: (B, *spatial, C) channels-last, like spconv.  Index discovery uses torch.nonzero (not on the hot path)."""
        spatial = list(x.shape[1:-1])
        mask = (x != 0).any(dim=-1)
        idx = mask.nonzero()
        feats = x[mask]
        return cls(feats.contiguous(), idx.int().contiguous(), spatial, x.shape[0])

    # ---- internals ----
    def _grid4(self):
        if len(self.spatial_shape) == 3:
            return (self.batch_size, *self.spatial_shape)
        if len(self.spatial_shape) == 2:
            return (self.batch_size, 1, *self.spatial_shape)
        raise QlidarError("only 2-D and 3-D sparse tensors are supported")

    def indices4(self) -> torch.Tensor:
        """[b, z, y, x] view of the indices (2-D tensors get z = 0)."""
        if self.indices.shape[1] == 4:
            return self.indices if self.indices.is_contiguous() else self.indices.contiguous()
        b = self.indices
        return torch.stack([b[:, 0], torch.zeros_like(b[:, 0]), b[:, 1], b[:, 2]], dim=1).contiguous()

    def table(self) -> torch.Tensor:
        # VoxelNeXt mutates `.indices` in place / rebinds it (spconv_backbone_voxelnext.py:194-197): rebuild when stale
        src = (self.indices.data_ptr(), self.indices._version, tuple(self.indices.shape))
        if self._table is None or self._table_src != src:
            self._table = ops.hash_build(self.indices4(), self._n_dev, self._grid4())
            self._table_src = src
        return self._table


class PendingCounts:
    """The (kept, found) row counts of one engine forward, on their way to the host: an asynchronous copy into pinned memory plus the
    event that marks its arrival.  counts() waits for the event (usually long past) and returns the kept rows per stage; it raises
    if the forward overflowed a stage capacity or a per-frame voxel cap -- the synchronous plugin call handles those by re-running,
    a deferred one no longer can."""

    def __init__(self, host: torch.Tensor, event: torch.cuda.Event, n_stages: int, frame_cap: int = 0):
        self.host, self.event, self.n_stages, self.frame_cap = host, event, n_stages, int(frame_cap)
        self._n: Optional[List[int]] = None

    def counts(self) -> List[int]:
        if self._n is None:
            self.event.synchronize()
            c = self.host[:2 * self.n_stages].view(self.n_stages, 2)
            if bool((c[1:, 1] > c[1:, 0]).any()) or (self.frame_cap and bool((self.host[2 * self.n_stages:] > self.frame_cap).any())):
                raise QlidarError("an engine capacity or per-frame voxel cap was exceeded in a deferred-count call: run this batch with "
                                  "engine_lazy_counts = False once (the synchronous call re-sizes the engine) and switch it back on")
            self._n = [int(v) for v in c[:, 0].tolist()]
        return self._n


class LazySparseConvTensor(SparseConvTensor):
    """An engine result whose row count is still on the device (backbone.engine_lazy_counts = True): `features` / `indices` are cut out
    of the engine's capacity-sized buffers the first time anyone reads them, which is when the host first needs the count.  A
    CenterPoint forward never does (the dense BEV map has a fixed shape), so the whole plugin chain stays free of host synchronisation
    and the host can queue the next batch while this one runs.  capacity_features / capacity_indices are the engine's own buffers:
    they hold this call's rows until the next call of the same backbone overwrites them."""

    def __init__(self, pending: PendingCounts, stage_i: int, feats_full: torch.Tensor, coords_full: torch.Tensor, spatial_shape, batch_size: int,
                 surface_dtype=None, index_cols=None):
        self._pending, self._stage_i = pending, stage_i
        self.capacity_features, self.capacity_indices, self._index_cols = feats_full, coords_full, index_cols
        self._lazy_f = self._lazy_i = None
        self._features_as = None
        self._surface_dtype = surface_dtype if (surface_dtype is not None and surface_dtype != feats_full.dtype) else None
        self.spatial_shape = [int(v) for v in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = None
        self.voxel_num = None
        self.benchmark = False
        self._table = None
        self._table_src = None
        self._n_dev = None

    def num_rows(self) -> int:
        return self._pending.counts()[self._stage_i]

    @property
    def _features(self):
        if self._lazy_f is None:
            self._lazy_f = self.capacity_features[:self.num_rows()]
        return self._lazy_f

    @_features.setter
    def _features(self, value):
        self._lazy_f = value

    @property
    def indices(self):
        if self._lazy_i is None:
            idx = self.capacity_indices[:self.num_rows()]
            self._lazy_i = idx if self._index_cols is None else idx[:, self._index_cols].contiguous()
        return self._lazy_i

    @indices.setter
    def indices(self, value):
        self._lazy_i = value


class SparseModule(nn.Module):
    """Marker base ([EXT] spconv.pytorch.modules.SparseModule; quant/quant.py:3)."""
    pass


def is_spconv_module(m) -> bool:
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    """[EXT] spconv SparseSequential: sparse modules get the tensor, dense modules (BatchNorm1d, ReLU) get `.features`."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError(f"index {idx} is out of range")
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        self.add_module(name if name is not None else str(len(self._modules)), module)

    def forward(self, input):
        for m in self._modules.values():
            if is_spconv_module(m):
                input = m(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(m(input.features))
            else:
                input = m(input)
        return input


class SparseConvolution(SparseModule):
    """Base of the conv modules ([EXT] spconv.conv.SparseConvolution; pcdet/utils/spconv_utils.py:23 discovers
    checkpoint keys through it)."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, indice_key=None, algo=None, **_unused):
        super().__init__()
        if groups != 1 or (dilation != 1 and tuple(_ntuple(dilation, ndim)) != (1,) * ndim):
            raise QlidarError("groups/dilation != 1 are not used by the reference backbones and are not supported")
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = list(_ntuple(kernel_size, ndim))
        self.stride = list(_ntuple(stride, ndim))
        self.padding = list(_ntuple(padding, ndim))
        self.dilation = [1] * ndim
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5 ** 0.5)
        self._pack_cache = {}

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, subm={self.subm}, bias={self.bias is not None}, indice_key={self.indice_key}")

    # -- geometry as zyx triples (2-D convs run as D=1 slabs) --
    def _k3(self):
        return tuple(self.kernel_size) if self.ndim == 3 else (1, *self.kernel_size)

    def _s3(self):
        return tuple(self.stride) if self.ndim == 3 else (1, *self.stride)

    def _p3(self):
        return tuple(self.padding) if self.ndim == 3 else (0, *self.padding)

    def get_rulebook(self, x: SparseConvTensor) -> IndiceData:
        data = x.find_indice_pair(self.indice_key)
        if data is not None and (not self.subm or data.in_indices is x.indices or data.in_indices.data_ptr() == x.indices.data_ptr()):
            return data
        grid = x._grid4()
        idx4 = x.indices4()
        n = idx4.shape[0]
        if self.subm:
            nbr, kmask = ops.rulebook_subm(idx4, x._n_dev, grid, self._k3(), x.table(), with_mask=True)
            data = IndiceData(nbr, n, x._n_dev, x.indices, x.table(), grid, self._k3(), (1, 1, 1), None, True, x.indices, kmask)
        else:
            k, s, p = self._k3(), self._s3(), self._p3()
            od, oh, ow = ops.conv_out_shape(grid[1:], k, s, p)
            per_in = 1
            for d in range(3):
                per_in *= -(-k[d] // s[d])
            cap = max(1, min(n * per_in, grid[0] * od * oh * ow))
            out_c, n_out_dev, out_table, nbr, ogrid, kmask = ops.rulebook_strided(idx4, x._n_dev, grid, k, s, p, cap)
            n_out = int(n_out_dev[0].item())          # module API returns exact shapes (one sync per strided rulebook)
            out_c = out_c[:n_out]
            nbr = nbr[:ops.num_tiles(max(n_out, 1))]
            out_idx = out_c if self.ndim == 3 else out_c[:, [0, 2, 3]].contiguous()
            data = IndiceData(nbr, n_out, None, out_idx, out_table, ogrid, k, s, p, False, x.indices, kmask)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = data
        return data

    def _packed_weight(self, dev):
        """fp16 weights in the kernel's shared-memory image; channels zero-padded to the MMA granularity."""
        key = (self.weight._version, self.weight.data_ptr(), str(dev))
        hit = self._pack_cache.get("w")
        if hit is not None and hit[0] == key:
            return hit[1:]
        oc, ic = self.out_channels, self.in_channels
        K = 1
        for k in self.kernel_size:
            K *= k
        ic_p, oc_p = _round_up(ic, 8), _round_up(oc, 16)
        w = torch.zeros((oc_p, K, ic_p), dtype=torch.float16)
        w[:oc, :, :ic] = self.weight.detach().reshape(oc, K, ic).to(torch.float16).cpu()
        packed = ops.pack_weights(w).to(dev)
        self._pack_cache["w"] = (key, packed, ic_p, oc_p)
        return packed, ic_p, oc_p

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        if not isinstance(x, SparseConvTensor):
            raise QlidarError("sparse conv modules take a SparseConvTensor")
        rb = self.get_rulebook(x)
        f = x.features
        in_dtype = f.dtype
        if in_dtype == torch.float32 and self.in_channels <= 8 and self.out_channels in (16, 32) and self.ndim == 3:
            # the un-quantised stem on raw point features (conv_input.0: metre-scale coordinates, intensities up to 255): fp32 SIMT
            # conv like the engine's, not the fp16 tensor-core path -- a half cast here would round the inputs (and overflow above
            # 65504) before calibration ever sees them
            key = ("stem", self.weight._version, self.weight.data_ptr(), str(f.device))
            hit = self._pack_cache.get("stem")
            if hit is None or hit[0] != key:
                K = 1
                for k in self.kernel_size:
                    K *= k
                w_kio = self.weight.detach().float().reshape(self.out_channels, K, self.in_channels).permute(1, 2, 0).contiguous().to(f.device)
                hit = (key, w_kio)
                self._pack_cache["stem"] = hit
            scale = torch.ones(self.out_channels, dtype=torch.float32, device=f.device)
            shift = self.bias.detach().float().to(f.device) if self.bias is not None else torch.zeros(self.out_channels, dtype=torch.float32, device=f.device)
            y = ops.stem_conv(f.contiguous(), rb.nbr, rb.n_out, rb.n_out_dev, hit[1], scale, shift, relu=False, out_dtype=torch.float32, kmask=rb.kmask)
            return _make_output(x, rb, y, self.ndim)
        packed, ic_p, oc_p = self._packed_weight(f.device)
        fh = f if f.dtype == torch.float16 else f.to(torch.float16)
        if ic_p != self.in_channels:
            fh = torch.nn.functional.pad(fh, (0, ic_p - self.in_channels))
        fh = fh.contiguous()
        scale = torch.ones(oc_p, dtype=torch.float32, device=f.device)
        shift = torch.zeros(oc_p, dtype=torch.float32, device=f.device)
        if self.bias is not None:
            shift[:self.out_channels] = self.bias.detach().float()
        out_dtype = torch.float32 if in_dtype == torch.float32 else torch.float16
        y = ops.spconv_mma(fh, rb.nbr, rb.n_out, rb.n_out_dev, oc_p, packed, scale, shift, out_dtype=out_dtype, kmask=rb.kmask)
        if oc_p != self.out_channels:
            y = y[:, :self.out_channels].contiguous()
        return _make_output(x, rb, y, self.ndim)


def _make_output(x: SparseConvTensor, rb: IndiceData, feats: torch.Tensor, ndim: int) -> SparseConvTensor:
    if rb.subm:
        return x.replace_feature(feats)
    shape = list(rb.out_grid[1:]) if ndim == 3 else list(rb.out_grid[2:])
    out = SparseConvTensor(feats, rb.out_indices, shape, x.batch_size, x.grid, x.voxel_num, x.indice_dict, x.benchmark)
    out._table = rb.out_table
    out._table_src = (rb.out_indices.data_ptr(), rb.out_indices._version, tuple(rb.out_indices.shape))
    out._n_dev = rb.n_out_dev
    return out


def _ntuple(v, n):
    if isinstance(v, (list, tuple)):
        if len(v) != n:
            raise ValueError(f"expected {n} values, got {v}")
        return [int(a) for a in v]
    return [int(v)] * n


def _round_up(v, m):
    return (v + m - 1) // m * m


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True, indice_key, algo)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False, indice_key, algo)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True, indice_key, algo)


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kw):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False, indice_key, algo)


class SparseInverseConv3d(SparseModule):
    """Only pcdet's UNet backbone uses it (spconv_unet.py) -- outside this path."""

    def __init__(self, *a, **kw):
        super().__init__()
        raise NotImplementedError("SparseInverseConv3d is out of scope for the CenterPoint/SECOND/VoxelNeXt backbone path")


def replace_feature(out, new_features):
    """pcdet/utils/spconv_utils.py:32-38."""
    return out.replace_feature(new_features)
