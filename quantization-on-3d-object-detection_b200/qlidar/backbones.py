"""Drop-in mirrors of the pcdet plugins on the hot path, with the same constructor signatures, attributes, module
tree (=> identical state_dict keys and `named_children` paths for q_conv3d's no_list) and batch_dict contract:

  MeanVFE                      pcdet/models/backbones_3d/vfe/mean_vfe.py:6-31
  DynamicMeanVFE               pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:38-76
  VoxelBackBone8x              pcdet/models/backbones_3d/spconv_backbone.py:70-181
  VoxelResBackBone8x           pcdet/models/backbones_3d/spconv_backbone.py:184-295
  VoxelResBackBone8xVoxelNeXt  pcdet/models/backbones_3d/spconv_backbone_voxelnext.py:69-225
  HeightCompression            pcdet/models/backbones_2d/map_to_bev/height_compression.py:4-26

`forward(batch_dict)` runs the module tree op by op (each conv = one fused kernel).  The graph-captured, sync-free
whole-backbone schedule lives in qlidar.engine.BackboneEngine (built from these modules)."""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import sparse as spconv
from .sparse import SparseConvTensor, replace_feature


class Cfg(dict):
    """Minimal EasyDict stand-in (pcdet/config.py uses easydict): attribute access + .get()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _cfg(c):
    return c if isinstance(c, Cfg) else Cfg(c or {})


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0, conv_type='subm', norm_fn=None):
    """spconv_backbone.py:8-27."""
    if conv_type == 'subm':
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == 'spconv':
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                                   indice_key=indice_key)
    elif conv_type == 'inverseconv':
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
    else:
        raise NotImplementedError
    return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


class SparseBasicBlock(spconv.SparseModule):
    """spconv_backbone.py:30-67 (res-block convs carry a bias: `bias = norm_fn is not None`, :37-38)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, bias=None, norm_fn=None, downsample=None, indice_key=None):
        super().__init__()
        assert norm_fn is not None
        if bias is None:
            bias = norm_fn is not None
        self.conv1 = spconv.SubMConv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=bias, indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=bias, indice_key=indice_key)
        self.bn2 = norm_fn(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.conv1(x)
        out = replace_feature(out, self.bn1(out.features))
        out = replace_feature(out, self.relu(out.features))
        out = self.conv2(out)
        out = replace_feature(out, self.bn2(out.features))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = replace_feature(out, out.features + identity.features)
        out = replace_feature(out, self.relu(out.features))
        return out


class _BackboneBase(nn.Module):
    """forward(batch_dict) -- the reference's plugin call (spconv_backbone.py:243-295) -- is the FAST path: the first call compiles
    the module tree into a qlidar.engine.BackboneEngine (BN folded into the conv epilogues, weights packed, every buffer allocated
    once, the whole forward captured in one CUDA graph) and later calls replay it; the engine is rebuilt when a parameter, a
    buffer (calibrated amax, BN statistics) or a quantiser switch changes, or when the input outgrows its capacity.  The op-by-op
    module tree below still runs -- every conv one fused kernel, BatchNorm / ReLU / residual as separate ATen kernels -- while
    calibrating (collect_stats disables the quantisers), in training mode, for module trees the engine cannot schedule, and
    when `use_engine` is False.  One host synchronisation per call remains: the reference API returns exactly-sized tensors.
    Row order: the engine returns every stage in ascending-key order (x_conv1 included: the reference keeps the voxeliser's
    order there); coordinates travel with the features, so consumers that index by `indices` are unaffected."""
    use_engine = True
    # True: once a synchronous call has sized the engine for the kind of batch it sees, later calls return WITHOUT reading the row counts
    # back (the call's only host synchronisation): the published sparse tensors are LazySparseConvTensor objects that resolve their shape
    # on first access, their rows live in the engine's buffers until the next call.  See sparse.LazySparseConvTensor.
    engine_lazy_counts = False
    engine_bev_dtype = None                   # set by HeightCompression.attach(): the BEV map is then produced inside the graph

    def _input_tensor(self, batch_dict):
        voxel_features, voxel_coords = batch_dict['voxel_features'], batch_dict['voxel_coords']
        # load_data_to_gpu casts coords to float32 (pcdet/models/__init__.py:36); `.int()` undoes it (spconv_backbone.py:258)
        idx = voxel_coords if voxel_coords.dtype == torch.int32 else voxel_coords.int()
        return SparseConvTensor(features=voxel_features, indices=idx.contiguous(), spatial_shape=self.sparse_shape,
                                batch_size=batch_dict['batch_size'])

    @staticmethod
    def _publish(batch_dict, out, taps):
        batch_dict.update({'encoded_spconv_tensor': out, 'encoded_spconv_tensor_stride': 8})
        batch_dict.update({'multi_scale_3d_features': taps})
        batch_dict.update({'multi_scale_3d_strides': {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8}})
        return batch_dict

    # ------------------------------------------------------------------ engine-backed plugin call
    def _snapshot(self):
        """What an engine was compiled from: every (container, key, object) of the module tree -- sub-modules, parameters, buffers,
        quantiser amax -- and the tensors' version counters.  _unchanged() re-checks it in ~0.1 ms per call (identity + integer
        compares, no hashing of data): surgery (q_conv3d swaps _modules entries), load_state_dict / optimiser steps (version bumps),
        calibration (new _amax buffers, quantiser switches) all invalidate it."""
        from .tensor_quant import TensorQuantizer
        entries, tensors, quants = [], [], []
        for m in self.modules():
            for d in (m._modules, m._parameters, m._buffers):
                for k, v in d.items():
                    entries.append((d, k, v))
                    if isinstance(v, torch.Tensor):
                        tensors.append(v)
            if isinstance(m, TensorQuantizer):
                quants.append(m)
        return dict(entries=entries, sizes=[(m, len(m._modules), len(m._parameters), len(m._buffers)) for m in self.modules()],
                    tensors=tensors, versions=[t._version for t in tensors], quants=quants,
                    qstate=[(q._disabled, q._if_quant, q._if_calib, q.num_bits) for q in quants])

    @staticmethod
    def _unchanged(snap) -> bool:
        for d, k, v in snap["entries"]:
            if d.get(k) is not v:
                return False
        for m, a, b, c in snap["sizes"]:
            if len(m._modules) != a or len(m._parameters) != b or len(m._buffers) != c:
                return False
        for t, ver in zip(snap["tensors"], snap["versions"]):
            if t._version != ver:
                return False
        for q, stt in zip(snap["quants"], snap["qstate"]):
            if (q._disabled, q._if_quant, q._if_calib, q.num_bits) != stt:
                return False
        return True

    def _engine_eligible(self, batch_dict) -> bool:
        if not self.use_engine or self.training or torch.is_grad_enabled() or getattr(self, "_engine_unsupported", False):
            return False
        src = batch_dict.get('_ql_points') if batch_dict.get('voxel_features') is None else batch_dict['voxel_features']
        if src is None or not src.is_cuda:
            return False
        st = getattr(self, "_engine_state", None)
        quants = st["snap"]["quants"] if st is not None else None
        if quants is None:
            from .tensor_quant import TensorQuantizer
            quants = [m for m in self.modules() if isinstance(m, TensorQuantizer)]
        for m in quants:
            if m._if_calib or m._disabled or not m._if_quant:
                return False                                    # calibration / quantisers off: the eager tree collects the statistics
        return True

    def _engine_for(self, batch_dict, stage_caps=None):
        from .engine import BackboneEngine
        from ._lib import QlidarError
        B = int(batch_dict['batch_size'])
        pts = batch_dict.get('_ql_points') if batch_dict.get('voxel_features') is None else None
        vfe = getattr(self, "_ql_vfe", None) if pts is not None else None
        if pts is not None:
            V, P = B * vfe.max_voxels, int(pts.shape[0])            # per-frame MAX_NUMBER_OF_VOXELS: the batch capacity cannot overflow
            dev = pts.device
        else:
            V, P = int(batch_dict['voxel_features'].shape[0]), None
            dev = batch_dict['voxel_features'].device
        st = getattr(self, "_engine_state", None)
        same = st is not None and self._unchanged(st["snap"]) and st["B"] == B and st["bev"] == self.engine_bev_dtype and (st["P"] is not None) == (P is not None)
        if stage_caps is None and same and V <= st["cap"] and (P is None or P <= st["P"]):
            return st["eng"]
        cap = max(V if pts is not None else int(V * 1.25) + 1024, st["cap"] if same else 0)
        if stage_caps is None and same:
            # the same model on a bigger input: scale the stage capacities learnt so far
            stage_caps = [int(c * cap / st["cap"]) + 128 for c in st["eng_caps"]]
        if stage_caps is not None:
            stage_caps = [max(cap, stage_caps[0])] + list(stage_caps[1:])
            cap = stage_caps[0]
        max_points = None if P is None else max(int(P * 1.25) + 1024, st["P"] if same and st["P"] else 0)
        self._engine_state = None                                   # release the old engine's buffers before allocating the new ones
        kw = {}
        if pts is not None:
            kw = dict(max_points=max_points, pc_range=vfe.point_cloud_range, voxel_size=vfe.voxel_size, max_pts_per_voxel=vfe.max_points_per_voxel,
                      max_voxels_per_frame=vfe.max_voxels, n_point_features=vfe.num_point_features,
                      sorted_voxelizer=not getattr(self, "_ql_frame_cap_hit", False))
        try:
            eng = BackboneEngine(self, B, cap, stage_cap_ratio=1.3, stage_caps=stage_caps, bev=self.engine_bev_dtype is not None,
                                 bev_dtype=self.engine_bev_dtype or torch.float16, device=dev, **kw)
        except QlidarError:
            self._engine_unsupported = True                     # a module tree the engine cannot schedule: the eager tree serves it
            return None
        self._engine_state = dict(snap=self._snapshot(), B=B, cap=cap, eng=eng, bev=self.engine_bev_dtype, eng_caps=[s.cap for s in eng.stages], P=max_points)
        return eng

    def _engine_forward(self, batch_dict):
        eng = self._engine_for(batch_dict)
        if eng is None:
            return None
        pts = batch_dict.get('_ql_points') if batch_dict.get('voxel_features') is None else None
        vf, vc = (pts, None) if pts is not None else (batch_dict['voxel_features'], batch_dict['voxel_coords'])
        B = int(batch_dict['batch_size'])
        if self.engine_lazy_counts and self._engine_state.get("settled"):
            return self._engine_forward_lazy(batch_dict, eng, pts, vf, vc, B)
        for attempt in range(10):
            out = eng.forward_points(pts) if pts is not None else eng.forward_voxels(vf, vc)
            call = eng.counts_all.cpu()                                         # the one synchronisation of the call
            counts = call[:2 * len(eng.stages)].view(len(eng.stages), 2)
            if pts is not None and eng.frame_cap_exceeded(call):
                # a frame has more voxels than MAX_NUMBER_OF_VOXELS: only the first-touch voxeliser drops the reference's choice of them
                eng.use_hash_voxelizer()
                self._ql_frame_cap_hit = True                                   # engines compiled later for this model start on the hash path
                continue
            over = (counts[:, 1] > counts[:, 0]).tolist()
            if not any(over[1:]):
                break
            # A stage outgrew its capacity (a denser scene than any before; 5^3 stride-2 convs multiply the site count by up to
            # 5): size that stage from the number of sites FOUND, scale the later ones (they saw a truncated input) and run again.
            first = over.index(True, 1)
            caps = [s.cap for s in eng.stages]
            grow = float(counts[first, 1]) / max(float(counts[first, 0]), 1.0) * 1.2
            caps = caps[:first] + [int(c * grow) + 1024 for c in caps[first:]]
            eng = self._engine_for(batch_dict, stage_caps=caps)
        else:
            raise RuntimeError("engine stage capacities did not converge")
        self._engine_state["settled"] = True                                    # capacities and the voxeliser choice fit this kind of batch
        n = [int(v) for v in counts[:, 0].tolist()]
        if pts is not None:
            # the fused voxeliser's products, as the VFE plugin would have published them (rows in ascending-key order)
            batch_dict['voxel_features'] = eng.vox_feats[:n[0], :eng.nfeat]
            batch_dict['voxel_coords'] = eng.stages[0].coords[:n[0]]

        def tensor(feats, stage_i):
            stg = eng.stages[stage_i]
            t = SparseConvTensor(features=feats[:n[stage_i]], indices=stg.coords[:n[stage_i]], spatial_shape=list(stg.grid[1:]), batch_size=B)
            if vf.dtype != feats.dtype:
                t._surface_dtype = vf.dtype                         # fp32 in -> fp32 out, converted when read
            return t

        last = eng.layers[-1]
        enc = tensor(last.out, last.stage_out)
        if getattr(eng, "merge_stage", None) is not None:                        # VoxelNeXt: the encoded tensor is 2-D, [b, y, x]
            enc = SparseConvTensor(features=enc._features, indices=enc.indices[:, [0, 2, 3]].contiguous(),
                                   spatial_shape=enc.spatial_shape[1:], batch_size=B)
            if vf.dtype != enc._features.dtype:
                enc._surface_dtype = vf.dtype
        if eng.bev:
            enc._ql_spatial_features = eng.spatial_features                     # HeightCompression takes the fused map
        taps = {L.tap: tensor(L.out, L.stage_out) for L in eng.layers if L.tap}
        return self._publish(batch_dict, enc, taps)


def _engine_forward_lazy(self, batch_dict, eng, pts, vf, vc, B):
    """The plugin call without its host synchronisation (engine_lazy_counts): the graph is replayed, the counts start their way to
    the host (pinned ring slot + event) and the published sparse tensors resolve their shapes on first access.  Only after one
    synchronous call has sized the engine for this kind of batch ("settled"); overflows then surface as errors at resolve time."""
    from .sparse import LazySparseConvTensor, PendingCounts
    eng.forward_points(pts) if pts is not None else eng.forward_voxels(vf, vc)
    ring = self.__dict__.setdefault("_ql_count_ring", [])
    k = self.__dict__.get("_ql_count_k", 0)
    self._ql_count_k = k + 1
    if len(ring) < 16:
        ring.append([torch.empty(eng.counts_all.numel(), dtype=torch.int32).pin_memory(), None])
        slot = ring[-1]
    else:
        slot = ring[k % 16]
        if slot[1] is not None and slot[1]._n is None and slot[0].numel() == eng.counts_all.numel():
            try:
                slot[1].counts()                                                # 16 calls old and never read: settle it before its slot is reused
            except Exception:
                pass
        if slot[0].numel() != eng.counts_all.numel():
            slot[0] = torch.empty(eng.counts_all.numel(), dtype=torch.int32).pin_memory()
    slot[0].copy_(eng.counts_all, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    pend = PendingCounts(slot[0], ev, len(eng.stages), eng.max_voxels_per_frame if (pts is not None and eng.sorted_voxelizer) else 0)
    slot[1] = pend
    if pts is not None:
        # capacity-sized: the first voxel_count rows are valid (this mode's one deviation from the plugin contract)
        batch_dict['voxel_features'] = eng.vox_feats[:, :eng.nfeat]
        batch_dict['voxel_coords'] = eng.stages[0].coords
        batch_dict['voxel_count'] = eng.stages[0].n_dev

    def tensor(feats, stage_i, cols=None, shape=None):
        stg = eng.stages[stage_i]
        return LazySparseConvTensor(pend, stage_i, feats, stg.coords, shape if shape is not None else list(stg.grid[1:]), B,
                                    surface_dtype=vf.dtype, index_cols=cols)

    last = eng.layers[-1]
    if getattr(eng, "merge_stage", None) is not None:                            # VoxelNeXt: the encoded tensor is 2-D, [b, y, x]
        enc = tensor(last.out, last.stage_out, cols=[0, 2, 3], shape=list(eng.stages[last.stage_out].grid[2:]))
    else:
        enc = tensor(last.out, last.stage_out)
    if eng.bev:
        enc._ql_spatial_features = eng.spatial_features
    taps = {L.tap: tensor(L.out, L.stage_out) for L in eng.layers if L.tap}
    return self._publish(batch_dict, enc, taps)


_BackboneBase._engine_forward_lazy = _engine_forward_lazy


class VoxelBackBone8x(_BackboneBase):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = [int(v) for v in (np.asarray(grid_size)[::-1] + [1, 0, 0])]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key='subm1'), norm_fn(16), nn.ReLU())
        block = post_act_block
        self.conv1 = spconv.SparseSequential(block(16, 16, 3, norm_fn=norm_fn, padding=1, indice_key='subm1'))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv2', conv_type='spconv'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv3', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'))
        self.conv4 = spconv.SparseSequential(
            block(64, 64, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key='spconv4', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'))
        last_pad = self.model_cfg.get('last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False, indice_key='spconv_down2'),
            norm_fn(128), nn.ReLU())
        self.num_point_features = 128
        self.backbone_channels = {'x_conv1': 16, 'x_conv2': 32, 'x_conv3': 64, 'x_conv4': 64}

    def forward(self, batch_dict):
        if self._engine_eligible(batch_dict):
            done = self._engine_forward(batch_dict)
            if done is not None:
                return done
        x = self.conv_input(self._input_tensor(batch_dict))
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        return self._publish(batch_dict, out, {'x_conv1': x_conv1, 'x_conv2': x_conv2, 'x_conv3': x_conv3, 'x_conv4': x_conv4})


class VoxelResBackBone8x(_BackboneBase):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        use_bias = self.model_cfg.get('USE_BIAS', None)
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = [int(v) for v in (np.asarray(grid_size)[::-1] + [1, 0, 0])]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key='subm1'), norm_fn(16), nn.ReLU())
        block = post_act_block
        self.conv1 = spconv.SparseSequential(
            SparseBasicBlock(16, 16, bias=use_bias, norm_fn=norm_fn, indice_key='res1'),
            SparseBasicBlock(16, 16, bias=use_bias, norm_fn=norm_fn, indice_key='res1'))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv2', conv_type='spconv'),
            SparseBasicBlock(32, 32, bias=use_bias, norm_fn=norm_fn, indice_key='res2'),
            SparseBasicBlock(32, 32, bias=use_bias, norm_fn=norm_fn, indice_key='res2'))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv3', conv_type='spconv'),
            SparseBasicBlock(64, 64, bias=use_bias, norm_fn=norm_fn, indice_key='res3'),
            SparseBasicBlock(64, 64, bias=use_bias, norm_fn=norm_fn, indice_key='res3'))
        self.conv4 = spconv.SparseSequential(
            block(64, 128, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key='spconv4', conv_type='spconv'),
            SparseBasicBlock(128, 128, bias=use_bias, norm_fn=norm_fn, indice_key='res4'),
            SparseBasicBlock(128, 128, bias=use_bias, norm_fn=norm_fn, indice_key='res4'))
        last_pad = self.model_cfg.get('last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(128, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False, indice_key='spconv_down2'),
            norm_fn(128), nn.ReLU())
        self.num_point_features = 128
        self.backbone_channels = {'x_conv1': 16, 'x_conv2': 32, 'x_conv3': 64, 'x_conv4': 128}

    def forward(self, batch_dict):
        if self._engine_eligible(batch_dict):
            done = self._engine_forward(batch_dict)
            if done is not None:
                return done
        x = self.conv_input(self._input_tensor(batch_dict))
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        return self._publish(batch_dict, out, {'x_conv1': x_conv1, 'x_conv2': x_conv2, 'x_conv3': x_conv3, 'x_conv4': x_conv4})


class _VoxelNeXtBasicBlock(SparseBasicBlock):
    """spconv_backbone_voxelnext.py:30-66 (no `bias` argument: always biased)."""

    def __init__(self, inplanes, planes, stride=1, norm_fn=None, downsample=None, indice_key=None):
        super().__init__(inplanes, planes, stride=stride, bias=None, norm_fn=norm_fn, downsample=downsample, indice_key=indice_key)


class VoxelResBackBone8xVoxelNeXt(_BackboneBase):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        ks = self.model_cfg.get('SPCONV_KERNEL_SIZES', [3, 3, 3, 3])
        ch = self.model_cfg.get('CHANNELS', [16, 32, 64, 128, 128])
        out_channel = self.model_cfg.get('OUT_CHANNEL', 128)
        self.sparse_shape = [int(v) for v in (np.asarray(grid_size)[::-1] + [1, 0, 0])]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, ch[0], 3, padding=1, bias=False, indice_key='subm1'), norm_fn(ch[0]), nn.ReLU())
        block, BB = post_act_block, _VoxelNeXtBasicBlock
        self.conv1 = spconv.SparseSequential(BB(ch[0], ch[0], norm_fn=norm_fn, indice_key='res1'), BB(ch[0], ch[0], norm_fn=norm_fn, indice_key='res1'))
        stages = [(ch[0], ch[1], ks[0]), (ch[1], ch[2], ks[1]), (ch[2], ch[3], ks[2]), (ch[3], ch[4], ks[3]), (ch[4], ch[4], ks[3])]
        for i, (ci, co, k) in enumerate(stages, start=2):
            setattr(self, f'conv{i}', spconv.SparseSequential(
                block(ci, co, k, norm_fn=norm_fn, stride=2, padding=int(k // 2), indice_key=f'spconv{i}', conv_type='spconv'),
                BB(co, co, norm_fn=norm_fn, indice_key=f'res{i}'), BB(co, co, norm_fn=norm_fn, indice_key=f'res{i}')))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv2d(ch[3], out_channel, 3, stride=1, padding=1, bias=False, indice_key='spconv_down2'),
            norm_fn(out_channel), nn.ReLU())
        self.shared_conv = spconv.SparseSequential(
            spconv.SubMConv2d(out_channel, out_channel, 3, stride=1, padding=1, bias=True), nn.BatchNorm1d(out_channel), nn.ReLU(True))
        self.forward_ret_dict = {}
        self.num_point_features = out_channel
        self.backbone_channels = {'x_conv1': ch[0], 'x_conv2': ch[1], 'x_conv3': ch[2], 'x_conv4': ch[3]}

    def bev_out(self, x_conv):
        """spconv_backbone_voxelnext.py:149-164: drop z, merge duplicate (b,y,x) rows by summation -- one fused CUDA op
        (ql_bev_merge2d: bitmap numbering in ascending (b,y,x) order == torch.unique(dim=0), fp32 atomic sums == index_add_)."""
        spatial_shape = x_conv.spatial_shape[1:]
        feats = x_conv.features.contiguous()
        if feats.dtype not in (torch.float16, torch.float32):
            feats = feats.float()
        f, c3, n_out = ops.bev_merge2d(feats, x_conv.indices.contiguous(), None, (x_conv.batch_size, spatial_shape[0], spatial_shape[1]))
        n = int(n_out[0].item())                                     # module (eager) path: the row count comes back to the host
        return SparseConvTensor(features=f[:n], indices=c3[:n], spatial_shape=spatial_shape, batch_size=x_conv.batch_size)

    def forward(self, batch_dict):
        if self._engine_eligible(batch_dict):
            done = self._engine_forward(batch_dict)
            if done is not None:
                return done
        x = self.conv_input(self._input_tensor(batch_dict))
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        x_conv5 = self.conv5(x_conv4)
        x_conv6 = self.conv6(x_conv5)
        x_conv5.indices[:, 1:] *= 2
        x_conv6.indices[:, 1:] *= 4
        x_conv4 = x_conv4.replace_feature(torch.cat([x_conv4.features, x_conv5.features, x_conv6.features]))
        x_conv4.indices = torch.cat([x_conv4.indices, x_conv5.indices, x_conv6.indices])
        out = self.bev_out(x_conv4)
        out = self.conv_out(out)
        out = self.shared_conv(out)
        return self._publish(batch_dict, out, {'x_conv1': x_conv1, 'x_conv2': x_conv2, 'x_conv3': x_conv3, 'x_conv4': x_conv4})


# ---------------------------------------------------------------------------------------------------------- VFE
class MeanVFE(nn.Module):
    def __init__(self, model_cfg=None, num_point_features=4, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, batch_dict, **kwargs):
        voxels, num = batch_dict['voxels'], batch_dict['voxel_num_points']
        if num.dtype not in (torch.float32, torch.int32):
            num = num.int()
        batch_dict['voxel_features'] = ops.mean_vfe(voxels.contiguous(), num.contiguous())
        return batch_dict


class DynamicMeanVFE(nn.Module):
    """No caps, mean over all points of a voxel.  Output order is first-touch (the reference's torch.unique order is
    key-sorted; compare as sets)."""

    def __init__(self, model_cfg=None, num_point_features=4, voxel_size=None, grid_size=None, point_cloud_range=None, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_point_features = num_point_features
        self.voxel_size = [float(v) for v in voxel_size]
        self.grid_size = [int(v) for v in grid_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]

    def get_output_feature_dim(self):
        return self.num_point_features

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        points = batch_dict['points'].contiguous()
        cap = int(batch_dict.get('max_voxels', points.shape[0]))
        feats, coords, npts, n_dev, table = ops.voxelize_mean(points, self.point_cloud_range, self.voxel_size, self.grid_size,
                                                              batch_dict['batch_size'], 0, max(cap, 1))
        n = int(n_dev[0].item())
        batch_dict['voxel_features'] = feats[:n]
        batch_dict['voxel_coords'] = coords[:n]
        return batch_dict


class VoxelizeMeanVFE(nn.Module):
    """Hard voxelisation + MeanVFE on the GPU behind the batch_dict contract: the work of DataProcessor.transform_points_to_voxels
    (pcdet/datasets/processor/data_processor.py:133-180 -> [EXT] Point2VoxelCPU3d: MAX_POINTS_PER_VOXEL, per-frame MAX_NUMBER_OF_VOXELS,
    first-touch order) + dataset.py:237-244 (batch column) + MeanVFE (mean_vfe.py:25-29), fused -- the (V, T, F) padded tensor is never
    materialised.  Reads batch_dict['points'] (sum P, 1+F) with the batch index in column 0 (frames contiguous and ascending, as
    collate_batch leaves them), writes voxel_features (V, F), voxel_coords (V, 4) int32 [b, z, y, x], voxel_num_points.
    attach(backbone): the voxeliser then runs INSIDE the backbone's CUDA graph (points -> voxels -> rulebooks -> convs [-> BEV] in one
    replay, no intermediate synchronisation); forward() only hands the points over and the backbone publishes voxel_features /
    voxel_coords afterwards."""

    def __init__(self, model_cfg=None, num_point_features=4, voxel_size=None, point_cloud_range=None, max_points_per_voxel=5,
                 max_voxels=40000, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_point_features = int(num_point_features)
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.max_points_per_voxel, self.max_voxels = int(max_points_per_voxel), int(max_voxels)
        r = np.asarray(self.point_cloud_range, dtype=np.float64)
        self.grid_size = np.round((r[3:6] - r[0:3]) / np.asarray(self.voxel_size, dtype=np.float64)).astype(np.int64).tolist()
        self._fused = False

    def get_output_feature_dim(self):
        return self.num_point_features

    def attach(self, backbone):
        backbone._ql_vfe = self
        self._fused = True
        return self

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        points = batch_dict['points']
        if self._fused:
            batch_dict['_ql_points'] = points if points.is_contiguous() else points.contiguous()
            batch_dict['voxel_features'] = None                    # published by the backbone call (one graph replay for both)
            batch_dict['voxel_coords'] = None
            return batch_dict
        B = int(batch_dict['batch_size'])
        feats, coords, npts, n_dev, _ = ops.voxelize_mean(points.contiguous(), self.point_cloud_range, self.voxel_size, self.grid_size, B,
                                                          self.max_points_per_voxel, B * self.max_voxels, max_voxels_per_frame=self.max_voxels)
        n = int(n_dev[0].item())
        batch_dict['voxel_features'], batch_dict['voxel_coords'], batch_dict['voxel_num_points'] = feats[:n], coords[:n], npts[:n]
        return batch_dict


class VoxelGeneratorWrapper:
    """GPU counterpart of pcdet/datasets/processor/data_processor.py:16-61 (same ctor keywords, `generate(points)`
    returns (voxel mean features, zyx coordinates, num_points) for ONE frame; the (V,T,F) padded tensor is never
    materialised because MeanVFE is fused)."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels):
        self.vsize, self.range = [float(v) for v in vsize_xyz], [float(v) for v in coors_range_xyz]
        self.nfeat, self.max_pts, self.max_voxels = int(num_point_features), int(max_num_points_per_voxel), int(max_num_voxels)
        r = np.asarray(self.range, dtype=np.float64)
        self.grid = np.round((r[3:6] - r[0:3]) / np.asarray(self.vsize, dtype=np.float64)).astype(np.int64).tolist()

    def generate(self, points: torch.Tensor):
        feats, coords, npts, n_dev, _ = ops.voxelize_mean(points.contiguous(), self.range, self.vsize, self.grid, 1, self.max_pts,
                                                          self.max_voxels, has_batch_col=False, n_feat=self.nfeat)
        n = int(n_dev[0].item())
        return feats[:n], coords[:n, 1:], npts[:n]


# ---------------------------------------------------------------------------------------------------------- BEV
class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES

    def attach(self, backbone, dtype=torch.float32):
        """Optional: let `backbone`'s engine produce the BEV map inside its CUDA graph (one kernel, no memset + scatter + permute);
        forward() then returns that buffer.  dtype float32 = the reference's; float16 halves the largest write of the path."""
        backbone.engine_bev_dtype = dtype
        return self

    def forward(self, batch_dict):
        t = batch_dict['encoded_spconv_tensor']
        fused = getattr(t, "_ql_spatial_features", None)
        if fused is not None:
            batch_dict['spatial_features'] = fused
            batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
            return batch_dict
        spatial_features = t.dense()
        N, C, D, H, W = spatial_features.shape
        batch_dict['spatial_features'] = spatial_features.view(N, C * D, H, W)
        batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
        return batch_dict
