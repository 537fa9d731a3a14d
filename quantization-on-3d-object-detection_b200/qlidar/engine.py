"""Sync-free, CUDA-graph-captured schedule of the whole hot path:

    points (P, 1+F) --voxelize+meanVFE--> stem --> [rulebook, fused conv]* --> BEV densify --> spatial_features

built from an (optionally q_conv3d-quantised) VoxelResBackBone8x / VoxelBackBone8x module tree.  Row counts never
leave the device (every kernel takes a capacity + a device-side count), buffers are allocated once, BatchNorm is
folded into the conv epilogue, and the ~45 launches of one forward replay as one graph.

Reference schedule being replaced: Detector.forward's module loop (pcdet/models/detectors/centerpoint.py:9-11)
over MeanVFE (mean_vfe.py:25-29), VoxelResBackBone8x.forward (spconv_backbone.py:243-295; per conv QConvNd.forward,
quant/quant.py:36-58, then BatchNorm1d/ReLU/add as separate ATen kernels) and HeightCompression
(height_compression.py:20-24), with voxelisation done on the CPU by DataLoader workers (data_processor.py:151-153)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import gc
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import QlidarError
from .quant import QConvNd, SQConv3d
from .tensor_quant import quant_scale
from .sparse import SparseConvolution, SparseSequential, SparseConvTensor, _round_up
from .backbones import SparseBasicBlock


@dataclass
class Layer:
    name: str
    kind: str                      # 'stem' | 'f16' | 'i8' | 'cw' | 'row'
    cin: int
    cout: int
    ksize: tuple
    stride: tuple
    pad: tuple
    subm: bool
    relu: bool
    residual: bool                 # add the block input (SparseBasicBlock.forward, spconv_backbone.py:64)
    block_input: bool              # this layer's input is a block input (keep it for the residual)
    tap: Optional[str] = None      # publish the output as multi_scale_3d_features[tap]
    w: Optional[torch.Tensor] = None
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    act_bits: int = 16
    act_amax: Optional[torch.Tensor] = None      # calibrated amax (static mode) or None (dynamic)
    stage_in: int = 0
    stage_out: int = 0
    rb_key: tuple = ()
    out: Optional[torch.Tensor] = None
    q_buf: Optional[torch.Tensor] = None
    act_scale: Optional[torch.Tensor] = None
    in_absmax: Optional[torch.Tensor] = None
    out_absmax: Optional[torch.Tensor] = None
    out_q: Optional[torch.Tensor] = None         # static mode: this layer's epilogue also writes the NEXT layer's int8 codes here
    out_qscale: Optional[torch.Tensor] = None    # ... with these per-channel quantisation scales (bound / calibrated amax)
    fused_q: bool = False                        # this layer's codes come from the previous layer's epilogue
    merge_srcs: Optional[list] = None            # 'merge' (VoxelNeXt bev_out): the layers whose outputs are merged, and their
    merge_scales: Optional[list] = None          # coordinate scales onto the target grid
    ws: Optional[torch.Tensor] = None
    smooth: Optional[torch.Tensor] = None        # SmoothQuant: per-input-channel smoothing scale (device)
    sq_dynamic: bool = False                     # ... re-derived (with the weight image and the scale) inside the graph per forward
    sq_alpha: float = 0.5
    w_f32: Optional[torch.Tensor] = None
    w_ic: Optional[torch.Tensor] = None
    bn_a: Optional[torch.Tensor] = None


@dataclass
class Stage:
    grid: tuple                    # (B, D, H, W)
    cap: int
    coords: Optional[torch.Tensor] = None
    n_dev: Optional[torch.Tensor] = None
    table: Optional[torch.Tensor] = None
    rank: Optional[object] = None  # ops.RankIndex of a key-sorted stage (every stage a strided conv produced)


def _bn_fold(bn: nn.BatchNorm1d):
    """Eval-mode BatchNorm1d as y = a*x + b, on the HOST in fp32 (the oracle's mirror reproduces these bits: every scale that feeds
    an int8 quantiser is computed with correctly rounded host arithmetic, never with a device reciprocal-multiply)."""
    g, be = bn.weight.detach().float().cpu(), bn.bias.detach().float().cpu()
    mu, var = bn.running_mean.detach().float().cpu(), bn.running_var.detach().float().cpu()
    a = g / torch.sqrt(var + bn.eps)                                                          # eps read from the module
    return a, be - a * mu


def _unwrap(m):
    if isinstance(m, SQConv3d):
        return m.spconv3d, m
    if isinstance(m, QConvNd):
        return m.module, m
    if isinstance(m, SparseConvolution):
        return m, None
    raise QlidarError(f"unsupported module in a conv slot: {type(m).__name__}")


class BackboneEngine:
    def __init__(self, backbone: nn.Module, batch_size: int, max_voxels: int, *, max_points: Optional[int] = None,
                 pc_range=None, voxel_size=None, max_pts_per_voxel: int = 5, n_point_features: Optional[int] = None,
                 bev: bool = True, bev_dtype=torch.float16, use_graph: bool = True, stage_cap_ratio: float = 1.0, stage_caps=None,
                 device="cuda", max_voxels_per_frame: int = 0, group_rows="auto", overlap_rulebooks: bool = True, sort_stage1: bool = True,
                 sorted_voxelizer: Optional[bool] = None):
        self.dev = torch.device(device)
        self.B = int(batch_size)
        self.max_voxels = int(max_voxels)                 # capacity, total over the batch
        self.max_voxels_per_frame = int(max_voxels_per_frame)   # the reference's MAX_NUMBER_OF_VOXELS (0 = only the batch capacity)
        self._sorted_voxelizer_arg = sorted_voxelizer
        self.max_points = max_points
        self.pc_range, self.voxel_size, self.max_pts = pc_range, voxel_size, int(max_pts_per_voxel)
        self.bev, self.bev_dtype, self.use_graph = bev, bev_dtype, use_graph
        # Grouped submanifold rulebooks on the ranked stages (rows binned by line key: fewer live (tile, offset) slabs, but the
        # gathers lose the L1 locality of key-consecutive rows).  True / False, or "auto": only where the layers STREAM their
        # weights (every skipped slab then also saves a 16-32 KB weight copy, which is what paces those launches) -- measured
        # +9-13 % on the C >= 64 submanifold layers, -7 % on the C = 32 residual layers.  The binning itself runs on
        # the side stream with the other rulebook work.
        self.group_rows = group_rows
        # Rulebooks depend on coordinates only, never on features: in graph mode every build after the first runs on a side
        # stream (a fork/join inside the captured graph) while the main stream runs the stem and the convs of earlier stages
        self.overlap_rulebooks = bool(overlap_rulebooks)
        self._side = None
        # The voxeliser numbers voxels in first-touch order (the reference's CPU voxeliser order, random for shuffled points): a
        # 128-row tile then touches all 27 kernel offsets and stage 1 needs the hash.  sort_stage1 renumbers the kept voxels in
        # ascending-key order (1x1x1 bitmap build + row permutation), which makes stage 1 a ranked stage like the others: its
        # tiles are runs along x, its rulebooks come from the rank index, and x_conv1 rows are returned in key order.
        self.sort_stage1 = bool(sort_stage1)
        self.sparse_shape = list(backbone.sparse_shape)
        self.grid_xyz = [self.sparse_shape[2], self.sparse_shape[1], self.sparse_shape[0] - 1]
        self.layers: List[Layer] = []
        self.stages: List[Stage] = [Stage((self.B, *self.sparse_shape), self.max_voxels)]
        self.rulebooks: Dict[tuple, torch.Tensor] = {}
        self._cap_ratio = stage_cap_ratio
        self._stage_caps = list(stage_caps) if stage_caps is not None else None
        self._last_from_points = False
        self._graph = None
        self._timing = None
        self.kernels_per_forward = 0
        self._compile(backbone)
        self.nfeat = n_point_features or self.layers[0].cin
        self._allocate()

    # ------------------------------------------------------------------ compile the module tree into a layer list
    def _add_conv(self, name, convm, bn, relu, residual=False, block_input=False):
        conv, qw = _unwrap(convm)
        a, b = _bn_fold(bn)
        K = int(np.prod(conv.kernel_size))
        cin, cout = conv.in_channels, conv.out_channels
        bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout)
        # 2-D convs (VoxelNeXt's tail) run as D = 1 slabs: zyx triples with a unit z axis
        L = Layer(name=name, kind="f16", cin=cin, cout=cout, ksize=tuple(conv._k3()), stride=tuple(conv._s3()),
                  pad=tuple(conv._p3()) if not conv.subm else tuple(k // 2 for k in conv._k3()), subm=conv.subm, relu=relu,
                  residual=residual, block_input=block_input)
        sq = isinstance(qw, SQConv3d)
        if sq:
            # SmoothQuant W8A8 (SQConv3d): static (calibrated per-channel amax) -> smoothed codes prepared once on the host;
            # dynamic -> re-prepared on the device inside the graph from the producing layer's abs-max (ql_sq_prepare_weights);
            # the scalar variant (no scaling_factor) cancels in both quantisers and is plain W8A8 per-tensor
            L.act_bits = qw.act_quant.num_bits
            L.act_amax = None if qw.act_amax is None else qw.act_amax.detach().float().reshape(-1).cpu()
            codes = None
        elif qw is not None:
            codes, amax_w, bound = qw.weight_codes()
            w_scale = amax_w.float().cpu() / bound                     # host: a true division
            L.act_bits = qw.act_quant.num_bits
            L.act_amax = None if qw.act_quant.amax is None else qw.act_quant.amax.detach().float().reshape(-1).cpu()
        else:
            codes, w_scale = None, torch.ones(cout)
        if cin < 8:
            # raw point features must stay fp32 (metres in fp16 lose centimetres): SIMT stem, weights fake-quantised if wrapped
            if qw is not None and L.act_bits <= 8:
                raise QlidarError("8-bit activation quantisation of conv_input is served by the module path (QConvNd), not the engine")
            if cout not in (16, 32):
                raise QlidarError("stem conv supports 16/32 output channels")
            wf = conv.weight.detach().float().reshape(cout, K, cin) if codes is None else codes.cpu() * w_scale.view(-1, 1, 1)
            L.kind = "stem"
            L.w = wf.permute(1, 2, 0).contiguous().to(self.dev)                       # (K, cin, cout)
            L.scale = a.cpu().to(self.dev)
            L.shift = (bias.cpu() * a.cpu() + b.cpu()).to(self.dev)
        else:
            if cin % 16 or cout % 16:
                raise QlidarError("engine layers need channel counts that are multiples of 16")
            if sq:
                if cin % 16 or cout % 16:
                    raise QlidarError("SQConv3d engine layers need channel counts that are multiples of 16")
                L.kind = "sq"
                L.shift = (bias.cpu() * a.cpu() + b.cpu()).to(self.dev)
                if qw.scaling_factor is None or L.act_amax is not None:
                    s_host = None if qw.scaling_factor is None else qw.smoothing_scale(L.act_amax.expand(cin) if L.act_amax.numel() == 1 else L.act_amax)
                    packed, _, _, w_sc, _ = qw._prepare(torch.device("cpu"), s_host)
                    L.w = packed.to(self.dev)
                    L.scale = (w_sc.cpu() * a.cpu()).to(self.dev)
                    L.smooth = None if s_host is None else s_host.float().contiguous().to(self.dev)
                    L.sq_dynamic = False
                else:
                    L.sq_dynamic = True
                    L.sq_alpha = float(qw.scaling_factor)
                    wf = conv.weight.detach().float().reshape(cout, K, cin).contiguous().to(self.dev)
                    L.w_f32, L.w_ic = wf, wf.abs().amax(dim=(0, 1)).contiguous()
                    L.bn_a = a.cpu().contiguous().to(self.dev)
                    L.w = torch.zeros(int(ops.lib().ql_packed_weight_bytes(cin, cout, K, ops.QL_S8)), dtype=torch.uint8, device=self.dev)
                    L.scale = torch.zeros(cout, dtype=torch.float32, device=self.dev)
                    L.smooth = torch.ones(cin, dtype=torch.float32, device=self.dev)
                self.layers_sq = getattr(self, "layers_sq", 0) + 1
                L.stage_in = len(self.stages) - 1
                self._finish_layer(L, conv, K)
                return L
            if qw is None or L.act_bits > 8:
                L.kind = "f16"
                wt = conv.weight.detach().float().reshape(cout, K, cin).cpu() if codes is None else codes.cpu()
                L.w = ops.pack_weights(wt.to(torch.float16)).to(self.dev)
            elif qw.per_row:
                L.kind = "row"                     # GQConv3d: per-voxel-row fake-quant (no abs-max input), then the cw kernel
                L.w = ops.pack_weights(codes.cpu().to(torch.float16)).to(self.dev)
            elif qw.cw:
                L.kind = "cw"
                L.w = ops.pack_weights(codes.cpu().to(torch.float16)).to(self.dev)
            else:
                L.kind = "i8"
                L.w = ops.pack_weights(codes.cpu().to(torch.int8)).to(self.dev)
            L.scale = (w_scale * a.cpu()).to(self.dev)
            L.shift = (bias.cpu() * a.cpu() + b.cpu()).to(self.dev)
        L.stage_in = len(self.stages) - 1
        self._finish_layer(L, conv, K)
        return L

    def _finish_layer(self, L, conv, K):
        """Stage bookkeeping: which stage the layer reads / writes and which rulebook it uses."""
        cin, cout = L.cin, L.cout
        if conv.subm:
            L.stage_out = L.stage_in
            # grouped (rows binned by line key) or plain rulebook; the SIMT stem always takes the plain one, so a grouped
            # stage 1 builds two rulebooks (the second on the side stream, under the stem)
            grp = self.group_rows
            if isinstance(grp, str):
                streamed = L.kind != "stem" and ops.weights_streamed(cin, cout, K, torch.int8 if L.kind in ("i8", "sq") else torch.float16)
                grp = streamed                                  # "auto": only where the weights are streamed
            ranked = L.stage_in > 0 or self.sort_stage1
            grp = bool(grp) and ranked and L.kind != "stem" and L.ksize[2] <= 31
            L.rb_key = ("subm", L.stage_in, L.ksize, "grouped" if grp else "plain")
        else:
            g = self.stages[-1].grid
            od, oh, ow = ops.conv_out_shape(g[1:], L.ksize, L.stride, L.pad)
            cap = max(128, int(self.stages[-1].cap * self._cap_ratio))
            if self._stage_caps is not None and len(self.stages) < len(self._stage_caps):
                cap = int(self._stage_caps[len(self.stages)])
            elif self.merge_stage is not None and L.stage_in >= self.merge_stage:
                # a regular (dilating) 2-D conv on the merged sites: at most k*k outputs per input, never more than the grid
                cap = int(min(g[0] * od * oh * ow, self.stages[-1].cap * max(2.0, self._cap_ratio)))
            self.stages.append(Stage((g[0], od, oh, ow), cap))
            L.stage_out = len(self.stages) - 1
            L.rb_key = ("strided", L.stage_in, L.ksize, L.stride, L.pad)
        self.layers.append(L)

    def _compile(self, bb):
        def seq_conv_bn_relu(prefix, seq):
            mods = list(seq._modules.values())
            if not (len(mods) == 3 and isinstance(mods[1], nn.BatchNorm1d) and isinstance(mods[2], nn.ReLU)):
                raise QlidarError(f"{prefix}: expected conv, BatchNorm1d, ReLU")
            return self._add_conv(prefix + ".0", mods[0], mods[1], True)

        seq_conv_bn_relu("conv_input", bb.conv_input)
        stage_last = {}
        self.merge_stage = None
        i = 1
        while hasattr(bb, f"conv{i}"):
            top = getattr(bb, f"conv{i}")
            last = None
            for cname, child in top._modules.items():
                p = f"conv{i}.{cname}"
                if isinstance(child, SparseBasicBlock):
                    if child.downsample is not None:
                        raise QlidarError("downsample branches are not used by the reference backbones")
                    self._add_conv(p + ".conv1", child.conv1, child.bn1, True, block_input=True)
                    last = self._add_conv(p + ".conv2", child.conv2, child.bn2, True, residual=True)
                elif isinstance(child, SparseSequential):
                    last = seq_conv_bn_relu(p, child)
                else:
                    raise QlidarError(f"{p}: unsupported child {type(child).__name__}")
            if i <= 4:
                last.tap = f"x_conv{i}"
            stage_last[i] = last
            i += 1
        if hasattr(bb, "shared_conv"):
            # VoxelResBackBone8xVoxelNeXt.forward (spconv_backbone_voxelnext.py:194-225): stages 5 / 6 land on the stage-4 grid
            # (indices * 2 / * 4), the three site lists are merged in 2-D (z dropped, duplicates summed), then the 2-D tail
            if not all(k in stage_last for k in (4, 5, 6)):
                raise QlidarError("VoxelNeXt backbone: conv4, conv5 and conv6 expected")
            self._add_merge([stage_last[4], stage_last[5], stage_last[6]], [1, 2, 4])
            seq_conv_bn_relu("conv_out", bb.conv_out)
            mods = list(bb.shared_conv._modules.values())
            if not (len(mods) == 3 and isinstance(mods[1], nn.BatchNorm1d) and isinstance(mods[2], nn.ReLU)):
                raise QlidarError("shared_conv: expected conv, BatchNorm1d, ReLU")
            self._add_conv("shared_conv.0", mods[0], mods[1], True)
            self.bev = False                                            # fully sparse: there is no dense BEV hand-off
        else:
            seq_conv_bn_relu("conv_out", bb.conv_out)
        self.num_convs = len(self.layers)

    def _add_merge(self, srcs, scales):
        s4 = self.stages[srcs[0].stage_out]
        B, _, H, W = s4.grid
        cap = sum(self.stages[L.stage_out].cap for L in srcs)
        L = Layer(name="bev_out", kind="merge", cin=srcs[0].cout, cout=srcs[0].cout, ksize=(1, 1, 1), stride=(1, 1, 1), pad=(0, 0, 0),
                  subm=False, relu=False, residual=False, block_input=False)
        L.merge_srcs, L.merge_scales = list(srcs), list(scales)
        L.stage_in = srcs[0].stage_out
        self.stages.append(Stage((B, 1, H, W), cap))
        L.stage_out = len(self.stages) - 1
        L.rb_key = None
        self.merge_stage = L.stage_out
        self.layers.append(L)

    # ------------------------------------------------------------------ static buffers
    def _allocate(self):
        dev = self.dev
        z = lambda *s, dt=torch.float16: torch.zeros(s, dtype=dt, device=dev)
        s0 = self.stages[0]
        # every stage's (kept, found) row counts in ONE tensor: the host reads them with a single copy
        # (kept, found) per stage, then the voxels found per frame by the key-sorted voxeliser: ONE tensor, one D2H read per forward
        self.counts_all = z(2 * len(self.stages) + self.B, dt=torch.int32)
        self.counts_dev = self.counts_all[:2 * len(self.stages)].view(len(self.stages), 2)
        self.frame_counts = self.counts_all[2 * len(self.stages):]
        s0.coords = z(s0.cap, 4, dt=torch.int32)
        s0.n_dev = self.counts_dev[0]
        # rows padded to 8 floats (32 bytes): the stem conv fetches a neighbour row with one 256-bit load
        self.vox_feats = z(s0.cap, 8 if self.nfeat <= 8 else self.nfeat, dt=torch.float32)
        self.vox_npts = z(s0.cap, dt=torch.int32)
        if self.max_points is not None:
            self.points = torch.full((self.max_points, 1 + self.nfeat), 1e30, dtype=torch.float32, device=dev)
            self.points[:, 0] = 0
            s0.table = z(ops.hash_capacity(self.max_points), dt=torch.int64)
            self.vox_ws = z(int(ops.lib().ql_voxelize_workspace_bytes(self.max_points, s0.cap, self.nfeat, self.max_pts)), dt=torch.uint8)
        else:
            s0.table = z(ops.hash_capacity(s0.cap), dt=torch.int64)
        if self.sort_stage1:
            self.vox_feats_ft = torch.zeros_like(self.vox_feats)           # first-touch order, as the voxeliser writes them
            self.coords_ft = z(s0.cap, 4, dt=torch.int32)
            self.n_ft = z(2, dt=torch.int32)
            w0 = z(ops.rulebook_strided_workspace_bytes(s0.grid, 1, 1, 0), dt=torch.uint8)
            s0.rank = ops.rulebook_strided_index(s0.grid, 1, 1, 0, w0)
            self.sort_src = z(s0.cap, dt=torch.int32)                      # sorted row -> first-touch row
        # key-sorted voxelisation straight from the points (csrc/rulebook.cu, ql_voxelize_sorted_*): the default from-points front end.
        # It cannot apply the per-frame voxel cap (a first-touch notion): frame_counts is checked after every forward and the engine
        # falls back to the hash voxeliser + renumbering for good once a frame exceeds it (frame_cap_exceeded()).
        # sorted_voxelizer=None: on when no per-frame cap was asked for (nothing to fall back from); True: on, and the CALLER checks
        # frame_cap_exceeded() after a forward (the plugin path and bench.py do); False: always the hash voxeliser.
        want = (not self.max_voxels_per_frame) if self._sorted_voxelizer_arg is None else bool(self._sorted_voxelizer_arg)
        self.sorted_voxelizer = bool(want and self.sort_stage1 and self.max_points is not None and self.max_pts > 0 and
                                     os.environ.get("QL_SORTED_VOXELIZER", "1") != "0")
        if self.sorted_voxelizer:
            self.vs_ws = z(int(ops.lib().ql_voxelize_sorted_workspace_bytes(self.max_points, s0.cap, self.max_pts)), dt=torch.uint8)
        for i, st in enumerate(self.stages[1:], start=1):
            st.coords = z(st.cap, 4, dt=torch.int32)
            st.n_dev = self.counts_dev[i]
            st.table = None                                   # key-sorted stages are indexed by rank (bitmap + prefix), not by hash
        # one strided-rulebook workspace per produced stage: its bitmap + prefix is that stage's rank index, used by the
        # submanifold rulebooks (and the BEV hand-off) that follow instead of a hash table
        for L in self.layers:
            if L.kind == "merge":
                continue                                          # its stage's rank index lives in the merge workspace (below)
            if not L.subm and self.stages[L.stage_out].rank is None:
                gi = self.stages[L.stage_in].grid
                w = z(ops.rulebook_strided_workspace_bytes(gi, L.ksize, L.stride, L.pad), dt=torch.uint8)
                self.stages[L.stage_out].rank = ops.rulebook_strided_index(gi, L.ksize, L.stride, L.pad, w)
        self.kmasks: Dict[tuple, torch.Tensor] = {}
        self.row_perms: Dict[tuple, Optional[torch.Tensor]] = {}
        self.group_ws = None
        n_abs = sum(L.cout for L in self.layers) + 256
        self.absmax_pool = z(n_abs, dt=torch.float32)
        off = 0
        prev_absmax = None
        for L in self.layers:
            so = self.stages[L.stage_out]
            if L.kind == "merge":
                L.ws = z(int(ops.lib().ql_bev_merge2d_workspace_bytes(so.grid[0], so.grid[2], so.grid[3], so.cap, L.cout, ops.QL_F16)), dt=torch.uint8)
                # the merge numbers its sites with the same bitmap + popcount prefix as a strided build (key (b*H + y)*W + x == the
                # 3-D key with D = 1): that pair, at the head of its workspace, is the merged stage's rank index
                n_words = (so.grid[0] * so.grid[2] * so.grid[3] + 31) // 32
                so.rank = ops.RankIndex(L.ws, L.ws.data_ptr(), L.ws.data_ptr() + ((n_words * 4 + 255) & ~255), n_words)
            elif L.rb_key not in self.rulebooks:
                K = int(np.prod(L.ksize))
                self.rulebooks[L.rb_key] = z(ops.num_tiles(so.cap), K, ops.TILE_M, dt=torch.int32)
                self.kmasks[L.rb_key] = z(ops.num_tiles(so.cap), ops.mask_words(K), dt=torch.int32)
                self.row_perms[L.rb_key] = None
                if L.rb_key[-1] == "grouped":
                    self.row_perms[L.rb_key] = z(ops.num_tiles(so.cap) * ops.TILE_M, dt=torch.int32)
                    nb = int(ops.lib().ql_rulebook_group_workspace_bytes(so.cap))
                    if self.group_ws is None or self.group_ws.numel() < nb:
                        self.group_ws = z(nb, dt=torch.uint8)
            L.out = ops.zero_led_rows(so.cap, L.cout, torch.float16, dev)          # gathered by the next conv: zero-row contract
            L.out_absmax = self.absmax_pool[off:off + L.cout]
            off += L.cout
            L.in_absmax = prev_absmax
            prev_absmax = L.out_absmax
            if L.kind in ("i8", "sq"):
                L.q_buf = ops.zero_led_rows(self.stages[L.stage_in].cap, L.cin, torch.int8, dev)
                L.act_scale = z(1, dt=torch.float32)
            elif L.kind in ("cw", "row"):
                L.q_buf = ops.zero_led_rows(self.stages[L.stage_in].cap, L.cin, torch.float16, dev)
            if L.act_amax is not None:
                L.act_amax = (L.act_amax.expand(L.cin) if L.act_amax.numel() == 1 else L.act_amax).contiguous().to(dev)
        # static calibration (collect_stats / compute_amax, quant/quantize.py:175-207): a per-tensor-quantised layer whose input
        # is the previous layer's output gets its int8 codes straight from that layer's epilogue (out_q) -- no quantise pass,
        # no abs-max pass; its de-quantisation scale amax/bound is a constant
        for i, L in enumerate(self.layers):
            L.fused_q = False
            if L.kind == "i8" and L.act_amax is not None:
                bound = float(2 ** (L.act_bits - 1) - 1)
                amax = L.act_amax.float().cpu()
                L.act_scale.copy_((amax.max() / bound).reshape(1))                 # host: a true division
                prev = self.layers[i - 1] if i > 0 else None
                if prev is not None and prev.kind in ("f16", "i8", "cw", "row") and L.act_bits == 8:
                    prev.out_qscale = quant_scale(amax, bound).contiguous().to(dev)
                    prev.out_q = L.q_buf
                    L.fused_q = True
        last = self.stages[-1]
        if self.bev:
            B, D, H, W = last.grid
            self.spatial_features = z(B, self.layers[-1].cout * D, H, W, dt=self.bev_dtype)
            self.bev_ws = z(int(ops.lib().ql_bev_densify_workspace_bytes(B, D, H, W)), dt=torch.uint8)
        self._need_absmax = [i for i, L in enumerate(self.layers[:-1])
                             if self.layers[i + 1].kind in ("i8", "cw", "sq") and self.layers[i + 1].act_amax is None]

    # ------------------------------------------------------------------ the schedule
    def _op(self, label, n_kernels, fn, *a, **kw):
        """Run one C-ABI op; when timing is on, bracket it with CUDA events on the launching (current) stream."""
        self.kernels_per_forward += n_kernels
        if self._timing is None:
            return fn(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        self._timing.append((label, e0, e1))
        return r

    def _build_rulebook(self, L):
        si, so = self.stages[L.stage_in], self.stages[L.stage_out]
        nbr, kmask, perm = self.rulebooks[L.rb_key], self.kmasks[L.rb_key], self.row_perms[L.rb_key]
        if perm is not None:
            # kernels: line keys + histogram, bin scan, slot scatter, ranked pairs
            self._op("rulebook_subm:" + L.name, 4, ops.rulebook_subm_ranked_grouped, si.coords, si.n_dev, si.grid, L.ksize, si.rank,
                     nbr=nbr, kmask=kmask, row_perm=perm, workspace=self.group_ws)
        elif L.subm and si.rank is not None:
            self._op("rulebook_subm:" + L.name, 1, ops.rulebook_subm_ranked, si.coords, si.n_dev, si.grid, L.ksize, si.rank, nbr=nbr, kmask=kmask)
        elif L.subm:
            self._op("rulebook_subm:" + L.name, 1, ops.rulebook_subm, si.coords, si.n_dev, si.grid, L.ksize, si.table, nbr=nbr, kmask=kmask)
        else:
            # kernels: mark, popc, scan, emit + (fill, scatter, kmask | ranked pairs)
            self._op("rulebook_strided:" + L.name, 5 if si.rank is not None else 7, ops.rulebook_strided, si.coords, si.n_dev, si.grid, L.ksize, L.stride, L.pad,
                     so.cap, out=(so.coords, so.n_dev, None, nbr), workspace=so.rank.workspace, kmask=kmask, in_index=si.rank)

    def _fork_rulebooks(self, with_first=False):
        """Issue every rulebook build except the first layer's on the side stream, in layer order (a strided build produces the
        stage its successors index); returns {rb_key: event}.  with_first (the from-points schedule): the fork happens right
        after the voxeliser's coordinate passes, and the side stream also renumbers stage 1 (coordinates only; event
        "sorted") and builds the first rulebook, all under the voxeliser's feature passes on the main stream."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        fork = torch.cuda.Event()
        fork.record(main)
        self._side.wait_event(fork)
        ready, seen = {}, (set() if with_first else {self.layers[0].rb_key})
        with torch.cuda.stream(self._side):
            if with_first and not self.sorted_voxelizer:
                s0 = self.stages[0]
                self._op("sort_stage1", 5, ops.renumber_by_key, self.coords_ft, self.n_ft, s0.grid, s0.rank.workspace, out_coords=s0.coords,
                         n_out_dev=s0.n_dev, src_row=self.sort_src)
                ev = torch.cuda.Event()
                ev.record(self._side)
                ready["sorted"] = ev
            for L in self.layers:
                if L.rb_key is None or L.rb_key in seen or (self.merge_stage is not None and L.stage_in >= self.merge_stage):
                    continue                  # (the 2-D tail's rulebooks depend on the merge, which runs on the main stream)
                seen.add(L.rb_key)
                self._build_rulebook(L)
                ev = torch.cuda.Event()
                ev.record(self._side)
                ready[L.rb_key] = ev
        return ready

    def _run_backbone(self, ready=None, sorted_input=False):
        built = set()
        if self.sort_stage1 and ready is None and not sorted_input:
            s0 = self.stages[0]
            # kernels: mark, popc, scan, prefix, assign (coordinates + feature rows move to their rank)
            self._op("sort_stage1", 5, self._sort_stage1, s0)
        x = self.vox_feats
        block_in = None
        if self._need_absmax:
            self.absmax_pool.zero_()
        overlap = self.overlap_rulebooks and self._timing is None and len({L.rb_key for L in self.layers if L.rb_key is not None}) > 1
        if ready is None:
            ready = self._fork_rulebooks() if overlap else {}
        else:
            overlap = True
        for i, L in enumerate(self.layers):
            si, so = self.stages[L.stage_in], self.stages[L.stage_out]
            absmax = L.out_absmax if i in self._need_absmax else None
            if L.kind == "merge":
                # VoxelNeXt bev_out: stages 4 / 5 / 6 onto the stage-4 grid, z dropped, duplicates summed (kernels: 3 marks, popc, scan,
                # emit, 3 adds, fp32 -> fp16)
                segs = [(m.out, self.stages[m.stage_out].coords, self.stages[m.stage_out].n_dev, sc) for m, sc in zip(L.merge_srcs, L.merge_scales)]
                self._op("bev_merge2d", 10, ops.bev_merge2d_multi, segs, (so.grid[0], so.grid[2], so.grid[3]), so.cap, L.out, so.coords, so.n_dev, L.ws)
                if absmax is not None:
                    self._op("absmax:" + L.name, 1, ops.absmax_cols, L.out, so.n_dev, absmax)
                x = L.out
                continue
            nbr, kmask, perm = self.rulebooks[L.rb_key], self.kmasks[L.rb_key], self.row_perms[L.rb_key]
            if L.rb_key not in built:
                if L.rb_key in ready:
                    torch.cuda.current_stream().wait_event(ready[L.rb_key])      # built on the side stream
                else:
                    self._build_rulebook(L)
                built.add(L.rb_key)
            if L.block_input:
                block_in = x
            res = block_in if L.residual else None
            if L.kind == "stem":
                if perm is not None:
                    raise QlidarError("the stem conv does not take a grouped rulebook")
                self._op("stem:" + L.name, 1, ops.stem_conv, x, nbr, so.cap, so.n_dev, L.w, L.scale, L.shift, relu=L.relu, out=L.out, absmax=absmax, kmask=kmask)
            elif L.kind == "f16":
                self._op("conv:" + L.name, 1, ops.spconv_mma, x, nbr, so.cap, so.n_dev, L.cout, L.w, L.scale, L.shift, residual=res,
                         relu=L.relu, out=L.out, absmax=absmax, kmask=kmask, out_q=L.out_q, out_qscale=L.out_qscale, row_perm=perm)
            elif L.kind == "i8":
                if not L.fused_q:
                    am = L.act_amax if L.act_amax is not None else L.in_absmax
                    self._op("quantize:" + L.name, 1, ops.quantize_rows, x, am, ops.QL_Q_CODES_PER_TENSOR, L.act_bits, si.n_dev, out=L.q_buf,
                             act_scale=L.act_scale)
                self._op("conv:" + L.name, 1, ops.spconv_mma, L.q_buf, nbr, so.cap, so.n_dev, L.cout, L.w, L.scale, L.shift, act_scale=L.act_scale, residual=res,
                               relu=L.relu, out=L.out, absmax=absmax, kmask=kmask, out_q=L.out_q, out_qscale=L.out_qscale, row_perm=perm)
            elif L.kind == "sq":
                am = L.act_amax if L.act_amax is not None else L.in_absmax
                if L.sq_dynamic:
                    self._op("sq_prepare:" + L.name, 2, ops.sq_prepare_weights, L.w_f32, L.w_ic, am, L.sq_alpha, bn_scale=L.bn_a, smooth=L.smooth,
                             packed=L.w, scale=L.scale)
                self._op("quantize:" + L.name, 1, ops.quantize_rows, x, am, ops.QL_Q_CODES_PER_TENSOR, L.act_bits, si.n_dev, smooth=L.smooth, out=L.q_buf,
                         act_scale=L.act_scale)
                self._op("conv:" + L.name, 1, ops.spconv_mma, L.q_buf, nbr, so.cap, so.n_dev, L.cout, L.w, L.scale, L.shift, act_scale=L.act_scale, residual=res,
                               relu=L.relu, out=L.out, absmax=absmax, kmask=kmask, row_perm=perm)
            elif L.kind in ("cw", "row"):
                if L.kind == "row":
                    self._op("quantize:" + L.name, 1, ops.quantize_rows, x, None, ops.QL_Q_FAKE_PER_ROW, L.act_bits, si.n_dev, out=L.q_buf)
                else:
                    am = L.act_amax if L.act_amax is not None else L.in_absmax
                    self._op("quantize:" + L.name, 1, ops.quantize_rows, x, am, ops.QL_Q_FAKE_PER_CHANNEL, L.act_bits, si.n_dev, out=L.q_buf)
                self._op("conv:" + L.name, 1, ops.spconv_mma, L.q_buf, nbr, so.cap, so.n_dev, L.cout, L.w, L.scale, L.shift, residual=res, relu=L.relu, out=L.out,
                               absmax=absmax, kmask=kmask, out_q=L.out_q, out_qscale=L.out_qscale, row_perm=perm)
            x = L.out
        if overlap:
            torch.cuda.current_stream().wait_stream(self._side)               # join (every event was already waited for)
        if self.bev:
            last = self.stages[-1]
            if last.rank is not None:
                self._op("bev_densify", 1, ops.bev_densify_ranked, x, last.rank, last.n_dev, last.grid, out=self.spatial_features, workspace=self.bev_ws)
            else:
                self._op("bev_densify", 2, ops.bev_densify, x, last.table, last.grid, out=self.spatial_features, workspace=self.bev_ws)

    def _sort_stage1(self, s0):
        ops.renumber_by_key(self.coords_ft, self.n_ft, s0.grid, s0.rank.workspace, out_coords=s0.coords, n_out_dev=s0.n_dev,
                            src_row=self.sort_src, rows_in=self.vox_feats_ft, rows_out=self.vox_feats)

    def _run_from_points(self):
        s0 = self.stages[0]
        self.kernels_per_forward = 0
        out = (self.vox_feats_ft, self.coords_ft, self.vox_npts, self.n_ft, s0.table) if self.sort_stage1 else \
              (self.vox_feats, s0.coords, self.vox_npts, s0.n_dev, s0.table)
        vox = lambda label, nk, phase: self._op(label, nk, ops.voxelize_mean, self.points, self.pc_range, self.voxel_size, self.grid_xyz, self.B,
                                                 self.max_pts, s0.cap, out=out, workspace=self.vox_ws,
                                                 max_voxels_per_frame=self.max_voxels_per_frame, phase=phase)
        if self.sorted_voxelizer:
            vs = lambda label, nk, phase: self._op(label, nk, ops.voxelize_sorted, self.points, self.pc_range, self.voxel_size, self.grid_xyz, self.B,
                                                    self.max_pts, s0.cap, s0.rank.workspace, out=(self.vox_feats, s0.coords, self.vox_npts, s0.n_dev),
                                                    workspace=self.vs_ws, frame_counts=self.frame_counts, phase=phase)
            # kernels: mark, popc, scan, prefix, rank (+ frames) | select, mean
            vs("voxelize_mean", 6, "coords")
            if self.overlap_rulebooks and self._timing is None:
                ready = self._fork_rulebooks(with_first=True)
                vs("voxelize_mean", 2, "features")
                self._run_backbone(ready=ready)
            else:
                vs("voxelize_mean", 2, "features")
                self._run_backbone(sorted_input=True)
        elif self.sort_stage1 and self.overlap_rulebooks and self._timing is None:
            # coordinates are final after the numbering passes: renumbering, the first rulebook and every later rulebook run on the
            # side stream under the voxeliser's feature passes (point selection + means), then the feature rows move to their rank
            vox("voxelize_mean", 5 if self.max_voxels_per_frame else 4, "coords")
            ready = self._fork_rulebooks(with_first=True)
            vox("voxelize_mean", 3, "features")
            torch.cuda.current_stream().wait_event(ready["sorted"])
            self._op("sort_stage1", 1, ops.permute_rows, self.vox_feats_ft, self.sort_src, s0.n_dev, out=self.vox_feats)
            self._run_backbone(ready=ready)
        else:
            vox("voxelize_mean", 8 if self.max_voxels_per_frame else 7, "mean")
            self._run_backbone()

    def _replay(self, fn):
        if not self.use_graph:
            fn()
            return
        if self._graph is None or self._graph[0] is not fn.__func__:
            fn()                                                  # warm-up (first-use attribute setup) outside capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # no garbage collection while the stream is capturing: a collected engine of an earlier call frees pinned host buffers and
            # CUDA graphs, and the allocator's event bookkeeping for them is not capturable (seen once as
            # cudaErrorStreamCaptureInvalidated in the plugin tests, where engines of several backbones die in one process)
            gc_was_on = gc.isenabled()
            gc.collect()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    fn()
            finally:
                if gc_was_on:
                    gc.enable()
            self._graph = (fn.__func__, g)
        self._graph[1].replay()

    # ------------------------------------------------------------------ public entry points
    def set_points(self, points: torch.Tensor):
        """Copy a collated (P, 1+F) fp32 point array (host or device) into the static input buffer; rows past P are
        parked outside the point-cloud range so the voxeliser skips them."""
        if self.max_points is None:
            raise QlidarError("engine was built without max_points")
        P = points.shape[0]
        if P > self.max_points:
            raise QlidarError("more points than max_points")
        self.points[:P].copy_(points, non_blocking=True)
        prev = getattr(self, "_points_valid", self.max_points)      # rows >= prev are parked already (the buffer starts out parked)
        if P < prev:
            self.points[P:prev, 1:4] = 1e30
        self._points_valid = P

    def forward_points(self, points: Optional[torch.Tensor] = None):
        if points is not None:
            self.set_points(points)
        self._last_from_points = True
        self._replay(self._run_from_points)
        return self.outputs()

    def frame_cap_exceeded(self, counts_all_host: Optional[torch.Tensor] = None) -> bool:
        """True when the key-sorted voxeliser found more voxels in some frame than max_voxels_per_frame.  The reference drops such a
        frame's surplus voxels in FIRST-TOUCH order (its CPU voxeliser, data_processor.py:45-61), which only the hash voxeliser
        reproduces: call use_hash_voxelizer() and run the batch again.  counts_all_host: a host copy of self.counts_all the caller
        already made (otherwise one D2H read here)."""
        if not (self.sorted_voxelizer and self.max_voxels_per_frame and self._last_from_points):
            return False
        fc = (counts_all_host if counts_all_host is not None else self.counts_all.cpu())[2 * len(self.stages):]
        return bool((fc > int(self.max_voxels_per_frame)).any())

    def use_hash_voxelizer(self):
        """Switch the from-points front end back to the first-touch hash voxeliser + renumbering (and re-capture the graph)."""
        self.sorted_voxelizer = False
        self._graph = None

    def forward_voxels(self, voxel_features: torch.Tensor, voxel_coords: torch.Tensor):
        """batch_dict-style entry: already voxelised input (voxel_features (V,F) fp32, voxel_coords (V,4) int32/float)."""
        s0 = self.stages[0]
        V = voxel_features.shape[0]
        if V > s0.cap:
            raise QlidarError("more voxels than the engine capacity")
        vc = voxel_coords.int() if voxel_coords.dtype != torch.int32 else voxel_coords
        self._last_from_points = False
        if self.sort_stage1:
            self.vox_feats_ft[:V, :voxel_features.shape[1]].copy_(voxel_features)
            self.coords_ft[:V].copy_(vc)
            self.n_ft.fill_(V)
        else:
            self.vox_feats[:V, :voxel_features.shape[1]].copy_(voxel_features)
            s0.coords[:V].copy_(vc)
            s0.n_dev.fill_(V)
            ops.hash_build(s0.coords, s0.n_dev, s0.grid, table=s0.table)
        self._replay(self._run_backbone)
        return self.outputs()

    def outputs(self):
        out = {"stage_counts": [st.n_dev for st in self.stages], "encoded_features": self.layers[-1].out,
               "encoded_coords": self.stages[-1].coords, "encoded_grid": self.stages[-1].grid}
        if self.bev:
            out["spatial_features"] = self.spatial_features
        out["taps"] = {L.tap: (L.out, self.stages[L.stage_out]) for L in self.layers if L.tap}
        return out

    def counts(self) -> List[int]:
        return [int(v) for v in self.counts_dev[:, 0].cpu().tolist()]

    def overflowed(self) -> bool:
        """True when a stage found more active sites than its capacity (rows were dropped): raise the capacities."""
        c = self.counts_dev.cpu()
        if self.sort_stage1 and not (self.sorted_voxelizer and self._last_from_points):
            c[0] = self.n_ft.cpu()                                   # (kept, found) of the voxeliser, not of the renumbering build
        over = c[:, 1] > c[:, 0]
        if self.max_voxels_per_frame and self.stages[0].cap >= self.B * self.max_voxels_per_frame:
            over[0] = False           # voxels beyond a frame's MAX_NUMBER_OF_VOXELS are dropped by design; the batch capacity cannot overflow
        return bool(over.any().item())

    def profile_ops(self, from_points: bool = True, iters: int = 5, flush=None):
        """Eager (no graph) passes with every op bracketed by CUDA events on the launching stream.  A long spin kernel is
        queued first so the CPU enqueues the whole pass while the GPU is busy and the events see back-to-back execution.
        Returns {label: mean milliseconds}."""
        acc: Dict[str, float] = {}
        order: List[str] = []
        for it in range(iters + 1):
            if flush is not None:
                flush()
            self._timing = []
            torch.cuda._sleep(int(3e7))
            (self._run_from_points if from_points else self._run_backbone)()
            torch.cuda.synchronize()
            if it > 0:                                   # first pass is a warm-up
                for label, e0, e1 in self._timing:
                    if label not in acc:
                        acc[label] = 0.0
                        order.append(label)
                    acc[label] += e0.elapsed_time(e1) / iters
            self._timing = None
        return {k: acc[k] for k in order}

    # ------------------------------------------------------------------ accounting for the roofline report
    def layer_accounting(self):
        """Algorithmic bytes / flops per launch: call after a forward; reads counts and pair counts back."""
        counts = self.counts()
        acct = []
        pairs = {}
        for L in self.layers:
            if L.kind == "merge":
                continue
            n_in, n_out = counts[L.stage_in], counts[L.stage_out]
            K = int(np.prod(L.ksize))
            tiles = ops.num_tiles(n_out)
            # rows past n_out in the last tile are -1 already; tiles past the live count hold stale data -> recount live ones
            if L.rb_key not in pairs:
                pairs[L.rb_key] = (int((ops.expand_rulebook(self.rulebooks[L.rb_key][:tiles], self.kmasks[L.rb_key][:tiles]) >= 0).sum().item()),
                                   int(sum(bin(int(w) & 0xFFFFFFFF).count("1") for w in self.kmasks[L.rb_key][:tiles].flatten().cpu().tolist())))
            P, live_slabs = pairs[L.rb_key]
            b_act = 4 if L.kind == "stem" else (1 if L.kind == "i8" else 2)
            b_w = 4 if L.kind == "stem" else (1 if L.kind == "i8" else 2)
            by = n_in * L.cin * b_act + n_out * L.cout * 2 + K * L.cin * L.cout * b_w + 4 * P + (n_out * L.cout * 2 if L.residual else 0)
            acct.append(dict(name=L.name, kind=L.kind, n_in=n_in, n_out=n_out, pairs=P, cin=L.cin, cout=L.cout, K=K,
                             bytes_alg=by, flops_alg=2 * P * L.cin * L.cout, rulebook_bytes_read=live_slabs * 512, live_slabs=live_slabs, tiles=tiles))
        return acct
