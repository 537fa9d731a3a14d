"""the VoxelNeXt sparse head, inference path, behind the reference's interface
(pcdet/models/dense_heads/voxelnext_head.py:13-47 SeparateHead, :50-107 constructor, :418-488 generate_predicted_boxes, :523-559
forward; centernet_utils.py:243-354 _topk_1d / gather_feat_idx / decode_bbox_from_voxels_nuscenes).

Same constructor arguments, attribute names (`heads_list`, `class_id_mapping_each_head`, `separate_head_cfg`, ...), state-dict keys
(`heads_list.i.<name>.j.k.*`) and data_dict contract (`encoded_spconv_tensor` in, `final_box_dicts` out: one dict per frame with
pred_boxes / pred_scores / pred_labels, labels 1-based).  Training (target assignment, losses) and DOUBLE_FLIP test-time
augmentation are outside the path and raise.

What runs where: every branch of a SeparateHead is SubMConv2d(3x3) + BatchNorm1d + ReLU -> SubMConv2d(1x1); in eval mode the first
three are ONE sparse-conv launch (BN folded into the kernel's scale / shift, ReLU in its epilogue; all branches share one 3x3
rulebook -- they see the same coordinates) and the 1x1 is a second one with fp32 output.  Post-processing: per head
ql_voxelhead_decode (per-frame top-K over the (voxel, class) scores + decode + masks, 2 launches); then either the class-agnostic
rotated NMS of the CenterHead path or, with IOU_BRANCH, ql_voxelhead_class_split + one ql_nms_rotated per class
(rotate_class_specific_nms_iou) -- no boolean-mask indexing, no host sweep; the one host sync is the read of the kept counts when the
variable-length result tensors are cut."""
from __future__ import annotations

import copy
from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .center_head import _cfg
from .sparse import SparseConvTensor, SparseSequential, SubMConv2d


def _bn1d_affine(bn: nn.BatchNorm1d):
    a = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return a, bn.bias.detach().float() - a * bn.running_mean.detach().float()


def _conv_bn_relu(block: SparseSequential, x: SparseConvTensor) -> SparseConvTensor:
    """SparseSequential(SubMConv2d, BatchNorm1d, ReLU) in one launch (eval mode)."""
    conv, bn = block[0], block[1]
    rb = conv.get_rulebook(x)
    packed, ic_p, oc_p = conv._packed_weight(x.features.device)
    f = x.features if x.features.dtype == torch.float16 else x.features.to(torch.float16)
    if ic_p != conv.in_channels:
        f = torch.nn.functional.pad(f, (0, ic_p - conv.in_channels))
    a, b = _bn1d_affine(bn)
    scale = torch.ones(oc_p, dtype=torch.float32, device=f.device)
    shift = torch.zeros(oc_p, dtype=torch.float32, device=f.device)
    scale[:conv.out_channels] = a
    shift[:conv.out_channels] = b if conv.bias is None else a * conv.bias.detach().float() + b
    y = ops.spconv_mma(f.contiguous(), rb.nbr, rb.n_out, rb.n_out_dev, oc_p, packed, scale, shift, relu=True, out_dtype=torch.float16, kmask=rb.kmask)
    if oc_p != conv.out_channels:
        y = y[:, :conv.out_channels].contiguous()
    return x.replace_feature(y)


class SeparateHead(nn.Module):
    def __init__(self, input_channels, sep_head_dict, kernel_size, init_bias=-2.19, use_bias=False):
        super().__init__()
        self.sep_head_dict = sep_head_dict
        for cur_name in self.sep_head_dict:
            output_channels = self.sep_head_dict[cur_name]['out_channels']
            num_conv = self.sep_head_dict[cur_name]['num_conv']
            fc_list = [SparseSequential(SubMConv2d(input_channels, input_channels, kernel_size, padding=int(kernel_size // 2), bias=use_bias,
                                                   indice_key=cur_name),
                                        nn.BatchNorm1d(input_channels), nn.ReLU()) for _ in range(num_conv - 1)]
            fc_list.append(SubMConv2d(input_channels, output_channels, 1, bias=True, indice_key=cur_name + 'out'))
            fc = nn.Sequential(*fc_list)
            if 'hm' in cur_name:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, SubMConv2d):
                        nn.init.kaiming_normal_(m.weight.data)
                        if m.bias is not None:
                            nn.init.constant_(m.bias, 0)
            self.__setattr__(cur_name, fc)

    def forward(self, x):
        ret_dict = {}
        for cur_name in self.sep_head_dict:
            fc = getattr(self, cur_name)
            y = x
            for m in fc:
                fusable = (isinstance(m, SparseSequential) and len(m) == 3 and isinstance(m[0], SubMConv2d) and isinstance(m[1], nn.BatchNorm1d)
                           and isinstance(m[2], nn.ReLU) and not self.training and type(m[0]) is SubMConv2d)
                y = _conv_bn_relu(m, y) if fusable else m(y)
            ret_dict[cur_name] = y.features
        return ret_dict


class VoxelNeXtHead(nn.Module):
    def __init__(self, model_cfg, input_channels, num_class, class_names, grid_size, point_cloud_range, voxel_size,
                 predict_boxes_when_training=False):
        super().__init__()
        self.model_cfg = cfg = _cfg(dict(model_cfg))
        self.num_class = num_class
        self.grid_size = grid_size
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.voxel_size = [float(v) for v in voxel_size]
        self.feature_map_stride = cfg.TARGET_ASSIGNER_CONFIG.get('FEATURE_MAP_STRIDE', None)
        self.class_names = list(class_names)
        self.iou_branch = cfg.get('IOU_BRANCH', False) or False
        if self.iou_branch:
            self.rectifier = list(cfg.get('RECTIFIER'))
            n = cfg.POST_PROCESSING.NMS_CONFIG
            self.nms_configs = [_cfg(dict(NMS_TYPE=n.NMS_TYPE, NMS_THRESH=n.NMS_THRESH[i], NMS_PRE_MAXSIZE=n.NMS_PRE_MAXSIZE[i],
                                          NMS_POST_MAXSIZE=n.NMS_POST_MAXSIZE[i])) for i in range(num_class)]
        self.double_flip = cfg.get('DOUBLE_FLIP', False) or False
        self.class_names_each_head, self._class_maps = [], []
        for cur in cfg.CLASS_NAMES_EACH_HEAD:
            self.class_names_each_head.append([x for x in cur if x in self.class_names])
            self._class_maps.append(np.array([self.class_names.index(x) for x in cur if x in self.class_names], dtype=np.int32))
        assert sum(len(x) for x in self.class_names_each_head) == len(self.class_names), f'class_names_each_head={self.class_names_each_head}'
        self.separate_head_cfg = cfg.SEPARATE_HEAD_CFG
        self.heads_list = nn.ModuleList()
        for cur in self.class_names_each_head:
            head_dict = copy.deepcopy(dict(self.separate_head_cfg.HEAD_DICT))
            head_dict['hm'] = dict(out_channels=len(cur), num_conv=cfg.NUM_HM_CONV)
            self.heads_list.append(SeparateHead(input_channels=cfg.get('SHARED_CONV_CHANNEL', 128) or 128, sep_head_dict=head_dict,
                                                kernel_size=cfg.get('KERNEL_SIZE_HEAD', 3) or 3, init_bias=-2.19,
                                                use_bias=cfg.get('USE_BIAS_BEFORE_NORM', False) or False))
        self.predict_boxes_when_training = predict_boxes_when_training
        self.forward_ret_dict = {}
        self._dev_cache = {}

    @property
    def class_id_mapping_each_head(self):
        dev = next(self.parameters()).device
        key = str(dev)
        if key not in self._dev_cache:
            self._dev_cache[key] = [torch.from_numpy(m).to(dev) for m in self._class_maps]
        return self._dev_cache[key]

    # ------------------------------------------------------------------
    def generate_predicted_boxes(self, batch_size, pred_dicts: List[Dict[str, torch.Tensor]], voxel_indices, spatial_shape=None):
        if self.double_flip:
            raise NotImplementedError("DOUBLE_FLIP test-time augmentation is outside the accelerated path")
        p = self.model_cfg.POST_PROCESSING
        K = int(p.MAX_OBJ_PER_SAMPLE)
        idx = voxel_indices.to(torch.int32).contiguous()
        f = lambda t: t.float().contiguous()
        order = self.separate_head_cfg.HEAD_ORDER
        decoded = []
        for h, pd in enumerate(pred_dicts):
            vel = f(pd['vel']) if 'vel' in order and 'vel' in pd else None
            iou = f(pd['iou']) if self.iou_branch else None
            decoded.append(ops.voxelhead_decode(f(pd['hm']), f(pd['center']), f(pd['center_z']), f(pd['dim']), f(pd['rot']), vel, iou, idx, None,
                                                batch_size, K, self.feature_map_stride, self.voxel_size, self.point_cloud_range,
                                                p.POST_CENTER_LIMIT_RANGE, p.SCORE_THRESH, class_map=self.class_id_mapping_each_head[h]))
        ret = []
        if not self.iou_branch:
            n = p.NMS_CONFIG
            per_head = [ops.nms_rotated(b, s, l, c, float(n.NMS_THRESH), int(n.NMS_PRE_MAXSIZE), int(n.NMS_POST_MAXSIZE), label_offset=1,
                                        box_dim=b.shape[2]) for (b, s, l, _, c) in decoded]
            counts = torch.stack([o["keep_count"] for o in per_head], 0).cpu()                  # the one host sync
            for k in range(batch_size):
                parts = [(o["boxes"][k, :int(counts[i, k])], o["scores"][k, :int(counts[i, k])], o["labels"][k, :int(counts[i, k])].long())
                         for i, o in enumerate(per_head)]
                ret.append({'pred_boxes': torch.cat([q[0] for q in parts], 0), 'pred_scores': torch.cat([q[1] for q in parts], 0),
                            'pred_labels': torch.cat([q[2] for q in parts], 0)})
            return ret
        if len(decoded) != 1:
            # several heads: their decoded rows are merged per frame first (device-side, padded to the summed capacity)
            decoded = [self._merge_heads(decoded)]
        boxes, scores, labels, ious, count = decoded[0]
        if boxes.shape[1] > 1024:
            raise ops.QlidarError("IOU_BRANCH post-processing holds at most 1024 decoded boxes per frame")
        rect = torch.tensor(self.rectifier, dtype=torch.float32, device=boxes.device)
        cb, cs, cl, cc = ops.voxelhead_class_split(boxes, scores, labels, ious, count, rect)
        per_cls = [ops.nms_rotated(cb[c], cs[c], cl[c], cc[c], float(self.nms_configs[c].NMS_THRESH), int(self.nms_configs[c].NMS_PRE_MAXSIZE),
                                   int(self.nms_configs[c].NMS_POST_MAXSIZE), label_offset=1, box_dim=boxes.shape[2]) for c in range(self.num_class)]
        counts = torch.stack([o["keep_count"] for o in per_cls], 0).cpu()                       # the one host sync
        for k in range(batch_size):
            parts = [(o["boxes"][k, :int(counts[c, k])], o["scores"][k, :int(counts[c, k])], o["labels"][k, :int(counts[c, k])].long())
                     for c, o in enumerate(per_cls)]
            ret.append({'pred_boxes': torch.cat([q[0] for q in parts], 0), 'pred_scores': torch.cat([q[1] for q in parts], 0),
                        'pred_labels': torch.cat([q[2] for q in parts], 0)})
        return ret

    @staticmethod
    def _merge_heads(decoded):
        B = decoded[0][0].shape[0]
        tot = sum(d[0].shape[1] for d in decoded)
        dev = decoded[0][0].device
        bd = decoded[0][0].shape[2]
        mb = torch.zeros((B, tot, bd), dtype=torch.float32, device=dev)
        ms = torch.zeros((B, tot), dtype=torch.float32, device=dev)
        ml = torch.zeros((B, tot), dtype=torch.int32, device=dev)
        mi = torch.zeros((B, tot), dtype=torch.float32, device=dev)
        kc = torch.zeros((B,), dtype=torch.int32, device=dev)
        for (b, s, l, i, c) in decoded:
            P = b.shape[1]
            ar = torch.arange(P, device=dev)[None, :]
            sel = ar < c[:, None]
            dst = (kc[:, None] + ar).long().clamp(max=tot - 1)
            bidx = torch.arange(B, device=dev)[:, None].expand(B, P)
            mb[bidx[sel], dst[sel]] = b[sel]; ms[bidx[sel], dst[sel]] = s[sel]; ml[bidx[sel], dst[sel]] = l[sel]; mi[bidx[sel], dst[sel]] = i[sel]
            kc = kc + c
        return mb, ms, ml, mi, kc

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError("VoxelNeXtHead here is the inference path (target assignment and losses are outside it)")
        x = data_dict['encoded_spconv_tensor']
        pred_dicts = [head(x) for head in self.heads_list]
        self.forward_ret_dict['pred_dicts'] = pred_dicts
        self.forward_ret_dict['voxel_indices'] = x.indices
        data_dict['final_box_dicts'] = self.generate_predicted_boxes(data_dict['batch_size'], pred_dicts, x.indices, x.spatial_shape)
        return data_dict
