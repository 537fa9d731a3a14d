"""CenterHead post-processing behind the reference's interface, every step on the device.

Mirrors `CenterHead.generate_predicted_boxes(batch_size, pred_dicts)` (pcdet/models/dense_heads/center_head.py:297-365): same
arguments, same return value (a list with one dict per frame: pred_boxes (n, 7|9), pred_scores (n,), pred_labels (n,) 1-based), same
POST_PROCESSING keys (SCORE_THRESH, POST_CENTER_LIMIT_RANGE, MAX_OBJ_PER_SAMPLE, NMS_CONFIG.{NMS_TYPE, NMS_THRESH, NMS_PRE_MAXSIZE,
NMS_POST_MAXSIZE}, USE_IOU_TO_RECTIFY_SCORE / IOU_RECTIFIER).  The reference runs torch.topk twice, ~15 gather / index kernels, a boolean
mask per frame (a host sync each), and an NMS that cudaMallocs, copies its bit mask to the host synchronously and sweeps it in a serial
CPU loop; here one head is 4 kernel launches (csrc/centerhead.cu) and the only host sync is the read of the per-frame box counts when
the variable-length result tensors are cut at the very end (`lazy=True` skips even that and returns padded tensors + counts).

NMS_TYPE: nms_gpu (class-agnostic, the CenterPoint configs) and class_specific_nms (per-class thresholds, the VoxelNeXt / IoU-head
configs; model_nms_utils.py:68-107) are built; circle_nms raises like the reference (center_head.py:347-348)."""
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from ._lib import QlidarError


class _Cfg(dict):
    """dict with attribute access and .get, like the reference's EasyDict configs"""
    __getattr__ = dict.get


def _cfg(d):
    if isinstance(d, dict) and not isinstance(d, _Cfg):
        return _Cfg({k: _cfg(v) for k, v in d.items()})
    return d


class CenterHeadPostProcessor:
    """generate_predicted_boxes of a CenterHead.  Construct from the head's attributes:
       class_names, class_names_each_head (-> class_id_mapping_each_head, center_head.py:65-74), point_cloud_range, voxel_size,
       feature_map_stride (model_cfg.TARGET_ASSIGNER_CONFIG.FEATURE_MAP_STRIDE), post_processing (model_cfg.POST_PROCESSING),
       head_order (SEPARATE_HEAD_CFG.HEAD_ORDER; 'vel' in it selects 9-value boxes)."""

    def __init__(self, class_names: Sequence[str], class_names_each_head: Sequence[Sequence[str]], point_cloud_range, voxel_size,
                 feature_map_stride, post_processing: Dict, head_order: Sequence[str] = ("center", "center_z", "dim", "rot"),
                 device="cuda"):
        self.class_names = list(class_names)
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.voxel_size = [float(v) for v in voxel_size]
        self.feature_map_stride = feature_map_stride
        self.post = _cfg(dict(post_processing))
        self.head_order = list(head_order)
        self.device = torch.device(device)
        self.class_id_mapping_each_head = [
            torch.tensor([self.class_names.index(x) for x in names if x in self.class_names], dtype=torch.int32, device=self.device)
            for names in class_names_each_head]
        self._ws = {}

    # ------------------------------------------------------------------
    def _decode_head(self, idx: int, pred_dict: Dict[str, torch.Tensor]):
        p = self.post
        f = lambda t: t.float().contiguous()
        vel = f(pred_dict["vel"]) if "vel" in self.head_order and "vel" in pred_dict else None
        iou = f(pred_dict["iou"]) if "iou" in pred_dict else None
        return ops.centerhead_decode(f(pred_dict["hm"]), f(pred_dict["center"]), f(pred_dict["center_z"]), f(pred_dict["dim"]),
                                     f(pred_dict["rot"]), vel, iou, int(p.MAX_OBJ_PER_SAMPLE), self.feature_map_stride, self.voxel_size,
                                     self.point_cloud_range, p.POST_CENTER_LIMIT_RANGE, p.SCORE_THRESH,
                                     class_map=self.class_id_mapping_each_head[idx])

    def _rectify(self, scores, labels, iou, count):
        """USE_IOU_TO_RECTIFY_SCORE (center_head.py:332-335): score^(1-a) * iou^a with a per class; the rows are re-sorted by the new
        score because the NMS walks them in score order (nms_gpu sorts, iou3d_nms_utils.py:127)."""
        rect = torch.tensor(self.post.IOU_RECTIFIER, dtype=torch.float32, device=scores.device)
        a = rect[labels.long()]
        new = torch.pow(scores, 1 - a) * torch.pow(torch.clamp(iou, min=0, max=1.0), a)
        K = scores.shape[1]
        valid = torch.arange(K, device=scores.device)[None, :] < count[:, None]
        order = torch.argsort(torch.where(valid, new, torch.full_like(new, -1.0)), dim=1, descending=True, stable=True)
        return new, order

    def generate_predicted_boxes(self, batch_size: int, pred_dicts: List[Dict[str, torch.Tensor]], lazy: bool = False):
        p = self.post
        nms = p.NMS_CONFIG
        if nms.NMS_TYPE == "circle_nms":
            raise NotImplementedError
        per_head = []
        for idx, pred_dict in enumerate(pred_dicts):
            boxes, scores, labels, iou, count = self._decode_head(idx, pred_dict)
            if p.get("USE_IOU_TO_RECTIFY_SCORE", False) and iou is not None:
                scores, order = self._rectify(scores, labels, iou, count)
                boxes = torch.gather(boxes, 1, order[:, :, None].expand_as(boxes)).contiguous()
                scores = torch.gather(scores, 1, order).contiguous()
                labels = torch.gather(labels, 1, order).contiguous()
            if nms.NMS_TYPE == "class_specific_nms":
                per_head.append(self._class_specific(boxes, scores, labels, count, nms))
            else:
                per_head.append(ops.nms_rotated(boxes, scores, labels, count, float(nms.NMS_THRESH), int(nms.NMS_PRE_MAXSIZE),
                                                int(nms.NMS_POST_MAXSIZE), label_offset=1, box_dim=boxes.shape[2]))
        if lazy:
            return per_head
        counts = torch.stack([h["keep_count"] for h in per_head], 0).cpu()          # the one host sync: variable-length outputs
        ret = []
        for k in range(batch_size):
            bs, ss, ls = [], [], []
            for hi, h in enumerate(per_head):
                n = int(counts[hi, k])
                bs.append(h["boxes"][k, :n]); ss.append(h["scores"][k, :n]); ls.append(h["labels"][k, :n].long())
            ret.append({"pred_boxes": torch.cat(bs, 0), "pred_scores": torch.cat(ss, 0), "pred_labels": torch.cat(ls, 0)})
        return ret

    def _class_specific(self, boxes, scores, labels, count, nms):
        """model_nms_utils.class_specific_nms (model_nms_utils.py:68-107): one NMS per class with that class's threshold and
        NMS_PRE_MAXSIZE; the kept boxes are concatenated class by class.  NMS_POST_MAXSIZE is passed to nms_gpu as `post_max_size`,
        a keyword nms_gpu swallows in **kwargs (iou3d_nms_utils.py:120): it has no effect in the reference and none here.  A class's
        rows are moved to the front (stable, so still in score order) and handed to the same device NMS."""
        B, K, bd = boxes.shape
        thr = nms.NMS_THRESH if isinstance(nms.NMS_THRESH, (list, tuple)) else [nms.NMS_THRESH] * len(self.class_names)
        pre = nms.NMS_PRE_MAXSIZE if isinstance(nms.NMS_PRE_MAXSIZE, (list, tuple)) else [nms.NMS_PRE_MAXSIZE] * len(thr)
        post = [K] * len(thr)
        st = nms.get("SCORE_THRESH", None)
        valid = torch.arange(K, device=boxes.device)[None, :] < count[:, None]
        outs = []
        for c in range(len(thr)):
            m = valid & (labels == c)
            if st is not None:
                m = m & (scores > float(st[c] if isinstance(st, (list, tuple)) else st))          # model_nms_utils.py:81-84
            order = torch.argsort((~m).to(torch.int8), dim=1, stable=True)
            cb = torch.gather(boxes, 1, order[:, :, None].expand_as(boxes)).contiguous()
            cs = torch.gather(scores, 1, order).contiguous()
            cl = torch.gather(labels, 1, order).contiguous()
            cc = m.sum(1).to(torch.int32)
            outs.append(ops.nms_rotated(cb, cs, cl, cc, float(thr[c]), int(pre[c]), int(post[c]), label_offset=1, box_dim=bd))
        tot = sum(o["boxes"].shape[1] for o in outs)
        mb = torch.zeros((B, tot, bd), dtype=torch.float32, device=boxes.device)
        ms = torch.zeros((B, tot), dtype=torch.float32, device=boxes.device)
        ml = torch.zeros((B, tot), dtype=torch.int32, device=boxes.device)
        kc = torch.zeros((B,), dtype=torch.int32, device=boxes.device)
        # concatenate the classes' kept rows per frame (device-side scatter by running offset)
        for o in outs:
            P = o["boxes"].shape[1]
            ar = torch.arange(P, device=boxes.device)[None, :]
            sel = ar < o["keep_count"][:, None]
            dst = (kc[:, None] + ar).long().clamp(max=tot - 1)
            bidx = torch.arange(B, device=boxes.device)[:, None].expand(B, P)
            mb[bidx[sel], dst[sel]] = o["boxes"][sel]
            ms[bidx[sel], dst[sel]] = o["scores"][sel]
            ml[bidx[sel], dst[sel]] = o["labels"][sel]
            kc = kc + o["keep_count"]
        return {"boxes": mb, "scores": ms, "labels": ml, "keep_count": kc, "keep": None}
