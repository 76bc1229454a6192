"""CPU oracle for the SoS-WSOD Stage-1 OICR+ ROI-head hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sos_wsod_b200/`` may import this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker / the CPU arm, never as the product path.

What it is: a restatement, in plain fp32 ``torch`` on the CPU, of the reference's glue code, calling
the SAME un-vendored library operators the reference calls for the arithmetic
(``torchvision.ops.roi_pool`` / ``torchvision.ops.nms``, ``torch.topk``, ``F.softmax``,
``F.binary_cross_entropy``, ``F.cross_entropy``).  detectron2 / fvcore are not importable in this
image, so the glue is restated line by line; every function cites the reference
file:line it follows.  Prefixes: ``W/`` = uwsod/projects/WSL/, ``U/`` = uwsod/.

Parity pin status -- every part is pinned:
  * pairwise_iou  -- the reference KAT U/tests/structures/test_boxes.py:150-173
  * Matcher       -- the reference KAT U/tests/modeling/test_matcher.py:19-27
  * Box2BoxTransform -- the round-trip property U/tests/modeling/test_box2box_transform.py:16-31
  * roi_pool / nms -- are the library ops themselves (torchvision 0.26 CPU kernels), cross-checked
    by the independent scalar C restatement in oracle/ref_kernels.c
  * pooler, box head, WSDDN scores/BCE, pseudo-GT mining, labelling, OICR losses, K-branch inference:
    the reference ships no test for them, so tests/golden/make_golden.py imports the REFERENCE'S OWN
    modules (read-only, through the stub importer tests/golden/ref_import.py) and commits their outputs
    on seeded inputs (tests/golden/oicr_plus_golden.pt); this file reproduces them bit for bit
    (tests/test_oracle.py)
  * the WHOLE training step (`train_step`): tests/golden/make_golden_step.py constructs the reference's own
    OICRPlusHeads and runs its `forward` (get_image_level_gt -> feature split -> `_forward_box`,
    roi_heads_oicrplus.py:149-430) + backward on seeded inputs, with dropout off and with the reference's own
    dropout keep-masks recorded -> tests/golden/step_golden.pt; `train_step` reproduces the losses to 1e-6 and every
    parameter / conv5 gradient to 1e-4 relative (VOC K=3 and COCO K=4 shapes), which pins the view averaging
    (:290-294, :390-395), the /4 combines (:288, :384-388) and the `2_flip` quirk (:381) by RUNNING them
  * TTA view generation + merge: tests/golden/make_golden_tta.py runs the reference's own
    DatasetMapperTTAAVG / GeneralizedRCNNWithTTAAVG (fvcore's Transform classes restated, the rest
    from the reference's files) -> tests/golden/tta_golden.pt, reproduced bit for bit
  * VOC / COCO writers and PGF: oracle/eval_ref.py, pinned the same way (tests/golden/eval_golden.json)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
import torchvision

# ----------------------------------------------------------------------------------------------
# d2 primitives
# ----------------------------------------------------------------------------------------------


def box_area(boxes: torch.Tensor) -> torch.Tensor:
    """U/detectron2/structures/boxes.py:172-181."""
    return (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])


def pairwise_iou(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """U/detectron2/structures/boxes.py:329-361.  [N,4],[M,4] -> [N,M]."""
    area1 = box_area(boxes1)
    area2 = box_area(boxes2)
    wh = torch.min(boxes1[:, None, 2:], boxes2[:, 2:]) - torch.max(boxes1[:, None, :2], boxes2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    iou = torch.where(
        inter > 0,
        inter / (area1[:, None] + area2 - inter),
        torch.zeros(1, dtype=inter.dtype),
    )
    return iou


def boxes_clip(boxes: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
    """U/detectron2/structures/boxes.py:183-196 (out of place).  image_size = (h, w)."""
    h, w = image_size
    out = boxes.clone()
    out[:, 0].clamp_(min=0, max=w)
    out[:, 1].clamp_(min=0, max=h)
    out[:, 2].clamp_(min=0, max=w)
    out[:, 3].clamp_(min=0, max=h)
    return out


class Matcher:
    """U/detectron2/modeling/matcher.py:8-111 (allow_low_quality_matches path included for the KAT)."""

    def __init__(self, thresholds: Sequence[float], labels: Sequence[int], allow_low_quality_matches=False):
        thresholds = list(thresholds)
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(lo <= hi for lo, hi in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in (-1, 0, 1) for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, q: torch.Tensor):
        assert q.dim() == 2
        if q.numel() == 0:
            m = q.new_full((q.size(1),), 0, dtype=torch.int64)
            l = q.new_full((q.size(1),), self.labels[0], dtype=torch.int8)
            return m, l
        assert torch.all(q >= 0)
        matched_vals, matches = q.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for l, low, high in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        if self.allow_low_quality_matches:
            highest, _ = q.max(dim=1)
            _, pred_inds = torch.nonzero(q == highest[:, None], as_tuple=True)
            match_labels[pred_inds] = 1
        return matches, match_labels


_SCALE_CLAMP = math.log(1000.0 / 16)


def get_deltas(src: torch.Tensor, tgt: torch.Tensor, weights=(10.0, 10.0, 5.0, 5.0)) -> torch.Tensor:
    """U/detectron2/modeling/box_regression.py:38-71."""
    sw = src[:, 2] - src[:, 0]
    sh = src[:, 3] - src[:, 1]
    sx = src[:, 0] + 0.5 * sw
    sy = src[:, 1] + 0.5 * sh
    tw = tgt[:, 2] - tgt[:, 0]
    th = tgt[:, 3] - tgt[:, 1]
    tx = tgt[:, 0] + 0.5 * tw
    ty = tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    dx = wx * (tx - sx) / sw
    dy = wy * (ty - sy) / sh
    dw = ww * torch.log(tw / sw)
    dh = wh * torch.log(th / sh)
    assert (sw > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
    return torch.stack((dx, dy, dw, dh), dim=1)


def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights=(10.0, 10.0, 5.0, 5.0)) -> torch.Tensor:
    """U/detectron2/modeling/box_regression.py:73-110.  deltas [N,4k], boxes [N,4] -> [N,4k]."""
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = deltas[:, 2::4] / ww
    dh = deltas[:, 3::4] / wh
    dw = torch.clamp(dw, max=_SCALE_CLAMP)
    dh = torch.clamp(dh, max=_SCALE_CLAMP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.zeros_like(deltas)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw
    out[:, 3::4] = pcy + 0.5 * ph
    return out


def nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """torchvision.ops.nms -- the library op the reference calls (U/detectron2/layers/nms.py:6-7,25)."""
    return torchvision.ops.nms(boxes, scores, thr)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, thr: float) -> torch.Tensor:
    """U/detectron2/layers/nms.py:10-29, canonicalised on the per-class ("vanilla") variant on un-offset
    fp32 boxes, i.e. the reference's own >=40000-candidate branch (:22-29) and torchvision's
    _batched_nms_vanilla.  Returned indices are score-descending."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    result_mask = torch.zeros_like(scores, dtype=torch.bool)
    for cid in torch.unique(idxs).tolist():
        mask = (idxs == cid).nonzero().view(-1)
        keep = nms(boxes[mask], scores[mask], thr)
        result_mask[mask[keep]] = True
    keep = result_mask.nonzero().view(-1)
    keep = keep[scores[keep].argsort(descending=True, stable=True)]
    return keep


# ----------------------------------------------------------------------------------------------
# (B) ROI pooling, (C) objectness scaling
# ----------------------------------------------------------------------------------------------


def boxes_to_pooler_format(box_lists: List[torch.Tensor]) -> torch.Tensor:
    """W/wsl/modeling/poolers.py:67-108: [M,5] = (batch index as float, x1, y1, x2, y2)."""
    out = []
    for i, b in enumerate(box_lists):
        idx = torch.full((len(b), 1), i, dtype=b.dtype)
        out.append(torch.cat((idx, b), dim=1))
    return torch.cat(out, dim=0)


def roi_pool(feat: torch.Tensor, rois5: torch.Tensor, out_size: int = 7, scale: float = 1.0 / 8):
    """W/wsl/modeling/poolers.py:183-186,263-270 -> torchvision RoIPool.  Returns (out, argmax int32)
    straight from the library kernel (``torch.ops.torchvision.roi_pool``)."""
    out, argmax = torch.ops.torchvision.roi_pool(feat, rois5, scale, out_size, out_size)
    return out, argmax.to(torch.int32)


def roi_pool_scaled(feat, boxes, obj_logits, out_size=7, scale=1.0 / 8):
    """Rows B + C: pool, then multiply by (objectness_logits + 1)
    (W/wsl/modeling/roi_heads/roi_heads_oicrplus.py:195-221)."""
    rois5 = boxes_to_pooler_format([boxes])
    out = torchvision.ops.roi_pool(feat, rois5, (out_size, out_size), scale)
    return out * (obj_logits + 1).view(-1, 1, 1, 1)


# ----------------------------------------------------------------------------------------------
# Parameters of the head (names = the reference's checkpoint keys)
# ----------------------------------------------------------------------------------------------


@dataclass
class HeadParams:
    """fc1/fc2 = W/wsl/modeling/roi_heads/box_head.py:55-67; cls/det = fast_rcnn_wsddn.py:490-498;
    refine[k] = (cls_score.weight, cls_score.bias, bbox_pred.weight, bbox_pred.bias),
    fast_rcnn_oicr.py:457-468."""

    fc1_w: torch.Tensor
    fc1_b: torch.Tensor
    fc2_w: torch.Tensor
    fc2_b: torch.Tensor
    cls_w: torch.Tensor
    cls_b: torch.Tensor
    det_w: torch.Tensor
    det_b: torch.Tensor
    refine: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]] = field(default_factory=list)

    def tensors(self) -> List[torch.Tensor]:
        out = [self.fc1_w, self.fc1_b, self.fc2_w, self.fc2_b, self.cls_w, self.cls_b, self.det_w, self.det_b]
        for r in self.refine:
            out.extend(r)
        return out

    def requires_grad_(self, flag=True):
        for t in self.tensors():
            t.requires_grad_(flag)
        return self


def init_head_params(num_classes: int, refine_k: int, in_dim: int = 512 * 7 * 7, fc_dim: int = 4096,
                     generator: Optional[torch.Generator] = None) -> HeadParams:
    """Reference initialisers: fc N(0,0.005)/bias 0.1 (box_head.py:64-67); cls/det Xavier-uniform, bias 0
    (fast_rcnn_wsddn.py:495-498); refine cls_score N(0,0.01), bbox_pred N(0,0.001), bias 0
    (fast_rcnn_oicr.py:465-468)."""
    g = generator
    C = num_classes

    def normal(shape, std):
        return torch.empty(shape).normal_(0, std, generator=g)

    def xavier(shape):
        fan_out, fan_in = shape
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return torch.empty(shape).uniform_(-a, a, generator=g)

    p = HeadParams(
        fc1_w=normal((fc_dim, in_dim), 0.005), fc1_b=torch.full((fc_dim,), 0.1),
        fc2_w=normal((fc_dim, fc_dim), 0.005), fc2_b=torch.full((fc_dim,), 0.1),
        cls_w=xavier((C, fc_dim)), cls_b=torch.zeros(C),
        det_w=xavier((C, fc_dim)), det_b=torch.zeros(C),
    )
    for _ in range(refine_k):
        p.refine.append((normal((C + 1, fc_dim), 0.01), torch.zeros(C + 1),
                         normal((4 * C, fc_dim), 0.001), torch.zeros(4 * C)))
    return p


# ----------------------------------------------------------------------------------------------
# (D) box head, (E,F) WSDDN
# ----------------------------------------------------------------------------------------------


def box_head(x: torch.Tensor, p: HeadParams, drop_masks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """W/wsl/modeling/roi_heads/box_head.py:82-91.  ``drop_masks`` = explicit keep-masks (0/1) for the two
    dropout layers (scaled by 1/(1-p) = 2); None = eval-mode dropout."""
    x = torch.flatten(x, start_dim=1)
    x = F.relu(F.linear(x, p.fc1_w, p.fc1_b))
    if drop_masks is not None:
        x = x * drop_masks[0] * 2.0
    x = F.relu(F.linear(x, p.fc2_w, p.fc2_b))
    if drop_masks is not None:
        x = x * drop_masks[1] * 2.0
    return x


def wsddn_scores(x: torch.Tensor, p: HeadParams) -> torch.Tensor:
    """W/wsl/modeling/roi_heads/fast_rcnn_wsddn.py:558-567 (one image per call)."""
    Cl = F.linear(x, p.cls_w, p.cls_b)
    Dl = F.linear(x, p.det_w, p.det_b)
    return F.softmax(Cl, dim=1) * F.softmax(Dl, dim=0)


def wsddn_scores_from_logits(Cl: torch.Tensor, Dl: torch.Tensor) -> torch.Tensor:
    return F.softmax(Cl, dim=1) * F.softmax(Dl, dim=0)


def wsddn_img_scores(scores: torch.Tensor) -> torch.Tensor:
    """fast_rcnn_wsddn.py:360-375: sum over proposals then clamp to [1e-6, 1-1e-6]."""
    return torch.clamp(torch.sum(scores, dim=0, keepdim=True), min=1e-6, max=1.0 - 1e-6)


def wsddn_loss(scores: torch.Tensor, gt_oh: torch.Tensor) -> torch.Tensor:
    """fast_rcnn_wsddn.py:340-358 with MEAN_LOSS=True: BCE(mean over C) / N_img (N_img = 1)."""
    return F.binary_cross_entropy(wsddn_img_scores(scores), gt_oh, reduction="mean") / gt_oh.size(0)


def image_level_gt(gt_classes: torch.Tensor, num_classes: int):
    """W/wsl/modeling/roi_heads/roi_heads.py:144-164 for one image."""
    gt_int = torch.unique(gt_classes, sorted=True).to(torch.int64)
    oh = torch.zeros((1, num_classes), dtype=torch.float).scatter_(1, gt_int.unsqueeze(0), 1)
    return gt_int, oh


# ----------------------------------------------------------------------------------------------
# (H,I) pseudo ground-truth mining, (J,K,L) labelling
# ----------------------------------------------------------------------------------------------


@dataclass
class Seeds:
    boxes: torch.Tensor    # [M,4]
    classes: torch.Tensor  # [M] int64
    scores: torch.Tensor   # [M] (== weights, roi_heads_oicrplus.py:599-600)
    index: torch.Tensor    # [M] int64 proposal index


def pgt_top_k(boxes: torch.Tensor, prev_scores: torch.Tensor, gt_int: torch.Tensor,
              top_k: float = 0.10, thres: float = 0.05):
    """W/wsl/modeling/roi_heads/roi_heads_oicrplus.py:607-757 with need_instance=False, need_weight=True,
    one image.  ``prev_scores`` is [R,C] (k=0) or [R,C+1] (k>=1); only the gt_int columns are read."""
    P = torch.index_select(prev_scores, 1, gt_int)                       # :646-649
    R = P.size(0)
    G = gt_int.numel()
    if top_k >= 1:
        kt = min(R, int(top_k))
    elif 0 < top_k < 1:
        kt = max(int(R * top_k), 1)                                       # :657-662
    else:
        kt = min(R, 1)
    sc, idx = torch.topk(P, kt, dim=0)                                    # :663-669
    bx = boxes[idx]                                                       # [kt,G,4]  :671-678
    cls = gt_int.unsqueeze(0).expand(kt, G)
    if thres > 0:
        mask = sc.ge(thres)
        mask = torch.cat([torch.full_like(mask[0:1, :], True), mask[1:, :]], dim=0)   # :698-704
        idx_f = torch.masked_select(idx, mask)
        sc_f = torch.masked_select(sc, mask)
        bx_f = torch.masked_select(bx, mask.unsqueeze(2).expand(kt, G, 4)).reshape(-1, 4)
        cls_f = torch.masked_select(cls, mask)
    else:
        idx_f, sc_f, bx_f, cls_f = idx.reshape(-1), sc.reshape(-1), bx.reshape(-1, 4), cls.reshape(-1)
    return sc_f, bx_f, cls_f, sc_f.clone(), idx_f


def pgt_mist(boxes: torch.Tensor, prev_scores: torch.Tensor, gt_int: torch.Tensor,
             top_pro: float = 0.10, thres: float = 0.05, nms_thr: float = 0.01) -> Seeds:
    """roi_heads_oicrplus.py:559-605: class-AGNOSTIC NMS(0.01) of the candidates (all idxs zero, :576);
    gt_weights are the kept SCORES (:599-600)."""
    sc, bx, cls, _w, idx = pgt_top_k(boxes, prev_scores, gt_int, top_k=top_pro, thres=thres)
    keep = batched_nms(bx, sc, torch.zeros_like(cls), nms_thr)
    return Seeds(boxes=bx[keep], classes=cls[keep], scores=sc[keep], index=idx[keep])


def label_proposals(boxes: torch.Tensor, seeds: Seeds, num_classes: int,
                    iou_thresholds=(0.5, 0.6), iou_labels=(0, -1, 1)):
    """W/wsl/modeling/roi_heads/roi_heads.py:266-375 + :225-257 (no sub-sampling) with the Matcher of
    :218-222.  Returns per-proposal (gt_classes int64, gt_weights fp32, gt_index int64, matched_idxs,
    gt_boxes)."""
    q = pairwise_iou(seeds.boxes, boxes)                                   # [M,R]
    matched_idxs, matched_labels = Matcher(iou_thresholds, iou_labels, False)(q)
    gt_classes = seeds.classes[matched_idxs].clone()
    gt_classes[matched_labels == 0] = num_classes
    gt_classes[matched_labels == -1] = -1
    gt_weights = seeds.scores[matched_idxs].clone()
    gt_index = seeds.index[matched_idxs].clone()
    gt_boxes = seeds.boxes[matched_idxs].clone()
    return gt_classes, gt_weights, gt_index, matched_idxs, gt_boxes


# ----------------------------------------------------------------------------------------------
# (M,N,O) refinement branch
# ----------------------------------------------------------------------------------------------


def refine_forward(x: torch.Tensor, refine_params) -> Tuple[torch.Tensor, torch.Tensor]:
    """W/wsl/modeling/roi_heads/fast_rcnn_oicr.py:504-528 (REFINE_REG True)."""
    cw, cb, bw, bb = refine_params
    return F.linear(x, cw, cb), F.linear(x, bw, bb)


def oicr_cls_loss(logits: torch.Tensor, gt_classes: torch.Tensor, gt_weights: torch.Tensor) -> torch.Tensor:
    """fast_rcnn_oicr.py:219-220 + :258-273: weights of ignored rows zeroed, CE(reduction none,
    ignore_index=-1) * w, mean over ALL rows."""
    w = gt_weights.clone()
    w[gt_classes == -1] = 0.0
    loss = F.cross_entropy(logits, gt_classes, reduction="none", ignore_index=-1)
    return torch.mean(loss * w)


def oicr_box_loss(deltas: torch.Tensor, gt_classes: torch.Tensor, proposal_boxes: torch.Tensor,
                  gt_boxes: torch.Tensor, num_classes: int, weights=(10.0, 10.0, 5.0, 5.0)) -> torch.Tensor:
    """fast_rcnn_oicr.py:276-352, smooth_l1 branch with beta=0 (== plain L1, fvcore smooth_l1_loss),
    class-specific regression, normalised by R."""
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes), as_tuple=True)[0]
    cols = 4 * gt_classes[fg][:, None] + torch.arange(4)
    tgt = get_deltas(proposal_boxes, gt_boxes, weights)
    loss = torch.abs(deltas[fg[:, None], cols] - tgt[fg]).sum()
    return loss / gt_classes.numel()


def oicr_accuracy_counters(logits: torch.Tensor, gt_classes: torch.Tensor):
    """fast_rcnn_oicr.py:228-256: (num_instances, num_fg, num_accurate, fg_num_accurate, num_false_negative)."""
    pred = logits.argmax(dim=1)
    bg = logits.shape[1] - 1
    fg = (gt_classes >= 0) & (gt_classes < bg)
    return (gt_classes.numel(), int(fg.sum()), int((pred == gt_classes).sum()),
            int((pred[fg] == gt_classes[fg]).sum()), int((pred[fg] == bg).sum()))


def reference_accuracy_scalars(logits: torch.Tensor, gt_classes: torch.Tensor, num_classes: int):
    """The EventStorage scalars of fast_rcnn_oicr.py:228-256 (`fast_rcnn/cls_accuracy_r{k}`, `fg_cls_accuracy`,
    `false_negative`) from the counters above; the fg ratios are absent when there is no foreground row."""
    n, n_fg, acc, fg_acc, fn = oicr_accuracy_counters(logits, gt_classes)
    out = {}
    if n > 0:
        out["cls_accuracy"] = acc / n
        if n_fg > 0:
            out["fg_cls_accuracy"] = fg_acc / n_fg
            out["false_negative"] = fn / n_fg
    return out


# ----------------------------------------------------------------------------------------------
# Full training step and test-time forward
# ----------------------------------------------------------------------------------------------


@dataclass
class View:
    feat: torch.Tensor     # [1,512,h,w]
    boxes: torch.Tensor    # [R,4]
    obj: torch.Tensor      # [R]  objectness_logits
    image_size: Tuple[int, int] = (0, 0)   # (h, w) of the (resized) image


def train_step(views: Sequence[View], gt_classes: torch.Tensor, p: HeadParams, num_classes: int,
               refine_k: int, drop_masks: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None,
               mist_p: float = 0.10, mist_thre: float = 0.05, reproduce_flip_quirk: bool = True,
               prev_override: Optional[Sequence[torch.Tensor]] = None) -> Tuple[Dict[str, torch.Tensor], Dict]:
    """OICRPlusHeads._forward_box, W/wsl/modeling/roi_heads/roi_heads_oicrplus.py:190-430, for the shipped
    flags (REFINE_MIST True, MIST_TYPE nms, REFINE_REG True, BBOX_UPDATE False, POOLER_TYPE ROIPool).
    views = (1, 1_flip, 2, 2_flip).  Returns (loss dict, aux) where aux carries the intermediate integer
    results (seeds, labels, weights, indices) the parity tests compare bit-exactly.

    ``prev_override[k]``: if given, replaces the view-averaged scores fed to branch k's pseudo-GT mining
    (lets a test feed IDENTICAL fp32 scores to both implementations)."""
    assert len(views) == 4
    gt_int, gt_oh = image_level_gt(gt_classes, num_classes)
    xs, S = [], []
    for vi, v in enumerate(views):
        pooled = roi_pool_scaled(v.feat, v.boxes, v.obj)                    # :195-221
        x = box_head(pooled, p, None if drop_masks is None else drop_masks[vi])   # :229-232
        xs.append(x)
        S.append(wsddn_scores(x, p))                                        # :277-280
    losses: Dict[str, torch.Tensor] = {}
    losses["loss_cls"] = sum(wsddn_loss(s, gt_oh) for s in S) / 4.0         # :283-288
    prev = (S[0].detach() + S[1].detach() + S[2].detach() + S[3].detach()) / 4.0   # :290-294
    aux = {"gt_int": gt_int, "gt_oh": gt_oh, "wsddn_scores": [s.detach() for s in S], "branches": []}
    for k in range(refine_k):
        if prev_override is not None:
            prev = prev_override[k]
        seeds = pgt_mist(views[0].boxes, prev, gt_int, mist_p, mist_thre)   # :313-315
        y, w, gidx, matched, _ = label_proposals(views[0].boxes, seeds, num_classes)   # :326
        Z, Dl = [], []
        for x in xs:                                                         # :373-376
            z, d = refine_forward(x, p.refine[k])
            Z.append(z)
            Dl.append(d)
        lc, lb = [], []
        for vi, v in enumerate(views):                                       # :378-381
            src = 2 if (vi == 3 and reproduce_flip_quirk) else vi            # quirk: losses_k2_flip uses predictions_k2
            gtbox = v.boxes[gidx]                                            # :327-371
            lc.append(oicr_cls_loss(Z[src], y, w))
            lb.append(oicr_box_loss(Dl[src], y, v.boxes, gtbox, num_classes))
        losses[f"loss_cls_r{k}"] = sum(lc) / 4.0                             # :384-388
        losses[f"loss_box_reg_r{k}"] = sum(lb) / 4.0
        probs = [F.softmax(z, dim=-1).detach() for z in Z]                   # :390-395
        prev = (probs[0] + probs[1] + probs[2] + probs[3]) / 4.0
        aux["branches"].append({"seeds": seeds, "gt_classes": y, "gt_weights": w, "gt_index": gidx,
                                "matched": matched, "logits": [z.detach() for z in Z],
                                "deltas": [d.detach() for d in Dl], "next_prev": prev})
    aux["x"] = [x.detach() for x in xs]
    return losses, aux


def predict_probs_K(logits_K: Sequence[torch.Tensor]) -> torch.Tensor:
    """fast_rcnn_oicr.py:718-735."""
    probs = torch.zeros_like(logits_K[0])
    for z in logits_K:
        probs += F.softmax(z, dim=-1)
    return probs / len(logits_K)


def predict_boxes_K(deltas_K: Sequence[torch.Tensor], boxes: torch.Tensor) -> torch.Tensor:
    """fast_rcnn_oicr.py:674-700."""
    d = torch.zeros_like(deltas_K[0])
    for dk in deltas_K:
        d += dk
    d = d / len(deltas_K)
    return apply_deltas(d, boxes)


def fast_rcnn_inference_single_image(boxes: torch.Tensor, scores: torch.Tensor, image_shape: Tuple[int, int],
                                     score_thresh: float, nms_thresh: float, topk_per_image: int):
    """fast_rcnn_oicr.py:86-148.  boxes [R,4C], scores [R,C+1] -> (boxes [n,4], scores [n], classes [n],
    pred_inds [n])."""
    pred_inds = torch.arange(scores.size(0)).unsqueeze(1).repeat(1, scores.size(1))
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid.all():
        boxes, scores, pred_inds = boxes[valid], scores[valid], pred_inds[valid]
    scores = scores[:, :-1]
    C = boxes.shape[1] // 4
    boxes = boxes_clip(boxes.reshape(-1, 4), image_shape).view(-1, C, 4)
    pred_inds = pred_inds[:, :-1]
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    pred_inds = pred_inds[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    return boxes[keep], scores[keep], filter_inds[keep, 1], pred_inds[keep]


def test_forward(view: View, p: HeadParams, num_classes: int, refine_k: int):
    """OICRPlusHeads._forward_box_test, roi_heads_oicrplus.py:432-475, up to (probs [R,C+1], boxes [R,4C])."""
    x = box_head(roi_pool_scaled(view.feat, view.boxes, view.obj), p, None)
    ZK, DK = [], []
    for k in range(refine_k):
        z, d = refine_forward(x, p.refine[k])
        ZK.append(z)
        DK.append(d)
    return predict_probs_K(ZK), predict_boxes_K(DK, view.boxes)


def resize_shortest_edge(h: int, w: int, size: int, max_size: int) -> Tuple[int, int]:
    """ResizeShortestEdge.get_transform, U/detectron2/data/transforms/augmentation_impl.py:155-175 -> (new_h, new_w)."""
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def tta_view_sizes(h: int, w: int, min_sizes: Sequence[int], max_size: int, flip: bool):
    """DatasetMapperTTAAVG.__call__, W/wsl/modeling/test_time_augmentation_avg.py:175-182: the views in order
    -> [(new_h, new_w, flipped)]."""
    out = []
    for s in min_sizes:
        nh, nw = resize_shortest_edge(h, w, s, max_size)
        out.append((nh, nw, False))
        if flip:
            out.append((nh, nw, True))
    return out


def _f32(x: float) -> torch.Tensor:
    return torch.tensor(x, dtype=torch.float64).to(torch.float32)


def tta_transform_proposals(boxes: torch.Tensor, hw: Tuple[int, int], new_hw: Tuple[int, int], flipped: bool,
                            min_box_size: float = 0.0):
    """transform_proposals, test_time_augmentation_avg.py:29-71 (without the final row filter): TransformList([Resize,
    HFlip]).apply_box on fp32 numpy boxes (ResizeTransform.apply_coords U/detectron2/data/transforms/transform.py:
    123-126: x * (new_w * 1.0 / w) with the Python double rounded to fp32 by numpy; fvcore Transform.apply_box takes
    the min / max of the four transformed corners after every member; HFlipTransform: x -> new_w - x), then
    Boxes.clip(new_hw) and Boxes.nonempty(min_box_size) (U/detectron2/structures/boxes.py:183-210).
    -> (boxes [R,4] in view coordinates, keep bool [R])."""
    (h, w), (nh, nw) = hw, new_hw
    b = boxes.clone().to(torch.float32)
    xa, xb = b[:, 0] * _f32(nw * 1.0 / w), b[:, 2] * _f32(nw * 1.0 / w)
    ya, yb = b[:, 1] * _f32(nh * 1.0 / h), b[:, 3] * _f32(nh * 1.0 / h)
    x1, x2, y1, y2 = torch.minimum(xa, xb), torch.maximum(xa, xb), torch.minimum(ya, yb), torch.maximum(ya, yb)
    if flipped:
        xa, xb = nw - x1, nw - x2
        x1, x2 = torch.minimum(xa, xb), torch.maximum(xa, xb)
    out = torch.stack([x1.clamp(min=0, max=nw), y1.clamp(min=0, max=nh), x2.clamp(min=0, max=nw), y2.clamp(min=0, max=nh)], 1)
    keep = ((out[:, 2] - out[:, 0]) > min_box_size) & ((out[:, 3] - out[:, 1]) > min_box_size)
    return out, keep


def tta_inverse_boxes(boxes: torch.Tensor, scale_x: float, scale_y: float, flipped: bool, view_w: int,
                      post_scale: Optional[Tuple[float, float]] = None):
    """W/wsl/modeling/test_time_augmentation_avg.py:353-365 with fvcore TransformList([Resize, HFlip]).inverse():
    inverse flip first (x' = W_view - x; fvcore HFlipTransform.apply_coords, then apply_box takes the corner
    min/max), then inverse resize (x * (w/new_w), y * (h/new_h), the Python doubles rounded to fp32), then -- when the
    mapper composed a pre-transform, :169-173 -- the inverse of that resize (`post_scale`).  boxes [n,4]."""
    b = boxes.clone()
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    if flipped:
        xa, xb = view_w - x1, view_w - x2
        x1, x2 = torch.minimum(xa, xb), torch.maximum(xa, xb)
    for sx, sy in [(scale_x, scale_y)] + ([post_scale] if post_scale is not None else []):
        xa, xb, ya, yb = x1 * _f32(sx), x2 * _f32(sx), y1 * _f32(sy), y2 * _f32(sy)
        x1, x2, y1, y2 = torch.minimum(xa, xb), torch.maximum(xa, xb), torch.minimum(ya, yb), torch.maximum(ya, yb)
    return torch.stack([x1, y1, x2, y2], 1)


def tta_merge(all_boxes: Sequence[torch.Tensor], all_scores: Sequence[torch.Tensor]):
    """test_time_augmentation_avg.py:367-371: mean over views of boxes [V,R,4C] and scores [V,R,C+1]."""
    return torch.stack(list(all_boxes), 0).mean(dim=0), torch.stack(list(all_scores), 0).mean(dim=0)


def voc_detection_rows(image_id: int, boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor):
    """U/detectron2/evaluation/pascal_voc_evaluation.py:57-71 + :89-113: the string round trip of the VOC
    writer -> list of json rows {"image_id","category_id"(1-based),"score","bbox":[x1+1,y1+1,x2,y2]}."""
    per_cls: Dict[int, List[str]] = {}
    for box, score, cls in zip(boxes.numpy(), scores.tolist(), classes.tolist()):
        xmin, ymin, xmax, ymax = box
        xmin += 1
        ymin += 1
        per_cls.setdefault(cls, []).append(f"{image_id} {score:.3f} {xmin:.1f} {ymin:.1f} {xmax:.1f} {ymax:.1f}")
    rows = []
    for cls_id in sorted(per_cls):
        for line in per_cls[cls_id]:
            m = line.split(" ")
            rows.append({"image_id": int(m[0]), "category_id": cls_id + 1, "score": float(m[1]),
                         "bbox": [float(m[2]), float(m[3]), float(m[4]), float(m[5])]})
    return rows


# ----------------------------------------------------------------------------------------------
# Seeded synthetic inputs
# ----------------------------------------------------------------------------------------------


def synth_boxes(R: int, img_h: int, img_w: int, g: torch.Generator, min_size: int = 20) -> torch.Tensor:
    """Integer-pixel XYXY proposals (MCG/SS-like), de-duplicated with the reference's unique_boxes hash
    (U/detectron2/structures/boxes.py:214-226); re-draws until R unique boxes exist."""
    out = torch.zeros((0, 4))
    while out.size(0) < R:
        n = 2 * R
        x1 = torch.rand(n, generator=g) * (img_w - 32)
        y1 = torch.rand(n, generator=g) * (img_h - 32)
        w = min_size + torch.rand(n, generator=g) * (img_w - x1 - min_size)
        h = min_size + torch.rand(n, generator=g) * (img_h - y1 - min_size)
        b = torch.stack([x1, y1, (x1 + w).clamp(max=img_w - 1), (y1 + h).clamp(max=img_h - 1)], 1).round()
        b = b[((b[:, 2] - b[:, 0]) >= min_size) & ((b[:, 3] - b[:, 1]) >= min_size)]
        out = torch.cat([out, b], 0)
        hashes = (out.double() @ torch.tensor([1, 1e3, 1e6, 1e9], dtype=torch.double)).long()
        _, first = torch.unique(hashes, return_inverse=True)
        seen, keep = set(), []
        for i, hsh in enumerate(hashes.tolist()):
            if hsh not in seen:
                seen.add(hsh)
                keep.append(i)
        out = out[torch.tensor(keep, dtype=torch.long)]
    return out[:R].contiguous()


def flip_boxes(boxes: torch.Tensor, img_w: int) -> torch.Tensor:
    """HFlip of XYXY boxes: x1' = W - x2, x2' = W - x1."""
    b = boxes.clone()
    b[:, 0] = img_w - boxes[:, 2]
    b[:, 2] = img_w - boxes[:, 0]
    return b


def synth_views(R: int, sizes: Sequence[Tuple[int, int]], g: torch.Generator, channels: int = 512,
                stride: int = 8) -> List[View]:
    """Four training views (1, 1_flip, 2, 2_flip): same R proposals in the same order
    (U/detectron2/data/dataset_mapper.py:353-361); view 2 = view 1 rescaled; flips mirror x.
    ``sizes`` = [(h1,w1),(h2,w2)] image sizes of the two scales."""
    (h1, w1), (h2, w2) = sizes
    base = synth_boxes(R, h1, w1, g)
    obj = torch.sort(torch.rand(R, generator=g), descending=True).values
    sx, sy = w2 / w1, h2 / h1
    b2 = base.clone()
    b2[:, 0::2] *= sx
    b2[:, 1::2] *= sy
    views = []
    for (h, w, b) in ((h1, w1, base), (h1, w1, flip_boxes(base, w1)), (h2, w2, b2), (h2, w2, flip_boxes(b2, w2))):
        fh, fw = (h + stride - 1) // stride, (w + stride - 1) // stride
        feat = torch.relu(torch.randn((1, channels, fh, fw), generator=g))
        views.append(View(feat=feat, boxes=b.contiguous(), obj=obj.clone(), image_size=(h, w)))
    return views


def synth_prev_scores(R: int, ncols: int, g: torch.Generator) -> torch.Tensor:
    """softmax(N(0,2))-shaped, tie-free fp32 scores for the pseudo-GT parity tests."""
    s = F.softmax(torch.randn((R, ncols), generator=g) * 2.0, dim=1)
    # make every column tie-free: nudge duplicates by distinct multiples of 1 ulp-ish
    for c in range(ncols):
        col = s[:, c]
        while torch.unique(col).numel() != col.numel():
            col = col + torch.rand(R, generator=g) * 1e-7
        s[:, c] = col
    return s.contiguous()
