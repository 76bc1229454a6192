"""OICR refinement predictor (uwsod/projects/WSL/wsl/modeling/roi_heads/fast_rcnn_oicr.py:151-735, the parts on
the OICR+ path): `cls_score` (C+1) and `bbox_pred` (4C) Linear layers (names/init :457-468), weighted CE + L1
losses (:258-352), K-branch averaged inference (:584-735) with per-class NMS on device."""
from typing import List, Tuple

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import ops
from ..layers import linear_act
from ..structures import Boxes, Instances, ShapeSpec


class _OICRLoss(Function):
    """(logits|deltas [R, 5C+1], boxes [R,4], gt_classes, gt_weights, gt_index) -> (loss_cls, loss_box_reg)."""

    @staticmethod
    def forward(ctx, pred, boxes, gt_class, gt_weight, gt_index, C, weights):
        R = pred.shape[0]
        lg = pred.detach().float().contiguous()
        dl = torch.zeros_like(lg)
        losses, _, acc = ops.oicr_loss(lg, 0, 5 * C + 1, boxes.reshape(1, R, 4).float(), gt_class.int().contiguous(),
                                       gt_weight.float().contiguous(), gt_index.int().contiguous(), 1, R, C, 1,
                                       flip_quirk=False, weights=weights, dlogits=dl)
        ctx.save_for_backward(dl)
        ctx.C = C
        return losses[0, 0], losses[0, 1]

    @staticmethod
    @once_differentiable
    def backward(ctx, gc, gb):
        (dl,) = ctx.saved_tensors
        C = ctx.C
        g = dl.clone()
        g[:, :C + 1] *= gc
        g[:, C + 1:] *= gb
        return g, None, None, None, None, None, None


class OICROutputLayers(nn.Module):
    def __init__(self, cfg, input_shape, k: int = None):
        super().__init__()
        if isinstance(input_shape, int):
            input_shape = ShapeSpec(channels=input_shape)
        input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        C = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        assert not cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG, "OICR+ ships class-specific box regression"
        self.num_classes = C
        self.refine_k = k
        self.refine_reg = bool(cfg.WSL.REFINE_REG[k if k is not None else 0])
        assert self.refine_reg, "the released OICR+ configs set REFINE_REG: True for every branch"
        self.cls_score = nn.Linear(input_size, C + 1)
        self.bbox_pred = nn.Linear(input_size, 4 * C)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in [self.cls_score, self.bbox_pred]:
            nn.init.constant_(l.bias, 0)
        self.box_dim = 4
        self.bbox_reg_weights = tuple(cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)
        self.test_score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
        self.test_nms_thresh = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
        self.test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE

    def forward(self, x) -> Tuple[torch.Tensor, torch.Tensor]:
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        w = torch.cat([self.cls_score.weight, self.bbox_pred.weight], 0)
        b = torch.cat([self.cls_score.bias, self.bbox_pred.bias], 0)
        out = linear_act(x, w, b)
        C = self.num_classes
        return out[:, :C + 1], out[:, C + 1:]

    def losses(self, predictions, proposals: List[Instances]):
        """proposals carry proposal_boxes, gt_classes, gt_weights, gt_index (label_and_sample_proposals output);
        gt boxes are proposal_boxes[gt_index] (roi_heads_oicrplus.py:327-371)."""
        scores, deltas = predictions
        p = proposals[0]
        k = self.refine_k
        lc, lb = _OICRLoss.apply(torch.cat([scores, deltas], 1), p.proposal_boxes.tensor, p.gt_classes, p.gt_weights,
                                 p.gt_index, self.num_classes, self.bbox_reg_weights)
        return {f"loss_cls_r{k}": lc, f"loss_box_reg_r{k}": lb}

    def predict_probs(self, predictions, proposals=None):
        scores, _ = predictions
        return [torch.softmax(scores.detach(), dim=-1)]

    def inference(self, predictions_K, proposals: List[Instances]):
        """predict_probs_K / predict_boxes_K + fast_rcnn_inference for ONE image -> ([Instances], [row indices],
        all_scores [[1,R,C+1]], all_boxes [[1,R,4C]]) -- per-image lists like fast_rcnn_oicr.py:46-83."""
        C = self.num_classes
        if isinstance(predictions_K[0], tuple):
            preds = list(predictions_K)
        else:
            preds = [predictions_K]
        K = len(preds)
        L = torch.cat([torch.cat([s.detach().float(), d.detach().float()], 1) for s, d in preds], 1).contiguous()
        p = proposals[0]
        probs, pboxes = ops.predict(L, 0, 5 * C + 1, p.proposal_boxes.tensor, C, K, self.bbox_reg_weights)
        return _detections_to_instances(ops.detect(probs, pboxes, p.image_size, self.test_score_thresh,
                                                   self.test_nms_thresh, self.test_topk_per_image),
                                        p.image_size) + ([probs.unsqueeze(0)], [pboxes.unsqueeze(0)])


def _detections_to_instances(det, image_size):
    db, ds, dc, dr, nd = det
    n = int(nd.item())
    inst = Instances(image_size)
    inst.pred_boxes = Boxes(db[:n])
    inst.scores = ds[:n]
    inst.pred_classes = dc[:n].long()
    inst.pred_inds = dr[:n].long()
    return [inst], [dr[:n].long()]
