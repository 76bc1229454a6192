"""WSDDN predictor (uwsod/projects/WSL/wsl/modeling/roi_heads/fast_rcnn_wsddn.py:154-832, the parts on the OICR+
path): two Linear streams `cls` / `det` (names and Xavier init :490-498), scores = softmax_c * softmax_r
(:558-567), image-level BCE (:340-375, :658-681)."""

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import ops
from ..layers import linear_act
from ..structures import ShapeSpec


class _WSDDNScoreLoss(Function):
    """(logits [R, 2C], gt one-hot [C]) -> (scores [R,C], loss); one fused kernel for forward AND gradient."""

    @staticmethod
    def forward(ctx, logits, gt_onehot):
        R, C2 = logits.shape
        C = C2 // 2
        lg = logits.detach().float().contiguous()
        dl = torch.zeros_like(lg)
        scores, img, loss = ops.wsddn_forward(lg, 0, C, 1, R, C, gt_onehot, dlogits=dl)
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(scores)
        return scores[0], loss[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, _gscores, gloss):
        (dl,) = ctx.saved_tensors
        return dl * gloss, None


class WSDDNOutputLayers(nn.Module):
    def __init__(self, cfg, input_shape):
        super().__init__()
        if isinstance(input_shape, int):
            input_shape = ShapeSpec(channels=input_shape)
        input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        self.num_classes = num_classes
        self.cls = nn.Linear(input_size, num_classes)
        self.det = nn.Linear(input_size, num_classes)
        nn.init.xavier_uniform_(self.cls.weight)
        nn.init.xavier_uniform_(self.det.weight)
        for l in [self.cls, self.det]:
            nn.init.constant_(l.bias, 0)
        self.box_dim = 4
        self.num_bbox_reg_classes = 1 if cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG else num_classes
        self.mean_loss = cfg.WSL.MEAN_LOSS
        self._logits = None

    def forward(self, x, proposals=None):
        """-> (scores [R,C], zero deltas [R,4C]) for ONE image per call (fast_rcnn_wsddn.py:566-567 softmaxes over
        all rows).  The logits are kept for `losses`."""
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        w = torch.cat([self.cls.weight, self.det.weight], 0)
        b = torch.cat([self.cls.bias, self.det.bias], 0)
        self._logits = linear_act(x, w, b)
        C = self.num_classes
        lg = self._logits.detach().float().contiguous()
        scores = ops.wsddn_forward(lg, 0, C, 1, lg.shape[0], C, torch.zeros(C, device=lg.device))[0][0]
        deltas = torch.zeros(scores.shape[0], self.box_dim * self.num_bbox_reg_classes, dtype=scores.dtype,
                             device=scores.device)
        return scores, deltas

    def losses(self, predictions, proposals, gt_classes_img_oh):
        """{'loss_cls': BCE(clamp(sum_r scores), one-hot, mean over C) / N_img}  (N_img = 1)."""
        assert self._logits is not None, "call forward() first"
        _, loss = _WSDDNScoreLoss.apply(self._logits, gt_classes_img_oh.reshape(-1).float())
        return {"loss_cls": loss}
