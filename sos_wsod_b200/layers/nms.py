"""NMS wrappers with the reference's surface (uwsod/detectron2/layers/nms.py:10-29).  `batched_nms` is the
per-class variant on un-offset fp32 boxes (the reference's >= 40000-candidate branch),
results sorted by score descending."""
import torch

from .. import ops


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return ops.nms(boxes, scores, iou_threshold)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    assert boxes.shape[-1] == 4
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    result_mask = torch.zeros_like(scores, dtype=torch.bool)
    for cid in torch.unique(idxs).cpu().tolist():
        mask = (idxs == cid).nonzero().view(-1)
        keep = ops.nms(boxes[mask], scores[mask], iou_threshold)
        result_mask[mask[keep]] = True
    keep = result_mask.nonzero().view(-1)
    return keep[scores[keep].argsort(descending=True, stable=True)]
