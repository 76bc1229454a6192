"""PGF -- the consumer of the detection-results json (reference: tools/pgf.py).  Same dictionaries in and out as the
reference's functions (image_id -> list of prediction dicts); the per-image decisions run on the device through
soswsod_pgf (csrc/pgf.cu) in the reference's double precision, the host only packs and unpacks.

  class_filter(result, class_dict)                       tools/pgf.py:273-290
  pgf(result, t_con, t_keep, use_diff, diff_classes)     tools/pgf.py:221-270 with contain_cal :210-219
  pgf_voc_results(rows, class_dict, ...)                 the VOC flow of pgf_voc (:43-117): rows of the json ->
                                                         0-based categories, grouped by image, filtered."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .. import ops

VOC_DIFF_CLASSES = [4, 5, 6, 8, 9, 15, 16]   # tools/pgf.py:112


def class_filter(result: Dict[int, List[dict]], class_dict: Dict[int, Sequence[int]]) -> None:
    for img_id in result:
        gt_classes = class_dict[img_id]
        result[img_id] = [p for p in result[img_id] if p["category_id"] in gt_classes]


def pgf(result: Dict[int, List[dict]], t_con: float = 0.85, t_keep: float = 0.2, use_diff: bool = False,
        diff_classes: Optional[Sequence[int]] = None, device="cuda") -> None:
    """In place, like the reference: result[img_id] keeps only the surviving predictions (order preserved)."""
    img_ids = list(result.keys())
    offsets = [0]
    boxes, scores, cats = [], [], []
    for img_id in img_ids:
        for p in result[img_id]:
            boxes.append([float(v) for v in p["bbox"]])
            scores.append(float(p["score"]))
            cats.append(int(p["category_id"]))
        offsets.append(len(scores))
    if not scores:
        return
    keep = ops.pgf_filter(torch.tensor(boxes, dtype=torch.float64, device=device),
                          torch.tensor(scores, dtype=torch.float64, device=device),
                          torch.tensor(cats, dtype=torch.int32, device=device),
                          torch.tensor(offsets, dtype=torch.int32, device=device), float(t_con), float(t_keep),
                          bool(use_diff), list(diff_classes or [])).cpu().tolist()
    for i, img_id in enumerate(img_ids):
        lo = offsets[i]
        result[img_id] = [p for k, p in enumerate(result[img_id]) if keep[lo + k]]


def pgf_voc_results(rows: Sequence[dict], class_dict: Dict[int, Sequence[int]], t_con: float = 0.85, t_keep: float = 0.2,
                    use_diff: bool = False, diff_classes: Sequence[int] = tuple(VOC_DIFF_CLASSES), device="cuda"):
    result: Dict[int, List[dict]] = {}
    for m in rows:
        m = dict(m)
        m["category_id"] = m["category_id"] - 1
        if m["image_id"] not in class_dict:
            continue
        result.setdefault(m["image_id"], []).append(m)
    class_filter(result, class_dict)
    pgf(result, t_con, t_keep, use_diff, diff_classes, device=device)
    return result
