"""Timing harness of the CPU arm: the reference's CPU path -- torchvision roi_pool + the PyTorch
head, as restated in oracle/oicr_plus_ref.py -- on the box's own host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by bench.py's `cpu_baseline` and `--impl reference` legs, never by
the product.  Everything is fp32 torch on the CPU; every result carries the core count, the CPU model string and the
library versions it was measured with."""
from __future__ import annotations

import os
import statistics
import time
from typing import Dict, List, Optional, Sequence

import torch
import torchvision

from . import oicr_plus_ref as ref


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_info(threads: int) -> Dict:
    return {"cores": os.cpu_count() or 1, "threads_used": threads, "cpu_model": cpu_model(), "torch": torch.__version__,
            "torchvision": torchvision.__version__, "dtype": "f32"}


def _median_time(fn, warmup: int, reps: int) -> float:
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def forward_stage_breakdown(view: "ref.View", params: "ref.HeadParams", gt_classes: torch.Tensor, C: int, K: int,
                            warmup: int = 3, reps: int = 5) -> Dict[str, float]:
    """configs[0] (cfg1): the forward of one image-view, stage by stage (seconds, median of `reps` after `warmup`).
    Stages as in BASELINE.md §3: pool / x(obj+1) / fc6 / fc7 / WSDDN+BCE / K x (pseudo-GT, label, CE+L1) / test NMS."""
    out: Dict[str, float] = {}
    gt_int, gt_oh = ref.image_level_gt(gt_classes, C)
    rois5 = ref.boxes_to_pooler_format([view.boxes])
    with torch.no_grad():
        pooled, _ = ref.roi_pool(view.feat, rois5)
        out["pool"] = _median_time(lambda: ref.roi_pool(view.feat, rois5), warmup, reps)
        scale = (view.obj + 1).view(-1, 1, 1, 1)
        out["obj_scale"] = _median_time(lambda: pooled * scale, warmup, reps)
        x0 = torch.flatten(pooled * scale, start_dim=1)
        f6 = lambda: torch.relu(torch.nn.functional.linear(x0, params.fc1_w, params.fc1_b))
        h6 = f6()
        out["fc6"] = _median_time(f6, warmup, reps)
        f7 = lambda: torch.relu(torch.nn.functional.linear(h6, params.fc2_w, params.fc2_b))
        x = f7()
        out["fc7"] = _median_time(f7, warmup, reps)

        def wsddn():
            s = ref.wsddn_scores(x, params)
            return s, ref.wsddn_loss(s, gt_oh)

        s, _ = wsddn()
        out["wsddn_bce"] = _median_time(wsddn, warmup, reps)

        def branches():
            prev = s
            for k in range(K):
                seeds = ref.pgt_mist(view.boxes, prev, gt_int, 0.10, 0.05)
                y, w, gidx, _, _ = ref.label_proposals(view.boxes, seeds, C)
                z, d = ref.refine_forward(x, params.refine[k])
                ref.oicr_cls_loss(z, y, w)
                ref.oicr_box_loss(d, y, view.boxes, view.boxes[gidx], C)
                prev = torch.softmax(z, dim=-1)

        out["oicr_branches"] = _median_time(branches, warmup, reps)
        logits = [ref.refine_forward(x, params.refine[k]) for k in range(K)]
        probs = ref.predict_probs_K([z for z, _ in logits])
        boxes = ref.predict_boxes_K([d for _, d in logits], view.boxes)
        out["test_nms"] = _median_time(
            lambda: ref.fast_rcnn_inference_single_image(boxes, probs, view.image_size, 1e-6, 0.3, 100), warmup, reps)
    out["total_forward"] = sum(out.values())
    return out


def train_step_time(views: Sequence["ref.View"], gt_classes: torch.Tensor, params: "ref.HeadParams", C: int, K: int,
                    drop_masks=None, prev_override=None, lr: float = 0.0, momentum: float = 0.9, weight_decay: float = 5e-4,
                    warmup: int = 1, reps: int = 2):
    """configs[1] (cfg2): forward + backward of the 4-view step (+ when lr > 0 the reference's optimizer.step():
    torch.optim.SGD with momentum and weight decay, tools/train_net_multi.py:157-164).  Returns (seconds per step
    (median), losses of the FIRST run -- taken before any update -- as floats)."""
    first: Dict[str, float] = {}
    opt = torch.optim.SGD([t.requires_grad_(True) for t in params.tensors()], lr=lr, momentum=momentum,
                          weight_decay=weight_decay) if lr > 0 else None

    def step():
        for v in views:
            v.feat.requires_grad_(True)
            v.feat.grad = None
        for t in params.tensors():
            t.requires_grad_(True)
            t.grad = None
        losses, _ = ref.train_step(views, gt_classes, params, C, K, drop_masks=drop_masks, prev_override=prev_override)
        total = sum(losses.values())
        total.backward()
        if not first:
            first.update({k: float(v.detach()) for k, v in losses.items()})
        if opt is not None:
            opt.step()

    t = _median_time(step, warmup, reps)
    return t, first
