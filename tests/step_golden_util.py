"""Helpers for tests/golden/step_golden.pt -- one whole training step (forward + backward) of the REFERENCE'S OWN
`OICRPlusHeads.forward` (tests/golden/make_golden_step.py)."""
import os

import torch

from oracle import oicr_plus_ref as ref

STEP_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_golden.pt")


def load():
    return torch.load(STEP_GOLDEN, weights_only=False)


def head_params(case) -> "ref.HeadParams":
    p = case["params"]
    hp = ref.HeadParams(fc1_w=p["box_head.fc1.weight"].clone(), fc1_b=p["box_head.fc1.bias"].clone(),
                        fc2_w=p["box_head.fc2.weight"].clone(), fc2_b=p["box_head.fc2.bias"].clone(),
                        cls_w=p["box_predictor.cls.weight"].clone(), cls_b=p["box_predictor.cls.bias"].clone(),
                        det_w=p["box_predictor.det.weight"].clone(), det_b=p["box_predictor.det.bias"].clone())
    for k in range(case["K"]):
        hp.refine.append(tuple(p[f"box_refinery_{k}.{n}"].clone() for n in
                               ("cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias")))
    return hp


def views(case):
    return [ref.View(feat=v["feat"].clone(), boxes=v["boxes"].clone(), obj=v["obj"].clone(), image_size=v["image_size"])
            for v in case["views"]]


def grad_key_map(K):
    """reference parameter name -> position in HeadParams.tensors()."""
    names = ["box_head.fc1.weight", "box_head.fc1.bias", "box_head.fc2.weight", "box_head.fc2.bias",
             "box_predictor.cls.weight", "box_predictor.cls.bias", "box_predictor.det.weight", "box_predictor.det.bias"]
    for k in range(K):
        names += [f"box_refinery_{k}.{n}" for n in ("cls_score.weight", "cls_score.bias", "bbox_pred.weight", "bbox_pred.bias")]
    return names


def engine_key_map(K):
    """reference parameter name -> key of engine.TrainOutput.grads / HeadOperands.master."""
    m = {"box_head.fc1.weight": "fc1_w", "box_head.fc1.bias": "fc1_b", "box_head.fc2.weight": "fc2_w",
         "box_head.fc2.bias": "fc2_b", "box_predictor.cls.weight": "cls_w", "box_predictor.cls.bias": "cls_b",
         "box_predictor.det.weight": "det_w", "box_predictor.det.bias": "det_b"}
    for k in range(K):
        m.update({f"box_refinery_{k}.cls_score.weight": f"r{k}_cls_w", f"box_refinery_{k}.cls_score.bias": f"r{k}_cls_b",
                  f"box_refinery_{k}.bbox_pred.weight": f"r{k}_box_w", f"box_refinery_{k}.bbox_pred.bias": f"r{k}_box_b"})
    return m
