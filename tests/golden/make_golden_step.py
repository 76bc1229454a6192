"""Generates tests/golden/step_golden.pt by running the REFERENCE'S OWN `OICRPlusHeads.forward` (training branch:
get_image_level_gt -> feature split -> `_forward_box`, wsl/modeling/roi_heads/roi_heads_oicrplus.py:149-430) on
seeded synthetic inputs, then `.backward()` of the summed losses.  The class is constructed through its own
`__init__` with the reference's own sub-modules (ROIPooler, DiscriminativeAdaptionNeck, WSDDNOutputLayers,
OICROutputLayers, Matcher) -- nothing of this repo's product or oracle is on that path; only the oracle's seeded
input generator is used to make the inputs.  Run in the build container only:

    python tests/golden/make_golden_step.py

Two cases per shape: dropout off (box_head in eval mode, the head itself in training mode) and dropout on, where the
keep-masks the reference drew are recorded by wrapping `F.dropout` inside the box head's module (the wrapper calls
the real function and stores `out != 0 | in == 0`-free masks derived from its output/input ratio).

What this pins that make_golden.py does not (VERDICT r01 "missing" #2): the view averaging of the WSDDN scores
(:290-294) and of the refinement softmaxes (:390-395), the /4.0 loss combines (:288, :384-388), the `2_flip` quirk
(:381, losses_k2_flip computed from predictions_k2), the feature split of the [image, flip] batches (:176-181), the
order of the branches, and every parameter / feature gradient of the whole step.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402

ref_import.install()

from detectron2.layers import ShapeSpec  # noqa: E402
from detectron2.modeling.box_regression import Box2BoxTransform  # noqa: E402
from detectron2.modeling.matcher import Matcher  # noqa: E402
from detectron2.structures import Boxes, Instances  # noqa: E402
from detectron2.utils.events import EventStorage  # noqa: E402
from wsl.modeling.poolers import ROIPooler  # noqa: E402
from wsl.modeling.roi_heads import box_head as ref_box_head_mod  # noqa: E402
from wsl.modeling.roi_heads import fast_rcnn_oicr, fast_rcnn_wsddn  # noqa: E402
from wsl.modeling.roi_heads.box_head import DiscriminativeAdaptionNeck  # noqa: E402
from wsl.modeling.roi_heads.roi_heads_oicrplus import OICRPlusHeads  # noqa: E402

from oracle import oicr_plus_ref as ora  # noqa: E402  (only its seeded input generator is used here)


def build_reference_heads(C, K, ch, fc, seed):
    tfm = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    torch.manual_seed(seed)
    pooler = ROIPooler(output_size=7, scales=(1.0 / 8,), sampling_ratio=0, pooler_type="ROIPool")
    head = DiscriminativeAdaptionNeck(ShapeSpec(channels=ch, height=7, width=7), conv_dims=[], fc_dims=[fc, fc])
    wl = fast_rcnn_wsddn.WSDDNOutputLayers(input_shape=ShapeSpec(channels=fc), box2box_transform=tfm, num_classes=C,
                                           mean_loss=True)
    refinery = [fast_rcnn_oicr.OICROutputLayers(input_shape=ShapeSpec(channels=fc), box2box_transform=tfm, num_classes=C,
                                                test_score_thresh=1e-6, test_nms_thresh=0.3, test_topk_per_image=100,
                                                refine_k=k, refine_reg=[True] * K) for k in range(K)]
    # sharpen the tiny random heads so that the view-averaged scores spread out (several seeds per class, all three
    # matcher labels present) -- the synthetic step should exercise the labelling, not sit at uniform scores
    wl.cls.weight.data.mul_(6.0)
    wl.det.weight.data.mul_(6.0)
    for r in refinery:
        r.cls_score.weight.data.mul_(40.0)
        r.bbox_pred.weight.data.mul_(40.0)
    cfg = types.SimpleNamespace(WSL=types.SimpleNamespace(REFINE_REG=[True] * K),
                                OICRPLUS=types.SimpleNamespace(BBOX_UPDATE=False))
    heads = OICRPlusHeads(box_in_features=["plain5"], box_pooler=pooler, box_head=head, box_predictor=wl, vis_period=0,
                          refine_K=K, refine_mist=True, mist_p=0.10, mist_thre=0.05, mist_type="nms",
                          refine_reg=[True] * K, box_refinery=refinery, cls_agnostic_bbox_reg=False, pooler_type="ROIPool",
                          cfg=cfg, num_classes=C, batch_size_per_image=4096, positive_fraction=1.0,
                          proposal_matcher=Matcher([0.5, 0.6], [0, -1, 1], allow_low_quality_matches=False),
                          proposal_append_gt=False)
    return heads


def run_case(heads, views, gt_classes, dropout):
    """One reference training forward + backward.  Returns losses, gradients and (if dropout) the keep-masks in the
    order the reference drew them: view 1 fc1, view 1 fc2, view 1_flip fc1, ... (box_head is called per view)."""
    heads.train()
    heads.box_head.train(dropout)
    for p in heads.parameters():
        p.grad = None
    f1 = torch.cat([views[0].feat, views[1].feat], 0).clone().requires_grad_(True)
    f2 = torch.cat([views[2].feat, views[3].feat], 0).clone().requires_grad_(True)
    props = [[Instances(v.image_size, proposal_boxes=Boxes(v.boxes.clone()), objectness_logits=v.obj.clone())] for v in views]
    targets = [Instances(views[0].image_size, gt_classes=gt_classes.clone(), gt_boxes=Boxes(torch.zeros(len(gt_classes), 4)))]
    masks = []
    real_dropout = ref_box_head_mod.F.dropout

    def recording_dropout(x, p=0.5, training=True, inplace=False):
        y = real_dropout(x, p=p, training=training, inplace=inplace)
        if training:
            # kept elements are x / (1 - p); x >= 0 after ReLU, so a kept zero is indistinguishable from a dropped one
            # and irrelevant to the result -- record it as kept
            masks.append(((y != 0) | (x == 0)).to(torch.uint8))
        return y

    patched = types.SimpleNamespace(**{k: getattr(ref_box_head_mod.F, k) for k in dir(ref_box_head_mod.F) if not k.startswith("__")})
    patched.dropout = recording_dropout
    ref_box_head_mod.F = patched
    try:
        with EventStorage(0) as storage:
            _, losses = heads([None] * 4, [{"plain5": f1}, {"plain5": f2}], props, [targets, None, None, None])
            total = sum(losses.values())
            total.backward()
            scalars = {k: float(v[0] if isinstance(v, tuple) else v) for k, v in storage.latest().items()}
    finally:
        ref_box_head_mod.F = torch.nn.functional
    grads = {n: p.grad.detach().clone() for n, p in heads.named_parameters()}
    out = {"losses": {k: v.detach().clone() for k, v in losses.items()}, "grads": grads,
           "grad_feat1": f1.grad.detach().clone(), "grad_feat2": f2.grad.detach().clone(), "storage": scalars}
    if dropout:
        assert len(masks) == 8, len(masks)
        out["drop_masks"] = [(masks[2 * i], masks[2 * i + 1]) for i in range(4)]
    return out


def main():
    out = {}
    for name, (C, K, R, ch, fc, sizes, seed, gt) in {
        "voc_k3": (20, 3, 300, 16, 64, [(240, 320), (288, 384)], 20261018, [3, 3, 11]),
        "coco_k4": (80, 4, 250, 8, 48, [(224, 288), (256, 336)], 20261019, [5, 17, 17, 42, 63]),
    }.items():
        g = torch.Generator().manual_seed(seed)
        views = ora.synth_views(R, sizes, g, channels=ch)
        heads = build_reference_heads(C, K, ch, fc, seed)
        gt_classes = torch.tensor(gt)
        case = {"C": C, "K": K, "R": R, "ch": ch, "fc": fc,
                "views": [{"feat": v.feat, "boxes": v.boxes, "obj": v.obj, "image_size": v.image_size} for v in views],
                "gt_classes": gt_classes, "params": {n: p.detach().clone() for n, p in heads.named_parameters()}}
        case["eval_dropout"] = run_case(heads, views, gt_classes, dropout=False)
        torch.manual_seed(seed + 1)
        case["train_dropout"] = run_case(heads, views, gt_classes, dropout=True)
        out[name] = case
        print(name, {k: round(float(v), 6) for k, v in case["eval_dropout"]["losses"].items()})
        print(name, "dropout", {k: round(float(v), 6) for k, v in case["train_dropout"]["losses"].items()})
        print(name, "storage", case["eval_dropout"]["storage"])
    dst = os.path.join(os.environ.get("SOSWSOD_GOLDEN_OUT", HERE), "step_golden.pt")
    # fp16-free, but drop what the tests do not need to keep the fixture small
    torch.save(out, dst)
    print("wrote step_golden.pt", os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
