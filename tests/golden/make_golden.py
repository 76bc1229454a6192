"""Generates tests/golden/oicr_plus_golden.pt by running the REFERENCE'S OWN CODE (imported read-only from
/root/reference through tests/golden/ref_import.py) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

What is produced by the reference itself (fp32, CPU, torch 2.11 / torchvision 0.26):
  iou        detectron2.structures.pairwise_iou                        (uwsod/detectron2/structures/boxes.py:329-361)
  matcher    detectron2.modeling.matcher.Matcher, thresholds [0.5,0.6] (uwsod/detectron2/modeling/matcher.py:63-111)
  deltas     Box2BoxTransform.get_deltas / apply_deltas                (uwsod/detectron2/modeling/box_regression.py)
  pool       wsl ROIPooler(ROIPool, 7, 1/8)                            (wsl/modeling/poolers.py:221-270)
  head       DiscriminativeAdaptionNeck (eval)                         (wsl/modeling/roi_heads/box_head.py:82-91)
  wsddn      WSDDNOutputLayers.forward + WSDDNOutputs loss             (wsl/modeling/roi_heads/fast_rcnn_wsddn.py)
  pgt        OICRPlusHeads.get_pgt_mist (seed mining + NMS 0.01)       (wsl/modeling/roi_heads/roi_heads_oicrplus.py:559-757)
  labels     ROIHeads.label_and_sample_proposals                       (wsl/modeling/roi_heads/roi_heads.py:266-375)
  oicr       OICROutputs weighted CE + L1 box loss                     (wsl/modeling/roi_heads/fast_rcnn_oicr.py:157-352)
  infer      predict_probs_K / predict_boxes_K / fast_rcnn_inference   (wsl/modeling/roi_heads/fast_rcnn_oicr.py:46-148,674-735)
The known-answer tests of the reference's own unit tests are copied as DATA (inputs + expected values):
  U/tests/structures/test_boxes.py:150-173 and U/tests/modeling/test_matcher.py:19-27.
"""
import math
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
from detectron2.structures import Boxes, Instances, pairwise_iou  # noqa: E402
from detectron2.utils.events import EventStorage  # noqa: E402
from wsl.modeling.poolers import ROIPooler  # noqa: E402
from wsl.modeling.roi_heads import fast_rcnn_oicr, fast_rcnn_wsddn  # noqa: E402
from wsl.modeling.roi_heads.box_head import DiscriminativeAdaptionNeck  # noqa: E402
from wsl.modeling.roi_heads.roi_heads import ROIHeads, get_image_level_gt  # noqa: E402
from wsl.modeling.roi_heads.roi_heads_oicrplus import OICRPlusHeads  # noqa: E402

from oracle import oicr_plus_ref as ora  # noqa: E402  (only its seeded input generators are used here)


def main():
    out = {}
    g = torch.Generator().manual_seed(20261017)
    C, K, R, ch, fc = 20, 3, 300, 16, 64
    views = ora.synth_views(R, [(240, 320), (288, 384)], g, channels=ch)
    v0 = views[0]
    boxes = v0.boxes
    out["inputs"] = {"boxes": boxes, "feat": v0.feat, "obj": v0.obj, "C": C, "K": K, "image_size": v0.image_size,
                     "boxes_view2": views[2].boxes}

    # ---- reference KATs (data copied from the reference's unit tests) ----
    b1 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    b2 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 1.0, 0.5], [0.0, 0.0, 0.5, 0.5],
                       [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.5, 1.5]])
    exp = torch.tensor([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)], [1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]])
    out["kat_iou"] = {"boxes1": b1, "boxes2": b2, "expected": exp, "reference_output": pairwise_iou(Boxes(b1), Boxes(b2))}
    q = torch.tensor([[0.15, 0.45, 0.2, 0.6], [0.3, 0.65, 0.05, 0.1], [0.05, 0.4, 0.25, 0.4]])
    m, l = Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(q)
    out["kat_matcher"] = {"quality": q, "expected_matches": torch.tensor([1, 1, 2, 0]),
                          "expected_labels": torch.tensor([-1, 1, 0, 1], dtype=torch.int8), "ref_matches": m, "ref_labels": l}

    # ---- pairwise IoU / Matcher / deltas on the synthetic proposals ----
    seeds_b = boxes[torch.randperm(R, generator=g)[:17]]
    iou = pairwise_iou(Boxes(seeds_b), Boxes(boxes))
    mm, ll = Matcher([0.5, 0.6], [0, -1, 1], allow_low_quality_matches=False)(iou)
    out["iou"] = {"seeds": seeds_b, "iou": iou, "matches": mm, "labels": ll}
    tfm = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    tgt = boxes[torch.randperm(R, generator=g)]
    d = tfm.get_deltas(boxes, tgt)
    dd = torch.randn((R, 4 * C), generator=g) * 0.5
    out["deltas"] = {"target": tgt, "get_deltas": d, "rand_deltas": dd, "apply_deltas": tfm.apply_deltas(dd, boxes)}

    # ---- pooler + box head + WSDDN ----
    pooler = ROIPooler(output_size=7, scales=(1.0 / 8,), sampling_ratio=0, pooler_type="ROIPool")
    pooled = pooler([v0.feat], [Boxes(boxes)])
    out["pool"] = {"pooled": pooled}
    torch.manual_seed(7)
    head = DiscriminativeAdaptionNeck(ShapeSpec(channels=ch, height=7, width=7), conv_dims=[], fc_dims=[fc, fc]).eval()
    x = head(pooled * (v0.obj + 1).view(-1, 1, 1, 1))
    out["head"] = {"fc1_w": head.fc1.weight.detach(), "fc1_b": head.fc1.bias.detach(), "fc2_w": head.fc2.weight.detach(),
                   "fc2_b": head.fc2.bias.detach(), "x": x.detach()}
    wl = fast_rcnn_wsddn.WSDDNOutputLayers(input_shape=ShapeSpec(channels=fc), box2box_transform=tfm, num_classes=C,
                                           mean_loss=True)
    wl.cls.weight.data.mul_(4.0)
    wl.det.weight.data.mul_(4.0)
    props = [Instances(v0.image_size, proposal_boxes=Boxes(boxes), objectness_logits=v0.obj)]
    gt_targets = [Instances(v0.image_size, gt_classes=torch.tensor([3, 3, 11]), gt_boxes=Boxes(torch.zeros(3, 4)))]
    _, gt_int, gt_oh = get_image_level_gt(gt_targets, C)
    with EventStorage(0):
        preds = wl(x, props)
        wloss = wl.losses(preds, props, gt_oh)
    out["wsddn"] = {"cls_w": wl.cls.weight.detach(), "cls_b": wl.cls.bias.detach(), "det_w": wl.det.weight.detach(),
                    "det_b": wl.det.bias.detach(), "scores": preds[0].detach(), "loss_cls": wloss["loss_cls"].detach(),
                    "gt_int": gt_int[0], "gt_oh": gt_oh}

    # ---- pseudo-GT mining + labelling, by the reference's own methods on a stand-in `self` ----
    prev = [ora.synth_prev_scores(R, C, g)] + [ora.synth_prev_scores(R, C + 1, g) for _ in range(K - 1)]
    for p in prev:
        p[:, 3] *= 3.0
    fake = types.SimpleNamespace(num_classes=C, cls_agnostic_bbox_reg=False, gt_classes_img_int=gt_int,
                                 proposal_matcher=Matcher([0.5, 0.6], [0, -1, 1], allow_low_quality_matches=False),
                                 proposal_append_gt=False, batch_size_per_image=4096, positive_sample_fraction=1.0)
    fake.get_pgt_top_k = types.MethodType(OICRPlusHeads.get_pgt_top_k, fake)
    fake._sample_proposals = types.MethodType(ROIHeads._sample_proposals, fake)
    branches = []
    with EventStorage(0):
        for k in range(K):
            tg = OICRPlusHeads.get_pgt_mist(fake, [Boxes(boxes)], [prev[k]], props, top_pro=0.10, thres=0.05)
            lab = ROIHeads.label_and_sample_proposals(fake, [Instances(v0.image_size, proposal_boxes=Boxes(boxes),
                                                                       objectness_logits=v0.obj)], tg, suffix=f"_r{k}")
            t, p = tg[0], lab[0]
            branches.append({"prev": prev[k], "seed_boxes": t.gt_boxes.tensor, "seed_classes": t.gt_classes,
                             "seed_scores": t.gt_scores, "seed_index": t.gt_index, "gt_classes": p.gt_classes,
                             "gt_weights": p.gt_weights, "gt_index": p.gt_index, "gt_boxes": p.gt_boxes.tensor})
    out["branches"] = branches

    # ---- OICR losses + K-branch inference ----
    layers = []
    preds_K = []
    oicr = []
    with EventStorage(0):
        for k in range(K):
            torch.manual_seed(100 + k)
            ol = fast_rcnn_oicr.OICROutputLayers(input_shape=ShapeSpec(channels=fc), box2box_transform=tfm, num_classes=C,
                                                 test_score_thresh=1e-6, test_nms_thresh=0.3, test_topk_per_image=100,
                                                 refine_k=k, refine_reg=[True] * K)
            ol.cls_score.weight.data.mul_(30.0)
            ol.bbox_pred.weight.data.mul_(30.0)
            pk = ol(x)
            b = branches[k]
            pr = [Instances(v0.image_size, proposal_boxes=Boxes(boxes), objectness_logits=v0.obj, gt_classes=b["gt_classes"],
                            gt_weights=b["gt_weights"], gt_boxes=Boxes(b["gt_boxes"]), gt_index=b["gt_index"])]
            ls = ol.losses(pk, pr)
            layers.append(ol)
            preds_K.append(pk)
            oicr.append({"cls_w": ol.cls_score.weight.detach(), "cls_b": ol.cls_score.bias.detach(),
                         "box_w": ol.bbox_pred.weight.detach(), "box_b": ol.bbox_pred.bias.detach(),
                         "logits": pk[0].detach(), "deltas": pk[1].detach(),
                         "losses": {kk: vv.detach() for kk, vv in ls.items()}})
        with torch.no_grad():   # the reference runs inference under no_grad (Boxes.clip is in-place)
            inst, inds, all_scores, all_boxes = layers[-1].inference([(a.detach(), b.detach()) for a, b in preds_K], props)
    out["oicr"] = oicr
    out["infer"] = {"all_scores": all_scores[0].detach(), "all_boxes": all_boxes[0].detach(),
                    "pred_boxes": inst[0].pred_boxes.tensor.detach(), "scores": inst[0].scores.detach(),
                    "pred_classes": inst[0].pred_classes, "pred_inds": inds[0]}
    dst = os.path.join(os.environ.get("SOSWSOD_GOLDEN_OUT", HERE), "oicr_plus_golden.pt")
    torch.save(out, dst)
    sz = os.path.getsize(dst)
    print("wrote oicr_plus_golden.pt", sz, "bytes;", {k: (len(v) if hasattr(v, "__len__") else "") for k, v in out.items()})


if __name__ == "__main__":
    main()
