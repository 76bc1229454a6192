"""CPU suite, part 1: pins the oracle (oracle/oicr_plus_ref.py + oracle/ref_kernels.c) against
  (a) the reference's own known-answer tests (U/tests/structures/test_boxes.py:150-173,
      U/tests/modeling/test_matcher.py:19-27, the round trip of U/tests/modeling/test_box2box_transform.py:16-31),
  (b) tests/golden/oicr_plus_golden.pt -- outputs of the REFERENCE'S OWN CODE on seeded inputs
      (tests/golden/make_golden.py imports the reference modules from /root/reference), and
  (c) torchvision's CPU kernels for the scalar C restatement.
Nothing here needs a GPU or /root/reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oicr_plus_ref as ref
from oracle import ref_kernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oicr_plus_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


# ---------------------------------------------------------------- (a) reference KATs
def test_kat_pairwise_iou(gold):
    k = gold["kat_iou"]
    got = ref.pairwise_iou(k["boxes1"], k["boxes2"])
    assert torch.allclose(got, k["expected"])
    assert torch.equal(got, k["reference_output"])


def test_kat_matcher(gold):
    k = gold["kat_matcher"]
    m, l = ref.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(k["quality"])
    assert torch.equal(m, k["expected_matches"]) and torch.equal(l, k["expected_labels"])
    assert torch.equal(m, k["ref_matches"]) and torch.equal(l, k["ref_labels"])


def test_box2box_round_trip():
    g = torch.Generator().manual_seed(0)
    src = ref.synth_boxes(50, 300, 400, g)
    dst = ref.synth_boxes(50, 300, 400, g)
    d = ref.get_deltas(src, dst)
    back = ref.apply_deltas(d, src)
    assert torch.allclose(back, dst, atol=1e-3)


# ---------------------------------------------------------------- (b) outputs of the reference's own code
def test_iou_matcher_deltas_vs_reference(gold):
    boxes = gold["inputs"]["boxes"]
    i = gold["iou"]
    q = ref.pairwise_iou(i["seeds"], boxes)
    assert torch.equal(q, i["iou"]), "pairwise_iou must be bit-equal to detectron2's"
    m, l = ref.Matcher([0.5, 0.6], [0, -1, 1])(q)
    assert torch.equal(m, i["matches"]) and torch.equal(l, i["labels"])
    d = gold["deltas"]
    assert torch.equal(ref.get_deltas(boxes, d["target"]), d["get_deltas"])
    assert torch.equal(ref.apply_deltas(d["rand_deltas"], boxes), d["apply_deltas"])


def test_pool_head_wsddn_vs_reference(gold):
    inp = gold["inputs"]
    pooled, _ = ref.roi_pool(inp["feat"], ref.boxes_to_pooler_format([inp["boxes"]]))
    assert torch.equal(pooled, gold["pool"]["pooled"])
    h, w = gold["head"], gold["wsddn"]
    p = ref.HeadParams(h["fc1_w"], h["fc1_b"], h["fc2_w"], h["fc2_b"], w["cls_w"], w["cls_b"], w["det_w"], w["det_b"])
    x = ref.box_head(ref.roi_pool_scaled(inp["feat"], inp["boxes"], inp["obj"]), p)
    assert torch.equal(x, h["x"])
    s = ref.wsddn_scores(x, p)
    assert torch.equal(s, w["scores"])
    gt_int, gt_oh = ref.image_level_gt(torch.tensor([3, 3, 11]), inp["C"])
    assert torch.equal(gt_int, w["gt_int"]) and torch.equal(gt_oh, w["gt_oh"])
    assert torch.equal(ref.wsddn_loss(s, gt_oh), w["loss_cls"])


def test_pgt_mining_and_labels_vs_reference(gold):
    inp = gold["inputs"]
    gt_int = gold["wsddn"]["gt_int"]
    for b in gold["branches"]:
        seeds = ref.pgt_mist(inp["boxes"], b["prev"], gt_int, 0.10, 0.05)
        assert torch.equal(seeds.index, b["seed_index"]), "seed proposal indices"
        assert torch.equal(seeds.classes, b["seed_classes"])
        assert torch.equal(seeds.scores, b["seed_scores"])
        assert torch.equal(seeds.boxes, b["seed_boxes"])
        y, w, gi, _, gb = ref.label_proposals(inp["boxes"], seeds, inp["C"])
        assert torch.equal(y, b["gt_classes"]), "pseudo-labels"
        assert torch.equal(w, b["gt_weights"])
        assert torch.equal(gi, b["gt_index"])
        assert torch.equal(gb, b["gt_boxes"])
        assert int((y == -1).sum()) + int((y == inp["C"]).sum()) < y.numel(), "fixture must contain foreground rows"


def test_oicr_losses_and_inference_vs_reference(gold):
    inp = gold["inputs"]
    x = gold["head"]["x"]
    C, K = inp["C"], inp["K"]
    ZK, DK = [], []
    for k, (o, b) in enumerate(zip(gold["oicr"], gold["branches"])):
        z, d = ref.refine_forward(x, (o["cls_w"], o["cls_b"], o["box_w"], o["box_b"]))
        assert torch.equal(z, o["logits"]) and torch.equal(d, o["deltas"])
        lc = ref.oicr_cls_loss(z, b["gt_classes"], b["gt_weights"])
        lb = ref.oicr_box_loss(d, b["gt_classes"], inp["boxes"], inp["boxes"][b["gt_index"]], C)
        torch.testing.assert_close(lc, o["losses"][f"loss_cls_r{k}"], rtol=1e-6, atol=1e-8)
        torch.testing.assert_close(lb, o["losses"][f"loss_box_reg_r{k}"], rtol=1e-5, atol=1e-8)
        ZK.append(z)
        DK.append(d)
    inf = gold["infer"]
    probs, pboxes = ref.predict_probs_K(ZK), ref.predict_boxes_K(DK, inp["boxes"])
    assert torch.equal(probs, inf["all_scores"].reshape(-1, C + 1)), "mean_k softmax(logits_k)"
    assert torch.equal(pboxes, inf["all_boxes"].reshape(-1, 4 * C)), "apply_deltas(mean_k deltas_k)"
    eb, es, ec, er = ref.fast_rcnn_inference_single_image(pboxes, probs, inp["image_size"], 1e-6, 0.3, 100)
    assert torch.equal(er, inf["pred_inds"]), "detection keep-list"
    assert torch.equal(ec, inf["pred_classes"]) and torch.equal(es, inf["scores"]) and torch.equal(eb, inf["pred_boxes"])


# ---------------------------------------------------------------- (c) scalar C restatement vs torchvision
def _edge_rois(g, R, h, w):
    boxes = ref.synth_boxes(R, h * 8, w * 8, g)
    extra = torch.tensor([[0.0, 0.0, w * 8 - 1.0, h * 8 - 1.0], [4.0, 4.0, 4.0, 4.0], [12.0, 20.0, 28.0, 36.0],
                          [100.0, 50.0, 60.0, 30.0], [-50.0, -40.0, 30.0, 30.0], [w * 8 + 100.0, h * 8 + 100.0, w * 8 + 200.0, h * 8 + 300.0],
                          [3.7, 9.2, 200.3, 150.9]])
    return torch.cat([boxes, extra], 0)


@pytest.mark.parametrize("C,h,w,R", [(6, 30, 40, 60), (3, 17, 23, 25)])
def test_c_roi_pool_equals_torchvision(C, h, w, R):
    import torchvision

    g = torch.Generator().manual_seed(C)
    feat = torch.randn((2, C, h, w), generator=g).requires_grad_(True)
    b0, b1 = _edge_rois(g, R, h, w), _edge_rois(g, R, h, w)
    rois = ref.boxes_to_pooler_format([b0, b1])
    exp, exp_arg = ref.roi_pool(feat.detach(), rois)
    out, arg = ref_kernels.roi_pool_forward(feat.detach().numpy(), rois.numpy(), 7, 0.125)
    assert np.array_equal(out, exp.numpy()) and np.array_equal(arg, exp_arg.numpy())
    go = torch.randn(exp.shape, generator=g)
    torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125).backward(go)
    gf = ref_kernels.roi_pool_backward(go.numpy(), arg, rois.numpy(), feat.shape)
    np.testing.assert_allclose(gf, feat.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,thr", [(1, 0.5), (200, 0.3), (1500, 0.01), (1500, 0.7)])
def test_c_nms_and_iou_equal_library(n, thr):
    g = torch.Generator().manual_seed(n)
    boxes = ref.synth_boxes(n, 480, 640, g) if n > 1 else torch.tensor([[1.0, 1.0, 5.0, 5.0]])
    scores = torch.rand(n, generator=g)
    if n > 10:
        scores[3] = scores[7]
    keep = ref_kernels.nms(boxes.numpy(), scores.numpy(), thr)
    assert np.array_equal(keep, ref.nms(boxes, scores, thr).numpy())
    q = ref_kernels.pairwise_iou(boxes[:50].numpy(), boxes.numpy())
    assert np.array_equal(q, ref.pairwise_iou(boxes[:50], boxes).numpy())


def test_batched_nms_is_per_class_and_score_sorted():
    g = torch.Generator().manual_seed(5)
    boxes = ref.synth_boxes(400, 300, 300, g)
    scores = torch.rand(400, generator=g)
    idxs = torch.randint(0, 5, (400,), generator=g)
    keep = ref.batched_nms(boxes, scores, idxs, 0.3)
    assert (scores[keep][:-1] >= scores[keep][1:]).all()
    for c in range(5):
        m = (idxs == c).nonzero().view(-1)
        assert set(m[ref.nms(boxes[m], scores[m], 0.3)].tolist()) == set(keep[idxs[keep] == c].tolist())
    assert ref.batched_nms(boxes[:0], scores[:0], idxs[:0], 0.3).numel() == 0


def test_train_step_structure_and_quirk():
    """Loss keys of roi_heads_oicrplus.py:283-388; the 2_flip quirk (:381) changes only the refinement losses."""
    g = torch.Generator().manual_seed(1)
    C, K, R = 20, 3, 60
    views = ref.synth_views(R, [(160, 192), (192, 240)], g, channels=8)
    p = ref.init_head_params(C, K, in_dim=8 * 49, fc_dim=32, generator=g)
    for r in p.refine:
        r[0].mul_(30.0)
    gt = torch.tensor([2, 5])
    a, _ = ref.train_step(views, gt, p, C, K, reproduce_flip_quirk=True)
    b, _ = ref.train_step(views, gt, p, C, K, reproduce_flip_quirk=False)
    assert sorted(a) == sorted(["loss_cls"] + [f"loss_cls_r{k}" for k in range(K)] + [f"loss_box_reg_r{k}" for k in range(K)])
    assert torch.equal(a["loss_cls"], b["loss_cls"])
    assert any(not torch.equal(a[f"loss_cls_r{k}"], b[f"loss_cls_r{k}"]) for k in range(K))


def test_voc_writer_rows_and_tta_inverse():
    boxes = torch.tensor([[10.04, 20.06, 30.0, 40.0], [1.0, 2.0, 3.0, 4.0]])
    rows = ref.voc_detection_rows(17, boxes, torch.tensor([0.98765, 0.5]), torch.tensor([4, 0]))
    assert rows[0] == {"image_id": 17, "category_id": 1, "score": 0.5, "bbox": [2.0, 3.0, 3.0, 4.0]}
    assert rows[1] == {"image_id": 17, "category_id": 5, "score": 0.988, "bbox": [11.0, 21.1, 30.0, 40.0]}
    b = torch.tensor([[10.0, 5.0, 30.0, 25.0]])
    out = ref.tta_inverse_boxes(b, 0.5, 0.5, True, 100)
    assert torch.equal(out, torch.tensor([[35.0, 2.5, 45.0, 12.5]]))


# ---------------------------------------------------------------- (d) test-time augmentation (row U) vs the reference
TTA_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tta_golden.pt")


@pytest.fixture(scope="module")
def tta_gold():
    return torch.load(TTA_GOLDEN, weights_only=False)


def _tta_post(c):
    (h, w), (dh, dw) = c["stored_hw"], c["dataset_hw"]
    return None if (h, w) == (dh, dw) else (dw * 1.0 / w, dh * 1.0 / h)


def test_tta_views_match_reference_mapper(tta_gold):
    """DatasetMapperTTAAVG.__call__ of the reference: view order, sizes and the transformed proposals, bit for bit."""
    for c in tta_gold["cases"]:
        h, w = c["stored_hw"]
        sizes = ref.tta_view_sizes(h, w, c["min_sizes"], c["max_size"], c["flip"])
        assert len(sizes) == len(c["views"]), c["name"]
        boxes = c["boxes"][: c["topk"]]
        for (nh, nw, fl), v in zip(sizes, c["views"]):
            assert v["image_shape"] == (3, nh, nw) and v["image_size"] == (nh, nw)
            assert ("HFlipTransform" in v["transforms"]) == fl
            b, keep = ref.tta_transform_proposals(boxes, (h, w), (nh, nw), fl)
            assert bool(keep.all())
            assert torch.equal(b, v["proposal_boxes"]), c["name"]
            assert torch.equal(c["obj"][: c["topk"]], v["objectness_logits"])


def test_tta_merge_and_detections_match_reference(tta_gold):
    """GeneralizedRCNNWithTTAAVG._get_augmented_boxes / _merge_detections of the reference, bit for bit."""
    for c in tta_gold["cases"]:
        h, w = c["stored_hw"]
        C = c["C"]
        sizes = ref.tta_view_sizes(h, w, c["min_sizes"], c["max_size"], c["flip"])
        eb = [ref.tta_inverse_boxes(c["view_boxes"][i].reshape(-1, 4), w * 1.0 / nw, h * 1.0 / nh, fl, nw, _tta_post(c))
              .reshape(-1, 4 * C) for i, (nh, nw, fl) in enumerate(sizes)]
        mb, mp = ref.tta_merge(eb, [c["view_scores"][i] for i in range(len(sizes))])
        assert torch.equal(mb, c["merged_boxes"]) and torch.equal(mp, c["merged_scores"]), c["name"]
        db, ds, dc, _ = ref.fast_rcnn_inference_single_image(mb, mp, c["dataset_hw"], 1e-6, 0.3, 100)
        assert torch.equal(db, c["det_boxes"]) and torch.equal(ds, c["det_scores"]) and torch.equal(dc, c["det_classes"])


def test_tta_transform_edge_cases():
    # malformed (x2 < x1) boxes come out ordered (corner min / max); boxes outside the view are clipped to empty
    b = torch.tensor([[30.0, 10.0, 10.0, 40.0], [-50.0, -50.0, -10.0, -5.0], [5.0, 5.0, 5.0, 30.0]])
    out, keep = ref.tta_transform_proposals(b, (60, 80), (120, 160), True)
    assert torch.equal(out[0], torch.tensor([100.0, 20.0, 140.0, 80.0]))
    assert keep.tolist() == [True, False, False]
    out2, keep2 = ref.tta_transform_proposals(b[:1], (60, 80), (120, 160), False, min_box_size=40.0)
    assert keep2.tolist() == [False] and torch.equal(out2[0], torch.tensor([20.0, 20.0, 60.0, 80.0]))


def test_tta_scalar_c_restatement_matches_reference_fixtures(tta_gold):
    """oracle/ref_kernels.c (independent scalar restatement) against the reference-generated TTA fixtures and the
    torch oracle: transformed proposals bit for bit; inverse transforms bit for bit (so the mean over views is too)."""
    for c in tta_gold["cases"]:
        h, w = c["stored_hw"]
        C = c["C"]
        sizes = ref.tta_view_sizes(h, w, c["min_sizes"], c["max_size"], c["flip"])
        boxes = c["boxes"][: c["topk"]]
        post = None if c["stored_hw"] == c["dataset_hw"] else c["dataset_hw"]
        inv = []
        for i, ((nh, nw, fl), v) in enumerate(zip(sizes, c["views"])):
            out, keep = ref_kernels.tta_transform_proposals(boxes.numpy(), (h, w), (nh, nw), fl)
            assert keep.all() and np.array_equal(out, v["proposal_boxes"].numpy()), (c["name"], i)
            b = ref_kernels.tta_inverse_boxes(c["view_boxes"][i].reshape(-1, 4).numpy(), (h, w), (nh, nw), fl, post)
            e = ref.tta_inverse_boxes(c["view_boxes"][i].reshape(-1, 4), w * 1.0 / nw, h * 1.0 / nh, fl, nw, _tta_post(c))
            assert np.array_equal(b, e.numpy()), (c["name"], i)
            inv.append(torch.from_numpy(b).reshape(-1, 4 * C))
        assert torch.equal(torch.stack(inv).mean(0), c["merged_boxes"]), c["name"]
    # edge cases: malformed, outside, degenerate, min size
    b = np.array([[30.0, 10.0, 10.0, 40.0], [-50.0, -50.0, -10.0, -5.0], [5.0, 5.0, 5.0, 30.0]], np.float32)
    out, keep = ref_kernels.tta_transform_proposals(b, (60, 80), (120, 160), True)
    eo, ek = ref.tta_transform_proposals(torch.from_numpy(b), (60, 80), (120, 160), True)
    assert np.array_equal(out, eo.numpy()) and keep.tolist() == ek.tolist() == [True, False, False]


# ---------------------------------------------------------------- (d) one WHOLE step of the reference's own OICRPlusHeads.forward
@pytest.mark.parametrize("name", ["voc_k3", "coco_k4"])
@pytest.mark.parametrize("mode", ["eval_dropout", "train_dropout"])
def test_train_step_vs_reference_forward_box(name, mode):
    """oracle.train_step against tests/golden/step_golden.pt: losses, parameter gradients and conv5 gradients of
    `OICRPlusHeads.forward` + backward run from the reference's own class (roi_heads_oicrplus.py:149-430) -- pins the
    view averaging (:290-294, :390-395), the /4.0 combines (:288, :384-388) and the `2_flip` quirk (:381)."""
    import step_golden_util as sg

    case = sg.load()[name]
    exp = case[mode]
    p = sg.head_params(case).requires_grad_(True)
    vs = sg.views(case)
    for v in vs:
        v.feat.requires_grad_(True)
    masks = None
    if mode == "train_dropout":
        masks = [(a.float(), b.float()) for a, b in exp["drop_masks"]]
    losses, aux = ref.train_step(vs, case["gt_classes"], p, case["C"], case["K"], drop_masks=masks)
    assert set(losses) == set(exp["losses"])
    for k, v in exp["losses"].items():
        assert abs(float(losses[k]) - float(v)) <= 1e-6 * max(1.0, abs(float(v))), (k, float(losses[k]), float(v))
    sum(losses.values()).backward()
    for n, t in zip(sg.grad_key_map(case["K"]), p.tensors()):
        e = exp["grads"][n]
        assert torch.allclose(t.grad, e, rtol=1e-4, atol=1e-7 + 1e-5 * float(e.abs().max())), (n, float((t.grad - e).abs().max()))
    g1 = torch.cat([vs[0].feat.grad, vs[1].feat.grad], 0)
    g2 = torch.cat([vs[2].feat.grad, vs[3].feat.grad], 0)
    for got, e in ((g1, exp["grad_feat1"]), (g2, exp["grad_feat2"])):
        assert torch.allclose(got, e, rtol=1e-4, atol=1e-7 + 1e-5 * float(e.abs().max()))
    # without the quirk the step must differ (the fixture really exercises :381)
    l2, _ = ref.train_step(sg.views(case), case["gt_classes"], sg.head_params(case), case["C"], case["K"], drop_masks=masks,
                           reproduce_flip_quirk=False)
    assert any(abs(float(l2[k]) - float(exp["losses"][k])) > 1e-7 for k in l2 if k != "loss_cls")


def test_reference_metric_scalars():
    """The EventStorage scalars the reference's step logs (names + values of the last view written):
    roi_head/num_{fg,bg,ig}_samples_r{k} (roi_heads.py:364-373) and fast_rcnn/{cls_accuracy,fg_cls_accuracy,
    false_negative}_r{k} (fast_rcnn_oicr.py:228-256) -- the oracle's counters reproduce them."""
    import step_golden_util as sg

    case = sg.load()["voc_k3"]
    exp = case["eval_dropout"]["storage"]
    vs = sg.views(case)
    losses, aux = ref.train_step(vs, case["gt_classes"], sg.head_params(case), case["C"], case["K"])
    C = case["C"]
    for k, br in enumerate(aux["branches"]):
        y = br["gt_classes"]
        assert exp[f"roi_head/num_fg_samples_r{k}"] == float(((y >= 0) & (y < C)).sum())
        assert exp[f"roi_head/num_bg_samples_r{k}"] == float((y == C).sum())
        assert exp[f"roi_head/num_ig_samples_r{k}"] == float((y == -1).sum())
        # _log_accuracy is called per view; the storage keeps the LAST call = view 2_flip, which (quirk) uses view 2's logits
        m = ref.reference_accuracy_scalars(br["logits"][2], y, C)
        for key in ("cls_accuracy", "fg_cls_accuracy", "false_negative"):
            if f"fast_rcnn/{key}_r{k}" in exp:
                assert abs(exp[f"fast_rcnn/{key}_r{k}"] - m[key]) < 1e-6, (key, k, exp[f"fast_rcnn/{key}_r{k}"], m[key])
