"""GPU parity tests of the test-time-augmentation path against
  * tests/golden/tta_golden.pt -- outputs of the REFERENCE'S OWN DatasetMapperTTAAVG / GeneralizedRCNNWithTTAAVG code
    (tests/golden/make_golden_tta.py), and
  * the CPU oracle on seeded inputs with edge cases.
Bars: transformed proposals, keep masks and keep-lists bit-exact; merged boxes / scores to fp32 rounding of the
mean over views (the summation order of torch.mean is an implementation detail of the reference's torch build)."""
import os

import pytest
import torch

from oracle import oicr_plus_ref as ref

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tta_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from sos_wsod_b200 import ops as _ops

    return _ops


def _specs(c):
    from sos_wsod_b200.modeling.test_time_augmentation_avg import ViewSpec, resize_shortest_edge

    h, w = c["stored_hw"]
    orig = None if c["stored_hw"] == c["dataset_hw"] else c["dataset_hw"]
    out = []
    for s in c["min_sizes"]:
        nh, nw = resize_shortest_edge(h, w, s, c["max_size"])
        out.append(ViewSpec(h, w, nh, nw, False, orig))
        if c["flip"]:
            out.append(ViewSpec(h, w, nh, nw, True, orig))
    return out


def test_tta_views_vs_reference_mapper(ops, gold):
    for c in gold["cases"]:
        specs = _specs(c)
        boxes = c["boxes"][: c["topk"]].cuda()
        rois, keep, dropped = ops.tta_views(boxes, [s.params(float(i % 2)) for i, s in enumerate(specs)])
        V, R = len(specs), boxes.size(0)
        rois = rois.view(V, R, 5).cpu()
        assert dropped.cpu().tolist() == [0] * V and bool(keep.all())
        for i, v in enumerate(c["views"]):
            assert torch.equal(rois[i, :, 1:], v["proposal_boxes"]), (c["name"], i)
            assert torch.all(rois[i, :, 0] == float(i % 2))


def test_tta_merge_vs_reference_wrapper(ops, gold):
    for c in gold["cases"]:
        specs = _specs(c)
        mb, mp = ops.tta_merge(c["view_boxes"].cuda(), c["view_scores"].cuda(), [s.params() for s in specs])
        torch.testing.assert_close(mb.cpu(), c["merged_boxes"], rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(mp.cpu(), c["merged_scores"], rtol=1e-6, atol=1e-9)
        # the reference's final detections from the reference's merged fp32 values: bit-exact keep-lists
        db, ds, dc, dr, nd = ops.detect(c["merged_scores"].cuda(), c["merged_boxes"].cuda(), c["dataset_hw"], 1e-6, 0.3, 100)
        n = int(nd.item())
        assert n == len(c["det_scores"]), c["name"]
        assert torch.equal(ds[:n].cpu(), c["det_scores"]) and torch.equal(dc[:n].cpu().long(), c["det_classes"])
        assert torch.equal(db[:n].cpu(), c["det_boxes"])


def test_tta_views_edge_cases_vs_oracle(ops):
    g = torch.Generator().manual_seed(3)
    h, w = 375, 500
    boxes = ref.synth_boxes(500, h, w, g)
    extra = torch.tensor([[30.0, 10.0, 10.0, 40.0],       # malformed: comes out ordered
                          [-50.0, -50.0, -10.0, -5.0],    # outside: clipped to empty
                          [5.0, 5.0, 5.0, 30.0],          # zero width
                          [0.0, 0.0, float(w), float(h)],
                          [w - 0.25, 3.0, w + 40.0, 90.0]])
    boxes = torch.cat([boxes, extra], 0)
    from sos_wsod_b200.modeling.test_time_augmentation_avg import ViewSpec, resize_shortest_edge

    specs = []
    for s in (480, 576, 688, 864, 1200):
        nh, nw = resize_shortest_edge(h, w, s, 1400)
        specs += [ViewSpec(h, w, nh, nw, False), ViewSpec(h, w, nh, nw, True)]
    for min_size in (0.0, 20.0):
        rois, keep, dropped = ops.tta_views(boxes.cuda(), [s.params() for s in specs], min_size)
        rois = rois.view(len(specs), -1, 5).cpu()
        for i, s in enumerate(specs):
            eb, ek = ref.tta_transform_proposals(boxes, (h, w), s.image_size, s.flip, min_size)
            assert torch.equal(rois[i, :, 1:], eb) and torch.equal(keep[i].cpu(), ek)
            assert int(dropped[i]) == int((~ek).sum())
    assert int(dropped.sum()) > 0
    with pytest.raises(RuntimeError, match="1 <= V <= 32"):
        ops.tta_views(boxes.cuda(), [specs[0].params()] * 33)


def test_tta_merge_matches_per_view_accumulate_and_oracle(ops):
    """One-launch merge == the per-view running sum (soswsod_tta_accumulate) == oracle; V = 10 and 16 views at the
    BASELINE row count (R = 2000, C = 20)."""
    g = torch.Generator().manual_seed(5)
    R, C = 2000, 20
    h, w = 480, 640
    from sos_wsod_b200.modeling.test_time_augmentation_avg import ViewSpec, resize_shortest_edge

    for sizes in ((480, 576, 672, 768, 864), (480, 576, 672, 768, 864, 960, 1056, 1152)):
        specs = []
        for s in sizes:
            nh, nw = resize_shortest_edge(h, w, s, 4000)
            specs += [ViewSpec(h, w, nh, nw, False), ViewSpec(h, w, nh, nw, True)]
        V = len(specs)
        pb = torch.rand((V, R, 4 * C), generator=g) * 600
        pb = torch.cat([pb.view(V, R, C, 4)[..., :2], pb.view(V, R, C, 4)[..., :2] + 1 + pb.view(V, R, C, 4)[..., 2:]], -1)
        pb = pb.reshape(V, R, 4 * C).contiguous()
        pr = torch.softmax(torch.randn((V, R, C + 1), generator=g) * 2, -1)
        mb, mp = ops.tta_merge(pb.cuda(), pr.cuda(), [s.params() for s in specs])
        acc_b = torch.empty((R, 4 * C), device="cuda")
        acc_p = torch.empty((R, C + 1), device="cuda")
        for v, s in enumerate(specs):
            ops.tta_accumulate(pb[v].cuda(), pr[v].cuda(), s.w * 1.0 / s.new_w, s.h * 1.0 / s.new_h, s.flip, float(s.new_w),
                               v == 0, float(V) if v == V - 1 else 0.0, acc_b, acc_p)
        assert torch.equal(mb, acc_b) and torch.equal(mp, acc_p)
        eb = [ref.tta_inverse_boxes(pb[v].reshape(-1, 4), s.w * 1.0 / s.new_w, s.h * 1.0 / s.new_h, s.flip, s.new_w)
              .reshape(R, 4 * C) for v, s in enumerate(specs)]
        emb, emp = ref.tta_merge(eb, [pr[v] for v in range(V)])
        torch.testing.assert_close(mb.cpu(), emb, rtol=1e-6, atol=1e-4)
        torch.testing.assert_close(mp.cpu(), emp, rtol=1e-6, atol=1e-9)


class _FakeBackbone(torch.nn.Module):
    """conv5 stand-in: stride-8 average pooling + a fixed 1x1 projection to `ch` channels (post-ReLU)."""

    def __init__(self, ch):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.register_buffer("proj", torch.randn((ch, 3, 1, 1), generator=g))

    def forward(self, x):
        x = torch.nn.functional.avg_pool2d(x, 8, ceil_mode=True)
        return {"plain5": torch.relu(torch.nn.functional.conv2d(x, self.proj) / 64.0)}


class _FakeRCNN(torch.nn.Module):
    """The slice of MultiInputRCNN the TTA wrapper touches (U/detectron2/modeling/meta_arch/rcnn_multi.py:210-254)."""

    def __init__(self, heads, ch):
        super().__init__()
        self.backbone = _FakeBackbone(ch)
        self.roi_heads = heads
        self.calls = 0

    @property
    def device(self):
        return self.backbone.proj.device

    def preprocess_image_inference(self, batched_inputs):
        from sos_wsod_b200.structures import ImageList

        imgs = [x["image"].to(self.device).float() for x in batched_inputs]
        return ImageList(torch.stack(imgs, 0), [tuple(i.shape[-2:]) for i in imgs])

    def inference(self, batched_inputs, detected_instances=None, do_postprocess=True):
        assert not self.training and detected_instances is None and not do_postprocess
        self.calls += 1
        images = self.preprocess_image_inference(batched_inputs)
        features = self.backbone(images.tensor)
        proposals = [x["proposals"].to(self.device) for x in batched_inputs]
        results, _, all_scores, all_boxes = self.roi_heads(images, features, proposals, None)
        return results, all_scores, all_boxes


def test_tta_wrapper_drop_in(ops):
    """GeneralizedRCNNWithTTAAVG(cfg, model)(batched_inputs): the fused pass (all views through one head launch
    sequence) and the reference-shaped view-by-view pass give the same merged result, equal to the oracle's merge +
    inference of the per-view head outputs; views that clip a proposal to nothing are handled like the reference."""
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling import build_roi_heads
    from sos_wsod_b200.modeling.test_time_augmentation_avg import GeneralizedRCNNWithTTAAVG
    from sos_wsod_b200.structures import Boxes, Instances, ShapeSpec

    torch.manual_seed(2)
    cfg = get_cfg()
    cfg.MODEL.ROI_BOX_HEAD.DAN_DIM = [256, 256]
    cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, cfg.TEST.AUG.FLIP = (96, 120, 160), 400, True
    cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST = 150
    ch, C = 16, 20
    heads = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=ch, stride=8)}).cuda().eval()
    for k in range(heads.refine_K):
        heads.box_refinery[k].cls_score.weight.data.mul_(40.0)
        heads.box_refinery[k].bbox_pred.weight.data.mul_(40.0)
    model = _FakeRCNN(heads, ch).cuda().eval()
    g = torch.Generator().manual_seed(9)
    H, W = 120, 160
    image = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)
    boxes = ref.synth_boxes(180, H, W, g)
    obj = torch.sort(torch.rand(180, generator=g), descending=True).values
    dd = {"image": image, "height": H, "width": W, "image_id": 5,
          "proposals": Instances((H, W), proposal_boxes=Boxes(boxes), objectness_logits=obj)}

    fused = GeneralizedRCNNWithTTAAVG(cfg, model)
    out_f = fused([dd])[0]["instances"]
    assert model.calls == 0
    plain = GeneralizedRCNNWithTTAAVG(cfg, model, fuse_views=False)
    out_p = plain([dd])[0]["instances"]
    assert model.calls == 6
    # per-view head outputs through the reference-shaped path, merged by the oracle
    aug, tfms = plain._get_augmented_inputs(dd)
    _, all_scores, all_boxes = plain._batch_inference(aug)
    assert all(s.shape == (1, 150, C + 1) for s in all_scores) and all(b.shape == (1, 150, 4 * C) for b in all_boxes)
    eb = [ref.tta_inverse_boxes(b.cpu().reshape(-1, 4), t.w * 1.0 / t.new_w, t.h * 1.0 / t.new_h, t.flip, t.new_w)
          .reshape(150, 4 * C) for b, t in zip(all_boxes, tfms)]
    emb, emp = ref.tta_merge(eb, [s.cpu()[0] for s in all_scores])
    mb, mp, _ = plain._get_augmented_boxes(aug, tfms)
    torch.testing.assert_close(mb.cpu(), emb, rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(mp.cpu(), emp, rtol=1e-6, atol=1e-9)
    e = ref.fast_rcnn_inference_single_image(mb.cpu(), mp.cpu(), (H, W), 1e-6, 0.3, 100)
    assert torch.equal(out_p.pred_boxes.tensor.cpu(), e[0]) and torch.equal(out_p.scores.cpu(), e[1])
    assert torch.equal(out_p.pred_classes.cpu(), e[2])
    # fused == view by view up to the GEMM-derived tolerance (the stand-in backbone convolves a view and its flip as
    # one batch of 2 in the fused pass, which moves conv5 by an ulp and can flip a bf16 rounding downstream)
    fb, fp, _ = fused._get_augmented_boxes(*fused._get_augmented_inputs(dd))
    torch.testing.assert_close(fb, mb, rtol=1e-2, atol=5e-2)
    torch.testing.assert_close(fp, mp, rtol=1e-2, atol=1e-5)
    assert len(out_f) == len(out_p) or abs(len(out_f) - len(out_p)) <= 2

    # a degenerate proposal among the first top-k rows: the reference filters FIRST and then takes the first 150 survivors
    # (test_time_augmentation_avg.py:62-71), i.e. it back-fills with row 150 of the 180 loaded proposals
    boxes2 = boxes.clone()
    boxes2[7] = torch.tensor([40.0, 30.0, 40.0, 90.0])
    dd2 = dict(dd, proposals=Instances((H, W), proposal_boxes=Boxes(boxes2), objectness_logits=obj))
    out2 = fused([dd2])[0]["instances"]
    keep_rows = torch.tensor([i for i in range(151) if i != 7])
    dd3 = dict(dd, proposals=Instances((H, W), proposal_boxes=Boxes(boxes2[keep_rows]), objectness_logits=obj[keep_rows]))
    out3 = fused([dd3])[0]["instances"]
    assert torch.equal(out2.scores, out3.scores) and torch.equal(out2.pred_boxes.tensor, out3.pred_boxes.tensor)
