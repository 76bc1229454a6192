"""GPU path against tests/golden/oicr_plus_golden.pt -- outputs of the REFERENCE'S OWN CODE (see
tests/golden/make_golden.py).  Integer results bit-exact; fp32 kernels within 1e-5; no GEMM is involved here (the
fixture's fp32 logits are fed to the fused kernels), so nothing falls under the bf16 tolerance."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oicr_plus_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from sos_wsod_b200 import ops as _ops

    return _ops


def test_roi_pool_vs_reference_pooler(ops, gold):
    inp = gold["inputs"]
    rois = torch.cat([torch.zeros(len(inp["boxes"]), 1), inp["boxes"]], 1)
    out, _, _ = ops.roi_pool_forward(inp["feat"].cuda(), rois.cuda())
    assert torch.equal(out.cpu(), gold["pool"]["pooled"])


def test_wsddn_vs_reference_layers(ops, gold):
    w, x, C = gold["wsddn"], gold["head"]["x"], gold["inputs"]["C"]
    logits = torch.cat([x @ w["cls_w"].t() + w["cls_b"], x @ w["det_w"].t() + w["det_b"]], 1).contiguous()
    scores, img, loss = ops.wsddn_forward(logits.cuda(), 0, C, 1, logits.shape[0], C, w["gt_oh"].cuda())
    torch.testing.assert_close(scores[0].cpu(), w["scores"], rtol=1e-5, atol=1e-9)
    assert abs(loss[0].item() - w["loss_cls"].item()) < 1e-5


def test_mining_and_labels_vs_reference_methods(ops, gold):
    inp = gold["inputs"]
    C, K, R = inp["C"], inp["K"], len(inp["boxes"])
    prev = torch.zeros((K, R, C + 1))
    for k, b in enumerate(gold["branches"]):
        prev[k, :, : b["prev"].shape[1]] = b["prev"]
    out = ops.oicr_mine_label(prev.cuda(), inp["boxes"].cuda(), gold["wsddn"]["gt_int"].cuda(), C, max(int(R * 0.10), 1))
    for k, b in enumerate(gold["branches"]):
        M = int(out["seed_count"][k].item())
        assert M == b["seed_index"].numel()
        assert torch.equal(out["seed_index"][k, :M].cpu().long(), b["seed_index"])
        assert torch.equal(out["seed_class"][k, :M].cpu().long(), b["seed_classes"])
        assert torch.equal(out["seed_score"][k, :M].cpu(), b["seed_scores"])
        assert torch.equal(out["gt_class"][k].cpu().long(), b["gt_classes"])
        assert torch.equal(out["gt_index"][k].cpu().long(), b["gt_index"])
        assert torch.equal(out["gt_weight"][k].cpu(), b["gt_weights"])


def test_oicr_losses_vs_reference_outputs(ops, gold):
    inp = gold["inputs"]
    C, K, R = inp["C"], inp["K"], len(inp["boxes"])
    stride = 5 * C + 1
    L = torch.cat([torch.cat([o["logits"], o["deltas"]], 1) for o in gold["oicr"]], 1).contiguous()
    y = torch.stack([b["gt_classes"] for b in gold["branches"]]).int()
    w = torch.stack([b["gt_weights"] for b in gold["branches"]])
    gi = torch.stack([b["gt_index"] for b in gold["branches"]]).int()
    losses, _, _ = ops.oicr_loss(L.cuda(), 0, stride, inp["boxes"].view(1, R, 4).cuda(), y.cuda(), w.cuda(), gi.cuda(), 1, R,
                                 C, K, flip_quirk=False)
    for k, o in enumerate(gold["oicr"]):
        assert abs(losses[k, 0].item() - o["losses"][f"loss_cls_r{k}"].item()) < 1e-5
        assert abs(losses[k, 1].item() - o["losses"][f"loss_box_reg_r{k}"].item()) < 1e-5


def test_inference_vs_reference_fast_rcnn_inference(ops, gold):
    inp, inf = gold["inputs"], gold["infer"]
    C, K = inp["C"], inp["K"]
    L = torch.cat([torch.cat([o["logits"], o["deltas"]], 1) for o in gold["oicr"]], 1).contiguous()
    probs, pb = ops.predict(L.cuda(), 0, 5 * C + 1, inp["boxes"].cuda(), C, K)
    torch.testing.assert_close(probs.cpu(), inf["all_scores"].reshape(-1, C + 1), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(pb.cpu(), inf["all_boxes"].reshape(-1, 4 * C), rtol=1e-5, atol=1e-3)
    db, ds, dc, dr, nd = ops.detect(inf["all_scores"].reshape(-1, C + 1).cuda(), inf["all_boxes"].reshape(-1, 4 * C).cuda(),
                                    inp["image_size"], 1e-6, 0.3, 100)
    n = int(nd.item())
    assert n == inf["scores"].numel()
    assert torch.equal(dr[:n].cpu().long(), inf["pred_inds"]), "keep-list must equal the reference's"
    assert torch.equal(dc[:n].cpu().long(), inf["pred_classes"])
    assert torch.equal(ds[:n].cpu(), inf["scores"]) and torch.equal(db[:n].cpu(), inf["pred_boxes"])
