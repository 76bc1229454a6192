"""Generates tests/golden/tta_golden.pt by running the REFERENCE'S OWN test-time-augmentation code (imported
read-only from /root/reference through tests/golden/ref_import.py) on seeded synthetic inputs.  Build container only:

    python tests/golden/make_golden_tta.py

Produced by the reference itself (uwsod/projects/WSL/wsl/modeling/test_time_augmentation_avg.py):
  DatasetMapperTTAAVG.__call__                         (:142-197)  per-view image shape / checksum, transformed proposals
  GeneralizedRCNNWithTTAAVG._get_augmented_boxes       (:349-373)  merged boxes / scores (inverse transforms + mean)
  GeneralizedRCNNWithTTAAVG._merge_detections          (:375-387)  final detections at the original size
with ResizeShortestEdge / RandomFlip / ResizeTransform of the reference's forked detectron2
(uwsod/detectron2/data/transforms/) and the restated fvcore Transform / TransformList / HFlipTransform base classes
(ref_import._fvcore_transforms; fvcore is an un-vendored, unpinned dependency).  The per-view head outputs fed to
_get_augmented_boxes are seeded stand-ins for model.inference (the head itself is pinned by make_golden.py)."""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402

ref_import.install()

from detectron2.structures import Boxes, Instances  # noqa: E402
from wsl.modeling.test_time_augmentation_avg import DatasetMapperTTAAVG, GeneralizedRCNNWithTTAAVG  # noqa: E402

from oracle import oicr_plus_ref as ora  # noqa: E402  (seeded synthetic proposals only)


class _Node(dict):
    __getattr__ = dict.__getitem__

    def clone(self):
        return self


def make_cfg(min_sizes, max_size, flip, topk, score_thr=1e-6, nms_thr=0.3, dets=100):
    N = _Node
    return N(TEST=N(AUG=N(MIN_SIZES=min_sizes, MAX_SIZE=max_size, FLIP=flip), DETECTIONS_PER_IMAGE=dets),
             INPUT=N(FORMAT="BGR"), MODEL=N(LOAD_PROPOSALS=True, KEYPOINT_ON=False, MASK_ON=False,
                                            ROI_HEADS=N(SCORE_THRESH_TEST=score_thr, NMS_THRESH_TEST=nms_thr)),
             DATASETS=N(PRECOMPUTED_PROPOSAL_TOPK_TEST=topk, PRECOMPUTED_PROPOSAL_TOPK_TRAIN=topk))


def run_case(name, g, stored_hw, dataset_hw, min_sizes, max_size, flip, R, C, topk):
    cfg = make_cfg(min_sizes, max_size, flip, topk)
    H, W = stored_hw
    image = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)
    boxes = ora.synth_boxes(R, H, W, g)
    boxes[3] = torch.tensor([0.0, 0.0, float(W), float(H)])          # the whole image
    boxes[4] = torch.tensor([W - 21.0, H - 22.0, float(W), float(H)])  # touches the far corner
    obj = torch.sort(torch.rand(R, generator=g), descending=True).values
    props = Instances(stored_hw, proposal_boxes=Boxes(boxes.clone()), objectness_logits=obj.clone())
    dd = {"image": image, "height": dataset_hw[0], "width": dataset_hw[1], "image_id": 17, "proposals": props}

    mapper = DatasetMapperTTAAVG(cfg)
    aug = mapper(dd)
    tfms = [x.pop("transforms") for x in aug]
    views = []
    for x, t in zip(aug, tfms):
        img = x["image"]
        views.append({"image_shape": tuple(img.shape), "image_sha1": hashlib.sha1(img.numpy().tobytes()).hexdigest(),
                      "proposal_boxes": x["proposals"].proposal_boxes.tensor.clone(),
                      "objectness_logits": x["proposals"].objectness_logits.clone(),
                      "image_size": tuple(x["proposals"].image_size),
                      "transforms": [type(tt).__name__ for tt in t.transforms]})
    Rk = len(views[0]["proposal_boxes"])

    # stand-in head outputs per view: boxes around the view's proposals, softmax scores shared up to noise
    base = torch.randn((Rk, C + 1), generator=g) * 2.5
    all_scores, all_boxes = [], []
    for v in views:
        pb = v["proposal_boxes"]
        jitter = torch.randn((Rk, C, 4), generator=g) * 6.0
        b = (pb[:, None, :] + jitter)
        b = torch.stack([torch.minimum(b[..., 0], b[..., 2]), torch.minimum(b[..., 1], b[..., 3]),
                         torch.maximum(b[..., 0], b[..., 2]) + 1.0, torch.maximum(b[..., 1], b[..., 3]) + 1.0], -1)
        all_boxes.append(b.reshape(1, Rk, 4 * C).contiguous())
        all_scores.append(torch.softmax(base + 0.3 * torch.randn((Rk, C + 1), generator=g), -1).reshape(1, Rk, C + 1))
    fake = types.SimpleNamespace(cfg=cfg)
    fake._batch_inference = lambda inputs: (None, [s.clone() for s in all_scores], [b.clone() for b in all_boxes])
    mb, ms, _ = GeneralizedRCNNWithTTAAVG._get_augmented_boxes(fake, aug, tfms)
    # fast_rcnn_inference_single_image clips its `boxes` argument IN PLACE (Boxes(boxes.reshape(-1, 4)).clip aliases
    # it, U/detectron2/modeling/roi_heads/fast_rcnn.py:101-103): keep the un-clipped means for the fixture
    mb_unclipped, ms_in = mb.clone(), ms.clone()
    merged = GeneralizedRCNNWithTTAAVG._merge_detections(fake, mb, ms, None, dataset_hw)
    mb, ms = mb_unclipped, ms_in
    return {"name": name, "stored_hw": stored_hw, "dataset_hw": dataset_hw, "min_sizes": tuple(min_sizes),
            "max_size": max_size, "flip": flip, "topk": topk, "C": C, "image": image, "boxes": boxes, "obj": obj,
            "views": views, "view_scores": torch.cat(all_scores, 0), "view_boxes": torch.cat(all_boxes, 0),
            "merged_boxes": mb, "merged_scores": ms,
            "det_boxes": merged.pred_boxes.tensor, "det_scores": merged.scores, "det_classes": merged.pred_classes}


def main():
    g = torch.Generator().manual_seed(20261017)
    cases = [
        run_case("voc_3scales_flip", g, (60, 80), (60, 80), (48, 66, 90), 4000, True, 200, 20, 4000),
        run_case("portrait_maxsize_noflip", g, (90, 56), (90, 56), (40, 64, 80, 100, 120), 150, False, 257, 20, 200),
        run_case("stored_not_dataset_size", g, (48, 64), (96, 128), (48, 72), 4000, True, 180, 8, 4000),
        run_case("coco_width_16views", g, (52, 70), (52, 70), (48, 57, 67, 76, 86, 96, 105, 115), 4000, True, 40, 80, 4000),
    ]
    path = os.path.join(os.environ.get("SOSWSOD_GOLDEN_OUT", HERE), "tta_golden.pt")
    torch.save({"cases": cases, "numpy": np.__version__, "torch": str(torch.__version__)}, path)
    for c in cases:
        print(c["name"], "views", len(c["views"]), "R", len(c["views"][0]["proposal_boxes"]), "dets", len(c["det_scores"]),
              [v["image_shape"] for v in c["views"]][:4])
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
