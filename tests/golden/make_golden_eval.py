"""Generates tests/golden/eval_golden.json by running the REFERENCE'S OWN CODE (imported read-only from
/root/reference through tests/golden/ref_import.py) on seeded synthetic detections.  Build container only:

    python tests/golden/make_golden_eval.py

Produced by the reference itself:
  voc_json   PascalVOCDetectionEvaluator.process + the detection-result dump of .evaluate
             (uwsod/detectron2/evaluation/pascal_voc_evaluation.py:57-118) -- the exact bytes of the json file
  coco_rows  instances_to_coco_json                     (uwsod/detectron2/evaluation/coco_evaluation.py:316-375)
  pgf        class_filter + pgf (t_keep / t_con / use_diff / diff_classes variants)   (tools/pgf.py:221-290)
"""
import copy
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402

ref_import.install()
sys.path.insert(0, "/root/reference/tools")

import pgf as ref_pgf  # noqa: E402  (tools/pgf.py)
from detectron2.evaluation.coco_evaluation import instances_to_coco_json  # noqa: E402
from detectron2.evaluation.pascal_voc_evaluation import PascalVOCDetectionEvaluator  # noqa: E402
from detectron2.structures import Boxes, Instances  # noqa: E402


def synth_detections(g, n_img, num_classes, max_det, img_hw=(480, 640)):
    """Per image: up to max_det detections, score-descending (as fast_rcnn_inference returns them), with nested and
    duplicated boxes so that the containment rule of PGF fires."""
    out = []
    H, W = img_hw
    for i in range(n_img):
        n = int(torch.randint(0, max_det + 1, (1,), generator=g))
        x1 = torch.rand(n, generator=g) * (W - 40)
        y1 = torch.rand(n, generator=g) * (H - 40)
        w = torch.rand(n, generator=g) * (W - x1 - 8) + 8
        h = torch.rand(n, generator=g) * (H - y1 - 8) + 8
        boxes = torch.stack([x1, y1, x1 + w, y1 + h], 1)
        for k in range(1, n, 3):      # nest every third box inside its predecessor
            b = boxes[k - 1]
            shrink = torch.rand(4, generator=g) * 0.12
            boxes[k] = torch.stack([b[0] + shrink[0] * (b[2] - b[0]), b[1] + shrink[1] * (b[3] - b[1]),
                                    b[2] - shrink[2] * (b[2] - b[0]), b[3] - shrink[3] * (b[3] - b[1])])
        scores = torch.sort(torch.rand(n, generator=g) ** 2, descending=True).values
        classes = torch.randint(0, num_classes, (n,), generator=g)
        for k in range(1, n, 3):
            classes[k] = classes[k - 1]
        out.append({"image_id": 1000 + 7 * i, "boxes": boxes, "scores": scores, "classes": classes, "size": img_hw})
    return out


def main():
    g = torch.Generator().manual_seed(20261018)
    C = 20
    dets = synth_detections(g, 12, C, 30)
    gold = {"num_classes": C, "detections": [
        {"image_id": d["image_id"], "boxes": d["boxes"].tolist(), "scores": d["scores"].tolist(),
         "classes": d["classes"].tolist(), "size": list(d["size"])} for d in dets]}

    # ---- VOC detection-result json: the reference evaluator, constructed without its dataset metadata ----
    ev = object.__new__(PascalVOCDetectionEvaluator)
    ev._dataset_name = "voc_2007_test"
    ev._class_names = [f"c{k}" for k in range(C)]
    ev._cpu_device = torch.device("cpu")
    ev.save_detection_result = True
    tmp = tempfile.mkdtemp()
    ev.save_path = os.path.join(tmp, "det_{}.json")
    ev.reset()
    for d in dets:
        inst = Instances(d["size"])
        inst.pred_boxes = Boxes(d["boxes"].clone())
        inst.scores = d["scores"].clone()
        inst.pred_classes = d["classes"].clone()
        ev.process([{"image_id": d["image_id"]}], [{"instances": inst}])
    try:
        ev.evaluate()          # dumps the json, then goes on to the AP computation that needs the real dataset
    except Exception:
        pass
    with open(ev.save_path.format(ev._dataset_name)) as f:
        gold["voc_json"] = f.read()

    # ---- COCO rows ----
    coco = []
    for d in dets:
        inst = Instances(d["size"])
        inst.pred_boxes = Boxes(d["boxes"].clone())
        inst.scores = d["scores"].clone()
        inst.pred_classes = d["classes"].clone()
        coco.append({"image_id": d["image_id"], "instances": instances_to_coco_json(inst, d["image_id"])})
    gold["coco_json"] = json.dumps(coco)

    # ---- PGF: the VOC flow of tools/pgf.py:43-117 on the json above, with synthetic image-level GT classes ----
    voc_rows = json.loads(gold["voc_json"])
    result = {}
    for m in voc_rows:
        m = dict(m)
        m["category_id"] = m["category_id"] - 1
        result.setdefault(m["image_id"], []).append(m)
    class_dict = {}
    for d in dets:
        present = sorted(set(d["classes"].tolist()))
        keep = [c for k, c in enumerate(present) if k % 4 != 3]     # drop a quarter of the classes
        class_dict[d["image_id"]] = keep
    gold["pgf_class_dict"] = {str(k): v for k, v in class_dict.items()}
    cases = []
    for (t_con, t_keep, use_diff, diff) in [(0.85, 0.2, False, [4, 5, 6, 8, 9, 15, 16]), (0.85, 0.2, True, [4, 5, 6, 8, 9, 15, 16]),
                                            (0.5, 0.05, False, [1, 2]), (0.95, 0.6, True, None)]:
        r = copy.deepcopy(result)
        ref_pgf.class_filter(r, class_dict, "golden")
        mid = copy.deepcopy(r)
        ref_pgf.pgf(r, "golden", t_con, t_keep, use_diff, diff if diff is not None else [])
        cases.append({"t_con": t_con, "t_keep": t_keep, "use_diff": use_diff, "diff_classes": diff if diff is not None else [],
                      "after_class_filter": {str(k): v for k, v in mid.items()},
                      "after_pgf": {str(k): v for k, v in r.items()}})
    gold["pgf_cases"] = cases
    # contain_cal known answers (tools/pgf.py:210-219)
    gold["contain"] = [{"a": a, "b": b, "val": ref_pgf.contain_cal(a, b)} for a, b in
                       [([10.0, 10.0, 20.0, 20.0], [5.0, 5.0, 40.0, 40.0]), ([0.0, 0.0, 10.0, 10.0], [5.0, 5.0, 10.0, 10.0]),
                        ([3.5, 2.25, 0.0, 7.0], [0.0, 0.0, 50.0, 50.0]), ([100.0, 100.0, 5.0, 5.0], [0.0, 0.0, 50.0, 50.0])]]
    path = os.path.join(os.environ.get("SOSWSOD_GOLDEN_OUT", HERE), "eval_golden.json")
    with open(path, "w") as f:
        json.dump(gold, f)
    n_after = [sum(len(v) for v in c["after_pgf"].values()) for c in cases]
    print("wrote", path, os.path.getsize(path), "bytes;", len(voc_rows), "detections ->", n_after, "after pgf")


if __name__ == "__main__":
    main()
