"""Rows either side of the hot path's output: detection-results writers, image
sharding, PGF.  The oracle (oracle/eval_ref.py) and the product's host code are pinned to tests/golden/eval_golden.json,
which the reference's own evaluator / tools/pgf.py produced (tests/golden/make_golden_eval.py); the device PGF kernel is
checked against both (-m gpu)."""
import copy
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from oracle import eval_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "eval_golden.json")) as f:
        return json.load(f)


def _per_image(gold):
    return [{"image_id": d["image_id"], "boxes": torch.tensor(d["boxes"], dtype=torch.float32).reshape(-1, 4),
             "scores": torch.tensor(d["scores"], dtype=torch.float32), "classes": torch.tensor(d["classes"], dtype=torch.int64),
             "size": tuple(d["size"])} for d in gold["detections"]]


def _result_dict(gold):
    result = {}
    for m in json.loads(gold["voc_json"]):
        m = dict(m)
        m["category_id"] -= 1
        result.setdefault(m["image_id"], []).append(m)
    return result


# ---------------------------------------------------------------- oracle vs the reference's outputs
def test_oracle_writers_reproduce_reference_bytes(gold):
    per = [{"image_id": d["image_id"], "boxes": d["boxes"].numpy(), "scores": d["scores"].tolist(), "classes": d["classes"].tolist()}
           for d in _per_image(gold)]
    assert eval_ref.voc_json_text(per, gold["num_classes"]) == gold["voc_json"]
    assert json.dumps(eval_ref.coco_predictions(per)) == gold["coco_json"]


def test_oracle_pgf_reproduces_reference(gold):
    for c in gold["contain"]:
        assert eval_ref.contain_cal(c["a"], c["b"]) == c["val"]
    class_dict = {int(k): v for k, v in gold["pgf_class_dict"].items()}
    for case in gold["pgf_cases"]:
        r = _result_dict(gold)
        eval_ref.class_filter(r, class_dict)
        assert {str(k): v for k, v in r.items()} == case["after_class_filter"]
        eval_ref.pgf(r, case["t_con"], case["t_keep"], case["use_diff"], case["diff_classes"])
        assert {str(k): v for k, v in r.items()} == case["after_pgf"]


def test_inference_shard_is_the_reference_sampler():
    for size in (1, 7, 8, 9, 4952, 5011):
        for world in (1, 2, 3, 8):
            blocks = [list(eval_ref.inference_shard(size, r, world)) for r in range(world)]
            assert sum(blocks, []) == list(range(size))                      # exact cover, in order
            assert max(len(b) for b in blocks) == (size - 1) // world + 1     # contiguous blocks of ceil(size/world)
            from sos_wsod_b200.evaluation import inference_shard
            assert [list(inference_shard(size, r, world)) for r in range(world)] == blocks


# ---------------------------------------------------------------- product host code vs the reference's outputs
def _instances(d):
    from sos_wsod_b200.structures import Boxes, Instances
    return Instances(d["size"], pred_boxes=Boxes(d["boxes"]), scores=d["scores"], pred_classes=d["classes"])


def test_product_writers_are_byte_compatible(gold, tmp_path):
    from sos_wsod_b200.evaluation import COCODetectionWriter, PascalVOCDetectionWriter

    voc = PascalVOCDetectionWriter("voc_2007_test", [f"c{k}" for k in range(gold["num_classes"])], str(tmp_path / "oicr_plus_{}.json"))
    coco = COCODetectionWriter("coco_2014_train", str(tmp_path / "oicr_plus_{}.json"))
    for d in _per_image(gold):
        voc.process([{"image_id": d["image_id"]}], [{"instances": _instances(d)}])
        coco.process([{"image_id": d["image_id"]}], [{"instances": _instances(d)}])
    assert open(voc.save()).read() == gold["voc_json"]
    assert open(coco.save()).read() == gold["coco_json"]
    # the block path (one device->host copy for many images, fixed-size detect outputs) writes the same bytes
    import numpy as np

    per = list(_per_image(gold))
    topk = max(len(d["scores"]) for d in per) + 3
    boxes = np.zeros((len(per), topk, 4), np.float32)
    scores = np.zeros((len(per), topk), np.float32)
    classes = np.zeros((len(per), topk), np.int32)
    counts = np.zeros((len(per),), np.int32)
    for i, d in enumerate(per):
        inst = _instances(d)
        n = len(inst.scores)
        counts[i] = n
        boxes[i, :n] = inst.pred_boxes.tensor.numpy()
        scores[i, :n] = inst.scores.numpy()
        classes[i, :n] = inst.pred_classes.numpy()
    voc2 = PascalVOCDetectionWriter("voc_2007_test", [f"c{k}" for k in range(gold["num_classes"])], str(tmp_path / "blk_{}.json"))
    voc2.process_arrays([d["image_id"] for d in per], boxes, scores, classes, counts)
    assert open(voc2.save()).read() == gold["voc_json"]
    # the directly assembled text is what json.dumps renders (incl. float reprs like 1e-05, inf-free by construction)
    import json as _json

    voc3 = PascalVOCDetectionWriter("x", ["a", "b"], str(tmp_path / "t_{}.json"))
    voc3._predictions[0] += ["7 0.000 1.0 2.5 1e-05 123456789.1", "8 1.000 0.1 0.2 0.3 0.4"]
    voc3._predictions[1] += ["9 0.500 10.0 20.0 30.0 40.0"]
    assert voc3.json_text() == _json.dumps(voc3.rows())
    assert PascalVOCDetectionWriter("x", ["a"], "y").json_text() == "[]"
    # COCO: per-rank encoded pieces joined == json.dump of the chained entries (an empty rank contributes nothing)
    import itertools

    per_rank = [COCODetectionWriter("c", "p") for _ in range(3)]
    for i, d in enumerate(per):
        per_rank[0 if i < 5 else 2].process([{"image_id": d["image_id"]}], [{"instances": _instances(d)}])
    joined = "[" + ", ".join(t for t in (w.json_piece() for w in per_rank) if t) + "]"
    assert joined == _json.dumps(list(itertools.chain(*[w._predictions for w in per_rank]))) == gold["coco_json"]


_SHARD_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from sos_wsod_b200.evaluation import PascalVOCDetectionWriter, generate_detection_results
from sos_wsod_b200.structures import Boxes, Instances
port, rank, out = sys.argv[2], int(sys.argv[3]), sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
dist.init_process_group("gloo", rank=rank, world_size=2)
gold = json.load(open(os.path.join(sys.argv[1], "tests", "golden", "eval_golden.json")))
dets = gold["detections"]
def detect(i):
    d = dets[i]
    inst = Instances(tuple(d["size"]), pred_boxes=Boxes(torch.tensor(d["boxes"], dtype=torch.float32).reshape(-1, 4)),
                     scores=torch.tensor(d["scores"], dtype=torch.float32), pred_classes=torch.tensor(d["classes"], dtype=torch.int64))
    return {"image_id": d["image_id"], "instances": inst}
w = PascalVOCDetectionWriter("voc_2007_test", [f"c{k}" for k in range(gold["num_classes"])], out)
mine = generate_detection_results(detect, len(dets), w, rank, 2)
assert list(mine) == list(range(rank * 6, min(12, rank * 6 + 6)))
path = w.save()
assert (path is not None) == (rank == 0)
dist.destroy_process_group()
'''


def test_sharded_generation_gloo_world2_equals_single_process(gold, tmp_path):
    """Two ranks, contiguous image blocks, no collective but the final gather of the rows: rank 0's file equals the
    single-process file byte for byte (images are contiguous per rank, so the class-major order is preserved)."""
    script = tmp_path / "shard.py"
    script.write_text(_SHARD_SCRIPT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "det_{}.json")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), out], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert open(out.format("voc_2007_test")).read() == gold["voc_json"]


# ---------------------------------------------------------------- device PGF
@pytest.mark.gpu
def test_pgf_kernel_reproduces_reference(cuda_lib, gold):
    from sos_wsod_b200.evaluation import class_filter, pgf, pgf_voc_results

    class_dict = {int(k): v for k, v in gold["pgf_class_dict"].items()}
    for case in gold["pgf_cases"]:
        r = _result_dict(gold)
        class_filter(r, class_dict)
        assert {str(k): v for k, v in r.items()} == case["after_class_filter"]
        pgf(r, case["t_con"], case["t_keep"], case["use_diff"], case["diff_classes"])
        assert {str(k): v for k, v in r.items()} == case["after_pgf"]
    case = gold["pgf_cases"][0]
    r = pgf_voc_results(json.loads(gold["voc_json"]), class_dict, case["t_con"], case["t_keep"], case["use_diff"], case["diff_classes"])
    assert {str(k): v for k, v in r.items()} == case["after_pgf"]


@pytest.mark.gpu
def test_pgf_kernel_vs_oracle_large(cuda_lib):
    """5000-image scale of BASELINE configs[4] is covered by a 600-image random set here (the oracle is an O(n^2) Python
    loop per image) plus edge cases: empty images, one detection, zero-area boxes, identical boxes, scores exactly at
    t_keep, categories >= 64."""
    from sos_wsod_b200.evaluation import pgf

    g = torch.Generator().manual_seed(5)
    result = {}
    for i in range(600):
        n = int(torch.randint(0, 101, (1,), generator=g)) if i % 50 else (0 if i % 100 else 1)
        preds = []
        for k in range(n):
            x, y = float(torch.rand(1, generator=g)) * 500, float(torch.rand(1, generator=g)) * 400
            w, h = float(torch.rand(1, generator=g)) * 200, float(torch.rand(1, generator=g)) * 200
            if k % 17 == 0:
                w = 0.0
            if k % 5 == 1 and preds:
                px, py, pw_, ph_ = preds[-1]["bbox"]
                x, y, w, h = px + 0.05 * pw_, py + 0.05 * ph_, 0.85 * pw_, 0.9 * ph_
            if k % 23 == 2 and preds:
                x, y, w, h = preds[-1]["bbox"]
            score = round(float(torch.rand(1, generator=g)), 3)
            if k % 11 == 3:
                score = 0.2
            cat = int(torch.randint(0, 80, (1,), generator=g)) if k % 5 != 1 or not preds else preds[-1]["category_id"]
            preds.append({"image_id": i, "category_id": cat, "score": score, "bbox": [x, y, w, h]})
        result[i] = preds
    for t_con, t_keep, use_diff, diff in [(0.85, 0.2, False, [4, 5, 70]), (0.85, 0.2, True, []), (0.3, 0.5, False, [])]:
        exp = copy.deepcopy(result)
        eval_ref.pgf(exp, t_con, t_keep, use_diff, diff)
        got = copy.deepcopy(result)
        pgf(got, t_con, t_keep, use_diff, diff)
        assert got == exp
