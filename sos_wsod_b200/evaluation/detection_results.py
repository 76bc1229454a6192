"""Detection-results writers and the sharded generation loop (BASELINE.json configs[4]).

PascalVOCDetectionWriter  -- the `process` / json-dump half of PascalVOCDetectionEvaluator
    (uwsod/detectron2/evaluation/pascal_voc_evaluation.py:57-71, 78-118): each detection goes through the same string
    round trip (`score:.3f`, coordinates `:.1f`, +1 on xmin/ymin in float32), rows are class-major, category ids
    1-based, and the file is byte-identical to what the reference's evaluator writes to WSODEVAL.SAVE_PATH.
COCODetectionWriter       -- COCOEvaluator.process + dump (coco_evaluation.py:106-140) with instances_to_coco_json
    (:316-375): one {"image_id", "instances": [...]} entry per image, XYWH boxes converted in float32.
inference_shard           -- InferenceSampler (uwsod/detectron2/data/samplers/distributed_sampler.py:191-194):
    contiguous blocks of ceil(size / world) images per rank.
generate_detection_results -- the test-time loop: every rank runs its block of images through the head engine
    (TTA views -> mean scores / boxes -> threshold -> per-class NMS -> top-k, all on the device), the writers collect
    the rows; there is NO collective on the compute path, only the final gather of the rows to rank 0
    (comm.gather in the reference, torch.distributed.gather_object here)."""
from __future__ import annotations

import json
from collections import defaultdict
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def inference_shard(size: int, rank: int, world: int) -> range:
    assert size > 0 and 0 <= rank < world
    shard = (size - 1) // world + 1
    return range(shard * rank, min(shard * (rank + 1), size))


_HOST_GROUP = None


def host_gather_group():
    """The process group host objects are gathered through: the default group if it is a CPU backend, else a gloo group
    over the same ranks, created once (a collective call: every rank must reach it together) -- the reference does the
    same (detectron2 comm._get_global_gloo_group, utils/comm.py:89-99; comm.gather :196-232).  Pickled rows pushed through
    NCCL go host -> device -> wire -> device -> host and pay NCCL's lazy connection set-up for the gather pattern (1.3 s
    of a 6.7 s job at 8 GPUs); call this once before a timed region to keep the group's creation out of it."""
    global _HOST_GROUP
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    world_pg = dist.group.WORLD
    if _HOST_GROUP is None or _HOST_GROUP[0] is not world_pg:      # first use, or the default group was re-created
        backend = str(dist.get_backend()).lower()
        group = dist.new_group(backend="gloo") if "nccl" in backend else world_pg
        probe = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
        dist.gather_object(dist.get_rank(), probe, dst=0, group=group)      # connects the pairs
        _HOST_GROUP = (world_pg, group)
    return _HOST_GROUP[1]


def _gather(obj, dst: int = 0):
    """comm.gather: list of every rank's object on `dst`, [] elsewhere; [obj] without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst, group=host_gather_group())
    return out if dist.get_rank() == dst else []


def _is_main(dst: int = 0) -> bool:
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == dst


def _instances_fields(instances):
    boxes = instances.pred_boxes.tensor if hasattr(instances.pred_boxes, "tensor") else instances.pred_boxes
    return (boxes.detach().to("cpu", torch.float32).numpy(), instances.scores.detach().to("cpu").tolist(),
            instances.pred_classes.detach().to("cpu").tolist())


class PascalVOCDetectionWriter:
    def __init__(self, dataset_name: str, class_names: Sequence[str], save_path: str):
        self._dataset_name = dataset_name
        self._class_names = list(class_names)
        self.save_path = save_path
        self.reset()

    def reset(self) -> None:
        self._predictions: Dict[int, List[str]] = defaultdict(list)   # class id -> prediction lines

    def process(self, inputs: Sequence[dict], outputs: Sequence[dict]) -> None:
        for inp, out in zip(inputs, outputs):
            image_id = inp["image_id"]
            boxes, scores, classes = _instances_fields(out["instances"])
            for box, score, cls in zip(boxes, scores, classes):
                xmin, ymin, xmax, ymax = box
                xmin += 1     # float32, the inverse of the VOC loader's -1 (pascal_voc_evaluation.py:66-68)
                ymin += 1
                self._predictions[cls].append(f"{image_id} {score:.3f} {xmin:.1f} {ymin:.1f} {xmax:.1f} {ymax:.1f}")

    def process_arrays(self, image_ids: Sequence[int], boxes: np.ndarray, scores: np.ndarray, classes: np.ndarray,
                       counts: np.ndarray) -> None:
        """`process` for a block of images whose detections were brought to the host in ONE copy: boxes float32
        [n, topk, 4], scores float32 [n, topk], classes int [n, topk], counts int [n] (soswsod_detect's fixed-size
        outputs).  Same float32 `+1` and the same string formatting as `process`."""
        boxes = np.asarray(boxes, dtype=np.float32)
        for i, image_id in enumerate(image_ids):
            n = int(counts[i])
            sc = [float(x) for x in scores[i, :n]]
            for box, score, cls in zip(boxes[i, :n], sc, classes[i, :n].tolist()):
                xmin, ymin, xmax, ymax = box
                xmin += 1
                ymin += 1
                self._predictions[cls].append(f"{image_id} {score:.3f} {xmin:.1f} {ymin:.1f} {xmax:.1f} {ymax:.1f}")

    def rows(self, all_predictions: Optional[List[dict]] = None) -> List[dict]:
        merged: Dict[int, List[str]] = {}
        for p in (all_predictions if all_predictions is not None else [self._predictions]):
            for key in list(p.keys()):
                merged[key] = merged.get(key, []) + p[key]
        rows = []
        for cls_id, _ in enumerate(self._class_names):
            for line in merged.get(cls_id, []):
                m = line.split(" ")
                rows.append({"image_id": int(m[0]), "category_id": cls_id + 1, "score": float(m[1]),
                             "bbox": [float(m[2]), float(m[3]), float(m[4]), float(m[5])]})
        return rows

    @staticmethod
    def _json_objects(lines: List[str], category_id: int) -> str:
        """The json objects of one class's prediction lines, ", "-joined, exactly as json.dump renders the row dicts of
        `rows()`: an int with str(), a float with float.__repr__, keys in insertion order, separators ", " and ": "."""
        fr = float.__repr__
        out = []
        for line in lines:
            m = line.split(" ")
            out.append('{"image_id": %d, "category_id": %d, "score": %s, "bbox": [%s, %s, %s, %s]}' % (
                int(m[0]), category_id, fr(float(m[1])), fr(float(m[2])), fr(float(m[3])), fr(float(m[4])), fr(float(m[5]))))
        return ", ".join(out)

    def json_text(self, all_predictions: Optional[List[dict]] = None) -> str:
        """The bytes `json.dump(self.rows(...), f)` writes, assembled directly from the prediction lines (json.dump of
        half a million small dicts costs seconds); tests compare with json.dumps(rows)."""
        preds = all_predictions if all_predictions is not None else [self._predictions]
        parts = []
        for cls_id, _ in enumerate(self._class_names):
            for p in preds:
                if p.get(cls_id):
                    parts.append(self._json_objects(p[cls_id], cls_id + 1))
        return "[" + ", ".join(parts) + "]"

    def save(self, dst: int = 0) -> Optional[str]:
        """Every rank renders its own rows (the string round trip is the expensive part and shards like the images),
        the per-class text pieces are gathered on `dst` in rank order (comm.gather's order) and concatenated class by
        class -- the file is byte-identical to the reference evaluator's `json.dump` of the gathered rows."""
        mine = {c: self._json_objects(lines, c + 1) for c, lines in self._predictions.items() if lines}
        gathered = _gather(mine, dst)
        if not _is_main(dst):
            return None
        parts = []
        for cls_id, _ in enumerate(self._class_names):
            for p in gathered:
                if p.get(cls_id):
                    parts.append(p[cls_id])
        path = self.save_path.format(self._dataset_name)
        with open(path, "w") as f:
            f.write("[" + ", ".join(parts) + "]")
        return path


class COCODetectionWriter:
    def __init__(self, dataset_name: str, save_path: str):
        self._dataset_name = dataset_name
        self.save_path = save_path
        self.reset()

    def reset(self) -> None:
        self._predictions: List[dict] = []

    @staticmethod
    def instances_to_coco_json(instances, img_id) -> List[dict]:
        boxes, scores, classes = _instances_fields(instances)
        if len(scores) == 0:
            return []
        boxes = np.array(boxes, dtype=np.float32, copy=True)
        boxes[:, 2] -= boxes[:, 0]      # XYXY_ABS -> XYWH_ABS in float32 (structures/boxes.py:113-115)
        boxes[:, 3] -= boxes[:, 1]
        bl = boxes.tolist()
        return [{"image_id": img_id, "category_id": classes[k], "bbox": bl[k], "score": scores[k]} for k in range(len(scores))]

    def process(self, inputs: Sequence[dict], outputs: Sequence[dict]) -> None:
        for inp, out in zip(inputs, outputs):
            self._predictions.append({"image_id": inp["image_id"],
                                      "instances": self.instances_to_coco_json(out["instances"], inp["image_id"])})

    def json_piece(self) -> str:
        """This rank's entries as json.dump renders them, without the enclosing brackets (the C encoder, on every rank)."""
        return json.dumps(self._predictions)[1:-1]

    def save(self, dst: int = 0) -> Optional[str]:
        """Every rank encodes its own entries; the text pieces are gathered on `dst` in rank order (comm.gather's order)
        and joined -- byte-identical to `json.dump(list(itertools.chain(*gathered)), f)` of the reference
        (coco_evaluation.py:126-140) without pickling half a million dicts to one rank and encoding them there."""
        gathered = _gather(self.json_piece(), dst)
        if not _is_main(dst):
            return None
        path = self.save_path.format(self._dataset_name)
        with open(path, "w") as f:
            f.write("[" + ", ".join(t for t in gathered if t) + "]")
        return path


def generate_detection_results(detect_image: Callable[[int], dict], num_images: int, writer, rank: int = 0, world: int = 1,
                               progress: Optional[Callable[[int, int], None]] = None) -> Iterable[int]:
    """Runs `detect_image(index)` for the images of this rank's InferenceSampler block and feeds the writer.
    detect_image returns {"image_id": ..., "instances": Instances-like with pred_boxes / scores / pred_classes}
    (e.g. OICRPlusHeadEngine.tta_detect wrapped by the caller).  Returns the indices processed."""
    mine = inference_shard(num_images, rank, world)
    for n, idx in enumerate(mine):
        out = detect_image(idx)
        writer.process([{"image_id": out["image_id"]}], [{"instances": out["instances"]}])
        if progress is not None:
            progress(n + 1, len(mine))
    return mine
