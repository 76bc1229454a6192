"""DatasetMapperTTAAVG / GeneralizedRCNNWithTTAAVG -- the reference's test-time augmentation wrapper
(uwsod/projects/WSL/wsl/modeling/test_time_augmentation_avg.py:127-387) as a drop-in, with the per-view numpy
round trips moved onto the device:

  reference (per image, V = len(MIN_SIZES) * (2 if FLIP else 1) views)      this module
  ----------------------------------------------------------------------    --------------------------------------
  transform_proposals: boxes -> CPU numpy -> apply_box -> clip -> nonempty   ops.tta_views: ONE launch for all views
     once per view (:57-71)
  model.inference once per view (:281-286)                                   all views through ONE head pass
                                                                             (engine.test_forward, M = V*R rows)
  tfm.inverse().apply_box on CPU numpy + H2D, once per view (:353-365)       ops.tta_merge: ONE launch
  torch.mean over views, fast_rcnn_inference_single_image (:367-387)         same launch / ops.detect

Same class names, constructor arguments, `__call__` input / output format and method names, so
`GeneralizedRCNNWithTTAAVG(cfg, model)` replaces the reference's wrapper in tools/train_net_multi.py's test path.
The image pixels themselves (PIL resize + flip, host side) feed the backbone, which is outside the hot path.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from itertools import count
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from ..engine import ViewBatch
from ..structures import Boxes, Instances
from .fast_rcnn_oicr import _detections_to_instances

__all__ = ["ViewSpec", "resize_shortest_edge", "DatasetMapperTTAAVG", "GeneralizedRCNNWithTTAAVG"]


def resize_shortest_edge(h: int, w: int, size: int, max_size: int) -> Tuple[int, int]:
    """ResizeShortestEdge.get_transform (uwsod/detectron2/data/transforms/augmentation_impl.py:155-175):
    (new_h, new_w) in the reference's own Python-float arithmetic."""
    if size == 0:
        return h, w
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


@dataclass(frozen=True)
class ViewSpec:
    """One augmented view = TransformList([ResizeTransform(h, w, new_h, new_w), HFlipTransform(new_w) if flip]).
    Plays the role of the `transforms` entry of the reference's augmented dataset dicts."""
    h: int
    w: int
    new_h: int
    new_w: int
    flip: bool
    # pre_tfm of the reference (:169-173): the dataset's (height, width) when the stored image is not at that size;
    # the inverse then ends with ResizeTransform(stored -> original), a second fp32 multiply
    orig_hw: Optional[Tuple[int, int]] = None

    @property
    def image_size(self) -> Tuple[int, int]:
        return (self.new_h, self.new_w)

    def params(self, batch_index: float = 0.0) -> List[float]:
        """The 10 floats of include/soswsod_b200.h (SOSWSOD_TTA_VIEW_PARAMS).  The scale factors are the reference's
        Python doubles (transform.py:123-126: `new_w * 1.0 / w`), rounded to fp32 at the C boundary exactly as numpy
        rounds them when it multiplies the fp32 box array."""
        post_x = post_y = 1.0
        if self.orig_hw is not None:
            post_x, post_y = self.orig_hw[1] * 1.0 / self.w, self.orig_hw[0] * 1.0 / self.h
        return [self.new_w * 1.0 / self.w, self.new_h * 1.0 / self.h, 1.0 if self.flip else 0.0, float(self.new_w),
                float(self.new_h), float(batch_index), self.w * 1.0 / self.new_w, self.h * 1.0 / self.new_h, post_x, post_y]

    def apply_image(self, img: np.ndarray) -> np.ndarray:
        """ResizeTransform.apply_image (transform.py:101-122, PIL bilinear on uint8 HWC) then HFlipTransform."""
        from PIL import Image

        assert img.shape[:2] == (self.h, self.w), (img.shape, self.h, self.w)
        if img.dtype == np.uint8:
            out = np.asarray(Image.fromarray(img).resize((self.new_w, self.new_h), Image.BILINEAR))
        else:
            t = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1)[None].float()
            t = torch.nn.functional.interpolate(t, (self.new_h, self.new_w), mode="bilinear", align_corners=False)
            out = t[0].permute(1, 2, 0).numpy().astype(img.dtype)
        if self.flip:
            out = np.flip(out, axis=1)
        return out


class DatasetMapperTTAAVG:
    """test_time_augmentation_avg.py:127-197.  `__call__(dataset_dict)` returns the list of augmented dataset dicts in
    the reference's order (for every MIN_SIZE: resized, then resized + flipped); each carries "image" (CHW tensor),
    "transforms" (a ViewSpec) and, when proposals are loaded, "proposals" transformed for that view -- all views by one
    device launch instead of one numpy pass per view."""

    def __init__(self, cfg):
        self.min_sizes = cfg.TEST.AUG.MIN_SIZES
        self.max_size = cfg.TEST.AUG.MAX_SIZE
        self.flip = cfg.TEST.AUG.FLIP
        self.image_format = cfg.INPUT.FORMAT
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.proposal_topk = None
        if cfg.MODEL.LOAD_PROPOSALS:
            self.proposal_topk = cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST

    def view_specs(self, h: int, w: int) -> List[ViewSpec]:
        out = []
        for min_size in self.min_sizes:
            nh, nw = resize_shortest_edge(h, w, min_size, self.max_size)
            out.append(ViewSpec(h, w, nh, nw, False))
            if self.flip:
                out.append(ViewSpec(h, w, nh, nw, True))
        return out

    def transform_proposals(self, proposals: Instances, specs: Sequence[ViewSpec], min_box_size: float = 0.0):
        """transform_proposals (:29-71) for every view at once.  Returns (list of per-view Instances on the device,
        dropped int32 [V] on the device).

        The reference filters the empty boxes of a view FIRST and then keeps `boxes[keep][:topk]` (:62-71).  When no
        row among the first `topk` becomes empty in any view -- the normal case: integer proposals scaled up stay
        non-empty -- that is exactly the first `topk` rows, which is what this fast path returns without a host
        synchronisation; `dropped[v]` counts the empty rows among them, and a non-zero count sends the caller to
        `transform_proposals_exact` (the reference's back-filled selection)."""
        boxes = proposals.proposal_boxes.tensor
        logits = proposals.objectness_logits
        if self.proposal_topk is not None:
            boxes, logits = boxes[: self.proposal_topk], logits[: self.proposal_topk]
        boxes = boxes.to(self.device, non_blocking=True)
        logits = logits.to(self.device, non_blocking=True)
        rois, _keep, dropped = ops.tta_views(boxes, [s.params(0.0) for s in specs], min_box_size)
        R = boxes.size(0)
        rois = rois.view(len(specs), R, 5)
        out = []
        for v, s in enumerate(specs):
            inst = Instances(s.image_size)
            inst.proposal_boxes = Boxes(rois[v, :, 1:5])
            inst.objectness_logits = logits
            out.append(inst)
        return out, dropped

    def transform_proposals_exact(self, proposals: Instances, specs: Sequence[ViewSpec], min_box_size: float = 0.0):
        """The reference's order of operations (:57-71) when some row does become empty: per view, transform + clip ALL
        rows, drop the empty ones, THEN take the first `topk` survivors -- so a view back-fills from the rows behind
        `topk` and the views may hold different proposals at the same row index (the reference averages them row by row
        all the same, :367-372).  One host synchronisation (the selections have data-dependent sizes)."""
        boxes = proposals.proposal_boxes.tensor.to(self.device)
        logits = proposals.objectness_logits.to(self.device)
        rois, keep, _ = ops.tta_views(boxes, [s.params(0.0) for s in specs], min_box_size)
        R = boxes.size(0)
        rois = rois.view(len(specs), R, 5)
        out = []
        for v, s in enumerate(specs):
            sel = keep[v].nonzero().flatten()
            if self.proposal_topk is not None:
                sel = sel[: self.proposal_topk]
            inst = Instances(s.image_size)
            inst.proposal_boxes = Boxes(rois[v, sel, 1:5].contiguous())
            inst.objectness_logits = logits[sel]
            out.append(inst)
        return out

    def __call__(self, dataset_dict):
        numpy_image = dataset_dict["image"].permute(1, 2, 0).numpy()
        shape = numpy_image.shape
        orig_shape = (dataset_dict["height"], dataset_dict["width"])
        # the reference composes a pre-transform when the stored image is not at the dataset's size (:169-173); the
        # proposals are transformed from the stored image's frame either way (only `tfms` is applied to them, :193)
        specs = self.view_specs(shape[0], shape[1])
        props = dropped = None
        if self.proposal_topk is not None and "proposals" in dataset_dict:
            props, dropped = self.transform_proposals(dataset_dict["proposals"], specs)
        ret = []
        for v, s in enumerate(specs):
            new_image = s.apply_image(np.copy(numpy_image))
            dic = {k: val for k, val in dataset_dict.items() if k not in ("image", "proposals")}
            dic = copy.deepcopy(dic)
            dic["image"] = torch.from_numpy(np.ascontiguousarray(new_image.transpose(2, 0, 1)))
            # pre_tfm + tfms (:169-173, 189): predicted boxes go back to the dataset's (height, width)
            dic["transforms"] = s if shape[:2] == orig_shape else ViewSpec(s.h, s.w, s.new_h, s.new_w, s.flip, orig_shape)
            if props is not None:
                dic["proposals"] = props[v]
                dic["proposals_dropped"] = dropped
            ret.append(dic)
        return ret


class GeneralizedRCNNWithTTAAVG(nn.Module):
    """test_time_augmentation_avg.py:200-387.  `model` is the reference's MultiInputRCNN / GeneralizedRCNNWSL (anything
    with `.inference(batched_inputs, detected_instances=None, do_postprocess=False)` returning
    (results, all_scores, all_boxes)) whose `roi_heads` is this package's OICRPlusHeads.

    fuse_views=True (default) runs the backbone per view (a view and its flip batched, as rcnn_multi.py:174-175 does in
    training) and then ALL views through one head pass; False calls `model.inference` view by view like the
    reference.  Either way the inverse transforms, the mean over views and the final thresholding / NMS run on the
    device."""

    def __init__(self, cfg, model, tta_mapper=None, batch_size: int = 1, fuse_views: bool = True):
        super().__init__()
        if isinstance(model, nn.parallel.DistributedDataParallel):
            model = model.module
        self.cfg = cfg.clone()
        self.model = model
        if tta_mapper is None:
            tta_mapper = DatasetMapperTTAAVG(cfg)
        self.tta_mapper = tta_mapper
        self.batch_size = batch_size
        self.fuse_views = fuse_views

    # ---- reference-shaped helpers ----
    def _batch_inference(self, batched_inputs, detected_instances=None):
        """:254-286, for models without the fused path."""
        if detected_instances is None:
            detected_instances = [None] * len(batched_inputs)
        outputs, all_scores, all_boxes = [], [], []
        inputs, instances = [], []
        for idx, inp, instance in zip(count(), batched_inputs, detected_instances):
            inputs.append(inp)
            instances.append(instance)
            if len(inputs) == self.batch_size or idx == len(batched_inputs) - 1:
                output, all_score, all_box = self.model.inference(
                    inputs, instances if instances[0] is not None else None, do_postprocess=False)
                outputs.extend(output)
                all_scores.extend(all_score)
                all_boxes.extend(all_box)
                inputs, instances = [], []
        return outputs, all_scores, all_boxes

    def __call__(self, batched_inputs):
        return [self._inference_one_image(self._complete(x)) for x in batched_inputs]

    @staticmethod
    def _complete(dataset_dict):
        ret = copy.copy(dataset_dict)
        if "image" not in ret:
            raise RuntimeError("GeneralizedRCNNWithTTAAVG: the dataset dict must carry the decoded image (CHW uint8); "
                               "file reading belongs to the data loader")
        if "height" not in ret and "width" not in ret:
            ret["height"] = ret["image"].shape[1]
            ret["width"] = ret["image"].shape[2]
        return ret

    def _get_augmented_inputs(self, input):
        augmented_inputs = self.tta_mapper(input)
        tfms = [x.pop("transforms") for x in augmented_inputs]
        return augmented_inputs, tfms

    def _fused_head_outputs(self, augmented_inputs):
        """All views through one head pass -> (probs [V,R,C+1], pred_boxes [V,R,4C]) in view coordinates."""
        model = self.model
        heads = model.roi_heads
        feats, groups = [], []
        v = 0
        V = len(augmented_inputs)
        while v < V:
            # a view and its flip have the same size: one backbone call, one feature tensor of batch 2
            n = 2 if (v + 1 < V and augmented_inputs[v + 1]["image"].shape == augmented_inputs[v]["image"].shape) else 1
            chunk = augmented_inputs[v:v + n]
            images = model.preprocess_image_inference(chunk)
            f = model.backbone(images.tensor)
            if isinstance(f, dict):
                f = f[heads.box_in_features[-1]]
            feats.append(f.float().contiguous())
            groups.append([x["proposals"] for x in chunk])
            v += n
        rois, obj = heads._view_rois(groups)
        vb = ViewBatch(feats, rois, obj, len(groups[0][0]))
        return heads.engine().test_forward(vb)

    def _get_augmented_boxes(self, augmented_inputs, tfms):
        """:349-373 -> (all_boxes [R,4C] in original-image coordinates, all_scores [R,C+1], None)."""
        heads = getattr(self.model, "roi_heads", None)
        fused = (self.fuse_views and heads is not None and hasattr(heads, "engine") and hasattr(self.model, "backbone")
                 and hasattr(self.model, "preprocess_image_inference") and "proposals" in augmented_inputs[0])
        if fused:
            probs, pboxes = self._fused_head_outputs(augmented_inputs)
        else:
            _, all_scores, all_boxes = self._batch_inference(augmented_inputs)
            probs = torch.cat([s.reshape(1, *s.shape[-2:]) for s in all_scores], 0)
            pboxes = torch.cat([b.reshape(1, *b.shape[-2:]) for b in all_boxes], 0)
        mean_boxes, mean_probs = ops.tta_merge(pboxes, probs, [t.params() for t in tfms])
        return mean_boxes, mean_probs, None

    def _merge_detections(self, all_boxes, all_scores, all_classes, shape_hw):
        """:375-387: fast_rcnn_inference_single_image on the merged boxes / scores at the original size."""
        rh = self.cfg.MODEL.ROI_HEADS
        det = ops.detect(all_scores, all_boxes, shape_hw, rh.SCORE_THRESH_TEST, rh.NMS_THRESH_TEST,
                         self.cfg.TEST.DETECTIONS_PER_IMAGE)
        inst, _ = _detections_to_instances(det, shape_hw)
        return inst[0]

    def _inference_one_image(self, input):
        orig_shape = (input["height"], input["width"])
        augmented_inputs, tfms = self._get_augmented_inputs(input)
        dropped = augmented_inputs[0].get("proposals_dropped")
        all_boxes, all_scores, _ = self._get_augmented_boxes(augmented_inputs, tfms)
        merged_instances = self._merge_detections(all_boxes, all_scores, None, orig_shape)
        # _detections_to_instances has synchronised with the device; reading the counter now costs nothing extra
        if dropped is not None and int(dropped.sum().item()) != 0:
            return self._inference_with_dropped_proposals(input, tfms)
        return {"instances": merged_instances}

    def _inference_with_dropped_proposals(self, input, tfms):
        """Some proposal among the rows in use became empty in some view: redo the image with the reference's exact
        selection (filter per view, then top-k, :57-71).  Views that end up with different row counts cannot be
        averaged by the reference either (torch.cat / mean of unequal shapes, :367)."""
        mapper = self.tta_mapper
        props = mapper.transform_proposals_exact(input["proposals"], tfms)
        counts = {len(p) for p in props}
        if len(counts) != 1:
            raise RuntimeError("TTA views keep different numbers of proposals after clip/nonempty: the reference cannot "
                               f"average them either (test_time_augmentation_avg.py:367); rows per view = {sorted(counts)}")
        augmented_inputs, tfms2 = self._get_augmented_inputs(input)
        for x, p in zip(augmented_inputs, props):
            x["proposals"] = p
            x.pop("proposals_dropped", None)
        all_boxes, all_scores, _ = self._get_augmented_boxes(augmented_inputs, tfms2)
        return {"instances": self._merge_detections(all_boxes, all_scores, None, (input["height"], input["width"]))}
