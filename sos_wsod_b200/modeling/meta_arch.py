"""MultiInputRCNN -- the caller on the other side of the head (uwsod/detectron2/modeling/meta_arch/rcnn_multi.py:23-291,
SURVEY.md §8f rank 3) as a thin mirror: same constructor arguments, `forward(batched_inputs)` contract (one image per
GPU, four views: image1 / image1_flip / image2 / image2_flip with their proposals), `inference`,
`preprocess_image(_inference)` and `_postprocess`.  The backbone is whatever module the caller passes (the reference's
VGG16 on cuDNN stays the backbone, BASELINE.json north_star); this class only does what the reference does around it:
normalise, batch an image with its flip (one backbone call per scale), hand both feature maps and the four proposal
sets to the ROI head, which runs all four views as one batched pass.  The per-view Python loop and the unused clones of
roi_heads_oicrplus.py:223-226 do not exist here, nor does the trainer's torch.cuda.empty_cache() per step."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from ..structures import ImageList, Instances

__all__ = ["MultiInputRCNN", "detector_postprocess"]


def detector_postprocess(results: Instances, output_height: int, output_width: int) -> Instances:
    """uwsod/detectron2/modeling/postprocessing.py:10-71 (boxes only): rescale to the output resolution, clip, drop
    empty boxes."""
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances((output_height, output_width), **results.get_fields())
    boxes = results.pred_boxes if results.has("pred_boxes") else results.proposal_boxes
    boxes.scale(scale_x, scale_y)
    boxes.clip(results.image_size)
    return results[boxes.nonempty()]


class MultiInputRCNN(nn.Module):
    def __init__(self, *, backbone: nn.Module, proposal_generator: Optional[nn.Module], roi_heads: nn.Module,
                 pixel_mean: Tuple[float, ...], pixel_std: Tuple[float, ...], input_format: Optional[str] = None,
                 vis_period: int = 0):
        super().__init__()
        self.backbone = backbone
        self.proposal_generator = proposal_generator
        self.roi_heads = roi_heads
        self.input_format = input_format
        self.vis_period = vis_period
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean, dtype=torch.float32).view(-1, 1, 1))
        self.register_buffer("pixel_std", torch.tensor(pixel_std, dtype=torch.float32).view(-1, 1, 1))
        assert self.pixel_mean.shape == self.pixel_std.shape

    @property
    def device(self):
        return self.pixel_mean.device

    @property
    def size_divisibility(self) -> int:
        return int(getattr(self.backbone, "size_divisibility", 0))

    # ---- rcnn_multi.py:255-274 ----
    def _normalised(self, batched_inputs, key: str) -> ImageList:
        images = [(x[key].to(self.device) - self.pixel_mean) / self.pixel_std for x in batched_inputs]
        return ImageList.from_tensors(images, self.size_divisibility)

    def preprocess_image(self, batched_inputs):
        return tuple(self._normalised(batched_inputs, k) for k in ("image1", "image2", "image1_flip", "image2_flip"))

    def preprocess_image_inference(self, batched_inputs) -> ImageList:
        return self._normalised(batched_inputs, "image")

    # ---- rcnn_multi.py:131-208 ----
    def forward(self, batched_inputs: List[Dict]):
        assert len(batched_inputs) == 1, "now, MultiInputRCNN only support the setting -> imgs_per_gpu=1"
        if not self.training:
            return self.inference(batched_inputs)
        images1, images2, images1_flip, images2_flip = self.preprocess_image(batched_inputs)
        # one backbone call per scale: the image and its flip have the same size
        features1 = self.backbone(torch.cat([images1.tensor, images1_flip.tensor], 0))
        features2 = self.backbone(torch.cat([images2.tensor, images2_flip.tensor], 0))
        suffixes = ("1", "1_flip", "2", "2_flip")
        for sfx in suffixes:
            assert "proposals" + sfx in batched_inputs[0], "precomputed proposals of all four views are required"
        proposals_list = [[x["proposals" + sfx].to(self.device) for x in batched_inputs] for sfx in suffixes]
        gt_list = [[x["instances" + sfx].to(self.device) for x in batched_inputs] if "instances" + sfx in batched_inputs[0]
                   else None for sfx in suffixes]
        images_list = [images1, images1_flip, images2, images2_flip]
        _, detector_losses = self.roi_heads(images_list, [features1, features2], proposals_list, gt_list)
        losses = {}
        losses.update(detector_losses)
        return losses

    # ---- rcnn_multi.py:210-254 ----
    def inference(self, batched_inputs, detected_instances=None, do_postprocess: bool = True):
        assert not self.training
        if detected_instances is not None:
            raise NotImplementedError("forward_with_given_boxes (mask / keypoint heads) is not part of the OICR+ path")
        images = self.preprocess_image_inference(batched_inputs)
        features = self.backbone(images.tensor)
        assert "proposals" in batched_inputs[0]
        proposals = [x["proposals"].to(self.device) for x in batched_inputs]
        targets = [x["instances"].to(self.device) for x in batched_inputs] if "instances" in batched_inputs[0] else None
        results, _, all_scores, all_boxes = self.roi_heads(images, features, proposals, targets)
        if do_postprocess:
            return MultiInputRCNN._postprocess(results, batched_inputs, images.image_sizes)
        return results, all_scores, all_boxes

    @staticmethod
    def _postprocess(instances, batched_inputs, image_sizes):
        out = []
        for results_per_image, input_per_image, image_size in zip(instances, batched_inputs, image_sizes):
            height = input_per_image.get("height", image_size[0])
            width = input_per_image.get("width", image_size[1])
            out.append({"instances": detector_postprocess(results_per_image, height, width)})
        return out
