"""Seeded synthetic inputs of the hot path: the proposal boxes, conv5 maps and objectness scores that
bench.py and the examples feed to the head.  Plain torch on the CPU, nothing of the test infrastructure and no CUDA
involved; the parity tests check that this generator and the test suite's own (the one the committed golden fixtures
were made with) produce identical tensors for the same seed."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch


@dataclass
class SynthView:
    feat: torch.Tensor                      # [1, channels, ceil(h/stride), ceil(w/stride)] fp32, post-ReLU (>= 0)
    boxes: torch.Tensor                     # [R, 4] XYXY fp32, integer pixel coordinates of the (resized) image
    obj: torch.Tensor                       # [R] objectness_logits in [0, 1), sorted descending
    image_size: Tuple[int, int] = (0, 0)    # (h, w)


def _first_occurrences(keys: torch.Tensor) -> torch.Tensor:
    """Indices of the first occurrence of every distinct key, ascending."""
    order = torch.argsort(keys, stable=True)
    sk = keys[order]
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    return torch.sort(order[first]).values


def synth_boxes(R: int, img_h: int, img_w: int, g: torch.Generator, min_size: int = 20) -> torch.Tensor:
    """R distinct integer-pixel boxes like MCG / selective-search proposals: x1 ~ U(0, W-32), y1 ~ U(0, H-32),
    w ~ U(min, W-x1), h ~ U(min, H-y1), rounded, sides >= min_size (MODEL.PROPOSAL_GENERATOR.MIN_SIZE), duplicates
    dropped by the coordinate hash of U/detectron2/structures/boxes.py:214-226 (first occurrence kept)."""
    out = torch.zeros((0, 4))
    hash_w = torch.tensor([1, 1e3, 1e6, 1e9], dtype=torch.double)
    while out.size(0) < R:
        n = 2 * R
        x1 = torch.rand(n, generator=g) * (img_w - 32)
        y1 = torch.rand(n, generator=g) * (img_h - 32)
        w = min_size + torch.rand(n, generator=g) * (img_w - x1 - min_size)
        h = min_size + torch.rand(n, generator=g) * (img_h - y1 - min_size)
        b = torch.stack([x1, y1, (x1 + w).clamp(max=img_w - 1), (y1 + h).clamp(max=img_h - 1)], 1).round()
        b = b[((b[:, 2] - b[:, 0]) >= min_size) & ((b[:, 3] - b[:, 1]) >= min_size)]
        out = torch.cat([out, b], 0)
        out = out[_first_occurrences((out.double() @ hash_w).long())]
    return out[:R].contiguous()


def hflip_boxes(boxes: torch.Tensor, img_w: int) -> torch.Tensor:
    b = boxes.clone()
    b[:, 0] = img_w - boxes[:, 2]
    b[:, 2] = img_w - boxes[:, 0]
    return b


def synth_views(R: int, sizes: Sequence[Tuple[int, int]], g: torch.Generator, channels: int = 512,
                stride: int = 8) -> List[SynthView]:
    """The four training views of one image (1, 1_flip, 2, 2_flip; rcnn_multi.py:152-199): the same R proposals in
    the same order in every view (U/detectron2/data/dataset_mapper.py:353-361), view 2 = view 1 rescaled, flips
    mirror x.  sizes = [(h1, w1), (h2, w2)]."""
    (h1, w1), (h2, w2) = sizes
    base = synth_boxes(R, h1, w1, g)
    obj = torch.sort(torch.rand(R, generator=g), descending=True).values
    b2 = base.clone()
    b2[:, 0::2] *= w2 / w1
    b2[:, 1::2] *= h2 / h1
    views = []
    for (h, w, b) in ((h1, w1, base), (h1, w1, hflip_boxes(base, w1)), (h2, w2, b2), (h2, w2, hflip_boxes(b2, w2))):
        fh, fw = (h + stride - 1) // stride, (w + stride - 1) // stride
        feat = torch.relu(torch.randn((1, channels, fh, fw), generator=g))
        views.append(SynthView(feat=feat, boxes=b.contiguous(), obj=obj.clone(), image_size=(h, w)))
    return views


def pack_views(views: Sequence[SynthView]):
    """Four views -> the head's batched inputs: feats = 2 x [2, ch, h, w] (image + flip per scale), rois = 2 x [2R, 5]
    (index inside the pair, x1, y1, x2, y2), obj = [4R]."""
    R = views[0].boxes.size(0)
    feats = [torch.cat([views[0].feat, views[1].feat], 0), torch.cat([views[2].feat, views[3].feat], 0)]
    rois = []
    for a, b in ((0, 1), (2, 3)):
        r0 = torch.cat([torch.zeros(R, 1), views[a].boxes], 1)
        r1 = torch.cat([torch.ones(R, 1), views[b].boxes], 1)
        rois.append(torch.cat([r0, r1], 0))
    obj = torch.cat([v.obj for v in views])
    return feats, rois, obj


def training_image(index: int, rank: int = 0, R: int = 2000, sizes: Sequence[Tuple[int, int]] = ((480, 640), (576, 768)),
                   num_classes: int = 20, channels: int = 512, cfg_id: int = 2):
    """The i-th synthetic training image of a rank for bench.py / the whole-step parity tests (seed
    1234 + 100 * cfg_id + ...): four views + the image-level GT classes (1..4 distinct classes, sorted)."""
    g = torch.Generator().manual_seed(1234 + 100 * cfg_id + rank * 17 + index)
    views = synth_views(R, list(sizes), g, channels=channels)
    ng = int(torch.randint(1, 5, (1,), generator=g))
    gt = torch.sort(torch.randperm(num_classes, generator=g)[:ng]).values
    return views, gt
