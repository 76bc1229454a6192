"""ROIPooler with the reference's surface (uwsod/projects/WSL/wsl/modeling/poolers.py:119-306), single-level
"ROIPool" only -- the configuration OICR+ ships (POOLER_TYPE: ROIPool, one dilated-C5 level, §8a row B)."""
from typing import List

import torch
from torch import nn

from ..layers import RoIPool


def convert_boxes_to_pooler_format(box_lists) -> torch.Tensor:
    """[M,5] = (batch index as float, x0, y0, x1, y1); poolers.py:74-108."""
    out = []
    for i, b in enumerate(box_lists):
        t = b.tensor
        out.append(torch.cat((torch.full((len(t), 1), i, dtype=t.dtype, device=t.device), t), dim=1))
    return torch.cat(out, dim=0)


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio=0, pooler_type="ROIPool", canonical_box_size=224,
                 canonical_level=4, use_range=False):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        self.output_size = output_size
        if pooler_type != "ROIPool":
            raise ValueError(f"pooler_type {pooler_type!r}: only 'ROIPool' (the OICR+ configuration) is built for sm_100a")
        if len(scales) != 1:
            raise ValueError("the OICR+ head pools from a single feature level (plain5)")
        self.level_poolers = nn.ModuleList(RoIPool(output_size, spatial_scale=s) for s in scales)
        self.pooler_type = pooler_type

    def forward(self, x: List[torch.Tensor], box_lists):
        """x: one NCHW feature map per level; box_lists: one Boxes per image -> [M, C, oh, ow]."""
        assert isinstance(x, list) and isinstance(box_lists, list) and len(x) == 1
        assert len(box_lists) == x[0].size(0), "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(
            x[0].size(0), len(box_lists))
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        return self.level_poolers[0](x[0], convert_boxes_to_pooler_format(box_lists))
