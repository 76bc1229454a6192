"""DiscriminativeAdaptionNeck: flatten -> fc1+ReLU+dropout -> fc2+ReLU+dropout
(uwsod/projects/WSL/wsl/modeling/roi_heads/box_head.py:15-103).  Parameter names fc1/fc2 and the N(0,0.005) /
0.1 init are the reference's, so released checkpoints load unchanged."""
import numpy as np
import torch
from torch import nn

from ..layers import linear_act
from ..registry import ROI_BOX_HEAD_REGISTRY
from ..structures import ShapeSpec


@ROI_BOX_HEAD_REGISTRY.register()
class DiscriminativeAdaptionNeck(nn.Module):
    def __init__(self, input_shape: ShapeSpec, *, conv_dims=(), fc_dims=(4096, 4096), conv_norm="", dropout=0.5):
        super().__init__()
        assert len(conv_dims) == 0, "OICR+ uses no conv layers in the box head (NUM_CONV: 0)"
        assert len(fc_dims) > 0
        self._output_size = (input_shape.channels, input_shape.height, input_shape.width)
        self.fcs = []
        for k, fc_dim in enumerate(fc_dims):
            fc = nn.Linear(int(np.prod(self._output_size)), fc_dim)
            self.add_module("fc{}".format(k + 1), fc)
            self.fcs.append(fc)
            self._output_size = fc_dim
        for layer in self.fcs:
            nn.init.normal_(layer.weight, std=0.005)
            nn.init.constant_(layer.bias, 0.1)
        self.dropout = dropout
        self._calls = 0

    @classmethod
    def from_config(cls, cfg, input_shape):
        return cls(input_shape, fc_dims=tuple(cfg.MODEL.ROI_BOX_HEAD.DAN_DIM),
                   dropout=cfg.MODEL.ROI_BOX_HEAD.get("DROPOUT", 0.5))

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        for i, layer in enumerate(self.fcs):
            self._calls += 1
            seed = torch.initial_seed() * 1000003 + self._calls
            x = linear_act(x, layer.weight, layer.bias, relu=True, dropout_p=self.dropout if self.training else 0.0,
                           seed=seed)
        return x

    @property
    def output_shape(self):
        o = self._output_size
        return ShapeSpec(channels=o) if isinstance(o, int) else ShapeSpec(channels=o[0], height=o[1], width=o[2])


def build_box_head(cfg, input_shape):
    """uwsod/detectron2/modeling/roi_heads/box_head.py:112-117."""
    return ROI_BOX_HEAD_REGISTRY.get(cfg.MODEL.ROI_BOX_HEAD.NAME).from_config(cfg, input_shape)
