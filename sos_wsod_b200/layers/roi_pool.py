"""ROI max-pool operator, wsl.layers style: an autograd.Function over the native forward/backward plus a thin
nn.Module (pattern of uwsod/projects/WSL/wsl/layers/roi_loop_pool.py:9-58).  API of torchvision.ops.RoIPool,
the op the reference's ROIPooler instantiates (wsl/modeling/poolers.py:183-186)."""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import ops


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, input, rois, output_size, spatial_scale):
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.input_shape = input.size()
        output, argmax, _ = ops.roi_pool_forward(input, rois, ctx.output_size, spatial_scale)
        ctx.save_for_backward(rois, argmax)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        grad_input = ops.roi_pool_backward(grad_output.contiguous().float(), argmax, rois, tuple(ctx.input_shape),
                                           ctx.output_size, spatial_scale=ctx.spatial_scale)
        return grad_input, None, None, None


roi_pool = _ROIPool.apply


class RoIPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois):
        """input: NCHW fp32; rois: Bx5 (batch index, x1, y1, x2, y2)."""
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_pool(input, rois, self.output_size, self.spatial_scale)

    def __repr__(self):
        return f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale})"
