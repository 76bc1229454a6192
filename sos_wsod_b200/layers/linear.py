"""nn.Linear (+ReLU, +dropout) on the tcgen05 GEMM, as an autograd.Function: the stand-alone form of the
fc6/fc7/head layers (wsl/modeling/roi_heads/box_head.py:82-91).  The fused head step (engine.py) does not go
through this; it is the module-by-module drop-in path."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import ops


class _LinearAct(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, dropout_p, seed):
        xb = x.detach()
        xb = xb if xb.dtype == torch.bfloat16 else ops.cast_f32_bf16(xb.float().contiguous())[0]
        wb = ops.cast_f32_bf16(weight.detach().contiguous())[0]
        y = ops.gemm_bf16(xb, wb, out_dtype=torch.bfloat16 if relu else torch.float32,
                          bias=None if bias is None else bias.detach().float(), relu=relu, dropout_p=dropout_p,
                          dropout_seed=seed)
        ctx.relu, ctx.p, ctx.has_bias, ctx.x_dtype = relu, dropout_p, bias is not None, x.dtype
        ctx.save_for_backward(xb, wb, y if relu else None)
        return y.float() if relu else y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xb, wb, y = ctx.saved_tensors
        g = gy.contiguous().float()
        m, n_out = g.shape
        gbuf = torch.zeros((m, (n_out + 7) // 8 * 8), dtype=torch.bfloat16, device=g.device)  # lda multiple of 8
        gb = gbuf[:, :n_out]
        # ReLU (+ dropout) backward fused into the fp32 -> bf16 cast of the incoming gradient: the saved activation is
        # positive exactly where the unit was kept and active
        ops.cast_f32_bf16(g, out=gb, mask_src=y if ctx.relu else None,
                          mask_scale=(1.0 / (1.0 - ctx.p) if ctx.p > 0 else 1.0))
        gx = ops.gemm_bf16(gb, wb, b_mn=True).to(ctx.x_dtype)
        gw = ops.gemm_bf16(gb, xb, a_mn=True, b_mn=True)
        gbias = ops.colsum(gb) if ctx.has_bias else None
        return gx, gw, gbias, None, None, None


def linear_act(x, weight, bias=None, relu=False, dropout_p=0.0, seed=0):
    return _LinearAct.apply(x, weight, bias, relu, float(dropout_p), int(seed))
