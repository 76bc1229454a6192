"""build_optimizer -- the parameter-group rules of the reference's solver (uwsod/detectron2/solver/build.py:143-218)
over the fused SGD kernel: per parameter lr / weight decay (bias lr x BIAS_LR_FACTOR with
WEIGHT_DECAY_BIAS, norm layers WEIGHT_DECAY_NORM, optional higher lr on the refinement branches), SGD with momentum
exactly as torch.optim.SGD(momentum=m, dampening=0, nesterov=False) computes it:
    buf = m * buf + (grad + wd * p);  p -= lr * buf
`B200SGD.step()` makes ONE pass over each parameter (soswsod_sgd_step reads p, grad, buf and writes p, buf) instead of
torch's 3-4 elementwise passes; the head's bf16 GEMM operands refresh themselves on the next forward (parameter
version counters).  Gradient clipping (SOLVER.CLIP_GRADIENTS) is off in every released OICR+ config and not built."""
from __future__ import annotations

from typing import Any, Dict, List, Set

import torch

from . import ops

_NORM_TYPES = (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d, torch.nn.SyncBatchNorm, torch.nn.GroupNorm,
               torch.nn.InstanceNorm1d, torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d, torch.nn.LayerNorm,
               torch.nn.LocalResponseNorm)


def get_optimizer_param_groups(cfg, model: torch.nn.Module) -> List[Dict[str, Any]]:
    """One group per parameter, in the reference's order, with its (lr, weight_decay)."""
    S = cfg.SOLVER
    params: List[Dict[str, Any]] = []
    memo: Set[torch.nn.Parameter] = set()
    if S.get("REFINE_SCALE_ON", False):
        # build.py:162-188: rules keyed on the parameter NAME ("refine" = the box_refinery_k branches)
        for key, value in dict(model.named_parameters()).items():
            if not value.requires_grad or value in memo:
                continue
            memo.add(value)
            lr, wd = S.BASE_LR, S.WEIGHT_DECAY
            refine, bias = "refine" in key, "bias" in key
            if "bn" in key.lower():
                wd = S.WEIGHT_DECAY_NORM
            elif refine and bias:
                lr, wd = S.BASE_LR * S.BIAS_LR_FACTOR * S.REFINE_LR_SCALE, S.WEIGHT_DECAY_BIAS
            elif refine:
                lr = S.BASE_LR * S.REFINE_LR_SCALE
            elif bias:
                lr, wd = S.BASE_LR * S.BIAS_LR_FACTOR, S.WEIGHT_DECAY_BIAS
            params.append({"params": [value], "lr": lr, "weight_decay": wd})
    else:
        # build.py:189-213: rules keyed on the owning module's type and the attribute name
        for module in model.modules():
            for key, value in module.named_parameters(recurse=False):
                if not value.requires_grad or value in memo:
                    continue
                memo.add(value)
                lr, wd = S.BASE_LR, S.WEIGHT_DECAY
                if isinstance(module, _NORM_TYPES):
                    wd = S.WEIGHT_DECAY_NORM
                elif key == "bias":
                    lr, wd = S.BASE_LR * S.BIAS_LR_FACTOR, S.WEIGHT_DECAY_BIAS
                params.append({"params": [value], "lr": lr, "weight_decay": wd})
    return params


class B200SGD(torch.optim.Optimizer):
    """torch.optim.SGD(momentum, dampening=0, nesterov=False) semantics; state key `momentum_buffer` like torch's."""

    def __init__(self, params, lr: float, momentum: float = 0.0, weight_decay: float = 0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("B200SGD steps CUDA parameters only (sm_100a kernel); there is no CPU fallback")
                st = self.state[p]
                if "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                w = p.detach()
                if not w.is_contiguous():
                    raise RuntimeError("B200SGD needs contiguous parameters")
                ops.sgd_step(w, grad, st["momentum_buffer"], group["lr"], group["momentum"], group["weight_decay"])
                # the kernel wrote through the raw pointer: a one-element in-place no-op on a view bumps the shared
                # version counter, which is what HeadOperands.refresh() watches to re-cast the bf16 GEMM operands
                w.view(-1)[:1].add_(0.0)
        return loss


def build_optimizer(cfg, model: torch.nn.Module) -> torch.optim.Optimizer:
    S = cfg.SOLVER
    if S.get("NESTEROV", False):
        raise NotImplementedError("SOLVER.NESTEROV is False in every OICR+ config; the fused SGD kernel has no Nesterov form")
    clip = S.get("CLIP_GRADIENTS", None)
    if clip is not None and clip.get("ENABLED", False):
        raise NotImplementedError("SOLVER.CLIP_GRADIENTS is not built (off in every released OICR+ config)")
    return B200SGD(get_optimizer_param_groups(cfg, model), S.BASE_LR, momentum=S.MOMENTUM)
