"""build_optimizer -- the parameter-group rules of the reference's solver (uwsod/detectron2/solver/build.py:143-218)
over the fused SGD kernel: per parameter lr / weight decay (bias lr x BIAS_LR_FACTOR with
WEIGHT_DECAY_BIAS, norm layers WEIGHT_DECAY_NORM, optional higher lr on the refinement branches), SGD with momentum
exactly as torch.optim.SGD(momentum=m, dampening=0, nesterov=False) computes it:
    buf = m * buf + (grad + wd * p);  p -= lr * buf
`B200SGD.step()` makes ONE pass over the parameters in ONE launch per 32 tensors (soswsod_sgd_multi reads p, grad, buf
and writes p, buf) instead of torch's 3-4 elementwise passes per tensor, and -- attached to an OICRPlusHeads -- writes
the head's bf16 GEMM operands in the same pass.  Gradient clipping (SOLVER.CLIP_GRADIENTS) is off in every released OICR+ config and not built."""
from __future__ import annotations

import contextlib
import os
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
    """torch.optim.SGD(momentum, dampening=0, nesterov=False) semantics; state key `momentum_buffer` like torch's.
    The whole step is ONE launch per 32 parameters (soswsod_sgd_multi).  With `attach_head(heads)` the same pass also
    writes the head's bf16 GEMM operands (and the fused head-bias vector), so the next forward does not re-cast 121 M
    weights (3 cast launches over 726 MB, engine.HeadOperands.refresh)."""

    def __init__(self, params, lr: float, momentum: float = 0.0, weight_decay: float = 0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._heads = []
        self.launches_last_step = 0
        # single GPU: run the update of an attached head UNDER the step's input-gradient GEMM and ROI backward (see
        # _overlapped_head_update); False = always on the caller's stream after everything
        self.overlap_update = True
        self.overlapped_last_step = False
        self._update_stream = None

    def attach_head(self, heads) -> "B200SGD":
        """heads: an OICRPlusHeads whose parameters this optimizer updates.  If the head exchanges its gradients
        (set_gradient_exchange) this optimizer becomes their consumer: backward() no longer blocks the compute stream."""
        self._heads.append(heads)
        ex = getattr(heads, "exchange", None)
        if ex is not None:
            ex.lazy_wait = True
        return self

    def sync_state(self) -> None:
        """Sharded data-parallel update: a rank's momentum buffers are current only on the rows it owns; this gathers
        the other rows (collective -- every rank calls it), e.g. before `state_dict()` for a checkpoint."""
        for h in self._heads:
            ex = getattr(h, "exchange", None)
            if ex is None or ex.world == 1:
                continue
            ex.sync_master()
            for key in sorted(ex.sharded):
                st = self.state.get(ex.master[key], {})
                if "momentum_buffer" in st:
                    ex.gather_rows(st["momentum_buffer"], key)

    def _overlapped_head_update(self, sinks) -> Set[int]:
        """Single-GPU heads.  The engine queues the weight gradient of fc6 BEFORE the input-gradient GEMM and the ROI
        backward (2.2 of the step's 7 ms, tensor- and shared-memory-bound), and publishes the events behind which every
        parameter gradient is complete.  The host reaches optimizer.step() while the device is still milliseconds
        behind, so the whole update of the head (HBM-bound, 2.7 GB) is queued on a second stream behind those events
        and runs under those kernels; the refreshed fc6 operand goes to a spare buffer (the input-gradient GEMM still
        reads the current one) and the two swap.  The caller's stream waits for the update before anything else it is
        given, so every later reader of the parameters is ordered as if the update had run in place.
        Only if every `.grad` is still EXACTLY the tensor the step produced (same address, same version counter: no
        accumulation, rescaling, clipping in between); otherwise nothing is done here and step() updates as usual.
        Returns the ids of the parameters updated."""
        done: Set[int] = set()
        self.overlapped_last_step = False
        for h in self._heads:
            ex = getattr(h, "exchange", None)
            eng = h.engine()
            early, eng.early_grads = eng.early_grads, None       # one use: a later step() without a new backward finds nothing
            if early is None or not self.overlap_update or (ex is not None and ex.world > 1):
                continue
            master = eng.op.master
            if any(p.grad is None or (p.grad.data_ptr(), p.grad._version) != early["ident"][k] or not p.grad.is_contiguous()
                   or not p.is_contiguous() for k, p in master.items()):
                continue
            group_of = {id(p): g for g in self.param_groups for p in g["params"]}
            if any(id(p) not in group_of for p in master.values()):
                continue
            if self._update_stream is None:
                self._update_stream = torch.cuda.Stream(device=master["fc1_w"].device)
            us = self._update_stream
            for ev in early["events"]:
                us.wait_event(ev)
            by_momentum = {}
            with torch.cuda.stream(us):
                for k, p in master.items():
                    g = group_of[id(p)]
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    p.grad.record_stream(us)       # `.grad` may be dropped by the host before the update ran
                    ob, of = sinks.get(id(p), (None, None))
                    if k == "fc1_w":
                        ob = eng.op.spare_w6()
                    by_momentum.setdefault(float(g["momentum"]), []).append(
                        (p.detach(), p.grad, st["momentum_buffer"], g["lr"], g["weight_decay"], ob, of))
                    done.add(id(p))
                for momentum, items in by_momentum.items():
                    self.launches_last_step += ops.sgd_multi(items, momentum)
                finished = torch.cuda.Event()
                finished.record()
            eng.op.swap_w6()
            torch.cuda.current_stream().wait_event(finished)
            self.overlapped_last_step = True
        return done

    def _sinks(self):
        out = {}
        for h in self._heads:
            out.update(h.engine().op.sinks())
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sinks = self._sinks()
        self.launches_last_step = 0
        updated_early = self._overlapped_head_update(sinks)
        # Data-parallel heads: the rest of the exchange runs on the exchange's UPDATE STREAM -- wait for the gradient
        # collectives, update (a sharded parameter only on the rows this rank owns: their averaged gradient arrived by
        # reduce-scatter), all-gather the bf16 operand rows -- while the compute stream goes on to the next step's ROI
        # pooling and waits at the engine's operand gate, right before its first GEMM.
        shard_rows, exchanges, key_of = {}, [], {}
        for h in self._heads:
            ex = getattr(h, "exchange", None)
            if ex is not None and ex.world > 1:
                exchanges.append((h, ex))
                for key, prm in ex.master.items():
                    key_of[id(prm)] = (ex, key)
        update_stream = exchanges[0][1].update_stream() if exchanges else None
        if update_stream is not None and os.environ.get("SOSWSOD_UPDATE_ON_MAIN"):      # diagnosis: serialise the update
            update_stream = torch.cuda.current_stream()
        ready = None
        if update_stream is not None:
            ready = torch.cuda.Event()
            ready.record()                       # every gradient kernel of the step is queued behind this point
        ctx = torch.cuda.stream(update_stream) if update_stream is not None else contextlib.nullcontext()
        nvls_items, nvls_params = [], set()
        for h, ex in exchanges:
            for key in sorted(ex.sharded):
                if ex.mode == "nvls":
                    nvls_params.add(id(ex.master[key]))
        if nvls_params:
            # nvls: the big matrices first, behind nothing but the kernels that produced their gradients (the host is
            # milliseconds ahead of the device: this is queued while the backward is still running, and the update kernel
            # shares the SMs with the input-gradient GEMM)
            with ctx:
                for group in self.param_groups:
                    for p in group["params"]:
                        if id(p) not in nvls_params or p.grad is None:
                            continue
                        ex, key = key_of[id(p)]
                        ex.check_gradient_buffer(key, p.grad)
                        st = self.state[p]
                        if "momentum_buffer" not in st:
                            st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        lo, hi = ex.owned_rows_nvls(key)
                        nvls_items.append((float(group["momentum"]), ex,
                                           (p.detach()[lo:hi], ex.multicast_address(f"g:{key}", lo), st["momentum_buffer"][lo:hi],
                                            ex.multicast_address(f"w{ex.operand_slot ^ 1}:{key}", lo), group["lr"], group["weight_decay"])))
                if nvls_items:
                    ex = nvls_items[0][1]
                    ex.wait_big_gradients()
                    ex.barrier()             # every rank's weight gradients are complete in the symmetric buffers
                    for momentum in sorted({m for m, _, _ in nvls_items}):
                        self.launches_last_step += ops.sgd_nvls([it for m, _, it in nvls_items if m == momentum], momentum, 1.0 / ex.world)
        if update_stream is not None:
            update_stream.wait_event(ready)
        with ctx:
            for h, ex in exchanges:
                ex.wait_gradients()
                for key in sorted(ex.sharded):
                    shard_rows[id(ex.master[key])] = ex.owned_rows(key)
            by_momentum, touched = {}, []          # the reference makes one group per parameter: batch across groups
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    if id(p) in updated_early:
                        touched.append(p)
                        continue
                    if not p.is_cuda:
                        raise RuntimeError("B200SGD steps CUDA parameters only (sm_100a kernel); there is no CPU fallback")
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    if not p.is_contiguous():
                        raise RuntimeError("B200SGD needs contiguous parameters")
                    if id(p) in key_of:
                        key_of[id(p)][0].check_gradient_buffer(key_of[id(p)][1], p.grad)
                    if update_stream is not None and id(p) not in nvls_params:
                        grad.record_stream(update_stream)     # `.grad` may be dropped by the host before the update ran
                    ob, of = sinks.get(id(p), (None, None))
                    items = by_momentum.setdefault(float(group["momentum"]), [])
                    pd, buf = p.detach(), st["momentum_buffer"]
                    if id(p) in nvls_params:       # done above by the fused reduce + update + broadcast kernel
                        touched.append(p)
                        continue
                    for lo, hi in shard_rows.get(id(p), [(0, p.size(0) if p.dim() else 1)]):
                        whole = (lo == 0 and hi == (p.size(0) if p.dim() else 1))
                        sl = (lambda t: t) if whole else (lambda t, lo=lo, hi=hi: None if t is None else t[lo:hi])
                        items.append((sl(pd), sl(grad), sl(buf), group["lr"], group["weight_decay"], sl(ob), sl(of)))
                    touched.append(p)
            for momentum, items in by_momentum.items():
                self.launches_last_step += ops.sgd_multi(items, momentum)
            if nvls_items:
                nvls_items[0][1].barrier()   # every rank's operand rows have landed everywhere; the gradient buffers are free
            for h, ex in exchanges:
                if ex.mode == "nvls":
                    ex.master_stale = True
                    if nvls_items:       # from the next launch on, the engine computes with the copies just written
                        ex.operand_slot ^= 1
                        op = h.engine().op
                        if "fc1_w" in ex.sharded:
                            op.w6 = ex.operand("fc1_w")
                        if "fc2_w" in ex.sharded:
                            op.w7 = ex.operand("fc2_w")
                else:
                    op = h.engine().op
                    ex.gather_operands({"fc1_w": op.w6, "fc2_w": op.w7})
                ex.mark_update_done()
        for p in touched:
            # the kernel wrote through raw pointers: bump the version counters like an in-place op would
            torch.autograd.graph.increment_version(p)
        for h in self._heads:
            op = h.engine().op
            if all(p.grad is not None for p in op.master.values()):
                op.mark_fresh()          # every operand was rewritten by this step
        return loss


def build_optimizer(cfg, model: torch.nn.Module) -> torch.optim.Optimizer:
    S = cfg.SOLVER
    if S.get("NESTEROV", False):
        raise NotImplementedError("SOLVER.NESTEROV is False in every OICR+ config; the fused SGD kernel has no Nesterov form")
    clip = S.get("CLIP_GRADIENTS", None)
    if clip is not None and clip.get("ENABLED", False):
        raise NotImplementedError("SOLVER.CLIP_GRADIENTS is not built (off in every released OICR+ config)")
    opt = B200SGD(get_optimizer_param_groups(cfg, model), S.BASE_LR, momentum=S.MOMENTUM)
    from .modeling.roi_heads_oicrplus import OICRPlusHeads

    for m in model.modules():
        if isinstance(m, OICRPlusHeads) and next(m.parameters()).is_cuda:
            opt.attach_head(m)
    return opt
