"""OICRPlusHeads -- the reference's Stage-1 ROI head (uwsod/projects/WSL/wsl/modeling/roi_heads/
roi_heads_oicrplus.py:36-757) as a drop-in: same registry name, constructor kwargs, sub-module and parameter
names (box_pooler, box_head.fc1/fc2, box_predictor.cls/det, box_refinery_{k}.cls_score/bbox_pred), same
forward signature and return structure.  Training and inference both run through the fused engine
(engine.py): all views batched through one ROI-pool/fc6/fc7/head pass with a hand-scheduled backward."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ..engine import HeadConfig, HeadOperands, OICRPlusHeadEngine, ViewBatch
from ..registry import ROI_HEADS_REGISTRY
from ..structures import Instances, ShapeSpec
from .box_head import build_box_head
from .fast_rcnn_oicr import OICROutputLayers, _detections_to_instances
from .fast_rcnn_wsddn import WSDDNOutputLayers
from .poolers import ROIPooler


def get_image_level_gt(targets: List[Instances], num_classes: int):
    """uwsod/projects/WSL/wsl/modeling/roi_heads/roi_heads.py:144-164 -> (list of per-image class tensors,
    unique sorted int64 classes per image, one-hot [N_img, C])."""
    if targets is None:
        return None, None, None
    gt_classes_img = [torch.unique(t.gt_classes, sorted=True) for t in targets]
    gt_classes_img_int = [gt.to(torch.int64) for gt in gt_classes_img]
    gt_classes_img_oh = torch.cat(
        [torch.zeros((1, num_classes), dtype=torch.float, device=gt.device).scatter_(1, torch.unsqueeze(gt, dim=0), 1)
         for gt in gt_classes_img_int], dim=0)
    return gt_classes_img, gt_classes_img_int, gt_classes_img_oh


class _FusedHeadStep(Function):
    """Runs engine.train_step in forward (which already produces every gradient of sum(losses)) and hands the
    stored gradients to autograd in backward.  Inputs: feats..., then the parameters in HeadOperands.master order."""

    @staticmethod
    def forward(ctx, engine, vb, gt_int, seeds, n_feats, loss_scale, *tensors):
        feats = tensors[:n_feats]
        need_fg = any(f.requires_grad for f in feats)
        gt_count = gt_onehot = None
        if isinstance(gt_int, tuple):      # (padded list, count, one-hot) kept on the device
            gt_int, gt_count, gt_onehot = gt_int
        out = engine.train_step(ViewBatch([f.detach() for f in feats], vb.rois, vb.obj, vb.R), gt_int, seeds,
                                need_feat_grad=need_fg, loss_scale=loss_scale, grad_hook=engine.grad_hook,
                                gt_count=gt_count, gt_onehot=gt_onehot)
        ctx.engine_out = out
        ctx.exchange = getattr(engine, "exchange", None)
        ctx.deferred = engine.deferred_scale_check
        ctx.loss_scale = float(loss_scale)
        ctx.n_feats = n_feats
        ctx.keys = list(engine.op.master.keys())
        ctx.loss_keys = list(out.losses.keys())
        engine.last_output = out
        return tuple(out.losses[k].clone() for k in ctx.loss_keys)

    @staticmethod
    @once_differentiable
    def backward(ctx, *gouts):
        out = ctx.engine_out
        if out is None:
            raise RuntimeError("OICRPlusHeads: backward ran twice through one fused head step (retain_graph=True / a second "
                               ".backward()); the step hands its gradient buffers to autograd once -- run forward again")
        if ctx.exchange is not None and ctx.exchange.world > 1:
            # The gradient collectives started from the engine's hook run on NCCL's stream.  DDP's contract: order this
            # stream behind them before autograd touches the buffers.  With an attached B200SGD (lazy_wait) the optimizer
            # consumes them on the exchange's update stream instead, and checks that `.grad` IS the reduced buffer.
            if ctx.exchange.lazy_wait:
                # the all-reduced tensors (biases, the fused head block whose row slices autograd may clone) are final
                # when backward returns; the reduce-scattered matrices stay with the update stream
                ctx.exchange.wait_allreduces()
            else:
                ctx.exchange.wait_gradients()
            ctx.exchange.expect_gradient_buffers({k: out.grads[k].data_ptr() for k in ctx.exchange.sharded})
        gs = torch.stack([x.reshape(()) for x in gouts])
        pre = ctx.loss_scale        # the gradients were produced for pre * sum(losses) (heads.expected_loss_scale)
        if ctx.deferred is not None:
            # upstream gradients are checked to equal the expected scale (1 for the reference trainer's
            # `sum(loss_dict.values()).backward()`, tools/train_net_multi.py:139) WITHOUT stalling the host: the values
            # go to pinned memory behind an event and OICRPlusHeads.check_deferred() / the next forward raises if they
            # were anything else
            ctx.deferred.push(gs, pre)
            s = 1.0
        else:
            g = gs.tolist()     # one tiny D2H read (host waits for the step)
            if any(abs(v - g[0]) > 1e-12 * max(1.0, abs(g[0])) for v in g):
                raise NotImplementedError(
                    "the fused OICR+ head step produces the gradient of s * sum(loss_dict.values()) -- the reference "
                    "trainer's objective (tools/train_net_multi.py:139); per-key loss weights are not supported")
            s = g[0] / pre if pre != 0.0 else 0.0
        def sc(t):
            return t if s == 1.0 else t * s
        gf = [sc(t) for t in out.grad_feats] if out.grad_feats else [None] * ctx.n_feats
        gp = [sc(out.grads[k]) for k in ctx.keys]
        # hand the buffers over: with no other reference left, autograd's AccumulateGrad adopts them as .grad instead
        # of cloning 484 MB of parameter gradients per step (engine.last_output keeps losses / aux only)
        ctx.engine_out = None
        out.grads = {}
        out.grad_feats = []
        return (None, None, None, None, None, None, *gf, *gp)


class _DeferredScaleCheck:
    """Upstream-gradient values of past backward calls, parked in pinned host memory behind CUDA events."""

    def __init__(self):
        self.pending = []

    def push(self, gs: torch.Tensor, expected: float) -> None:
        host = torch.empty(gs.shape, dtype=gs.dtype, pin_memory=True)
        host.copy_(gs, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((host, ev, float(expected)))

    def check(self, wait: bool) -> None:
        keep = []
        for host, ev, expected in self.pending:
            if not wait and not ev.query():
                keep.append((host, ev, expected))
                continue
            ev.synchronize()
            if not bool(torch.all((host - expected).abs() <= 1e-6 * max(1.0, abs(expected)))):
                self.pending = []
                raise NotImplementedError(
                    "loss_scale_check='deferred': a previous backward received the upstream gradients "
                    f"{host.tolist()} but the step was run for {expected} * sum(loss_dict.values()).  Set "
                    "heads.expected_loss_scale to the factor applied to the summed loss (AMP GradScaler scale, "
                    "1 / ITER_SIZE), or use loss_scale_check='sync'.  The gradients of that step were WRONGLY scaled.")
        self.pending = keep


@ROI_HEADS_REGISTRY.register()
class OICRPlusHeads(nn.Module):
    def __init__(self, *, box_in_features: List[str], box_pooler: ROIPooler, box_head: nn.Module,
                 box_predictor: nn.Module, vis_period: int = 0, refine_K: int = 4, refine_mist: bool = False,
                 mist_p: float = 0.10, mist_thre: float = 0.05, mist_type: str = "nms",
                 refine_reg: List[bool] = (False, False, False, False), box_refinery: List[nn.Module] = (None,) * 4,
                 cls_agnostic_bbox_reg: bool = False, pooler_type: str = "ROIPool", cfg=None, num_classes: int = 20,
                 **kwargs):
        super().__init__()
        assert mist_type in ["nms", "wetectron"], f"{mist_type} is wrong"
        if mist_type != "nms" or not refine_mist:
            raise NotImplementedError("only REFINE_MIST: True with MIST_TYPE: nms (the released OICR+ configuration) is "
                                      "built; the reference's 'wetectron' variant references an undefined attribute "
                                      "(roi_heads_oicrplus.py:555) and cannot run there either")
        assert pooler_type == "ROIPool"
        self.mist_type, self.mist_p, self.mist_thre = mist_type, mist_p, mist_thre
        self.cfg = cfg
        self.num_classes = num_classes
        self.in_features = self.box_in_features = box_in_features
        self.box_pooler = box_pooler
        self.box_head = box_head
        self.box_predictor = box_predictor
        self.pooler_type = pooler_type
        self.iter = 0
        self.iter_test = 0
        self.vis_period = vis_period
        self.refine_K = refine_K
        self.refine_mist = refine_mist
        self.refine_reg = list(refine_reg)
        self.box_refinery = list(box_refinery)
        for k in range(self.refine_K):
            self.add_module("box_refinery_{}".format(k), self.box_refinery[k])
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.reproduce_flip_quirk = True     # roi_heads_oicrplus.py:381; set False to use the flipped logits
        self._engine: Optional[OICRPlusHeadEngine] = None
        # data-parallel: grad_hook(key, grad, row0) starts a gradient's collective as soon as it is produced; normally
        # installed by set_gradient_exchange()
        self.grad_hook = None
        self.exchange = None
        self.last_metrics: Dict[str, torch.Tensor] = {}
        # "sync": backward reads the upstream gradients on the host (exact, stalls the host once per step);
        # "deferred": assumes sum(loss_dict.values()).backward() and verifies it one step late, without a stall
        self.loss_scale_check = "sync"
        # the factor the caller applies to sum(loss_dict.values()) before .backward() (AMP GradScaler scale, 1 / ITER_SIZE
        # gradient accumulation): the fused step produces its gradients already multiplied by it (free: folded into the
        # logit-gradient cast); "sync" corrects any other value exactly, "deferred" verifies it one step late
        self.expected_loss_scale = 1.0
        self._deferred = _DeferredScaleCheck()
        # image-level labels of CUDA targets are derived on the device (soswsod_image_level_gt); False restores the
        # reference's torch.unique call (identical values, one host synchronisation per step)
        self.image_level_gt_on_device = True
        self._gt_dev = None

    # ---- construction from config (roi_heads_oicrplus.py:88-147) ----
    @staticmethod
    def check_config(cfg) -> None:
        """Every flag that changes the reference's arithmetic is either honoured or refused here, at construction:
        the fused kernels implement the released OICR+ configuration and nothing may silently fall back to it."""
        W, H, B = cfg.WSL, cfg.MODEL.ROI_HEADS, cfg.MODEL.ROI_BOX_HEAD

        def refuse(cond, what):
            if cond:
                raise NotImplementedError(f"OICRPlusHeads (sm_100a fused path): {what}")

        refuse(cfg.get("OICRPLUS", {}).get("BBOX_UPDATE", False),
               "OICRPLUS.BBOX_UPDATE: True (delta-averaged pseudo-box update, roi_heads_oicrplus.py:397-425) is not "
               "built; it is False in every released config")
        refuse(B.get("BBOX_REG_LOSS_TYPE", "smooth_l1") != "smooth_l1" or float(B.get("SMOOTH_L1_BETA", 0.0)) != 0.0,
               "the box-regression loss is the released smooth_l1 with SMOOTH_L1_BETA 0.0 (= L1, fast_rcnn_oicr.py:309-318); "
               f"got {B.get('BBOX_REG_LOSS_TYPE')} / beta {B.get('SMOOTH_L1_BETA')}")
        refuse(float(B.get("BBOX_REG_LOSS_WEIGHT", 1.0)) != 1.0, "BBOX_REG_LOSS_WEIGHT must be 1.0 (per-key loss weights are "
               "not supported by the fused backward)")
        refuse(not W.get("MEAN_LOSS", True), "WSL.MEAN_LOSS: False (sum-reduced BCE, fast_rcnn_wsddn.py:340-358) is not built")
        refuse(list(H.get("IOU_LABELS", [0, -1, 1])) != [0, -1, 1] or len(H.IOU_THRESHOLDS) != 2,
               "MODEL.ROI_HEADS.IOU_LABELS must be [0, -1, 1] with two IOU_THRESHOLDS (matcher.py:63-111)")
        refuse(H.get("PROPOSAL_APPEND_GT", False), "PROPOSAL_APPEND_GT: True is not built (False in Base-RCNN-DilatedC5.yaml:15)")
        refuse(B.get("CLS_AGNOSTIC_BBOX_REG", False), "class-agnostic box regression is not built")
        refuse(not W.REFINE_MIST or W.MIST_TYPE != "nms", "only REFINE_MIST: True with MIST_TYPE: nms is built")
        refuse(not all(bool(x) for x in list(W.REFINE_REG)[:W.REFINE_NUM]) or len(W.REFINE_REG) < W.REFINE_NUM,
               "WSL.REFINE_REG must be True for every refinement branch")
        refuse(B.POOLER_TYPE != "ROIPool", f"POOLER_TYPE {B.POOLER_TYPE} (only ROIPool is on the OICR+ path)")

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        cls.check_config(cfg)
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        res = cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        scale = 1.0 / input_shape[in_features[-1]].stride
        in_channels = input_shape[in_features[-1]].channels
        box_pooler = ROIPooler(output_size=res, scales=(scale,), sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                               pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE)
        box_head = build_box_head(cfg, ShapeSpec(channels=in_channels, height=res, width=res))
        box_predictor = WSDDNOutputLayers(cfg, box_head.output_shape)
        K = cfg.WSL.REFINE_NUM
        return cls(box_in_features=in_features, box_pooler=box_pooler, box_head=box_head, box_predictor=box_predictor,
                   vis_period=cfg.VIS_PERIOD, refine_K=K, refine_mist=cfg.WSL.REFINE_MIST, mist_p=cfg.WSL.MIST_P,
                   mist_thre=cfg.WSL.MIST_THRE, mist_type=cfg.WSL.MIST_TYPE, refine_reg=cfg.WSL.REFINE_REG,
                   box_refinery=[OICROutputLayers(cfg, box_head.output_shape, k) for k in range(K)],
                   cls_agnostic_bbox_reg=cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG,
                   pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE, cfg=cfg, num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES)

    # ---- engine plumbing ----
    def head_config(self) -> HeadConfig:
        cfg = self.cfg
        fc1 = self.box_head.fc1
        res = self.box_pooler.output_size[0]
        hc = HeadConfig(num_classes=self.num_classes, refine_k=self.refine_K, pooled=res,
                        spatial_scale=self.box_pooler.level_poolers[0].spatial_scale,
                        in_channels=fc1.in_features // (res * res), fc_dim=fc1.out_features,
                        dropout_p=self.box_head.dropout, mist_p=self.mist_p, mist_thre=self.mist_thre,
                        reproduce_flip_quirk=self.reproduce_flip_quirk)
        if cfg is not None:
            hc.iou_thresholds = tuple(cfg.MODEL.ROI_HEADS.IOU_THRESHOLDS)
            hc.bbox_reg_weights = tuple(cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)
            hc.score_thresh_test = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
            hc.nms_thresh_test = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
            hc.detections_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        return hc

    def engine(self) -> OICRPlusHeadEngine:
        dev = self.box_head.fc1.weight.device
        if self._engine is None or self._engine.op.w6.device != dev:
            hc = self.head_config()
            bh, bp = self.box_head, self.box_predictor
            refine = [(r.cls_score.weight, r.cls_score.bias, r.bbox_pred.weight, r.bbox_pred.bias)
                      for r in self.box_refinery[:self.refine_K]]
            op = HeadOperands(hc, bh.fc1.weight, bh.fc1.bias, bh.fc2.weight, bh.fc2.bias, bp.cls.weight, bp.cls.bias,
                              bp.det.weight, bp.det.bias, refine)
            self._engine = OICRPlusHeadEngine(hc, op)
        self._engine.cfg.reproduce_flip_quirk = self.reproduce_flip_quirk
        self._engine.grad_hook = self.grad_hook
        self._engine.exchange = self.exchange
        self._engine.operand_gate = self.exchange.operand_gate if self.exchange is not None else None
        self._engine.op.pre_refresh = self.exchange.sync_master if self.exchange is not None else None
        assert self.loss_scale_check in ("sync", "deferred")
        self._engine.deferred_scale_check = self._deferred if self.loss_scale_check == "deferred" else None
        return self._engine

    def set_gradient_exchange(self, mode: str = "sharded", group=None):
        """Data-parallel training (one image per GPU, tools/train_net_multi.py:75-78 wraps the model in DDP): installs a
        distributed.GradientExchange on this head -- its hook starts every gradient's collective from inside the
        backward, `backward()` returns only after ordering the stream behind them, and an attached solver.B200SGD
        finishes the exchange (sharded update + operand all-gather).  Returns the exchange."""
        from ..distributed import GradientExchange

        eng = self.engine()
        self.exchange = GradientExchange(eng.op.master, group=group, mode=mode)
        self.grad_hook = self.exchange.hook if self.exchange.world > 1 else None
        if mode == "nvls" and self.exchange.world > 1:
            ex = self.exchange
            ex.setup_nvls()
            # the engine writes the big weight gradients straight into the symmetric buffers (one launch each: no row
            # panels, nothing to overlap) and computes with the symmetric operands the fused update broadcasts into
            eng.external_grad_bufs.update({k: ex.symm_tensors[f"g:{k}"] for k in ex.sharded})
            eng.op.adopt_operand_storage(w6=ex.operand("fc1_w") if "fc1_w" in ex.sharded else None,
                                         w7=ex.operand("fc2_w") if "fc2_w" in ex.sharded else None)
            eng.fc1_wgrad_panels = 1
        self.engine()
        return self.exchange

    def state_dict(self, *args, **kwargs):
        if self.exchange is not None:
            self.exchange.sync_master()      # rows updated by other ranks (sharded optimizer) -> current fp32 values
        return super().state_dict(*args, **kwargs)

    def image_level_gt_lists(self):
        """(gt_classes_img, gt_classes_img_int, gt_classes_img_oh) exactly as get_image_level_gt returns them, for the
        last forward (reads the class count back when the labels were kept on the device)."""
        if self._gt_dev is not None:
            lst, cnt, oh = self._gt_dev
            g = lst[:int(cnt.item())].to(torch.int64)
            return [g], [g], oh[None]
        return self.gt_classes_img, self.gt_classes_img_int, self.gt_classes_img_oh

    def metric_scalars(self) -> Dict[str, float]:
        """The scalars the reference's step leaves in its EventStorage, under the reference's names, from the counters
        of the last training forward (one small device->host read; call it at the logging period):
          roi_head/num_{fg,bg,ig}_samples_r{k}            wsl/modeling/roi_heads/roi_heads.py:364-373
          fast_rcnn/{cls_accuracy,fg_cls_accuracy,false_negative}_r{k}   fast_rcnn_oicr.py:228-256
        `_log_accuracy` runs once per view and a storage keeps the LAST value written, i.e. the one of view "2_flip"
        (which, :381, is evaluated on view 2's logits) -- reproduced here."""
        if not self.last_metrics:
            return {}
        acc = self.last_metrics["acc_counts"].cpu()       # [K, V, 5]: rows, fg, accurate, fg accurate, false negatives
        lab = self.last_metrics["label_counts"].cpu()     # [K, 3]: fg, bg, ignored
        out: Dict[str, float] = {}
        for k in range(acc.size(0)):
            out[f"roi_head/num_fg_samples_r{k}"] = float(lab[k, 0])
            out[f"roi_head/num_bg_samples_r{k}"] = float(lab[k, 1])
            out[f"roi_head/num_ig_samples_r{k}"] = float(lab[k, 2])
            n, n_fg, n_acc, n_fgacc, n_fn = (int(x) for x in acc[k, -1])
            if n > 0:
                out[f"fast_rcnn/cls_accuracy_r{k}"] = n_acc / n
                if n_fg > 0:
                    out[f"fast_rcnn/fg_cls_accuracy_r{k}"] = n_fgacc / n_fg
                    out[f"fast_rcnn/false_negative_r{k}"] = n_fn / n_fg
        return out

    def log_metrics(self, storage) -> None:
        """storage: anything with put_scalar(name, value) -- detectron2's EventStorage."""
        for k, v in self.metric_scalars().items():
            storage.put_scalar(k, v)

    def check_deferred(self, wait: bool = True) -> None:
        """Raises if a backward run under loss_scale_check='deferred' saw a non-unit upstream gradient."""
        self._deferred.check(wait)

    def _param_list(self):
        return list(self.engine().op.master.values())

    @staticmethod
    def _view_rois(proposal_groups: List[List[Instances]]):
        """One roi tensor per feature tensor: views of a group are the batch entries 0..n-1 of that tensor."""
        rois, obj = [], []
        for group in proposal_groups:
            parts = []
            for i, p in enumerate(group):
                t = p.proposal_boxes.tensor
                parts.append(torch.cat((torch.full((len(t), 1), float(i), dtype=t.dtype, device=t.device), t), 1))
                obj.append(p.objectness_logits)
            rois.append(torch.cat(parts, 0).contiguous())
        return rois, torch.cat(obj, 0).float().contiguous()

    # ---- forward (roi_heads_oicrplus.py:149-188) ----
    def forward(self, images_list, features_list, proposals_list, targets_list=(None, None, None, None)):
        if not self.training:
            pred_instances, all_scores, all_boxes = self._forward_box_test(features_list, proposals_list, targets_list)
            return pred_instances, {}, all_scores, all_boxes
        features1, features2 = features_list
        proposals1, proposals1_flip, proposals2, proposals2_flip = proposals_list
        targets1 = targets_list[0]
        self._gt_dev = None
        if (self.image_level_gt_on_device and targets1 is not None and len(targets1) == 1
                and targets1[0].gt_classes.is_cuda and self.num_classes <= 128):
            # same result as get_image_level_gt, but the number of distinct classes never travels to the host
            # (torch.unique reads it back, which stalls the host until the previous step has drained)
            from .. import ops
            lst, cnt, oh = ops.image_level_gt(targets1[0].gt_classes, self.num_classes)
            self._gt_dev = (lst, cnt, oh)
            self.gt_classes_img_oh = oh[None]
            self.gt_classes_img = self.gt_classes_img_int = None      # materialised on demand: image_level_gt_lists()
        else:
            self.gt_classes_img, self.gt_classes_img_int, self.gt_classes_img_oh = get_image_level_gt(targets1, self.num_classes)
        f1 = features1[self.box_in_features[-1]]
        f2 = features2[self.box_in_features[-1]]
        losses = self._forward_box(f1, f2, proposals1, proposals1_flip, proposals2, proposals2_flip)
        self.iter = self.iter + 1
        return None, losses

    def _forward_box(self, features1, features2, proposals1, proposals1_flip, proposals2, proposals2_flip):
        """features1/2: [2,C,h,w] (image, flipped image) of the two scales; one image per GPU (rcnn_multi.py:148)."""
        assert len(proposals1) == 1, "the reference trains with one image per GPU (rcnn_multi.py:148)"
        eng = self.engine()
        self._deferred.check(wait=False)
        if self.exchange is not None:
            self.exchange.begin_step()
        rois, obj = self._view_rois([[proposals1[0], proposals1_flip[0]], [proposals2[0], proposals2_flip[0]]])
        R = len(proposals1[0])
        vb = ViewBatch([features1, features2], rois, obj, R)
        seeds = (torch.initial_seed() * 7919 + 2 * self.iter + 1, torch.initial_seed() * 7919 + 2 * self.iter + 2)
        gt_arg = self._gt_dev if self._gt_dev is not None else self.gt_classes_img_int[0]
        outs = _FusedHeadStep.apply(eng, vb, gt_arg, seeds, 2, float(self.expected_loss_scale), features1, features2,
                                    *self._param_list())
        out = eng.last_output
        self.last_metrics = {"acc_counts": out.aux["acc_counts"], "label_counts": out.aux["counts"]}
        return dict(zip(out.losses.keys(), outs))

    def _forward_box_test(self, features, proposals, targets=None):
        """roi_heads_oicrplus.py:432-475: single view -> ([Instances], all_scores, all_boxes) with, as in the reference
        (fast_rcnn_oicr.py:46-83 returns per-image lists), all_scores = [[1,R,C+1]] and all_boxes = [[1,R,4C]]."""
        eng = self.engine()
        if isinstance(features, dict):
            f = features[self.box_in_features[-1]]
        else:
            f = features[0] if isinstance(features, (list, tuple)) else features
            if isinstance(f, dict):
                f = f[self.box_in_features[-1]]
        assert len(proposals) == 1 and f.size(0) == 1, "inference runs one image (view) per call, as the reference's TTA does"
        rois, obj = self._view_rois([[proposals[0]]])
        vb = ViewBatch([f], rois, obj, len(proposals[0]))
        probs, pboxes = eng.test_forward(vb)
        inst, _ = _detections_to_instances(eng.detect(probs[0], pboxes[0], proposals[0].image_size), proposals[0].image_size)
        return inst, [probs], [pboxes]


def build_roi_heads(cfg, input_shape):
    """uwsod/detectron2/modeling/roi_heads/roi_heads.py:38-43."""
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME).from_config(cfg, input_shape)
