"""The fused OICR+ head step: all views of an image batched through ONE pass of
ROI pool -> fc6 -> fc7 -> [cls | det | K x (cls_score, bbox_pred)] -> WSDDN -> K x OICR, with a hand-scheduled
backward (no autograd graph) -- 23 kernel launches for a 4-view forward+backward instead of the reference's
several thousand.

What it follows in the reference (paths under uwsod/projects/WSL/wsl/modeling/roi_heads/):
  train step   roi_heads_oicrplus.py:190-430 (_forward_box)   -- OICRPlusHeadEngine.train_step
  test forward roi_heads_oicrplus.py:432-475 (_forward_box_test) + fast_rcnn_oicr.py:584-735 -- .test_forward
Layout of the head logits matrix L [V*R, ld] fp32 (one GEMM for every output layer):
  [0, C)            WSDDN classification stream   (box_predictor.cls)
  [C, 2C)           WSDDN detection stream        (box_predictor.det)
  2C + k*(5C+1) ..  branch k: C+1 class logits (box_refinery_k.cls_score) then 4C box deltas (.bbox_pred)
  zero padding up to a multiple of 64 columns
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops


@dataclass
class HeadConfig:
    num_classes: int = 20
    refine_k: int = 4                 # WSL.REFINE_NUM of every code_release yaml (voc07_oicr_plus.yaml:56-58)
    pooled: int = 7
    spatial_scale: float = 1.0 / 8
    in_channels: int = 512
    fc_dim: int = 4096
    dropout_p: float = 0.5            # box_head.py:88-90 (training only)
    mist_p: float = 0.10              # WSL.MIST_P
    mist_thre: float = 0.05           # WSL.MIST_THRE
    mist_nms: float = 0.01            # roi_heads_oicrplus.py:576-581
    iou_thresholds: Tuple[float, float] = (0.5, 0.6)   # MODEL.ROI_HEADS.IOU_THRESHOLDS
    bbox_reg_weights: Tuple[float, float, float, float] = (10.0, 10.0, 5.0, 5.0)
    reproduce_flip_quirk: bool = True  # roi_heads_oicrplus.py:381
    score_thresh_test: float = 1e-6
    nms_thresh_test: float = 0.3
    detections_per_image: int = 100

    @property
    def in_dim(self) -> int:
        return self.in_channels * self.pooled * self.pooled

    @property
    def ref_stride(self) -> int:
        return 5 * self.num_classes + 1

    @property
    def col_ref0(self) -> int:
        return 2 * self.num_classes

    @property
    def head_cols(self) -> int:
        return 2 * self.num_classes + self.refine_k * self.ref_stride

    @property
    def head_cols_padded(self) -> int:
        return (self.head_cols + 63) // 64 * 64

    def top_k(self, R: int) -> int:
        # roi_heads_oicrplus.py:657-662 (Python float arithmetic, then int())
        p = self.mist_p
        if p >= 1:
            return min(R, int(p))
        if 0 < p < 1:
            return max(int(R * p), 1)
        return min(R, 1)


class HeadOperands:
    """bf16 GEMM-operand copies of the fp32 master parameters (names = the reference's checkpoint keys:
    box_head.fc1/fc2, box_predictor.cls/det, box_refinery_k.cls_score/bbox_pred)."""

    def __init__(self, cfg: HeadConfig, fc1_w, fc1_b, fc2_w, fc2_b, cls_w, cls_b, det_w, det_b, refine):
        self.cfg = cfg
        self.master = {"fc1_w": fc1_w, "fc1_b": fc1_b, "fc2_w": fc2_w, "fc2_b": fc2_b, "cls_w": cls_w, "cls_b": cls_b,
                       "det_w": det_w, "det_b": det_b}
        for k, (cw, cb, bw, bb) in enumerate(refine):
            self.master.update({f"r{k}_cls_w": cw, f"r{k}_cls_b": cb, f"r{k}_box_w": bw, f"r{k}_box_b": bb})
        dev = fc1_w.device
        self.w6 = torch.empty(fc1_w.shape, dtype=torch.bfloat16, device=dev)
        self.w7 = torch.empty(fc2_w.shape, dtype=torch.bfloat16, device=dev)
        self.wh = torch.zeros((cfg.head_cols_padded, cfg.fc_dim), dtype=torch.bfloat16, device=dev)
        self.bh = torch.zeros((cfg.head_cols_padded,), dtype=torch.float32, device=dev)
        self._w6_spare = None       # second fc6 operand buffer, see spare_w6()
        self._versions = None
        self.pre_refresh = None     # callable run before a re-cast (a sharded optimizer brings the fp32 masters up to date)
        self.refresh()

    def head_slices(self) -> List[Tuple[str, str, int, int]]:
        """(weight key, bias key, first row, number of rows) of every output layer inside the fused head matrix."""
        C, K = self.cfg.num_classes, self.cfg.refine_k
        out = [("cls_w", "cls_b", 0, C), ("det_w", "det_b", C, C)]
        for k in range(K):
            c0 = self.cfg.col_ref0 + k * self.cfg.ref_stride
            out.append((f"r{k}_cls_w", f"r{k}_cls_b", c0, C + 1))
            out.append((f"r{k}_box_w", f"r{k}_box_b", c0 + C + 1, 4 * C))
        return out

    def _current_versions(self):
        # (version counter, storage address): in-place updates bump the first, `param.data = ...` / module._apply
        # (.float(), .to(), load with assign=True) change the second.  A write through `.data` IN PLACE changes neither:
        # call invalidate() after one (EMA / LARS-style optimizers, manual weight surgery).
        return tuple((t._version, t.data_ptr()) for t in self.master.values())

    def invalidate(self) -> None:
        """Forces the next refresh(force=False) to re-cast every operand."""
        self._versions = None

    def adopt_operand_storage(self, w6: Optional[torch.Tensor] = None, w7: Optional[torch.Tensor] = None) -> None:
        """Moves the fc6 / fc7 bf16 operands into caller-owned memory of the same shape (symmetric multicast memory of
        the nvls gradient exchange), keeping their current values."""
        for name, t in (("w6", w6), ("w7", w7)):
            if t is not None:
                cur = getattr(self, name)
                assert t.shape == cur.shape and t.dtype == cur.dtype and t.device == cur.device
                t.copy_(cur)
                setattr(self, name, t)

    def spare_w6(self) -> torch.Tensor:
        """A second buffer for the fc6 operand: an optimizer that runs UNDER the step's input-gradient GEMM (which still
        reads `w6`) writes the refreshed operand here and then calls swap_w6()."""
        if self._w6_spare is None:
            self._w6_spare = torch.empty_like(self.w6)
        return self._w6_spare

    def swap_w6(self) -> None:
        self.w6, self._w6_spare = self._w6_spare, self.w6

    def mark_fresh(self) -> None:
        """Declares the operands up to date with the master parameters as they are now -- called by an optimizer that
        wrote the bf16 copies itself in its update pass (solver.B200SGD)."""
        self._versions = self._current_versions()

    def sinks(self) -> Dict[int, Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]:
        """id(master parameter) -> (bf16 operand view | None, fp32 copy view | None): where an optimizer that updates
        the parameter in one pass has to put the refreshed GEMM operand (soswsod_sgd_multi's out_bf16 / out_f32)."""
        m = self.master
        out = {id(m["fc1_w"]): (self.w6, None), id(m["fc2_w"]): (self.w7, None),
               id(m["fc1_b"]): (None, None), id(m["fc2_b"]): (None, None)}
        for wk, bk, r0, n in self.head_slices():
            out[id(m[wk])] = (self.wh[r0:r0 + n], None)
            out[id(m[bk])] = (None, self.bh[r0:r0 + n])
        return out

    def refresh(self, force: bool = True) -> None:
        """Re-casts the operands if any master parameter changed (optimizer step, checkpoint load)."""
        v = self._current_versions()
        if not force and v == self._versions:
            return
        if self.pre_refresh is not None:
            self.pre_refresh()
        m = self.master
        with torch.no_grad():
            ops.cast_f32_bf16(m["fc1_w"].detach(), out=self.w6)
            ops.cast_f32_bf16(m["fc2_w"].detach(), out=self.w7)
            for wk, bk, r0, n in self.head_slices():
                ops.cast_f32_bf16(m[wk].detach(), out=self.wh[r0:r0 + n])
                self.bh[r0:r0 + n].copy_(m[bk].detach())
        self._versions = self._current_versions()


@dataclass
class ViewBatch:
    """Inputs of one head step.  `feats[i]` is an NCHW fp32 conv5 map holding n_i views (the reference batches an
    image with its flip, rcnn_multi.py:174-175); `rois[i]` is [n_i*R, 5] = (index inside feats[i], x1, y1, x2, y2)
    ordered view-major; `obj` is [V*R] objectness logits in the same global row order.  V = sum n_i."""
    feats: List[torch.Tensor]
    rois: List[torch.Tensor]
    obj: torch.Tensor
    R: int

    @property
    def num_views(self) -> int:
        return sum(int(f.size(0)) for f in self.feats)

    def view_boxes(self) -> torch.Tensor:
        """[V, R, 4] proposal boxes per view."""
        return torch.cat([r[:, 1:5] for r in self.rois], 0).reshape(self.num_views, self.R, 4).contiguous()


@dataclass
class TrainOutput:
    losses: Dict[str, torch.Tensor]
    grads: Dict[str, torch.Tensor]          # keyed like HeadOperands.master
    grad_feats: List[torch.Tensor]          # d loss / d feats[i]
    aux: Dict[str, torch.Tensor] = field(default_factory=dict)


class OICRPlusHeadEngine:
    def __init__(self, cfg: HeadConfig, operands: HeadOperands):
        self.cfg = cfg
        self.op = operands
        self.launches_last_step = 0
        self.grad_hook = None
        self.deferred_scale_check = None
        self.last_output: Optional[TrainOutput] = None
        self.fc1_wgrad_panels = 4          # only with a grad_hook (data-parallel): see train_step
        self.operand_gate = None           # callable: makes the current stream wait for in-flight operand updates
        self.fc1_wgrad_position = "first"  # "first" | "middle" | "last" among (dW6, dX, ROI backward); see train_step
        self.bias_on_side_stream = True    # bias-gradient column sums under the GEMMs (only without a grad_hook)
        # engine-level training loops may let the step write the big weight gradients into the same buffers every
        # step (the caller consumes them before the next step); never with autograd, which adopts the buffers as .grad
        self.persistent_grads = False
        self._grad_bufs: Dict[str, torch.Tensor] = {}
        self.external_grad_bufs: Dict[str, torch.Tensor] = {}     # "fc1_w" / "fc2_w" / "head_w" -> caller-owned fp32 buffers
        self._side_stream = None
        # set by every train_step that ran single-GPU with dW6 first: the events behind which EVERY parameter gradient
        # of the step is complete (while the input-gradient GEMM and the ROI backward are still running) and the identity
        # (address, version counter) of each gradient tensor -- solver.B200SGD runs the update under those kernels if
        # `.grad` is still exactly what the step produced
        self.publish_grad_events = True
        self.early_grads: Optional[dict] = None

    def _operands_ready(self):
        """Called between the ROI pooling (which needs no weights) and the first GEMM: an operand all-gather of the
        previous step's sharded update may still be in flight behind the pooling kernels (distributed.GradientExchange);
        then re-cast whatever master parameter changed outside a fused optimizer step."""
        if self.operand_gate is not None:
            self.operand_gate()
        self.op.refresh(force=False)

    def _side(self):
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.op.w6.device)
        return self._side_stream

    # -------------------------------------------------------------------------------------------
    def _pool(self, vb: ViewBatch, keep_argmax: bool):
        cfg = self.cfg
        V, R = vb.num_views, vb.R
        X = torch.empty((V * R, cfg.in_dim), dtype=torch.bfloat16, device=vb.obj.device)
        argmaxes = []
        row = 0
        for f, r in zip(vb.feats, vb.rois):
            m = r.size(0)
            u16 = f.size(2) * f.size(3) < 65535
            # per-call plan: rois grouped by image + bin bounds / scale / backward colouring per roi (7x7 only)
            plan = ops.roi_pool_plan(r, tuple(f.shape), (cfg.pooled, cfg.pooled), cfg.spatial_scale,
                                     row_scale=vb.obj[row:row + m], row_scale_bias=1.0) if u16 else None
            _, am, _ = ops.roi_pool_forward(f, r, (cfg.pooled, cfg.pooled), cfg.spatial_scale,
                                            row_scale=vb.obj[row:row + m], row_scale_bias=1.0, want_f32=False,
                                            argmax_u16=u16, out_bf16=X[row:row + m], plan=plan)
            argmaxes.append((am, plan) if keep_argmax else None)
            row += m
            self.launches_last_step += 1 + (plan is not None)
        return X, argmaxes

    def _trunk(self, X, train: bool, seeds: Tuple[int, int]):
        cfg, op = self.cfg, self.op
        p = cfg.dropout_p if train else 0.0
        H6 = ops.gemm_bf16(X, op.w6, out_dtype=torch.bfloat16, bias=op.master["fc1_b"].detach(), relu=True, dropout_p=p,
                           dropout_seed=seeds[0])
        H7 = ops.gemm_bf16(H6, op.w7, out_dtype=torch.bfloat16, bias=op.master["fc2_b"].detach(), relu=True, dropout_p=p,
                           dropout_seed=seeds[1])
        L = ops.gemm_bf16(H7, op.wh, out_dtype=torch.float32, bias=op.bh)
        self.launches_last_step += 3
        return H6, H7, L

    # -------------------------------------------------------------------------------------------
    def train_step(self, vb: ViewBatch, gt_classes_img: torch.Tensor, dropout_seeds: Tuple[int, int] = (1, 2),
                   need_feat_grad: bool = True, loss_scale: float = 1.0, grad_hook=None,
                   gt_count: Optional[torch.Tensor] = None, gt_onehot: Optional[torch.Tensor] = None) -> TrainOutput:
        """Forward + backward of the head for one image (V views).  gt_classes_img: int [G] sorted ascending
        (get_image_level_gt, roi_heads.py:144-164); or, with `gt_count` / `gt_onehot` from ops.image_level_gt, the
        padded device list [C] whose live length stays on the device (no host synchronisation anywhere in the step).
        Gradients are those of loss_scale * sum(all loss keys).
        grad_hook(key, grad, row0) is called as soon as the kernel producing a parameter gradient is queued -- key is
        "fc1_w" / "fc1_b" / "fc2_w" / "fc2_b" / "head_w" / "head_b" (the fused [cls | det | K x (cls_score, bbox_pred)]
        block), `grad` the gradient or, for fc1_w, a panel of rows starting at row0.  The data-parallel caller
        (distributed.GradientExchange) starts that tensor's collective there, so it overlaps the remaining backward
        kernels (the reference gets the same overlap from DDP buckets, tools/train_net_multi.py:75-78)."""
        cfg, op = self.cfg, self.op
        self.launches_last_step = 0
        C, K, V, R = cfg.num_classes, cfg.refine_k, vb.num_views, vb.R
        dev = vb.obj.device
        gt_int = gt_classes_img.to(device=dev, dtype=torch.int32)
        if gt_count is not None:
            assert gt_onehot is not None and gt_int.numel() == C
            gt_oh = gt_onehot
        else:
            gt_oh = torch.zeros((C,), dtype=torch.float32, device=dev)
            gt_oh[gt_int.long()] = 1.0
        boxes = vb.view_boxes()

        # ---------------- forward ----------------
        X, argmaxes = self._pool(vb, keep_argmax=need_feat_grad)
        self._operands_ready()
        H6, H7, L = self._trunk(X, True, dropout_seeds)
        ldp = cfg.head_cols_padded
        dL = torch.zeros((V * R, ldp), dtype=torch.float32, device=dev)
        scores, img_scores, wloss = ops.wsddn_forward(L, 0, C, V, R, C, gt_oh, dlogits=dL)
        prev = ops.oicr_avg_scores(scores, L, cfg.col_ref0, cfg.ref_stride, V, R, C, K)
        mined = ops.oicr_mine_label(prev, boxes[0], gt_int, C, cfg.top_k(R), cfg.mist_thre, cfg.mist_nms,
                                    cfg.iou_thresholds[0], cfg.iou_thresholds[1], gt_count=gt_count)
        olosses, view_losses, acc = ops.oicr_loss(L, cfg.col_ref0, cfg.ref_stride, boxes, mined["gt_class"],
                                                  mined["gt_weight"], mined["gt_index"], V, R, C, K,
                                                  flip_quirk=cfg.reproduce_flip_quirk and V == 4,
                                                  weights=cfg.bbox_reg_weights, dlogits=dL)
        self.launches_last_step += 6
        losses = {"loss_cls": wloss.mean()}                       # roi_heads_oicrplus.py:283-288
        for k in range(K):
            losses[f"loss_cls_r{k}"] = olosses[k, 0]
            losses[f"loss_box_reg_r{k}"] = olosses[k, 1]

        # ---------------- backward ----------------
        # wsddn_forward wrote d loss_v / d logits; loss_cls is the mean over views -> scale its two blocks by 1/V
        col_scale = torch.full((ldp,), float(loss_scale), dtype=torch.float32, device=dev)
        col_scale[:2 * C] = float(loss_scale) / V
        dLb, _ = ops.cast_f32_bf16(dL, col_scale=col_scale)
        mscale = 1.0 / (1.0 - cfg.dropout_p) if cfg.dropout_p > 0 else 1.0
        # Bias gradients (column sums, HBM-bound, ~110 us per step) ride on a side stream under the tensor-bound GEMMs when
        # nobody needs them before the end of the step (no gradient hook); the main stream joins the side stream at the end.
        side = self._side() if (grad_hook is None and self.bias_on_side_stream) else None

        def grad_buf(name, shape):
            ext = self.external_grad_bufs.get(name)
            if ext is not None:          # e.g. symmetric multicast memory of the nvls gradient exchange
                assert tuple(ext.shape) == tuple(shape) and ext.dtype == torch.float32
                # a fresh tensor object over the same memory: autograd adopts it as `.grad` (a second reference to one
                # tensor object would make it clone 478 MB per step)
                return torch.empty(0, dtype=ext.dtype, device=ext.device).set_(ext.untyped_storage(), ext.storage_offset(),
                                                                              ext.shape, ext.stride())
            if not self.persistent_grads:
                return None
            t = self._grad_bufs.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.device != dev:
                t = torch.empty(shape, dtype=torch.float32, device=dev)
                self._grad_bufs[name] = t
            return t

        def bias_grad(x):
            if side is None:
                return ops.colsum(x)
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                return ops.colsum(x)

        dWh = ops.gemm_bf16(dLb, H7, a_mn=True, b_mn=True, out=grad_buf("head_w", (ldp, cfg.fc_dim)))     # [ldp, fc]
        dbh_raw = bias_grad(dL)
        if grad_hook is not None:
            dbh = dbh_raw * col_scale
            grad_hook("head_w", dWh, 0)
            grad_hook("head_b", dbh, 0)
        dH7 = ops.gemm_bf16(dLb, op.wh, b_mn=True, out_dtype=torch.bfloat16, mask_src=H7, mask_scale=mscale)
        dW7 = ops.gemm_bf16(dH7, H6, a_mn=True, b_mn=True, out=grad_buf("fc2_w", (cfg.fc_dim, cfg.fc_dim)))
        db7 = bias_grad(dH7)
        if grad_hook is not None:
            grad_hook("fc2_w", dW7, 0)
            grad_hook("fc2_b", db7, 0)
        dH6 = ops.gemm_bf16(dH7, op.w7, b_mn=True, out_dtype=torch.bfloat16, mask_src=H6, mask_scale=mscale)
        # fc6: weight gradient (dW6, in row panels with a gradient hook), input gradient (dX) and the ROI backward.  The
        # order only matters data-parallel, where dW6 is 85 % of the exchanged bytes and its collectives share SMs / HBM
        # with whatever runs next; `fc1_wgrad_position` places the panels first / between dX and the ROI backward / last.
        grad_feats: List[torch.Tensor] = []

        def input_gradient():
            return ops.gemm_bf16(dH6, op.w6, b_mn=True, out_dtype=torch.bfloat16)

        def roi_backward(dX):
            row = 0
            for f, r, (am, plan) in zip(vb.feats, vb.rois, argmaxes):
                m = r.size(0)
                grad_feats.append(ops.roi_pool_backward(dX[row:row + m], am, r, tuple(f.shape), (cfg.pooled, cfg.pooled),
                                                        row_scale=vb.obj[row:row + m], row_scale_bias=1.0,
                                                        spatial_scale=cfg.spatial_scale, plan=plan))
                row += m
                self.launches_last_step += 1

        def weight_gradient():
            panels = self.fc1_wgrad_panels if grad_hook is not None else 1
            if panels > 1 and cfg.fc_dim % (128 * panels) == 0:
                # row panels: every panel's collective starts as soon as its GEMM is queued and overlaps what follows.
                # A panel = the same tiles the single launch would compute (bit-identical result).
                dW6 = grad_buf("fc1_w", (cfg.fc_dim, cfg.in_dim))
                if dW6 is None:
                    dW6 = torch.empty((cfg.fc_dim, cfg.in_dim), dtype=torch.float32, device=dev)
                db6 = ops.colsum(dH6)
                grad_hook("fc1_b", db6, 0)
                rows = cfg.fc_dim // panels
                for pi in range(panels):
                    m0 = pi * rows
                    ops.gemm_bf16(dH6[:, m0:m0 + rows], X, a_mn=True, b_mn=True, out=dW6[m0:m0 + rows])
                    grad_hook("fc1_w", dW6[m0:m0 + rows], m0)
                self.launches_last_step += panels - 1
            else:
                dW6 = ops.gemm_bf16(dH6, X, a_mn=True, b_mn=True, out=grad_buf("fc1_w", (cfg.fc_dim, cfg.in_dim)))
                db6 = bias_grad(dH6)
                if grad_hook is not None:
                    grad_hook("fc1_b", db6, 0)
                    grad_hook("fc1_w", dW6, 0)
            return dW6, db6

        pos = self.fc1_wgrad_position if need_feat_grad else "first"
        early_events = None
        self.early_grads = None
        if pos == "first":
            dW6, db6 = weight_gradient()
            if grad_hook is None and need_feat_grad and self.publish_grad_events:
                # every parameter gradient is queued (this stream up to here, bias sums on the side stream)
                if side is not None:
                    with torch.cuda.stream(side):
                        dbh = dbh_raw * col_scale
                        ev_side = torch.cuda.Event()
                        ev_side.record()
                else:
                    dbh = dbh_raw * col_scale
                ev_main = torch.cuda.Event()
                ev_main.record()
                early_events = [ev_main] + ([ev_side] if side is not None else [])
        if need_feat_grad:
            dX = input_gradient()
            self.launches_last_step += 1
            if pos == "middle":
                dW6, db6 = weight_gradient()
            roi_backward(dX)
            del dX
            if pos == "last":
                dW6, db6 = weight_gradient()
        self.launches_last_step += 9
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if grad_hook is None and early_events is None:
            dbh = dbh_raw * col_scale
        grads = {"fc1_w": dW6, "fc1_b": db6, "fc2_w": dW7, "fc2_b": db7}
        for wk, bk, r0, n in op.head_slices():
            grads[wk] = dWh[r0:r0 + n]
            grads[bk] = dbh[r0:r0 + n]
        if early_events is not None:
            self.early_grads = {"events": early_events, "ident": {k: (g.data_ptr(), g._version) for k, g in grads.items()}}
        aux = {"scores": scores, "img_scores": img_scores, "prev": prev, "logits": L, "x": H7, "acc_counts": acc,
               "view_losses": view_losses, "wsddn_view_losses": wloss}
        aux.update(mined)
        return TrainOutput(losses=losses, grads=grads, grad_feats=grad_feats, aux=aux)

    # -------------------------------------------------------------------------------------------
    def test_forward(self, vb: ViewBatch):
        """_forward_box_test for V views at once -> (probs [V,R,C+1], pred_boxes [V,R,4C]) in each view's own
        coordinates (predict_probs_K / predict_boxes_K, fast_rcnn_oicr.py:674-735)."""
        cfg, op = self.cfg, self.op
        self.launches_last_step = 0
        C, K, V, R = cfg.num_classes, cfg.refine_k, vb.num_views, vb.R
        X, _ = self._pool(vb, keep_argmax=False)
        self._operands_ready()
        _, _, L = self._trunk(X, False, (0, 0))
        boxes = vb.view_boxes().reshape(V * R, 4)
        probs, pred_boxes = ops.predict(L, cfg.col_ref0, cfg.ref_stride, boxes, C, K, cfg.bbox_reg_weights)
        self.launches_last_step += 1
        return probs.view(V, R, C + 1), pred_boxes.view(V, R, 4 * C)

    def detect(self, probs: torch.Tensor, pred_boxes: torch.Tensor, image_size: Tuple[int, int], workspace=None):
        """fast_rcnn_inference_single_image (fast_rcnn_oicr.py:86-148) on device."""
        cfg = self.cfg
        self.launches_last_step += 5
        return ops.detect(probs, pred_boxes, image_size, cfg.score_thresh_test, cfg.nms_thresh_test,
                          cfg.detections_per_image, workspace)

    def tta_detect(self, vb: ViewBatch, view_tfms: Sequence[Tuple[float, float, bool, float]], image_size: Tuple[int, int]):
        """GeneralizedRCNNWithTTAAVG._get_augmented_boxes + _merge_detections
        (test_time_augmentation_avg.py:349-387): per-view inference, inverse transform, mean over views, one NMS.
        view_tfms[v] = (scale_x, scale_y, flipped, view_width) mapping view v back to the original image."""
        cfg = self.cfg
        probs, pboxes = self.test_forward(vb)
        V, R, C = vb.num_views, vb.R, cfg.num_classes
        acc_b = torch.empty((R, 4 * C), dtype=torch.float32, device=probs.device)
        acc_p = torch.empty((R, C + 1), dtype=torch.float32, device=probs.device)
        for v, (sx, sy, fl, vw) in enumerate(view_tfms):
            ops.tta_accumulate(pboxes[v], probs[v], sx, sy, fl, vw, v == 0, float(V) if v == V - 1 else 0.0, acc_b, acc_p)
            self.launches_last_step += 1
        return self.detect(acc_p, acc_b, image_size), acc_p, acc_b
