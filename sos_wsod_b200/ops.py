"""Tensor-level wrappers over the C ABI (include/soswsod_b200.h).  Every function takes CUDA tensors, allocates
outputs/workspaces with torch (so the caching allocator and streams keep working), passes raw device pointers
and the current stream, and raises RuntimeError on any failure.  No function here has a CPU path."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ARGMAX_I32, ARGMAX_U16, DTYPE_BF16, DTYPE_F32, check


# Number of kernels of THIS library launched so far (bench.py reports the delta over its timed region).
COUNTERS = {"launches": 0}


def _count(n: int) -> None:
    COUNTERS["launches"] += n


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sos_wsod_b200 ops run on CUDA tensors only (sm_100a); there is no CPU fallback")


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise RuntimeError(f"unsupported dtype {t.dtype}")


# ------------------------------------------------------------------------------------------------
# (1) ROI max-pool
# ------------------------------------------------------------------------------------------------
def roi_pool_plan(rois: torch.Tensor, feat_shape, pooled: Tuple[int, int] = (7, 7), spatial_scale: float = 0.125,
                  row_scale: Optional[torch.Tensor] = None, row_scale_bias: float = 0.0) -> Optional[torch.Tensor]:
    """The per-call plan of the fast 7x7 kernels (rois grouped by image + every roi's bin bounds, scale factor and
    backward colouring), or None when the shape has no plan.  Pass it to roi_pool_forward / roi_pool_backward called
    with the SAME rois, feature shape, spatial_scale and row_scale."""
    _need_cuda(rois, row_scale)
    n, c, h, w = feat_shape
    m = rois.size(0)
    lib = _lib.load()
    nbytes = lib.soswsod_roi_pool_plan_bytes(m, pooled[0], pooled[1])
    if nbytes == 0 or n > 64 or h >= 65536 or w >= 65536:
        return None
    rois = rois.contiguous()
    if row_scale is not None:
        row_scale = row_scale.contiguous().float()
    plan = torch.empty((nbytes + 127) // 128 * 128, dtype=torch.uint8, device=rois.device)
    check(lib.soswsod_roi_pool_plan(_ptr(rois), m, n, h, w, pooled[0], pooled[1], float(spatial_scale), _ptr(row_scale),
                                    float(row_scale_bias), _ptr(plan), plan.numel(), _stream()), "roi_pool_plan")
    _count(1)
    return plan


def roi_pool_forward(feat: torch.Tensor, rois: torch.Tensor, pooled: Tuple[int, int] = (7, 7),
                     spatial_scale: float = 0.125, row_scale: Optional[torch.Tensor] = None,
                     row_scale_bias: float = 0.0, want_f32: bool = True, want_bf16: bool = False,
                     argmax_u16: bool = False, out_bf16: Optional[torch.Tensor] = None,
                     plan: Optional[torch.Tensor] = None):
    """feat fp32 [N,C,H,W], rois fp32 [M,5] -> (out_f32 [M,C,ph,pw] | None, argmax [M,C,ph,pw], out_bf16 [M, C*ph*pw] | None).
    out_bf16 = pooled * (row_scale + row_scale_bias), the fc6 GEMM operand.  `plan` (roi_pool_plan) selects the fast
    7x7 operand-mode kernel; results are identical without it."""
    _need_cuda(feat, rois, row_scale, plan)
    assert feat.dtype == torch.float32 and rois.dtype == torch.float32 and rois.dim() == 2 and rois.size(1) == 5
    feat = feat.contiguous()
    rois = rois.contiguous()
    n, c, h, w = feat.shape
    m = rois.size(0)
    ph, pw = pooled
    out = torch.empty((m, c, ph, pw), dtype=torch.float32, device=feat.device) if want_f32 else None
    if argmax_u16:
        argmax = torch.empty((m, c, ph, pw), dtype=torch.int16, device=feat.device)  # bits are uint16
    else:
        argmax = torch.empty((m, c, ph, pw), dtype=torch.int32, device=feat.device)
    obf = None
    ld = 0
    if want_bf16 or out_bf16 is not None:
        obf = out_bf16 if out_bf16 is not None else torch.empty((m, c * ph * pw), dtype=torch.bfloat16, device=feat.device)
        assert obf.dtype == torch.bfloat16 and obf.stride(-1) == 1
        ld = obf.stride(0)
    if row_scale is not None:
        row_scale = row_scale.contiguous().float()
    lib = _lib.load()
    check(lib.soswsod_roi_pool_forward(_ptr(feat), n, c, h, w, _ptr(rois), m, ph, pw, float(spatial_scale),
                                       _ptr(row_scale), float(row_scale_bias), _ptr(out), _ptr(argmax),
                                       ARGMAX_U16 if argmax_u16 else ARGMAX_I32, _ptr(obf), ld, _ptr(plan),
                                       0 if plan is None else plan.numel(), _stream()),
          "roi_pool_forward")
    _count(1)
    return out, argmax, obf


def roi_pool_backward(grad_out: torch.Tensor, argmax: torch.Tensor, rois: torch.Tensor, feat_shape,
                      pooled: Tuple[int, int] = (7, 7), row_scale: Optional[torch.Tensor] = None,
                      row_scale_bias: float = 0.0, spatial_scale: float = 0.125,
                      plan: Optional[torch.Tensor] = None) -> torch.Tensor:
    """grad_out [M, C*ph*pw] (or [M,C,ph,pw]) fp32/bf16 -> grad_feat fp32 [N,C,H,W] (overwritten, atomic-free).
    `argmax` must come from roi_pool_forward on the same rois and spatial_scale; `plan` as in roi_pool_forward."""
    _need_cuda(grad_out, argmax, rois, plan)
    n, c, h, w = feat_shape
    ph, pw = pooled
    m = rois.size(0)
    if grad_out.dim() == 4:
        grad_out = grad_out.reshape(m, -1)
    if grad_out.stride(-1) != 1:
        grad_out = grad_out.contiguous()
    rois = rois.contiguous()
    if m == 0:
        return torch.zeros((n, c, h, w), dtype=torch.float32, device=grad_out.device)
    argmax = argmax.contiguous()
    # The kernels stream grad / arg-max rows with TMA: 16-byte aligned rows.  Odd channel counts are padded with empty
    # channels (arg-max -1, zero gradient) instead of falling back to atomics -- there is no atomic path.
    row_a = c * ph * pw * argmax.element_size()
    row_g = grad_out.stride(0) * grad_out.element_size()
    if row_a % 16 or row_g % 16 or grad_out.data_ptr() % 16:
        cp = (c + 7) // 8 * 8
        g2 = torch.zeros((m, cp * ph * pw), dtype=grad_out.dtype, device=grad_out.device)
        g2[:, :c * ph * pw] = grad_out[:, :c * ph * pw]
        a2 = torch.full((m, cp, ph, pw), -1, dtype=argmax.dtype, device=argmax.device)
        a2[:, :c] = argmax.view(m, c, ph, pw)
        out = roi_pool_backward(g2, a2, rois, (n, cp, h, w), pooled, row_scale, row_scale_bias, spatial_scale, plan)
        return out[:, :c].contiguous()
    grad_feat = torch.empty((n, c, h, w), dtype=torch.float32, device=grad_out.device)
    a_dt = ARGMAX_U16 if argmax.dtype == torch.int16 else ARGMAX_I32
    if row_scale is not None:
        row_scale = row_scale.contiguous().float()
    lib = _lib.load()
    check(lib.soswsod_roi_pool_backward(_ptr(grad_out), _dt(grad_out), grad_out.stride(0), _ptr(argmax), a_dt,
                                        _ptr(rois), m, _ptr(row_scale), float(row_scale_bias), n, c, h, w, ph, pw,
                                        float(spatial_scale), _ptr(grad_feat), _ptr(plan),
                                        0 if plan is None else plan.numel(), _stream()), "roi_pool_backward")
    _count(1)
    return grad_feat


# ------------------------------------------------------------------------------------------------
# (2) GEMM + helpers
# ------------------------------------------------------------------------------------------------
_GEMM_SCHED = {}
GEMM_SCHED_SLOTS = 1     # > 1: launches on a stream rotate through that many scheduler workspaces (diagnosis only)


def _gemm_sched(device) -> torch.Tensor:
    """The GEMM's dynamic tile-scheduler workspace (8 bytes, zeroed once, re-armed by the kernel) of the current
    stream: launches on one stream are ordered and share it, other streams get their own."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _GEMM_SCHED.get(key)
    if t is None or t[0].size(0) != GEMM_SCHED_SLOTS:
        t = [torch.zeros((GEMM_SCHED_SLOTS, 4), dtype=torch.int32, device=device), 0]
        _GEMM_SCHED[key] = t
    t[1] = (t[1] + 1) % GEMM_SCHED_SLOTS
    return t[0][t[1]]


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
              out_dtype: torch.dtype = torch.float32, bias: Optional[torch.Tensor] = None, relu: bool = False,
              mask_src: Optional[torch.Tensor] = None, mask_scale: float = 1.0, dropout_p: float = 0.0,
              dropout_seed: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D[M,N] = epilogue(sum_k A[m,k] B[n,k]).  a: [M,K] (a_mn False) or [K,M] (a_mn True); b likewise with N."""
    _need_cuda(a, b, bias, mask_src, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        k, m = a.shape
    else:
        m, k = a.shape
    if b_mn:
        kb, n = b.shape
    else:
        n, kb = b.shape
    assert k == kb, f"K mismatch {k} vs {kb}"
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    assert out.shape == (m, n) and out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
        bias = bias.contiguous()
    ld_mask = 0
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == (m, n) and mask_src.stride(1) == 1
        ld_mask = mask_src.stride(0)
    lib = _lib.load()
    check(lib.soswsod_gemm_bf16(_ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), _ptr(out),
                                out.stride(0), _dt(out), m, n, k, _ptr(bias), int(relu), _ptr(mask_src), ld_mask,
                                float(mask_scale), float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF,
                                _ptr(_gemm_sched(a.device)), _stream()),
          "gemm_bf16")
    _count(1)
    return out


def dropout_mask(m: int, n: int, p: float, seed: int, device="cuda") -> torch.Tensor:
    mask = torch.empty((m, n), dtype=torch.uint8, device=device)
    check(_lib.load().soswsod_dropout_mask(_ptr(mask), m, n, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
          "dropout_mask")
    _count(1)
    return mask


def cast_f32_bf16(x: torch.Tensor, col_scale: Optional[torch.Tensor] = None, want_out: bool = True,
                  want_t: bool = False, out: Optional[torch.Tensor] = None, ld_pad: int = 8,
                  mask_src: Optional[torch.Tensor] = None, mask_scale: float = 1.0):
    """x fp32 [rows, cols] -> (bf16 [rows, cols] | None, bf16 transposed [cols, rows_padded][:, :rows] | None).
    mask_src (bf16 [rows, cols]): the result is multiplied by mask_scale where mask_src > 0, by 0 elsewhere."""
    _need_cuda(x, col_scale, mask_src)
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == x.shape and mask_src.stride(1) == 1
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    o = None
    if out is not None:
        o = out
    elif want_out:
        o = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    ot = None
    if want_t:
        rp = (rows + ld_pad - 1) // ld_pad * ld_pad
        ot = torch.zeros((cols, rp), dtype=torch.bfloat16, device=x.device)[:, :rows]
    check(_lib.load().soswsod_cast_f32_bf16(_ptr(x), x.stride(0), rows, cols, _ptr(col_scale), _ptr(o),
                                            0 if o is None else o.stride(0), _ptr(ot), 0 if ot is None else ot.stride(0),
                                            _ptr(mask_src), 0 if mask_src is None else mask_src.stride(0), float(mask_scale),
                                            _stream()), "cast_f32_bf16")
    _count(1)
    return o, ot


def transpose_bf16(x: torch.Tensor, ld_pad: int = 8) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    rp = (rows + ld_pad - 1) // ld_pad * ld_pad
    ot = torch.zeros((cols, rp), dtype=torch.bfloat16, device=x.device)[:, :rows]
    check(_lib.load().soswsod_transpose_bf16(_ptr(x), x.stride(0), rows, cols, _ptr(ot), ot.stride(0), _stream()),
          "transpose_bf16")
    _count(1)
    return ot


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = torch.empty((cols,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    nbytes = lib.soswsod_colsum_workspace_bytes(rows, cols, _dt(x))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device) if nbytes else None
    check(lib.soswsod_colsum(_ptr(x), _dt(x), x.stride(0), rows, cols, _ptr(out), _ptr(ws), nbytes, _stream()), "colsum")
    _count(2 if nbytes else 1)
    return out


def sgd_step(param: torch.Tensor, grad: torch.Tensor, buf: torch.Tensor, lr: float, momentum: float,
             weight_decay: float, grad_scale: float = 1.0, param_bf16: Optional[torch.Tensor] = None) -> None:
    _need_cuda(param, grad, buf, param_bf16)
    assert param.is_contiguous() and grad.is_contiguous() and buf.is_contiguous()
    assert param.dtype == grad.dtype == buf.dtype == torch.float32
    assert param_bf16 is None or (param_bf16.is_contiguous() and param_bf16.dtype == torch.bfloat16)
    check(_lib.load().soswsod_sgd_step(_ptr(param), _ptr(grad), _ptr(buf), param.numel(), float(lr), float(momentum),
                                       float(weight_decay), float(grad_scale), _ptr(param_bf16), _stream()), "sgd_step")
    _count(1)


def sgd_multi(items, momentum: float, grad_scale: float = 1.0) -> int:
    """One fused SGD(+momentum, weight decay) launch per 32 tensors.  items: iterable of
    (param, grad, momentum_buf, lr, weight_decay, out_bf16 | None, out_f32 | None); every tensor contiguous, fp32
    (out_bf16 bf16), same numel.  Shards are passed as views.  Returns the number of launches."""
    import ctypes

    items = list(items)
    lib = _lib.load()
    n_launch = 0
    for i0 in range(0, len(items), _lib.SGD_MAX_TENSORS):
        chunk = items[i0:i0 + _lib.SGD_MAX_TENSORS]
        arr = (_lib.SgdTensor * len(chunk))()
        for d, (p, g, buf, lr, wd, ob, of) in zip(arr, chunk):
            _need_cuda(p, g, buf, ob, of)
            n = p.numel()
            if not (p.is_contiguous() and g.is_contiguous() and buf.is_contiguous() and g.numel() == n and buf.numel() == n
                    and p.dtype == g.dtype == buf.dtype == torch.float32):
                raise RuntimeError("sgd_multi: param / grad / momentum_buf must be contiguous fp32 tensors of one size")
            if ob is not None and not (ob.is_contiguous() and ob.dtype == torch.bfloat16 and ob.numel() == n):
                raise RuntimeError("sgd_multi: out_bf16 must be a contiguous bf16 tensor of the parameter's size")
            if of is not None and not (of.is_contiguous() and of.dtype == torch.float32 and of.numel() == n):
                raise RuntimeError("sgd_multi: out_f32 must be a contiguous fp32 tensor of the parameter's size")
            d.param, d.grad, d.momentum_buf = p.data_ptr(), g.data_ptr(), buf.data_ptr()
            d.out_bf16, d.out_f32 = _ptr(ob), _ptr(of)
            d.n, d.lr, d.weight_decay = n, float(lr), float(wd)
        check(lib.soswsod_sgd_multi(ctypes.cast(arr, ctypes.c_void_p), len(chunk), float(momentum), float(grad_scale),
                                    _stream()), "sgd_multi")
        n_launch += 1
    _count(n_launch)
    return n_launch


NVLS_MAX_CTAS = 0      # persistent grid of the fused NVLS update; 0 = the library default (1 per SM)


def sgd_nvls(items, momentum: float, grad_scale: float) -> int:
    """The fused NVLS update (soswsod_sgd_nvls).  items: iterable of (param_rows, grad_mc_address, momentum_rows,
    operand_mc_address, lr, weight_decay) -- param / momentum rows are contiguous fp32 CUDA tensors (the rows this rank
    owns), the two addresses are multicast virtual addresses (int) of the same rows in the symmetric gradient / bf16
    operand buffers."""
    import ctypes

    items = list(items)
    if not items:
        return 0
    if len(items) > _lib.SGD_NVLS_MAX_TENSORS:
        raise RuntimeError(f"sgd_nvls: at most {_lib.SGD_NVLS_MAX_TENSORS} tensors per call")
    arr = (_lib.SgdNvlsTensor * len(items))()
    for d, (p, g_mc, buf, ob_mc, lr, wd) in zip(arr, items):
        _need_cuda(p, buf)
        if not (p.is_contiguous() and buf.is_contiguous() and p.dtype == buf.dtype == torch.float32 and p.numel() == buf.numel()):
            raise RuntimeError("sgd_nvls: param / momentum rows must be contiguous fp32 tensors of one size")
        d.param, d.grad_mc, d.momentum_buf, d.out_bf16_mc = p.data_ptr(), int(g_mc), buf.data_ptr(), int(ob_mc)
        d.n, d.lr, d.weight_decay = p.numel(), float(lr), float(wd)
    check(_lib.load().soswsod_sgd_nvls(ctypes.cast(arr, ctypes.c_void_p), len(items), float(momentum), float(grad_scale),
                                       int(NVLS_MAX_CTAS), _stream()), "sgd_nvls")
    _count(1)
    return 1


# ------------------------------------------------------------------------------------------------
# (3) WSDDN
# ------------------------------------------------------------------------------------------------
def wsddn_forward(logits: torch.Tensor, col_cls: int, col_det: int, num_views: int, R: int, C: int,
                  gt_onehot: torch.Tensor, dlogits: Optional[torch.Tensor] = None):
    """logits fp32 [V*R, ld] -> (scores [V,R,C], img_scores [V,C], loss [V]); if dlogits is given the unit-upstream
    gradient of each view's loss is written into its cls/det column blocks."""
    _need_cuda(logits, gt_onehot, dlogits)
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and logits.size(0) == num_views * R
    dev = logits.device
    scores = torch.empty((num_views, R, C), dtype=torch.float32, device=dev)
    img = torch.empty((num_views, C), dtype=torch.float32, device=dev)
    loss = torch.empty((num_views,), dtype=torch.float32, device=dev)
    gt = gt_onehot.reshape(-1).contiguous().float()
    assert gt.numel() == C
    check(_lib.load().soswsod_wsddn_forward(_ptr(logits), logits.stride(0), col_cls, col_det, num_views, R, C, _ptr(gt),
                                            _ptr(scores), _ptr(img), _ptr(loss), _ptr(dlogits),
                                            0 if dlogits is None else dlogits.stride(0), _stream()), "wsddn_forward")
    _count(1)
    return scores, img, loss


# ------------------------------------------------------------------------------------------------
# (4) OICR
# ------------------------------------------------------------------------------------------------
def oicr_avg_scores(wsddn_scores: torch.Tensor, logits: Optional[torch.Tensor], col_ref0: int, ref_stride: int,
                    num_views: int, R: int, C: int, K: int) -> torch.Tensor:
    _need_cuda(wsddn_scores, logits)
    prev = torch.empty((K, R, C + 1), dtype=torch.float32, device=wsddn_scores.device)
    check(_lib.load().soswsod_oicr_avg_scores(_ptr(wsddn_scores.contiguous()), _ptr(logits),
                                              0 if logits is None else logits.stride(0), col_ref0, ref_stride,
                                              num_views, R, C, K, _ptr(prev), _stream()), "oicr_avg_scores")
    _count(1)
    return prev


def image_level_gt(gt_classes: torch.Tensor, C: int):
    """get_image_level_gt on the device, without torch.unique's host read-back: gt_classes int64/int32 [n] ->
    (gt_list int32 [C] ascending distinct classes padded with -1, gt_count int32 [1], gt_onehot fp32 [C])."""
    _need_cuda(gt_classes)
    assert gt_classes.dtype in (torch.int64, torch.int32) and gt_classes.dim() == 1
    gt_classes = gt_classes.contiguous()
    dev = gt_classes.device
    lst = torch.empty((C,), dtype=torch.int32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    oh = torch.empty((C,), dtype=torch.float32, device=dev)
    check(_lib.load().soswsod_image_level_gt(_ptr(gt_classes), int(gt_classes.dtype == torch.int64), gt_classes.numel(), C,
                                             _ptr(lst), _ptr(cnt), _ptr(oh), _stream()), "image_level_gt")
    _count(1)
    return lst, cnt, oh


def oicr_mine_label(prev: torch.Tensor, boxes: torch.Tensor, gt_classes: torch.Tensor, C: int, top_k: int,
                    score_thr: float = 0.05, nms_thr: float = 0.01, iou_lo: float = 0.5, iou_hi: float = 0.6,
                    gt_count: Optional[torch.Tensor] = None):
    """prev fp32 [K,R,ld]; boxes fp32 [R,4]; gt_classes int32 [G] ascending (with gt_count, a device int32 scalar, only
    its first gt_count entries are live and G is the capacity).  Returns a dict of device tensors."""
    _need_cuda(prev, boxes, gt_classes, gt_count)
    assert gt_count is None or (gt_count.dtype == torch.int32 and gt_count.numel() == 1)
    assert prev.dtype == torch.float32 and prev.dim() == 3 and prev.is_contiguous()
    K, R, ld = prev.shape
    G = gt_classes.numel()
    gt_classes = gt_classes.to(torch.int32).contiguous()
    boxes = boxes.contiguous()
    dev = prev.device
    ms = top_k * G
    out = {
        "seed_count": torch.zeros((K,), dtype=torch.int32, device=dev),
        "seed_index": torch.zeros((K, ms), dtype=torch.int32, device=dev),
        "seed_class": torch.zeros((K, ms), dtype=torch.int32, device=dev),
        "seed_score": torch.zeros((K, ms), dtype=torch.float32, device=dev),
        "gt_class": torch.empty((K, R), dtype=torch.int32, device=dev),
        "gt_weight": torch.empty((K, R), dtype=torch.float32, device=dev),
        "gt_index": torch.empty((K, R), dtype=torch.int32, device=dev),
        "counts": torch.empty((K, 3), dtype=torch.int32, device=dev),
    }
    lib = _lib.load()
    wsb = lib.soswsod_oicr_mine_workspace_bytes(top_k, G, K)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    check(lib.soswsod_oicr_mine_label(_ptr(prev), ld, _ptr(boxes), _ptr(gt_classes), G, _ptr(gt_count), R, C, K, top_k, float(score_thr),
                                      float(nms_thr), float(iou_lo), float(iou_hi), _ptr(out["seed_count"]),
                                      _ptr(out["seed_index"]), _ptr(out["seed_class"]), _ptr(out["seed_score"]),
                                      _ptr(out["gt_class"]), _ptr(out["gt_weight"]), _ptr(out["gt_index"]),
                                      _ptr(out["counts"]), _ptr(ws), wsb, _stream()), "oicr_mine_label")
    _count(2)
    return out


def oicr_loss(logits: torch.Tensor, col_ref0: int, ref_stride: int, boxes: torch.Tensor, gt_class: torch.Tensor,
              gt_weight: torch.Tensor, gt_index: torch.Tensor, num_views: int, R: int, C: int, K: int,
              flip_quirk: bool = True, weights=(10.0, 10.0, 5.0, 5.0), dlogits: Optional[torch.Tensor] = None):
    """Returns (losses [K,2], view_losses [K,V,2], acc_counts [K,V,5])."""
    _need_cuda(logits, boxes, gt_class, gt_weight, gt_index, dlogits)
    dev = logits.device
    losses = torch.empty((K, 2), dtype=torch.float32, device=dev)
    view_losses = torch.zeros((K, num_views, 2), dtype=torch.float32, device=dev)
    acc = torch.zeros((K, num_views, 5), dtype=torch.int32, device=dev)
    boxes = boxes.contiguous()
    assert boxes.shape == (num_views, R, 4)
    check(_lib.load().soswsod_oicr_loss(_ptr(logits), logits.stride(0), col_ref0, ref_stride, _ptr(boxes),
                                        _ptr(gt_class), _ptr(gt_weight), _ptr(gt_index), num_views, R, C, K,
                                        int(flip_quirk), *[float(x) for x in weights], _ptr(losses), _ptr(view_losses),
                                        _ptr(acc), _ptr(dlogits), 0 if dlogits is None else dlogits.stride(0),
                                        _stream()), "oicr_loss")
    _count(2)
    return losses, view_losses, acc


# ------------------------------------------------------------------------------------------------
# (5) test-time
# ------------------------------------------------------------------------------------------------
def predict(logits: torch.Tensor, col_ref0: int, ref_stride: int, boxes: torch.Tensor, C: int, K: int,
            weights=(10.0, 10.0, 5.0, 5.0)):
    _need_cuda(logits, boxes)
    R = boxes.size(0)
    probs = torch.empty((R, C + 1), dtype=torch.float32, device=logits.device)
    pred_boxes = torch.empty((R, 4 * C), dtype=torch.float32, device=logits.device)
    check(_lib.load().soswsod_predict(_ptr(logits), logits.stride(0), col_ref0, ref_stride, _ptr(boxes.contiguous()), R, C,
                                      K, *[float(x) for x in weights], _ptr(probs), _ptr(pred_boxes), _stream()), "predict")
    _count(1)
    return probs, pred_boxes


def tta_accumulate(pred_boxes: torch.Tensor, probs: torch.Tensor, scale_x: float, scale_y: float, flipped: bool,
                   view_w: float, first: bool, finalize_div: float, acc_boxes: torch.Tensor, acc_probs: torch.Tensor):
    _need_cuda(pred_boxes, probs, acc_boxes, acc_probs)
    R = probs.size(0)
    C = probs.size(1) - 1
    check(_lib.load().soswsod_tta_accumulate(_ptr(pred_boxes.contiguous()), _ptr(probs.contiguous()), R, C, float(scale_x),
                                             float(scale_y), int(flipped), float(view_w), int(first), float(finalize_div),
                                             _ptr(acc_boxes), _ptr(acc_probs), _stream()), "tta_accumulate")
    _count(1)


TTA_VIEW_PARAMS = 10   # SOSWSOD_TTA_VIEW_PARAMS


def _view_params_host(view_params) -> "ctypes.Array":
    import ctypes

    flat = [float(x) for row in view_params for x in row]
    if len(flat) % TTA_VIEW_PARAMS != 0:
        raise RuntimeError("view_params: 10 floats per view (include/soswsod_b200.h, SOSWSOD_TTA_VIEW_PARAMS)")
    return (ctypes.c_float * len(flat))(*flat), len(flat) // TTA_VIEW_PARAMS


def tta_views(boxes: torch.Tensor, view_params, min_box_size: float = 0.0):
    """DatasetMapperTTAAVG's proposal path for all V views of an image in one launch.  boxes fp32 [R,4] in
    original-image coordinates; view_params: V rows of 10 floats (see the header).  Returns (rois [V*R,5] view-major,
    keep bool [V,R], dropped int32 [V]) -- device tensors, no host synchronisation."""
    _need_cuda(boxes)
    import ctypes

    boxes = boxes.contiguous().float()
    R = boxes.size(0)
    arr, V = _view_params_host(view_params)
    rois = torch.empty((V * R, 5), dtype=torch.float32, device=boxes.device)
    keep = torch.empty((V, R), dtype=torch.uint8, device=boxes.device)
    dropped = torch.empty((V,), dtype=torch.int32, device=boxes.device)
    check(_lib.load().soswsod_tta_views(_ptr(boxes), R, ctypes.cast(arr, ctypes.c_void_p), V, float(min_box_size),
                                        _ptr(rois), _ptr(keep), _ptr(dropped), _stream()), "tta_views")
    _count(1)
    return rois, keep.bool(), dropped


def tta_merge(pred_boxes: torch.Tensor, probs: torch.Tensor, view_params):
    """GeneralizedRCNNWithTTAAVG._get_augmented_boxes for all views in one launch: pred_boxes [V,R,4C] and probs
    [V,R,C+1] in each view's coordinates -> (mean boxes [R,4C] in original-image coordinates, mean probs [R,C+1])."""
    _need_cuda(pred_boxes, probs)
    import ctypes

    V, R, C1 = probs.shape
    C = C1 - 1
    arr, nv = _view_params_host(view_params)
    if nv != V or tuple(pred_boxes.shape) != (V, R, 4 * C):
        raise RuntimeError(f"tta_merge: {nv} view rows / boxes {tuple(pred_boxes.shape)} for probs {tuple(probs.shape)}")
    pred_boxes = pred_boxes.contiguous()
    probs = probs.contiguous()
    mb = torch.empty((R, 4 * C), dtype=torch.float32, device=probs.device)
    mp = torch.empty((R, C1), dtype=torch.float32, device=probs.device)
    check(_lib.load().soswsod_tta_merge(_ptr(pred_boxes), _ptr(probs), V, R, C, ctypes.cast(arr, ctypes.c_void_p), _ptr(mb),
                                        _ptr(mp), _stream()), "tta_merge")
    _count(1)
    return mb, mp


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thr: float) -> torch.Tensor:
    """torchvision.ops.nms drop-in: int64 indices of kept boxes, score-descending.  (One D2H read of the count.)"""
    _need_cuda(boxes, scores)
    n = boxes.size(0)
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    lib = _lib.load()
    keep = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    num = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    wsb = lib.soswsod_nms_workspace_bytes(n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=boxes.device)
    check(lib.soswsod_nms(_ptr(boxes), _ptr(scores), n, float(iou_thr), _ptr(keep), _ptr(num), _ptr(ws), wsb, _stream()),
          "nms")
    _count(4)
    return keep[: int(num.item())]


def detect(probs: torch.Tensor, pred_boxes: torch.Tensor, image_size: Tuple[float, float], score_thr: float,
           nms_thr: float, topk: int, workspace: Optional[torch.Tensor] = None):
    """fast_rcnn_inference_single_image on device.  Returns (det_boxes [topk,4], det_scores [topk], det_classes [topk],
    det_rows [topk], num_det [1]) -- all device tensors, no host synchronisation."""
    _need_cuda(probs, pred_boxes)
    R, C1 = probs.shape
    C = C1 - 1
    dev = probs.device
    lib = _lib.load()
    wsb = lib.soswsod_detect_workspace_bytes(R, C)
    if workspace is None or workspace.numel() < wsb:
        workspace = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    db = torch.zeros((topk, 4), dtype=torch.float32, device=dev)
    ds = torch.zeros((topk,), dtype=torch.float32, device=dev)
    dc = torch.zeros((topk,), dtype=torch.int32, device=dev)
    dr = torch.zeros((topk,), dtype=torch.int32, device=dev)
    nd = torch.zeros((1,), dtype=torch.int32, device=dev)
    check(lib.soswsod_detect(_ptr(probs.contiguous()), _ptr(pred_boxes.contiguous()), R, C, float(image_size[0]),
                             float(image_size[1]), float(score_thr), float(nms_thr), int(topk), _ptr(db), _ptr(ds),
                             _ptr(dc), _ptr(dr), _ptr(nd), _ptr(workspace), workspace.numel(), _stream()), "detect")
    _count(5)
    return db, ds, dc, dr, nd


# ------------------------------------------------------------------------------------------------
# (6) PGF (tools/pgf.py) on the device
# ------------------------------------------------------------------------------------------------
def pgf_filter(boxes_xywh: torch.Tensor, scores: torch.Tensor, cats: torch.Tensor, img_offsets: torch.Tensor,
               t_con: float, t_keep: float, use_diff: bool, diff_classes) -> torch.Tensor:
    """boxes float64 [n,4] XYWH, scores float64 [n], cats int32 [n], img_offsets int32 [I+1] -> keep bool [n]."""
    _need_cuda(boxes_xywh, scores, cats, img_offsets)
    assert boxes_xywh.dtype == torch.float64 and scores.dtype == torch.float64
    assert cats.dtype == torch.int32 and img_offsets.dtype == torch.int32
    boxes_xywh, scores, cats, img_offsets = (t.contiguous() for t in (boxes_xywh, scores, cats, img_offsets))
    n = scores.numel()
    keep = torch.empty((n,), dtype=torch.uint8, device=scores.device)
    lo = hi = 0
    for c in diff_classes:
        if not 0 <= int(c) < 128:
            raise RuntimeError("pgf_filter: diff classes must lie in [0, 128)")
        if c < 64:
            lo |= 1 << int(c)
        else:
            hi |= 1 << (int(c) - 64)
    check(_lib.load().soswsod_pgf(_ptr(boxes_xywh), _ptr(scores), _ptr(cats), _ptr(img_offsets), img_offsets.numel() - 1,
                                  float(t_con), float(t_keep), int(use_diff), lo, hi, _ptr(keep), _stream()), "pgf")
    _count(1)
    return keep.bool()
