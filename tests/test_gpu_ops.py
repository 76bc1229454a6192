"""GPU parity tests of every C-ABI entry point against the CPU oracle (oracle/oicr_plus_ref.py, which calls the
same torchvision/torch operators the reference calls).  Bars (BASELINE.json north_star):
  * bit-exact: pooled maxima, argmax indices, pseudo-GT seeds, labels / matched indices, NMS keep-lists
  * GEMM-derived values: <= 1e-2 relative; losses within 1e-3
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import oicr_plus_ref as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from sos_wsod_b200 import ops as _ops

    return _ops


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------
# (1) ROI pool
# ------------------------------------------------------------------------------------------------
def _rois_with_edge_cases(R, img_h, img_w, g):
    boxes = ref.synth_boxes(R, img_h, img_w, g)
    extra = torch.tensor([
        [0.0, 0.0, img_w - 1.0, img_h - 1.0],      # whole image
        [4.0, 4.0, 4.0, 4.0],                      # single point -> .5 rounding (4/8 = 0.5)
        [12.0, 20.0, 28.0, 36.0],                  # x.5 boundaries on both ends
        [100.0, 50.0, 60.0, 30.0],                 # malformed: x2 < x1, y2 < y1
        [-50.0, -40.0, 30.0, 30.0],                # partly outside (negative)
        [img_w + 100.0, img_h + 100.0, img_w + 200.0, img_h + 300.0],  # fully outside -> empty bins
        [img_w - 20.0, img_h - 20.0, img_w + 60.0, img_h + 60.0],      # sticks out at the far corner
        [3.7, 9.2, 200.3, 150.9],                  # non-integer coords
    ])
    return torch.cat([boxes, extra], 0)


@pytest.mark.parametrize("C,h,w,R", [(512, 60, 80, 300), (64, 37, 53, 257), (8, 96, 152, 64), (3, 20, 20, 33)])
def test_roi_pool_forward_bit_exact(ops, C, h, w, R):
    g = _gen(10 + C)
    feat = torch.relu(torch.randn((1, C, h, w), generator=g))
    boxes = _rois_with_edge_cases(R, h * 8, w * 8, g)
    rois = ref.boxes_to_pooler_format([boxes])
    exp_out, exp_arg = ref.roi_pool(feat, rois)
    obj = torch.rand(rois.size(0), generator=g)
    out, arg, obf = ops.roi_pool_forward(feat.cuda(), rois.cuda(), row_scale=obj.cuda(), row_scale_bias=1.0,
                                         want_bf16=True)
    assert torch.equal(out.cpu(), exp_out), "pooled maxima must be bit-exact"
    assert torch.equal(arg.cpu(), exp_arg), "argmax indices must be bit-exact"
    exp_bf = (exp_out * (obj + 1).view(-1, 1, 1, 1)).flatten(1).to(torch.bfloat16)
    assert torch.equal(obf.cpu(), exp_bf)
    if h * w < 65535:
        _, arg16, _ = ops.roi_pool_forward(feat.cuda(), rois.cuda(), want_f32=False, argmax_u16=True)
        a = arg16.cpu().to(torch.int32) & 0xFFFF
        a[a == 0xFFFF] = -1
        assert torch.equal(a, exp_arg)


def test_roi_pool_forward_batch_and_negative_features(ops):
    g = _gen(3)
    feat = torch.randn((2, 16, 30, 40), generator=g)  # negative values: an all-negative bin must still return its max
    b0 = ref.synth_boxes(40, 240, 320, g)
    b1 = ref.synth_boxes(50, 240, 320, g)
    rois = ref.boxes_to_pooler_format([b0, b1])
    perm = torch.randperm(rois.size(0), generator=g)
    rois = rois[perm].contiguous()
    exp_out, exp_arg = ref.roi_pool(feat, rois)
    out, arg, _ = ops.roi_pool_forward(feat.cuda(), rois.cuda())
    assert torch.equal(out.cpu(), exp_out)
    assert torch.equal(arg.cpu(), exp_arg)


def test_roi_pool_forward_large_plane_fallback(ops):
    g = _gen(4)
    feat = torch.relu(torch.randn((1, 4, 300, 260), generator=g))  # 78000 cells: does not fit shared memory
    boxes = ref.synth_boxes(64, 2400, 2080, g)
    rois = ref.boxes_to_pooler_format([boxes])
    exp_out, exp_arg = ref.roi_pool(feat, rois)
    out, arg, _ = ops.roi_pool_forward(feat.cuda(), rois.cuda())
    assert torch.equal(out.cpu(), exp_out)
    assert torch.equal(arg.cpu(), exp_arg)


def test_roi_pool_empty(ops):
    feat = torch.zeros((1, 8, 10, 10)).cuda()
    out, arg, _ = ops.roi_pool_forward(feat, torch.zeros((0, 5)).cuda())
    assert out.shape == (0, 8, 7, 7)
    gf = ops.roi_pool_backward(torch.zeros((0, 8 * 49)).cuda(), arg, torch.zeros((0, 5)).cuda(), (1, 8, 10, 10))
    assert float(gf.abs().sum()) == 0.0


@pytest.mark.parametrize("grad_bf16,argmax_u16", [(False, False), (True, True)])
def test_roi_pool_backward(ops, grad_bf16, argmax_u16):
    import torchvision

    g = _gen(5)
    C, h, w, R = 32, 45, 60, 200
    feat = torch.relu(torch.randn((1, C, h, w), generator=g)).requires_grad_(True)
    boxes = _rois_with_edge_cases(R, h * 8, w * 8, g)
    rois = ref.boxes_to_pooler_format([boxes])
    obj = torch.rand(rois.size(0), generator=g)
    pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125) * (obj + 1).view(-1, 1, 1, 1)
    go = torch.randn(pooled.shape, generator=g)
    if grad_bf16:
        go = go.to(torch.bfloat16).float()
    pooled.backward(go)
    _, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rois.cuda(), want_f32=False, argmax_u16=argmax_u16)
    go_dev = go.flatten(1).cuda()
    if grad_bf16:
        go_dev = go_dev.to(torch.bfloat16)
    gf = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (1, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0)
    # fp32 sums in a different association order: tolerance, not bit-exact
    torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("C,h,w,N", [(16, 150, 200, 1), (8, 300, 260, 1), (24, 72, 96, 2), (10, 20, 30, 2)])
def test_roi_pool_backward_plane_configs(ops, C, h, w, N):
    """Planes that need fewer channels per CTA / several row bands / the misaligned-input fallback (C=10), and a
    two-image batch with interleaved rois."""
    import torchvision

    g = _gen(50 + C)
    R = 120
    feat = torch.randn((N, C, h, w), generator=g).requires_grad_(True)
    bl = [ref.synth_boxes(R, h * 8, w * 8, g) for _ in range(N)]
    rois = ref.boxes_to_pooler_format(bl)
    if N > 1:
        rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125)
    go = torch.randn(pooled.shape, generator=g)
    pooled.backward(go)
    _, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rois.cuda(), want_f32=False)
    assert torch.equal(arg.cpu(), ref.roi_pool(feat.detach(), rois)[1])
    gf = ops.roi_pool_backward(go.flatten(1).cuda(), arg, rois.cuda(), (N, C, h, w))
    torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("pooled", [(7, 7), (3, 5), (2, 2), (6, 9)])
def test_roi_pool_backward_tiny_rois(ops, pooled):
    """rois only 1..6 cells high / wide repeat cells across many bins (the colour strides of the bin-ownership
    schedule grow beyond 2), all-zero features make every bin of a roi row-major-first ties, and pooled grids
    other than 7x7 take the generic template."""
    import torchvision

    g = _gen(77 + pooled[0])
    C, h, w, R = 16, 40, 50, 300
    feat = torch.relu(torch.randn((1, C, h, w), generator=g))
    feat[:, :4] = 0.0
    feat.requires_grad_(True)
    x1 = torch.rand(R, generator=g) * (w * 8 - 60)
    y1 = torch.rand(R, generator=g) * (h * 8 - 60)
    bw = torch.rand(R, generator=g) * 56 + 1
    bh = torch.rand(R, generator=g) * 56 + 1
    bw[::3] = torch.rand(R, generator=g)[::3] * 200 + 1
    boxes = torch.stack([x1, y1, x1 + bw, y1 + bh], 1).round()
    rois = ref.boxes_to_pooler_format([boxes])
    pooled_ref = torchvision.ops.roi_pool(feat, rois, pooled, 0.125)
    go = torch.randn(pooled_ref.shape, generator=g)
    pooled_ref.backward(go)
    out, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rois.cuda(), pooled)
    assert torch.equal(out.cpu(), pooled_ref.detach())
    gf = ops.roi_pool_backward(go.flatten(1).cuda(), arg, rois.cuda(), (1, C, h, w), pooled)
    torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=1e-4)


# ---- the planned 7x7 operand-mode kernels (roi_pool_fast.cu) ----
def _fast_fwd_check(ops, feat, rois, obj):
    """Forward with a plan: bf16 operand and uint16 arg-max must equal the oracle's bit for bit."""
    exp_out, exp_arg = ref.roi_pool(feat, rois)
    N, C, h, w = feat.shape
    plan = ops.roi_pool_plan(rois.cuda(), (N, C, h, w), row_scale=None if obj is None else obj.cuda(), row_scale_bias=1.0)
    assert plan is not None
    _, arg16, obf = ops.roi_pool_forward(feat.cuda(), rois.cuda(), row_scale=None if obj is None else obj.cuda(),
                                         row_scale_bias=1.0, want_f32=False, want_bf16=True, argmax_u16=True, plan=plan)
    a = arg16.cpu().to(torch.int32) & 0xFFFF
    a[a == 0xFFFF] = -1
    assert torch.equal(a, exp_arg), "argmax indices must be bit-exact"
    scale = torch.ones(rois.size(0)) if obj is None else obj + 1
    exp_bf = (exp_out * scale.view(-1, 1, 1, 1)).flatten(1).to(torch.bfloat16)
    assert torch.equal(obf.cpu().view(torch.int16), exp_bf.view(torch.int16)), "bf16 operand must be bit-exact (incl. the sign of zero)"
    return plan, arg16


@pytest.mark.parametrize("C,h,w,R", [(512, 60, 80, 300), (64, 72, 96, 257), (8, 37, 53, 300), (6, 96, 152, 200),
                                     (4, 20, 200, 150), (2, 150, 30, 100), (8, 108, 144, 300), (16, 96, 128, 300),
                                     (12, 84, 112, 200), (4, 90, 120, 200)])
def test_roi_pool_fast_forward_bit_exact(ops, C, h, w, R):
    """CI=4 and CI=2 interleaves, full window table (60x80 four channels; 72x96, 84x112, 90x120 two) and half table
    (96x128, 96x152, 108x144, two channels), direct / window / wide bin modes (bins up to 200/7 cells wide), clipped,
    malformed, out-of-image and empty rois."""
    g = _gen(300 + C + h)
    feat = torch.relu(torch.randn((1, C, h, w), generator=g))
    boxes = _rois_with_edge_cases(R, h * 8, w * 8, g)
    rois = ref.boxes_to_pooler_format([boxes])
    _fast_fwd_check(ops, feat, rois, torch.rand(rois.size(0), generator=g))


@pytest.mark.parametrize("h,w", [(45, 64), (72, 96), (100, 140)])
def test_roi_pool_fast_forward_ties_negative_and_batch(ops, h, w):
    """All-zero channels (every cell ties: the first cell in row-major order must win), negative features, signed
    zeros, tiny rois (bins narrower than a cell), three images with shuffled roi order and an image without rois; on the
    full-table kernel four (45x64) and two (72x96) channels at a time and on the half-table kernel (100x140)."""
    g = _gen(311)
    N, C = 4, 8
    feat = torch.randn((N, C, h, w), generator=g)
    feat[:, 0] = 0.0
    feat[:, 1] = -0.0
    feat[:, 2] = torch.relu(feat[:, 2])
    feat[:, 3] = -torch.relu(feat[:, 3])           # <= 0 with many -0.0
    feat[:, 4] = torch.round(feat[:, 4] * 2) / 2   # heavy ties
    bl = []
    for i in range(N):
        R = 0 if i == 2 else 150
        x1 = torch.rand(R, generator=g) * (w * 8 - 60)
        y1 = torch.rand(R, generator=g) * (h * 8 - 60)
        bw = torch.rand(R, generator=g) * 56 + 1
        bh = torch.rand(R, generator=g) * 56 + 1
        bw[::3] = torch.rand(R, generator=g)[::3] * 400 + 1
        bh[::4] = torch.rand(R, generator=g)[::4] * 300 + 1
        bl.append(torch.stack([x1, y1, x1 + bw, y1 + bh], 1).round())
    rois = ref.boxes_to_pooler_format(bl)
    rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    _fast_fwd_check(ops, feat, rois, None)


def test_roi_pool_fast_matches_general_kernel_at_bench_shape(ops):
    """BASELINE shape (2 x [512, 72, 96], 2 x 2000 proposals): planned kernel == general kernel, every element."""
    g = _gen(312)
    feat = torch.relu(torch.randn((2, 512, 72, 96), generator=g)).cuda()
    rois = ref.boxes_to_pooler_format([ref.synth_boxes(2000, 576, 768, g) for _ in range(2)]).cuda()
    obj = torch.rand(rois.size(0), generator=g).cuda()
    _, a0, x0 = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, want_bf16=True, argmax_u16=True)
    plan = ops.roi_pool_plan(rois, feat.shape, row_scale=obj, row_scale_bias=1.0)
    _, a1, x1 = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, want_bf16=True, argmax_u16=True, plan=plan)
    assert torch.equal(a0, a1)
    assert torch.equal(x0.view(torch.int16), x1.view(torch.int16))


@pytest.mark.parametrize("h,w,scale", [(96, 128, 1.6), (108, 144, 1.8)])
def test_roi_pool_fast_matches_general_kernel_at_tta_scales(ops, h, w, scale):
    """The two largest test-time views of BASELINE configs[4] (image + flip, 2 x 2000 proposals scaled with the view):
    half-table kernel == general kernel, every element."""
    g = _gen(313)
    feat = torch.relu(torch.randn((2, 128, h, w), generator=g)).cuda()
    boxes = [ref.synth_boxes(2000, 480, 640, g) * scale for _ in range(2)]
    rois = ref.boxes_to_pooler_format(boxes).cuda()
    obj = torch.rand(rois.size(0), generator=g).cuda()
    _, a0, x0 = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, want_bf16=True, argmax_u16=True)
    plan = ops.roi_pool_plan(rois, feat.shape, row_scale=obj, row_scale_bias=1.0)
    _, a1, x1 = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, want_bf16=True, argmax_u16=True, plan=plan)
    assert torch.equal(a0, a1)
    assert torch.equal(x0.view(torch.int16), x1.view(torch.int16))


@pytest.mark.parametrize("C,h,w,N,grad_bf16", [(32, 45, 60, 1, True), (24, 72, 96, 2, False), (7, 30, 40, 3, True),
                                               (16, 150, 200, 1, True), (8, 200, 300, 1, False)])
def test_roi_pool_fast_backward(ops, C, h, w, N, grad_bf16):
    """Planned backward (two planes per warp, colour steps from the plan) against torchvision's autograd: edge-case
    rois, tiny rois with strides > 2, shuffled multi-image batches, odd channel counts, planes needing row bands."""
    import torchvision

    g = _gen(400 + C + h)
    feat = torch.relu(torch.randn((N, C, h, w), generator=g))
    feat[:, :2] = 0.0
    feat.requires_grad_(True)
    bl = []
    for i in range(N):
        boxes = _rois_with_edge_cases(150, h * 8, w * 8, g)
        R = 100
        x1 = torch.rand(R, generator=g) * (w * 8 - 60)
        y1 = torch.rand(R, generator=g) * (h * 8 - 60)
        bw = torch.rand(R, generator=g) * 56 + 1
        bh = torch.rand(R, generator=g) * 56 + 1
        bl.append(torch.cat([boxes, torch.stack([x1, y1, x1 + bw, y1 + bh], 1).round()], 0))
    rois = ref.boxes_to_pooler_format(bl)
    if N > 1:
        rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    obj = torch.rand(rois.size(0), generator=g)
    pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125) * (obj + 1).view(-1, 1, 1, 1)
    go = torch.randn(pooled.shape, generator=g)
    if grad_bf16:
        go = go.to(torch.bfloat16).float()
    pooled.backward(go)
    plan = ops.roi_pool_plan(rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0)
    _, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rois.cuda(), want_f32=False, argmax_u16=True, want_bf16=True,
                                     row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    go_dev = go.flatten(1).cuda()
    if grad_bf16:
        go_dev = go_dev.to(torch.bfloat16)
    gf = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=1e-4)
    gf2 = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    assert torch.equal(gf, gf2), "the backward must be deterministic (odd channel counts are padded, never atomics)"


def test_roi_pool_fast_backward_bench_shape(ops):
    """BASELINE shape: planned backward == general backward up to fp32 summation order, and run-to-run identical."""
    g = _gen(401)
    feat = torch.relu(torch.randn((2, 512, 60, 80), generator=g)).cuda()
    rois = ref.boxes_to_pooler_format([ref.synth_boxes(2000, 480, 640, g) for _ in range(2)]).cuda()
    obj = torch.rand(rois.size(0), generator=g).cuda()
    plan = ops.roi_pool_plan(rois, feat.shape, row_scale=obj, row_scale_bias=1.0)
    _, arg, _ = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, want_bf16=True,
                                     argmax_u16=True, plan=plan)
    go = torch.randn((4000, 25088), generator=g).to(torch.bfloat16).cuda()
    g0 = ops.roi_pool_backward(go, arg, rois, feat.shape, row_scale=obj, row_scale_bias=1.0)
    g1 = ops.roi_pool_backward(go, arg, rois, feat.shape, row_scale=obj, row_scale_bias=1.0, plan=plan)
    g2 = ops.roi_pool_backward(go, arg, rois, feat.shape, row_scale=obj, row_scale_bias=1.0, plan=plan)
    assert torch.equal(g1, g2)
    torch.testing.assert_close(g1, g0, rtol=2e-4, atol=2e-3)


# ------------------------------------------------------------------------------------------------
# (2) GEMM
# ------------------------------------------------------------------------------------------------
def _gemm_ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 512, 200), (2000, 4096, 1024), (264, 344, 1000), (77, 40, 72)])
def test_gemm_majors(ops, a_mn, b_mn, m, n, k):
    g = _gen(m + n + k)
    pad = lambda v: (v + 7) // 8 * 8
    a = (torch.randn((k, pad(m)) if a_mn else (m, pad(k)), generator=g) * 0.5).to(torch.bfloat16)
    b = (torch.randn((k, pad(n)) if b_mn else (n, pad(k)), generator=g) * 0.5).to(torch.bfloat16)
    a_v = a[:, :m] if a_mn else a[:, :k]
    b_v = b[:, :n] if b_mn else b[:, :k]
    exp = _gemm_ref(a_v, b_v, a_mn, b_mn)
    ad, bd = a.cuda(), b.cuda()
    out = ops.gemm_bf16(ad[:, :m] if a_mn else ad[:, :k], bd[:, :n] if b_mn else bd[:, :k], a_mn=a_mn, b_mn=b_mn)
    torch.testing.assert_close(out.cpu(), exp, rtol=2e-3, atol=2e-3 * math.sqrt(k))


def test_gemm_epilogue_bias_relu_dropout_bf16(ops):
    g = _gen(7)
    m, n, k = 500, 768, 320
    a = (torch.randn((m, k), generator=g) * 0.3).to(torch.bfloat16)
    b = (torch.randn((n, k), generator=g) * 0.3).to(torch.bfloat16)
    bias = torch.randn(n, generator=g)
    seed = 0x1234ABCD5678
    out = ops.gemm_bf16(a.cuda(), b.cuda(), out_dtype=torch.bfloat16, bias=bias.cuda(), relu=True, dropout_p=0.5,
                        dropout_seed=seed)
    mask = ops.dropout_mask(m, n, 0.5, seed).cpu()
    keep_frac = mask.float().mean().item()
    assert 0.48 < keep_frac < 0.52
    exp = torch.relu(a.float() @ b.float().t() + bias) * mask * 2.0
    torch.testing.assert_close(out.float().cpu(), exp, rtol=1e-2, atol=2e-2)
    # dropped entries are exactly zero, kept positive entries are non-zero
    assert torch.equal(out.cpu() == 0, (exp.to(torch.bfloat16) == 0))


def test_gemm_epilogue_backward_mask(ops):
    g = _gen(8)
    m, n, k = 260, 512, 128
    a = (torch.randn((m, k), generator=g) * 0.3).to(torch.bfloat16)
    b = (torch.randn((k, n), generator=g) * 0.3).to(torch.bfloat16)   # MN-major B (dgrad form)
    y = torch.relu(torch.randn((m, n), generator=g)).to(torch.bfloat16)
    y[::3] = -y[::3]                                                   # negative / zero entries -> masked out
    out = ops.gemm_bf16(a.cuda(), b.cuda(), b_mn=True, out_dtype=torch.bfloat16, mask_src=y.cuda(), mask_scale=2.0)
    exp = (a.float() @ b.float()) * (y.float() > 0).float() * 2.0
    torch.testing.assert_close(out.float().cpu(), exp, rtol=1e-2, atol=2e-2)


def test_cast_transpose_colsum_sgd(ops):
    g = _gen(9)
    x = torch.randn((300, 1000), generator=g)
    s = torch.rand(1000, generator=g)
    o, ot = ops.cast_f32_bf16(x.cuda(), col_scale=s.cuda(), want_t=True)
    exp = (x * s).to(torch.bfloat16)
    assert torch.equal(o.cpu(), exp)
    assert torch.equal(ot.cpu(), exp.t())
    assert torch.equal(ops.transpose_bf16(o).cpu(), exp.t())
    torch.testing.assert_close(ops.colsum(x.cuda()).cpu(), x.sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(ops.colsum(o).cpu(), exp.float().sum(0), rtol=1e-4, atol=1e-3)
    # SGD step vs torch.optim.SGD
    p = torch.randn(5000, generator=g)
    grad = torch.randn(5000, generator=g)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pt], lr=1e-3, momentum=0.9, weight_decay=5e-4)
    pd, buf = p.cuda(), torch.zeros(5000).cuda()
    pbf = torch.empty(5000, dtype=torch.bfloat16).cuda()
    for _ in range(3):
        pt.grad = grad.clone()
        opt.step()
        ops.sgd_step(pd, grad.cuda(), buf, 1e-3, 0.9, 5e-4, 1.0, pbf)
    torch.testing.assert_close(pd.cpu(), pt.detach(), rtol=1e-6, atol=1e-7)
    assert torch.equal(pbf.cpu(), pd.cpu().to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------
# (3) WSDDN
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,C,V", [(2000, 20, 4), (777, 80, 2), (33, 5, 1)])
def test_wsddn_forward_backward(ops, R, C, V):
    g = _gen(R + C)
    ld = 2 * C + 24
    col_cls, col_det = 8, 8 + C + 3
    logits = torch.randn((V * R, ld), generator=g) * 2.0
    gt = torch.zeros(C)
    gt[torch.randperm(C, generator=g)[: max(1, C // 6)]] = 1.0
    lg = logits.clone().requires_grad_(True)
    exp_scores, exp_loss = [], []
    for v in range(V):
        blk = lg[v * R:(v + 1) * R]
        s = ref.wsddn_scores_from_logits(blk[:, col_cls:col_cls + C], blk[:, col_det:col_det + C])
        exp_scores.append(s)
        exp_loss.append(ref.wsddn_loss(s, gt.view(1, C)))
    torch.stack(exp_loss).sum().backward()
    dl = torch.zeros((V * R, ld)).cuda()
    scores, img, loss = ops.wsddn_forward(logits.cuda(), col_cls, col_det, V, R, C, gt.cuda(), dlogits=dl)
    for v in range(V):
        torch.testing.assert_close(scores[v].cpu(), exp_scores[v].detach(), rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(img[v].cpu(), ref.wsddn_img_scores(exp_scores[v].detach())[0], rtol=1e-4, atol=1e-7)
        assert abs(loss[v].item() - exp_loss[v].item()) < 1e-3 * max(1.0, abs(exp_loss[v].item()))
    torch.testing.assert_close(dl.cpu(), lg.grad, rtol=2e-3, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# (4) OICR: mining + labelling (bit-exact), loss
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,C,G,K,seed", [(2000, 20, 3, 3, 0), (2000, 20, 1, 4, 1), (1500, 80, 7, 3, 2), (50, 20, 2, 1, 3),
                                          (4000, 20, 4, 3, 4), (4000, 20, 4, 4, 5), (10000, 80, 8, 3, 6),
                                          (10000, 80, 2, 4, 7), (16384, 80, 1, 1, 8)])
def test_oicr_mine_label_bit_exact(ops, R, C, G, K, seed):
    """Shipped shapes included: K=4 (voc07_oicr_plus.yaml:56-58), R=4000 (PRECOMPUTED_PROPOSAL_TOPK_TRAIN of
    Base-RCNN-DilatedC5.yaml:4-10), R=10000 x C=80 (coco_oicr_plus.yaml:64-72), and the kernel's limit R=16384."""
    g = _gen(100 + seed)
    boxes = ref.synth_boxes(R, 480, 640, g)
    gt_int = torch.sort(torch.randperm(C, generator=g)[:G]).values
    prev = torch.stack([ref.synth_prev_scores(R, C + 1, g) for _ in range(K)])
    # make a few columns peaky so that the >= 0.05 threshold keeps many candidates in some classes, few in others
    prev[:, :, gt_int[0]] = prev[:, :, gt_int[0]] * 3.0
    top_k = max(int(R * 0.10), 1)
    out = ops.oicr_mine_label(prev.cuda(), boxes.cuda(), gt_int.cuda(), C, top_k)
    for k in range(K):
        seeds = ref.pgt_mist(boxes, prev[k], gt_int, 0.10, 0.05)
        y, w, gidx, matched, _ = ref.label_proposals(boxes, seeds, C)
        M = int(out["seed_count"][k].item())
        assert M == seeds.index.numel()
        assert torch.equal(out["seed_index"][k, :M].cpu().long(), seeds.index)
        assert torch.equal(out["seed_class"][k, :M].cpu().long(), seeds.classes)
        assert torch.equal(out["seed_score"][k, :M].cpu(), seeds.scores)
        assert torch.equal(out["gt_class"][k].cpu().long(), y), "pseudo-labels must be bit-exact"
        assert torch.equal(out["gt_index"][k].cpu().long(), gidx), "assignment indices must be bit-exact"
        assert torch.equal(out["gt_weight"][k].cpu(), w)
        cnt = out["counts"][k].cpu().tolist()
        assert cnt == [int(((y >= 0) & (y < C)).sum()), int((y == C).sum()), int((y == -1).sum())]


def test_image_level_gt_and_device_side_count(ops):
    """soswsod_image_level_gt == get_image_level_gt (sorted distinct classes, one-hot), and mining with the padded list +
    device count gives exactly what mining with the host-sized list gives."""
    g = _gen(91)
    for C, raw in [(20, [7, 3, 3, 12, 7]), (80, [79, 0, 64, 31, 32, 0]), (20, [5]), (128, list(range(127, -1, -3)))]:
        gt = torch.tensor(raw, dtype=torch.int64)
        lst, cnt, oh = ops.image_level_gt(gt.cuda(), C)
        e_int, e_oh = ref.image_level_gt(gt, C)
        n = int(cnt.item())
        assert n == e_int.numel() and torch.equal(lst[:n].cpu().long(), e_int) and bool((lst[n:] == -1).all())
        assert torch.equal(oh.cpu(), e_oh.reshape(-1))
        lst32, cnt32, _ = ops.image_level_gt(gt.to(torch.int32).cuda(), C)
        assert torch.equal(lst32, lst) and torch.equal(cnt32, cnt)
    R, C, K = 1200, 20, 3
    prev = torch.stack([ref.synth_prev_scores(R, C + 1, g) for _ in range(K)]).cuda()
    boxes = ref.synth_boxes(R, 480, 640, g).cuda()
    gt = torch.tensor([11, 2, 2, 17], dtype=torch.int64).cuda()
    lst, cnt, _ = ops.image_level_gt(gt, C)
    a = ops.oicr_mine_label(prev, boxes, torch.unique(gt).to(torch.int32), C, 120)
    b = ops.oicr_mine_label(prev, boxes, lst, C, 120, gt_count=cnt)
    for k in range(K):
        M = int(a["seed_count"][k])
        assert M == int(b["seed_count"][k])
        for key in ("seed_index", "seed_class", "seed_score"):
            assert torch.equal(a[key][k, :M], b[key][k, :M]), key
    for key in ("gt_class", "gt_weight", "gt_index", "counts"):
        assert torch.equal(a[key], b[key]), key


@pytest.mark.parametrize("R,C,K,quirk", [(2000, 20, 3, True), (500, 80, 2, True), (300, 20, 4, False)])
def test_oicr_loss_and_grad(ops, R, C, K, quirk):
    g = _gen(200 + R)
    V = 4
    views = ref.synth_views(R, [(480, 640), (576, 768)], g, channels=1)
    boxes = torch.stack([v.boxes for v in views])
    stride = 5 * C + 1
    col0 = 2 * C
    ld = (col0 + K * stride + 63) // 64 * 64
    logits = torch.randn((V * R, ld), generator=g)
    logits[:, col0:] *= 0.5
    gt_int = torch.tensor([1, 4])
    ys, ws, gis = [], [], []
    for k in range(K):
        prev = ref.synth_prev_scores(R, C + 1, g)
        seeds = ref.pgt_mist(views[0].boxes, prev, gt_int, 0.10, 0.05)
        y, w, gidx, _, _ = ref.label_proposals(views[0].boxes, seeds, C)
        ys.append(y), ws.append(w), gis.append(gidx)
    lg = logits.clone().requires_grad_(True)
    exp = torch.zeros((K, 2))
    exp_acc = []
    total = 0
    for k in range(K):
        c0 = col0 + k * stride
        lc, lb = [], []
        for vi in range(V):
            src = 2 if (vi == 3 and quirk) else vi
            Z = lg[src * R:(src + 1) * R, c0:c0 + C + 1]
            D = lg[src * R:(src + 1) * R, c0 + C + 1:c0 + stride]
            lc.append(ref.oicr_cls_loss(Z, ys[k], ws[k]))
            lb.append(ref.oicr_box_loss(D, ys[k], views[vi].boxes, views[vi].boxes[gis[k]], C))
            exp_acc.append(ref.oicr_accuracy_counters(Z.detach(), ys[k]))
        l0, l1 = sum(lc) / 4.0, sum(lb) / 4.0
        total = total + l0 + l1
        exp[k, 0], exp[k, 1] = l0.item(), l1.item()
    total.backward()
    dl = torch.zeros((V * R, ld)).cuda()
    losses, view_losses, acc = ops.oicr_loss(
        logits.cuda(), col0, stride, boxes.cuda(), torch.stack(ys).int().cuda(), torch.stack(ws).cuda(),
        torch.stack(gis).int().cuda(), V, R, C, K, flip_quirk=quirk, dlogits=dl)
    assert (losses.cpu() - exp).abs().max().item() < 1e-3, (losses.cpu(), exp)
    torch.testing.assert_close(dl.cpu()[:, col0:col0 + K * stride], lg.grad[:, col0:col0 + K * stride], rtol=2e-3, atol=1e-7)
    assert acc.cpu().view(-1, 5).tolist() == [list(a) for a in exp_acc]


def test_oicr_avg_scores(ops):
    g = _gen(11)
    V, R, C, K = 4, 700, 20, 3
    stride = 5 * C + 1
    col0 = 2 * C
    ld = 384
    S = torch.rand((V, R, C), generator=g) * 1e-3
    logits = torch.randn((V * R, ld), generator=g)
    prev = ops.oicr_avg_scores(S.cuda(), logits.cuda(), col0, stride, V, R, C, K).cpu()
    e0 = (S[0] + S[1] + S[2] + S[3]) / 4.0
    assert torch.equal(prev[0, :, :C], e0), "k=0 average follows the reference's summation order exactly"
    assert float(prev[0, :, C].abs().max()) == 0.0
    for k in range(1, K):
        c0 = col0 + (k - 1) * stride
        p = [F.softmax(logits[v * R:(v + 1) * R, c0:c0 + C + 1], dim=-1) for v in range(V)]
        torch.testing.assert_close(prev[k], (p[0] + p[1] + p[2] + p[3]) / 4.0, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------------
# (5) test-time: predict, TTA merge, NMS, detect
# ------------------------------------------------------------------------------------------------
def test_predict_and_tta(ops):
    g = _gen(12)
    R, C, K = 900, 20, 3
    stride, col0, ld = 5 * C + 1, 2 * C, 384
    logits = torch.randn((R, ld), generator=g)
    logits[:, col0:] *= 0.3
    boxes = ref.synth_boxes(R, 480, 640, g)
    ZK = [logits[:, col0 + k * stride: col0 + k * stride + C + 1] for k in range(K)]
    DK = [logits[:, col0 + k * stride + C + 1: col0 + (k + 1) * stride] for k in range(K)]
    probs, pb = ops.predict(logits.cuda(), col0, stride, boxes.cuda(), C, K)
    torch.testing.assert_close(probs.cpu(), ref.predict_probs_K(ZK), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(pb.cpu(), ref.predict_boxes_K(DK, boxes), rtol=1e-5, atol=1e-3)
    # TTA: 3 views (scale, flip) merged by running mean
    views = [(1.0, 1.0, False, 640.0), (0.8, 0.8, True, 800.0), (1.25, 1.25, True, 512.0)]
    acc_b = torch.empty((R, 4 * C)).cuda()
    acc_p = torch.empty((R, C + 1)).cuda()
    eb, ep = [], []
    for i, (sx, sy, fl, vw) in enumerate(views):
        b_i = pb.cpu() * (1.0 + 0.01 * i)
        p_i = probs.cpu() * (1.0 - 0.1 * i)
        ops.tta_accumulate(b_i.cuda(), p_i.cuda(), sx, sy, fl, vw, i == 0, float(len(views)) if i == len(views) - 1 else 0.0,
                           acc_b, acc_p)
        eb.append(ref.tta_inverse_boxes(b_i.reshape(-1, 4), sx, sy, fl, vw).reshape(R, 4 * C))
        ep.append(p_i)
    mb, mp = ref.tta_merge(eb, ep)
    torch.testing.assert_close(acc_b.cpu(), mb, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(acc_p.cpu(), mp, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,thr", [(1, 0.3), (63, 0.3), (64, 0.5), (65, 0.01), (2000, 0.3), (2000, 0.01), (5000, 0.7), (12000, 0.3)])
def test_nms_keep_list_bit_exact(ops, n, thr):
    g = _gen(300 + n)
    boxes = ref.synth_boxes(n, 600, 800, g) if n > 4 else torch.tensor([[1.0, 2.0, 30.0, 40.0]] * n)
    scores = torch.rand(n, generator=g)
    if n >= 64:
        scores[10] = scores[20]     # exact ties: lower index first
        scores[33] = scores[5]
        boxes[40] = boxes[41]       # duplicate boxes (IoU 1)
    exp = ref.nms(boxes, scores, thr)
    keep = ops.nms(boxes.cuda(), scores.cuda(), thr)
    assert torch.equal(keep.cpu(), exp), "NMS keep-list must be bit-exact (order included)"


def test_nms_degenerate_boxes(ops):
    # zero-area boxes give 0/0 = NaN IoU -> never suppressed (NaN > thr is false), as in torchvision
    boxes = torch.tensor([[10.0, 10.0, 10.0, 10.0], [10.0, 10.0, 10.0, 10.0], [0.0, 0.0, 5.0, 5.0], [0.0, 0.0, 5.0, 5.0]])
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6])
    exp = ref.nms(boxes, scores, 0.5)
    keep = ops.nms(boxes.cuda(), scores.cuda(), 0.5)
    assert torch.equal(keep.cpu(), exp)


@pytest.mark.parametrize("R,C,thr", [(2000, 20, 1e-6), (1000, 80, 1e-5), (300, 20, 0.05)])
def test_detect_matches_reference_inference(ops, R, C, thr):
    g = _gen(400 + R)
    boxes = ref.synth_boxes(R, 480, 640, g)
    deltas = torch.randn((R, 4 * C), generator=g) * 0.5
    pred_boxes = ref.apply_deltas(deltas, boxes)
    probs = F.softmax(torch.randn((R, C + 1), generator=g) * 3.0, dim=1)
    probs[5, 2] = float("nan")          # non-finite rows are dropped
    pred_boxes[9, 7] = float("inf")
    eb, es, ec, er = ref.fast_rcnn_inference_single_image(pred_boxes, probs, (480, 640), thr, 0.3, 100)
    db, ds, dc, dr, nd = ops.detect(probs.cuda(), pred_boxes.cuda(), (480, 640), thr, 0.3, 100)
    n = int(nd.item())
    assert n == es.numel()
    assert torch.equal(dr[:n].cpu().long(), er), "detection row indices (keep-list) must be bit-exact"
    assert torch.equal(dc[:n].cpu().long(), ec)
    assert torch.equal(ds[:n].cpu(), es)
    assert torch.equal(db[:n].cpu(), eb)


def test_roi_pool_backward_duplicate_argmax_merge(ops):
    """Neighbouring bins of a roi (up to 2 x 2 of them) can share their arg-max cell; the backward's colour steps exist to
    keep such updates ordered.  Sparse peaky planes make these groups the rule: every bin around a peak points at it.
    Checked against torchvision's autograd, the general kernel, and for run-to-run identity."""
    import torchvision

    g = _gen(977)
    N, C, h, w = 2, 64, 60, 80
    feat = torch.zeros((N, C, h, w))
    idx = torch.randint(0, h * w, (N, C, 60), generator=g)
    feat.view(N, C, -1).scatter_(2, idx, torch.rand((N, C, 60), generator=g) + 0.5)      # ~1 % of the cells are peaks
    feat[:, 5] = 0.0                                                                      # an all-zero plane (first-cell arg-max)
    feat[:, 6] = 1.0                                                                      # a constant plane
    feat.requires_grad_(True)
    rois = ref.boxes_to_pooler_format([ref.synth_boxes(600, h * 8, w * 8, g) for _ in range(N)])
    obj = torch.rand(rois.size(0), generator=g)
    pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125) * (obj + 1).view(-1, 1, 1, 1)
    go = torch.randn(pooled.shape, generator=g).to(torch.bfloat16).float()
    pooled.backward(go)
    plan = ops.roi_pool_plan(rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0)
    _, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rois.cuda(), want_f32=False, argmax_u16=True, want_bf16=True,
                                     row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    a = arg.view(torch.int16).cpu().view(-1, C, 49)
    dup = (a[:, :, :-1] == a[:, :, 1:]) & (a[:, :, 1:] != -1)
    assert dup.float().mean() > 0.05, "the fixture must contain many bins sharing a cell"
    go_dev = go.flatten(1).cuda().to(torch.bfloat16)
    gf = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=2e-4)
    gf2 = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0, plan=plan)
    assert torch.equal(gf, gf2)
    general = ops.roi_pool_backward(go_dev, arg, rois.cuda(), (N, C, h, w), row_scale=obj.cuda(), row_scale_bias=1.0)
    torch.testing.assert_close(gf, general, rtol=1e-4, atol=2e-4)
