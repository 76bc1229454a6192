#!/usr/bin/env python
"""Times the ROI-pool kernels at the bench shapes (CUDA events, L2 flushed between launches by cycling buffers)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sos_wsod_b200 import ops
from oracle import oicr_plus_ref as ref

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator().manual_seed(0)
for (h, w) in [(60, 80), (72, 96)]:
    R = 2000
    feats = [torch.relu(torch.randn((2, 512, h, w), generator=g)).cuda() for _ in range(2)]
    rois = ref.boxes_to_pooler_format([ref.synth_boxes(R, h * 8, w * 8, g) for _ in range(2)]).cuda()
    obj = torch.rand(2 * R, generator=g).cuda()
    X = [torch.empty((2 * R, 25088), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    state = {"i": 0}
    def fwd(plan):
        i = state["i"] = (state["i"] + 1) % 2
        return ops.roi_pool_forward(feats[i], rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, argmax_u16=True, out_bf16=X[i], plan=plan)
    plan = ops.roi_pool_plan(rois, feats[0].shape, row_scale=obj, row_scale_bias=1.0)
    t_plan = timeit(lambda: ops.roi_pool_plan(rois, feats[0].shape, row_scale=obj, row_scale_bias=1.0))
    t0 = timeit(lambda: fwd(None)); t1 = timeit(lambda: fwd(plan))
    _, am, _ = fwd(plan)
    go = [torch.randn((2 * R, 25088), device="cuda").to(torch.bfloat16) for _ in range(2)]
    def bwd(plan):
        i = state["i"] = (state["i"] + 1) % 2
        return ops.roi_pool_backward(go[i], am, rois, feats[0].shape, row_scale=obj, row_scale_bias=1.0, plan=plan)
    b0 = timeit(lambda: bwd(None)); b1 = timeit(lambda: bwd(plan))
    state['i'] = 0; ga = bwd(None).clone(); state['i'] = 0; gb = bwd(plan); d = ((ga - gb).abs().max() / ga.abs().max()).item()
    print(f"{h}x{w}: plan {t_plan*1e3:.1f} us | fwd general {t0*1e3:.1f} us, planned {t1*1e3:.1f} us | bwd general {b0*1e3:.1f} us, planned {b1*1e3:.1f} us (max diff {d:.3g})")
