#!/usr/bin/env python
"""A/B timing of the queued ROI backward at the bench shapes: duplicate merge on/off x poll sleep of the prep warps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sos_wsod_b200 import ops  # noqa: E402
from sos_wsod_b200.synthetic import synth_boxes  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator().manual_seed(0)
for (h, w) in [(60, 80), (72, 96)]:
    R = 2000
    feat_shape = (2, 512, h, w)
    feat = torch.relu(torch.randn(feat_shape, generator=g)).cuda()
    rois = torch.cat([torch.cat([torch.full((R, 1), float(i)), synth_boxes(R, h * 8, w * 8, g)], 1) for i in range(2)], 0).cuda()
    obj = torch.rand(2 * R, generator=g).cuda()
    plan = ops.roi_pool_plan(rois, feat_shape, row_scale=obj, row_scale_bias=1.0)
    _, am, _ = ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, argmax_u16=True, want_bf16=True, plan=plan)
    go = [torch.randn((2 * R, 25088), device="cuda").to(torch.bfloat16) for _ in range(2)]
    state = {"i": 0}

    def bwd():
        i = state["i"] = (state["i"] + 1) % 2
        return ops.roi_pool_backward(go[i], am, rois, feat_shape, row_scale=obj, row_scale_bias=1.0, plan=plan)

    base = None
    for dedup in ("0", "1"):
        for poll in ("0", "50", "100", "200", "400"):
            for p in ("2", "3", "4"):
                os.environ.update(SOSWSOD_BWDQ_DEDUP=dedup, SOSWSOD_BWDQ_POLL_NS=poll, SOSWSOD_BWDQ_P=p)
                t = timeit(bwd)
                state["i"] = 0
                out = bwd().clone()
                if base is None:
                    base = out
                d = ((out - base).abs().max() / base.abs().max()).item()
                print(f"{h}x{w} dedup={dedup} poll_ns={poll} P={p}: {t:.1f} us  (max rel diff vs first {d:.2g})", flush=True)
