#!/usr/bin/env python
"""Times the ROI-pool forward at the training maps and the five test-time scales of BASELINE configs[4] (image + flip,
2 x 2000 proposals scaled with the view): general kernel vs planned kernel (full / half window table as selected).
SOSWSOD_ROI_FWD_FULL2=1 prefers two channels + full table over four channels + half table (A/B of the 72x96 map)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sos_wsod_b200 import ops
from sos_wsod_b200.synthetic import synth_boxes


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
    return e0.elapsed_time(e1) / n


g = torch.Generator().manual_seed(0)
R = 2000
for (h, w, scale) in [(60, 80, 1.0), (72, 96, 1.2), (84, 112, 1.4), (96, 128, 1.6), (108, 144, 1.8)]:
    feats = [torch.relu(torch.randn((2, 512, h, w), generator=g)).cuda() for _ in range(2)]
    boxes = [synth_boxes(R, 480, 640, g) * scale for _ in range(2)]
    rois = torch.cat([torch.cat([torch.full((R, 1), float(i)), b], 1) for i, b in enumerate(boxes)], 0).cuda()
    obj = torch.rand(2 * R, generator=g).cuda()
    X = [torch.empty((2 * R, 25088), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    state = {"i": 0}

    def fwd(plan):
        i = state["i"] = (state["i"] + 1) % 2
        return ops.roi_pool_forward(feats[i], rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, argmax_u16=True,
                                    out_bf16=X[i], plan=plan)

    plan = ops.roi_pool_plan(rois, feats[0].shape, row_scale=obj, row_scale_bias=1.0)
    t0 = timeit(lambda: fwd(None))
    t1 = timeit(lambda: fwd(plan))
    print(f"{h}x{w}: forward general {t0 * 1e3:.1f} us, planned {t1 * 1e3:.1f} us", flush=True)
