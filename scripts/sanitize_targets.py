#!/usr/bin/env python
"""Small invocations of the two kernels with hand-rolled synchronisation -- the queued ROI backward (tagged shared-memory
queue between prep and accumulator warps) and the GEMM (mbarrier rings, TMEM hand-off, global tile counter re-armed by
the last CTA) -- for `compute-sanitizer --tool racecheck|memcheck|synccheck` (scripts/run_sanitizer.sh).  Two shapes
each; results are checked against torch so that a sanitizer-clean run is also a correct one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sos_wsod_b200 import ops  # noqa: E402
from sos_wsod_b200.synthetic import synth_boxes  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator().manual_seed(3)
if which in ("all", "roi"):
    import torchvision

    for (n, c, h, w, R) in [(2, 16, 30, 40, 160), (1, 24, 45, 60, 120)]:
        feat = torch.relu(torch.randn((n, c, h, w), generator=g)).requires_grad_(True)
        rois = torch.cat([torch.cat([torch.full((R, 1), float(i)), synth_boxes(R, h * 8, w * 8, g)], 1) for i in range(n)], 0)
        obj = torch.rand(n * R, generator=g)
        pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125) * (obj + 1).view(-1, 1, 1, 1)
        go = torch.randn(pooled.shape, generator=g).to(torch.bfloat16).float()
        pooled.backward(go)
        rd, od = rois.cuda(), obj.cuda()
        plan = ops.roi_pool_plan(rd, (n, c, h, w), row_scale=od, row_scale_bias=1.0)
        _, arg, _ = ops.roi_pool_forward(feat.detach().cuda(), rd, want_f32=False, argmax_u16=True, want_bf16=True, row_scale=od,
                                         row_scale_bias=1.0, plan=plan)
        gf = ops.roi_pool_backward(go.flatten(1).cuda().to(torch.bfloat16), arg, rd, (n, c, h, w), row_scale=od,
                                   row_scale_bias=1.0, plan=plan)
        torch.cuda.synchronize()
        torch.testing.assert_close(gf.cpu(), feat.grad, rtol=1e-4, atol=1e-4)
        print("roi backward ok", (n, c, h, w, R), flush=True)
if which in ("all", "gemm"):
    for (m, n, k, a_mn, b_mn) in [(128 * 20, 256 * 10, 192, False, False), (384, 1024, 520, True, True)]:
        a = (torch.randn((k, m) if a_mn else (m, k), generator=g) * 0.5).to(torch.bfloat16).cuda()
        b = (torch.randn((k, n) if b_mn else (n, k), generator=g) * 0.5).to(torch.bfloat16).cuda()
        for rep in range(2):     # the second launch runs on the counter re-armed by the first
            y = ops.gemm_bf16(a, b, a_mn=a_mn, b_mn=b_mn)
        torch.cuda.synchronize()
        exp = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
        torch.testing.assert_close(y, exp, rtol=2e-3, atol=2e-2)
        print("gemm ok", (m, n, k, a_mn, b_mn), flush=True)
