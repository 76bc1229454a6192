"""One planned ROI-forward launch per large test-time map (for an ncu capture of roi_pool_fwd_half_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sos_wsod_b200 import ops
from sos_wsod_b200.synthetic import synth_boxes

g = torch.Generator().manual_seed(0)
R = 2000
for (h, w, scale) in [(96, 128, 1.6), (108, 144, 1.8)]:
    feat = torch.relu(torch.randn((2, 512, h, w), generator=g)).cuda()
    boxes = [synth_boxes(R, 480, 640, g) * scale for _ in range(2)]
    rois = torch.cat([torch.cat([torch.full((R, 1), float(i)), b], 1) for i, b in enumerate(boxes)], 0).cuda()
    obj = torch.rand(2 * R, generator=g).cuda()
    X = torch.empty((2 * R, 25088), dtype=torch.bfloat16, device="cuda")
    plan = ops.roi_pool_plan(rois, feat.shape, row_scale=obj, row_scale_bias=1.0)
    for _ in range(2):
        ops.roi_pool_forward(feat, rois, row_scale=obj, row_scale_bias=1.0, want_f32=False, argmax_u16=True, out_bf16=X, plan=plan)
    torch.cuda.synchronize()
print("done")
