import sys, torch
sys.path.insert(0, '.')
from sos_wsod_b200 import ops
from oracle import oicr_plus_ref as ref
g = torch.Generator().manual_seed(0)
for (C,h,w,N,R) in [(16,150,200,1,120),(512,72,96,2,2000)]:
    feat = torch.relu(torch.randn((N,C,h,w), generator=g)).cuda()
    bl = [ref.synth_boxes(R, h*8, w*8, g) for _ in range(N)]
    rois = ref.boxes_to_pooler_format(bl).cuda()
    _, arg, x = ops.roi_pool_forward(feat, rois, want_f32=False, want_bf16=True, argmax_u16=True)
    torch.cuda.synchronize(); print('fwd ok', C,h,w)
    go = torch.randn((N*R, C*49), device='cuda').to(torch.bfloat16)
    gf = ops.roi_pool_backward(go, arg, rois, (N,C,h,w))
    torch.cuda.synchronize(); print('bwd ok', C,h,w, float(gf.abs().sum()))
