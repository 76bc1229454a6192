#!/usr/bin/env python
"""N-rank proof of the data-parallel path (run under torchrun on N B200s): the head trained with
`set_gradient_exchange(mode)` + the attached B200SGD -- reduce-scatter / owned-row update on the update stream / bf16
operand all-gather, through the plugin surface and autograd -- against the same head trained the way the reference does
it (tools/train_net_multi.py:75-78,137-164): every rank averages every gradient with a plain NCCL all-reduce and runs
torch.optim.SGD on all parameters.  Same images per rank, same dropout seeds, three steps in lockstep; then the fp32
parameters (after the checkpoint-time sync), the momentum buffers and the bf16 operands must agree on every rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sos_wsod_b200.solver import build_optimizer, get_optimizer_param_groups  # noqa: E402
from sos_wsod_b200.structures import Boxes, Instances  # noqa: E402
from sos_wsod_b200.synthetic import pack_views, training_image  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "sharded"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bench.init_dist(dev)
    R = 600
    a, cfg = bench.build_heads(dev)          # the exchange path
    b, _ = bench.build_heads(dev)            # the reference path (same seed -> same initial parameters)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)
    a.set_gradient_exchange(mode=mode)
    opt_a = build_optimizer(cfg, a)
    if os.environ.get("CHECK_SYNC_EACH_STEP"):
        print("synchronising the device after every optimizer step (serialised pipeline)", file=sys.stderr)
    opt_b = torch.optim.SGD(get_optimizer_param_groups(cfg, b), cfg.SOLVER.BASE_LR, momentum=cfg.SOLVER.MOMENTUM)
    lr_boost = 50.0                            # make three steps move the weights visibly
    for o in (opt_a, opt_b):
        for g in o.param_groups:
            g["lr"] *= lr_boost
    sizes = [(480, 640), (576, 768)]
    n_steps = int(os.environ.get("CHECK_STEPS", "1"))      # 1: identical weights on both paths (no chaotic feedback)
    for step in range(n_steps):
        views, gt = training_image(step, rank, R=R, sizes=sizes, num_classes=bench.NUM_CLASSES)
        feats, rois, obj = pack_views(views)
        feats = [f.to(dev) for f in feats]
        props = [[Instances(v.image_size, proposal_boxes=Boxes(v.boxes.to(dev)), objectness_logits=v.obj.to(dev))] for v in views]
        targets = [Instances(views[0].image_size, gt_classes=gt.to(dev))]
        for heads, opt in ((a, opt_a), (b, opt_b)):
            heads.iter = step
            opt.zero_grad(set_to_none=True)
            _, losses = heads(None, [{"plain5": feats[0]}, {"plain5": feats[1]}], props, [targets, None, None, None])
            sum(losses.values()).backward()
            if heads is b:                      # DDP's arithmetic, by hand
                for p in heads.parameters():
                    dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
            opt.step()
            if os.environ.get("CHECK_SYNC_EACH_STEP"):
                torch.cuda.synchronize()
    sd_a = a.state_dict()                       # collective: brings the rows owned by other ranks up to date
    opt_a.sync_state()
    errs = {}
    sd_b = b.state_dict()
    for (n, pa), (_, pb) in zip(sd_a.items(), sd_b.items()):
        # box_predictor.det.bias has an identically vanishing gradient (softmax over the proposals): both paths hold
        # rounding noise only -> measured against the scale of the classification stream's bias
        scale = sd_b["box_predictor.cls.bias"].abs().max() if n == "box_predictor.det.bias" else pb.abs().max()
        errs[n] = float((pa - pb).abs().max() / scale.clamp_min(1e-30))
    moved = float((b.box_head.fc1.weight - bench.build_heads(dev)[0].box_head.fc1.weight).abs().max())
    mom = {}
    names = [n for n, _ in a.named_parameters()]
    cls_b_scale = opt_b.state[b.box_predictor.cls.bias]["momentum_buffer"].abs().max()
    for n, pa, pb in zip(names, a.parameters(), b.parameters()):
        ma, mb = opt_a.state[pa]["momentum_buffer"], opt_b.state[pb]["momentum_buffer"]
        scale = cls_b_scale if n == "box_predictor.det.bias" else mb.abs().max()
        mom[n] = float((ma - mb).abs().max() / scale.clamp_min(1e-30))
    op = a.engine().op
    a.engine().operand_gate and a.engine().operand_gate()
    opnd_ok = bool(torch.equal(op.w6, a.box_head.fc1.weight.detach().to(torch.bfloat16)) and
                   torch.equal(op.w7, a.box_head.fc2.weight.detach().to(torch.bfloat16)))
    worst = torch.tensor([max(errs.values()), max(mom.values())], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    ok = torch.tensor([int(opnd_ok)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": f"exchange + B200SGD vs all-reduce(AVG) + torch.optim.SGD, {n_steps} step(s)", "mode": mode, "n_gpus": world,
                          "max_rel_err_params": float(worst[0]), "max_rel_err_momentum": float(worst[1]),
                          "bf16_operands_equal_cast_of_masters_on_every_rank": bool(ok.item()),
                          "fc1_weight_moved_by": moved, "rank0_param_errs": {k: float(f"{v:.3g}") for k, v in errs.items()},
                          "rank0_momentum_errs": {str(k): float(f"{v:.3g}") for k, v in mom.items()}, "tolerance": ("1 step from identical weights: parameters 1e-5, momentum 1e-5 of the tensor's largest magnitude. "
                                        "3 steps (learning rates x50): the two paths' fp32 rounding differences (summation order of the "
                                        "collective, FMA contraction of the fused update) feed back through bf16 re-casts of the weights "
                                        "and through the discontinuous L1 / label decisions -- every exchange mode lands on the SAME "
                                        "values, bar 5e-2") ,
                          "ok": bool(worst[0] < (1e-5 if n_steps == 1 else 5e-2) and worst[1] < (1e-5 if n_steps == 1 else 5e-2) and ok.item() == 1)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
