#!/usr/bin/env python
"""Microbenchmark of the fused NVLS update kernel (soswsod_sgd_nvls) under torchrun: full-size fc1 / fc2 matrices, sweep
of the persistent grid size.  Prints per setting the kernel time (CUDA events, max over ranks) and the implied rates."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sos_wsod_b200 import ops  # noqa: E402
from sos_wsod_b200.distributed import GradientExchange  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bench.init_dist(dev)
    master = {"fc1_w": torch.randn(4096, 25088, device=dev) * 0.01, "fc2_w": torch.randn(4096, 4096, device=dev) * 0.01}
    ex = GradientExchange(master, mode="nvls", min_shard_elems=1)
    ex.setup_nvls()
    bufs = {k: torch.zeros_like(v) for k, v in master.items()}
    for k in master:
        ex.symm_tensors[f"g:{k}"].copy_(torch.randn_like(master[k]) * (rank + 1))
    torch.cuda.synchronize()
    dist.barrier()
    items = []
    for k in sorted(ex.sharded):
        lo, hi = ex.owned_rows_nvls(k)
        items.append((master[k][lo:hi], ex.multicast_address(f"g:{k}", lo), bufs[k][lo:hi], ex.multicast_address(f"w:{k}", lo), 0.0, 0.0))
    n_own = sum(it[0].numel() for it in items)
    rows = []
    for ctas in [8, 16, 32, 64, 148, 296, 592, 1184, 2368]:
        ops.NVLS_MAX_CTAS = ctas
        ts = []
        for it in range(6):
            ex.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.sgd_nvls(items, 0.0, 1.0 / world)
            e1.record()
            ex.barrier()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ts[1:])], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rows.append({"ctas": ctas, "ms": ms, "reduced_gradient_GBps": n_own * 4 / ms / 1e6, "elements_per_us": n_own / ms / 1e3})
        if rank == 0:
            print(rows[-1], flush=True)
    # correctness of the reduction itself: buf == mean over ranks of the gradients (momentum 0, lr 0 -> buf = g / world)
    want = ex.symm_tensors["g:fc2_w"].clone()
    dist.all_reduce(want, op=dist.ReduceOp.AVG)
    lo, hi = ex.owned_rows_nvls("fc2_w")
    err = float((bufs["fc2_w"][lo:hi] - want[lo:hi]).abs().max() / want.abs().max())
    if rank == 0:
        print(json.dumps({"n_gpus": world, "owned_elements_per_rank": n_own, "sweep": rows, "mean_gradient_max_rel_err_vs_nccl": err}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
