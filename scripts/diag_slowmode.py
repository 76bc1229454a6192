#!/usr/bin/env python
"""Diagnosis of the bimodal step time of round 1 (fc6 wgrad 1.17 ms in one run, 2.25 ms in
another on the same commit).  From process start, with NO settle loop, times consecutive blocks of 10 fwd+bwd steps and
the fc6 weight-gradient GEMM inside them, rotating four configurations:

  base        the round-1 step (fresh torch.empty gradient buffers, bias-gradient column sums on a side stream,
              one tile-scheduler workspace per stream)
  noside      column sums on the main stream
  persist     the big weight gradients written into persistent buffers
  schedrot    a different tile-scheduler workspace for every GEMM launch

Writes gpurun_out/diag/slowmode.json: per block (config, ms/step, fc6 wgrad ms, fc6 dgrad ms, SM clock samples)."""
import json
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sos_wsod_b200 import ops  # noqa: E402
from sos_wsod_b200.engine import ViewBatch  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    t_start = time.perf_counter()
    smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    clock_lines = []
    threading.Thread(target=lambda: [clock_lines.append((time.perf_counter() - t_start, ln.strip())) for ln in smi.stdout],
                     daemon=True).start()
    heads, cfg = bench.build_heads(dev)
    eng = heads.engine()
    host = bench.make_host_images(3, 0)
    imgs = [{"feats": [f.to(dev) for f in im["feats"]], "rois": [r.to(dev) for r in im["rois"]], "obj": im["obj"].to(dev),
             "gt": im["gt"].to(dev)} for im in host]
    orig_gemm = ops.gemm_bf16
    marks = []

    def gemm(a, b, **kw):
        big = a.numel() * b.numel() > 1e15          # fc6-sized operands
        if not big:
            return orig_gemm(a, b, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_gemm(a, b, **kw)
        e1.record()
        kind = "wgrad" if kw.get("a_mn") else ("dgrad" if kw.get("b_mn") else "fwd")
        marks.append((kind, e0, e1))
        return out

    ops.gemm_bf16 = gemm
    configs = ["base", "noside", "persist", "schedrot"]

    def apply(name):
        eng.bias_on_side_stream = name != "noside"
        eng.persistent_grads = name == "persist"
        ops.GEMM_SCHED_SLOTS = 16 if name == "schedrot" else 1

    rows = []
    step = 0
    for rnd in range(int(os.environ.get("DIAG_ROUNDS", "10"))):
        for name in configs:
            apply(name)
            marks.clear()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter() - t_start
            e0.record()
            keep = []
            for i in range(10):
                im = imgs[step % 3]
                out = eng.train_step(ViewBatch(im["feats"], im["rois"], im["obj"], 2000), im["gt"], dropout_seeds=(2 * step + 1, 2 * step + 2))
                keep = [out]
                step += 1
                if i % 2 == 1:
                    torch.cuda.current_stream().synchronize()
            e1.record()
            torch.cuda.synchronize()
            per = {}
            for kind, a, b in marks:
                per.setdefault(kind, []).append(a.elapsed_time(b))
            rows.append({"t": round(w0, 3), "round": rnd, "config": name, "ms_per_step": e0.elapsed_time(e1) / 10,
                         **{k: {"min": min(v), "max": max(v), "mean": sum(v) / len(v)} for k, v in per.items()}})
            print(rows[-1], flush=True)
        if rnd == 4:
            time.sleep(2.0)       # an idle gap: do the clocks / the slow mode come back?
    smi.terminate()
    os.makedirs("gpurun_out/diag", exist_ok=True)
    with open("gpurun_out/diag/slowmode.json", "w") as f:
        json.dump({"rows": rows, "clocks": clock_lines}, f)


if __name__ == "__main__":
    main()
