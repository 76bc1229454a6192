#!/usr/bin/env python
"""BASELINE configs[4] timing: test-time detection-result generation for one rank's block of images --
V = 2 x #scales views per image (5 scales + h-flip by default) x 2000 proposals, VOC shape: on-device proposal
transform (tta_views) -> ROI pool + fc6/fc7 + heads for all views in one pass -> predict -> inverse transform + mean
over views (tta_merge) -> threshold, per-class NMS, top-100 (detect).  conv5 maps are synthetic (the VGG16 backbone
is outside the hot path).  Prints one JSON line; CUDA-event timed, every image uses fresh feature buffers cycled
through a pool larger than L2.

    python scripts/bench_detect.py [--images 40] [--scales 480 576 672 768 864] [--classes 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sos_wsod_b200 import ops  # noqa: E402
from sos_wsod_b200.engine import HeadConfig, HeadOperands, OICRPlusHeadEngine, ViewBatch  # noqa: E402
from sos_wsod_b200.modeling.test_time_augmentation_avg import ViewSpec, resize_shortest_edge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scales", type=int, nargs="+", default=[480, 576, 672, 768, 864])
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--refine-k", type=int, default=3)
    ap.add_argument("--proposals", type=int, default=2000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    C, K, R = a.classes, a.refine_k, a.proposals
    H, W = 480, 640
    g = torch.Generator().manual_seed(1234)
    cfg = HeadConfig(num_classes=C, refine_k=K, dropout_p=0.0, score_thresh_test=1e-6 if C == 20 else 1e-5)
    fc1_w = (torch.randn((4096, 25088), generator=g) * 0.005).to(dev)
    fc2_w = (torch.randn((4096, 4096), generator=g) * 0.005).to(dev)
    b01 = torch.full((4096,), 0.1, device=dev)
    small = lambda n: (torch.randn((n, 4096), generator=g) * 0.02).to(dev)  # noqa: E731
    refine = [(small(C + 1), torch.zeros(C + 1, device=dev), small(4 * C) * 0.05, torch.zeros(4 * C, device=dev)) for _ in range(K)]
    op = HeadOperands(cfg, fc1_w, b01, fc2_w, b01.clone(), small(C), torch.zeros(C, device=dev), small(C), torch.zeros(C, device=dev), refine)
    eng = OICRPlusHeadEngine(cfg, op)
    specs = []
    for s in a.scales:
        nh, nw = resize_shortest_edge(H, W, s, 4000)
        specs += [ViewSpec(H, W, nh, nw, False), ViewSpec(H, W, nh, nw, True)]
    V = len(specs)
    params = [sp.params(float(i % 2)) for i, sp in enumerate(specs)]
    # a pool of synthetic images: per image one [2,512,h,w] conv5 tensor per scale (image + flip) and its proposals
    def synth_boxes(n):
        x1 = torch.rand(n, generator=g) * (W - 32)
        y1 = torch.rand(n, generator=g) * (H - 32)
        bw = 20 + torch.rand(n, generator=g) * (W - x1 - 20)
        bh = 20 + torch.rand(n, generator=g) * (H - y1 - 20)
        return torch.stack([x1, y1, x1 + bw, y1 + bh], 1).round()

    pool = []
    for i in range(4):
        feats = [torch.relu(torch.randn((2, 512, (sp.new_h + 7) // 8, (sp.new_w + 7) // 8), generator=g)).to(dev) for sp in specs[::2]]
        boxes = synth_boxes(R).to(dev)
        obj = torch.sort(torch.rand(R, generator=g), descending=True).values.to(dev)
        pool.append((feats, boxes, obj))

    def one_image(i):
        feats, boxes, obj = pool[i % len(pool)]
        rois, keep, dropped = ops.tta_views(boxes, params)
        rois = rois.view(V // 2, 2 * R, 5)
        vb = ViewBatch(feats, [rois[j] for j in range(V // 2)], obj.repeat(V), R)
        probs, pboxes = eng.test_forward(vb)
        mb, mp = ops.tta_merge(pboxes, probs, params)
        return eng.detect(mp, mb, (H, W)), dropped

    for i in range(a.warmup):
        one_image(i)
    torch.cuda.synchronize()
    n0 = ops.COUNTERS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nd_total = 0
    for i in range(a.images):
        (db, ds, dc, dr, nd), dropped = one_image(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.images
    print(json.dumps({"metric": "detection-result generation (TTA) images/s", "value": 1000.0 / ms, "unit": "images/s",
                      "ms_per_image": ms, "views_per_image": V, "image_views_per_s": V * 1000.0 / ms,
                      "proposals_per_s": V * R * 1000.0 / ms, "n_gpus": 1, "images": a.images, "dtype": "bf16",
                      "config": {"workload": f"cfg5 per rank: {V} views ({len(a.scales)} scales x flip) x {R} proposals, C={C}, K={K}, "
                                             "480x640 base image, synthetic conv5 maps", "scales": a.scales},
                      "gpu_launches_per_image": (ops.COUNTERS["launches"] - n0) / a.images,
                      "detections_last_image": int(nd.item()), "dropped_last_image": int(dropped.sum().item())}))


if __name__ == "__main__":
    main()
