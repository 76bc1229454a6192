#!/usr/bin/env python
"""bench.py -- OICR+ head fwd+bwd throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1]): one training image of the reference = 4 image-views
(2 scales x h-flip, 480x640 and 576x768 -> conv5 60x80 and 72x96, 512 ch) x 2000 proposals, 20 classes, K=3
refinement branches, dropout 0.5 on, forward + backward including parameter gradients and the gradient w.r.t.
both conv5 maps.  Unit of the metric: image-views/s (1 image-view = one conv5 map + 2000 proposals); proposals/s
= 2000 x that; reference-style "training images/s" = that / 4.  With N > 1 each rank steps its own image (weak
scaling) and the parameter gradients are averaged with NCCL all-reduce, started per layer as soon as the
layer's gradient is produced.

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_PROPOSALS = 2000
NUM_CLASSES = 20
REFINE_K = 3
SIZES = [(480, 640), (576, 768)]
VIEWS = 4
WORKLOAD = ("cfg2: OICR+ head fwd+bwd, VOC07 shape, 1 training image = 4 image-views (480x640 + 576x768, each with "
            "its h-flip) x 2000 proposals, C=20, K=3, dropout 0.5, grads to all head params and both conv5 maps")
METRIC = "OICR+ head fwd+bwd images/s"
UNIT = "image-views/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0), "bf16_tflops": p.get("bf16_tflops", 1590.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# ----------------------------------------------------------------------------------------------------
def make_host_images(n_images: int, rank: int):
    """Pinned host tensors of n_images training images: feats (2 x [2,512,h,w]), rois (2 x [2R,5]), obj [4R], gt."""
    from oracle import oicr_plus_ref as ref   # only the seeded input generator is used here

    images = []
    for i in range(n_images):
        g = torch.Generator().manual_seed(1234 + 200 + rank * 17 + i)
        views = ref.synth_views(R_PROPOSALS, SIZES, g)
        feats = [torch.cat([views[0].feat, views[1].feat], 0), torch.cat([views[2].feat, views[3].feat], 0)]
        rois = []
        for a, b in ((0, 1), (2, 3)):
            r0 = torch.cat([torch.zeros(R_PROPOSALS, 1), views[a].boxes], 1)
            r1 = torch.cat([torch.ones(R_PROPOSALS, 1), views[b].boxes], 1)
            rois.append(torch.cat([r0, r1], 0))
        obj = torch.cat([v.obj for v in views])
        ng = int(torch.randint(1, 5, (1,), generator=g))
        gt = torch.sort(torch.randperm(NUM_CLASSES, generator=g)[:ng]).values
        pin = (lambda t: t.contiguous().pin_memory()) if torch.cuda.is_available() else (lambda t: t.contiguous())
        images.append({"feats": [pin(f) for f in feats], "rois": [pin(r) for r in rois], "obj": pin(obj), "gt": gt,
                       "views": views})
    return images


def synth_views_sizes():
    return [(SIZES[0][0] // 8, SIZES[0][1] // 8), (SIZES[1][0] // 8, SIZES[1][1] // 8)]


# ----------------------------------------------------------------------------------------------------
# clocks sampling
# ----------------------------------------------------------------------------------------------------
PRE_WARMUP = int(os.environ.get("SOSWSOD_PRE_WARMUP", "8"))   # extra untimed steps before the caller's --warmup steps
SETTLE_MAX_BLOCKS = int(os.environ.get("SOSWSOD_SETTLE_BLOCKS", "12"))   # 0 under a profiler (every launch is replayed)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("SOSWSOD_NO_SMI"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("SOSWSOD_SMI_MS", "100")],
                                         stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def wait_first_sample(self, timeout=15.0):
        """nvidia-smi takes a second or two to attach to the driver (and slows CUDA calls while it does): the
        timed region only starts once it is in its steady 100 ms polling loop."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, ln in self.lines:
            if ts < t_mark:
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's path (oracle port: torchvision roi_pool + PyTorch head) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step(views, gt, params, drop_masks=None):
    from oracle import oicr_plus_ref as ref

    for v in views:
        v.feat.requires_grad_(True)
        v.feat.grad = None
    for t in params.tensors():
        t.grad = None
    losses, _ = ref.train_step(views, gt, params, NUM_CLASSES, REFINE_K, drop_masks=drop_masks)
    sum(losses.values()).backward()
    return float(sum(losses.values()))


def cpu_sample_setup(target_seconds: float):
    """Chooses the proposals-per-view of the CPU sample so that one fwd+bwd of the 4-view step takes about
    `target_seconds` on this host (calibrated with a small fc6 GEMM)."""
    from oracle import oicr_plus_ref as ref

    torch.set_num_threads(os.cpu_count() or 1)
    a = torch.randn(256, 25088)
    w = torch.randn(4096, 25088)
    torch.mm(a, w.t())
    t0 = time.perf_counter()
    torch.mm(a, w.t())
    dt = time.perf_counter() - t0
    gflops = 2 * 256 * 25088 * 4096 / dt / 1e9
    # fwd+bwd fc FLOPs per proposal: 717.2 MFLOP; the rest of the path adds ~30 % on CPU
    per_prop = 717.2e6 * 1.3
    r = int(target_seconds * gflops * 1e9 / per_prop / VIEWS)
    r = max(50, min(R_PROPOSALS, r))
    g = torch.Generator().manual_seed(1234 + 100)
    views = ref.synth_views(r, SIZES, g)
    params = ref.init_head_params(NUM_CLASSES, REFINE_K, generator=g).requires_grad_(True)
    gt = torch.tensor([2, 7, 7, 14])
    return views, gt, params, r, gflops


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the whole --steps K --warmup W run stays within ~2-3 minutes of host work, one step at most the
    # full 2000-proposal workload (~7 s on 16 cores)
    n_warm = max(1, min(args.warmup, 2))
    budget = float(os.environ.get("SOSWSOD_REF_BUDGET_SECONDS", "100"))
    views, gt, params, r, gflops = cpu_sample_setup(min(6.0, budget / (args.steps + n_warm)))
    for _ in range(n_warm):
        cpu_reference_step(views, gt, params)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(views, gt, params)
    dt = (time.perf_counter() - t0) / args.steps
    value = VIEWS * (r / R_PROPOSALS) / dt          # image-views/s normalised to 2000 proposals per view
    cores = os.cpu_count() or 1
    sample = f"{VIEWS} views x {r} proposals per step (scaled to 2000/view), fwd+bwd, C={NUM_CLASSES}, K={REFINE_K}, fp32, eval-mode dropout"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "proposals_per_s": value * R_PROPOSALS,
            "config": {"workload": WORKLOAD, "sample": sample, "host_threads": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_fc6_gflops": gflops},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def build_heads(device):
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling import build_roi_heads
    from sos_wsod_b200.structures import ShapeSpec

    cfg = get_cfg()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = NUM_CLASSES
    cfg.WSL.REFINE_NUM = REFINE_K
    torch.manual_seed(1234)
    heads = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=512, stride=8)}).to(device)
    heads.train()
    return heads


def run_b200_arm(args):
    import torch.distributed as dist

    from sos_wsod_b200 import _lib, ops
    from sos_wsod_b200.engine import ViewBatch
    from sos_wsod_b200.structures import Boxes, Instances

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); there is no CPU fallback for the sm_100a path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()
    if world > 1:
        # stdout carries ONE json line: NCCL prints its version banner there at any debug level >= VERSION
        os.environ.pop("NCCL_DEBUG", None)
        if "SOSWSOD_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["SOSWSOD_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    peaks = _peaks()

    heads = build_heads(dev)
    eng = heads.engine()
    n_img = 3
    host = make_host_images(n_img, rank)
    dev_imgs = [{"feats": [f.to(dev) for f in im["feats"]], "rois": [r.to(dev) for r in im["rois"]],
                 "obj": im["obj"].to(dev), "gt": im["gt"].to(dev)} for im in host]

    # ---- data-parallel gradient averaging: per-layer async all-reduce on NCCL's stream ----
    works = []

    def grad_hook(name, tensors):
        if world > 1:
            for t in tensors:
                works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True))

    def wait_grads():
        for w in works:
            w.wait()
        works.clear()

    # ---- per-GEMM CUDA-event timing (roofline of the dominant kernel) ----
    gemm_events = []
    orig_gemm = ops.gemm_bf16
    record_gemm = {"on": False}

    def timed_gemm(a, b, **kw):
        if not record_gemm["on"]:
            return orig_gemm(a, b, **kw)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_gemm(a, b, **kw)
        e1.record()
        m, n = out.shape
        k = a.shape[0] if kw.get("a_mn") else a.shape[1]
        gemm_events.append((e0, e1, 2.0 * m * n * k, (m, n, k, bool(kw.get("a_mn")), bool(kw.get("b_mn")))))
        return out

    ops.gemm_bf16 = timed_gemm

    # ---- the ROI-pool kernels, timed the same way (secondary roofline: HBM-side compulsory bytes / duration) ----
    roi_events = {"fwd": [], "bwd": []}
    orig_fwd, orig_bwd = ops.roi_pool_forward, ops.roi_pool_backward

    def timed_roi(kind, orig):
        def fn(*a, **kw):
            if not record_gemm["on"]:
                return orig(*a, **kw)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(*a, **kw)
            e1.record()
            roi_events[kind].append((e0, e1))
            return out
        return fn

    ops.roi_pool_forward = timed_roi("fwd", orig_fwd)
    ops.roi_pool_backward = timed_roi("bwd", orig_bwd)

    def device_step(i):
        im = dev_imgs[i % n_img]
        vb = ViewBatch(im["feats"], im["rois"], im["obj"], R_PROPOSALS)
        out = eng.train_step(vb, im["gt"], dropout_seeds=(2 * i + 1, 2 * i + 2), need_feat_grad=True,
                             grad_hook=grad_hook if world > 1 else None)
        wait_grads()
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def settle(step_fn, block=3, max_blocks=SETTLE_MAX_BLOCKS, tol=0.03):
        """Untimed: repeats blocks of `block` steps until two consecutive blocks take the same time within `tol` (max
        over ranks), i.e. until transients that are not ours have died down (the previous process's memory still being
        scrubbed by the driver, clocks leaving idle).  Every rank takes the same decision.  Returns the blocks run."""
        prev = None
        for nb in range(1, max_blocks + 1):
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            sync_all()
            a.record()
            for j in range(block):
                step_fn(j)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cur = float(t.item())
            if prev is not None and abs(cur - prev) <= tol * prev:
                return nb
            prev = cur
        return max_blocks

    eng.fc1_wgrad_panels = int(os.environ.get("SOSWSOD_FC1_PANELS", eng.fc1_wgrad_panels))

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    # a fresh box / process runs its first steps slower (lazy module loads, allocator growth, clocks leaving idle):
    # PRE_WARMUP extra untimed steps come before the W warm-up steps the caller asked for
    for i in range(PRE_WARMUP + args.warmup):
        device_step(i)
    settle_blocks = settle(lambda j: device_step(j))
    sampler.wait_first_sample()
    sync_all()
    sampler.mark()
    launches0 = ops.COUNTERS["launches"]
    record_gemm["on"] = True
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    # the host stays at most two steps ahead of the device (it issues a step in ~2 ms, the device runs it in ~7 ms):
    # the same pacing the e2e loop gets from its lagged loss read, without ever starving the device
    pace = []
    for i in range(args.steps):
        if len(pace) >= 2:
            pace.pop(0).synchronize()
        out = device_step(PRE_WARMUP + args.warmup + i)
        ev = torch.cuda.Event()
        ev.record()
        pace.append(ev)
    t_end.record()
    sync_all()
    record_gemm["on"] = False
    clocks = sampler.stop()
    launches = ops.COUNTERS["launches"] - launches0
    ms = t_start.elapsed_time(t_end)
    loss_val = float(sum(v.item() for v in out.losses.values()))
    if not (loss_val == loss_val):
        raise RuntimeError("non-finite loss in the timed region")
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = world * VIEWS / (ms_per_step / 1e3)

    # roofline of the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs / event-timed launch duration
    tot_flops = sum(e[2] for e in gemm_events)
    tot_ms = sum(e[0].elapsed_time(e[1]) for e in gemm_events)
    by_shape = {}
    for e0, e1, fl, shp in gemm_events:
        d = by_shape.setdefault(shp, [0.0, 0.0, 0])
        d[0] += fl
        d[1] += e0.elapsed_time(e1)
        d[2] += 1
    detail = [{"m": s[0], "n": s[1], "k": s[2], "a_mn": s[3], "b_mn": s[4], "launches": d[2],
               "avg_ms": d[1] / d[2], "tflops": d[0] / d[1] / 1e9} for s, d in by_shape.items()]
    fc6 = max(detail, key=lambda d: d["m"] * d["n"] * d["k"] if not d["a_mn"] and not d["b_mn"] else 0)
    achieved = tot_flops / tot_ms / 1e9 if tot_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    n_gemm = len(gemm_events) // max(args.steps, 1)
    # DRAM traffic of the same launches from the committed `ncu --set full` capture (profiles/ncu_traffic.json);
    # averaged per launch like `achieved`.  Algorithmic bytes = operands read once + output written once.
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["gemm_bf16_kernel"]["dram_bytes_per_launch_avg"]
        traffic_src = tj["source"]
    alg_bytes = sum(2.0 * (s[0] * s[2] + s[1] * s[2]) + 2.0 * s[0] * s[1] for s in by_shape) / max(len(by_shape), 1)
    roofline = {"bound": "tensor", "kernel": f"gemm_bf16_kernel (tcgen05, all {n_gemm} launches of a step)", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": peaks["source"] + ", sustained",
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (avg over the step's GEMM launches)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                "gemm_share_of_step": tot_ms / ms, "fc6_fwd_tflops": fc6["tflops"], "detail": detail}

    # ROI pool: per step 2 forward + 2 backward launches (one per scale pair).  Compulsory HBM bytes per step:
    # forward = conv5 planes in + (bf16 operand + uint16 arg-max) out; backward = (bf16 grad + uint16 arg-max) in + fp32 planes out
    plane_bytes = sum(2 * 512 * h * w * 4 for (h, w) in synth_views_sizes())
    xa_bytes = VIEWS * R_PROPOSALS * 25088 * (2 + 2)
    roi = {}
    for kind in ("fwd", "bwd"):
        tot = sum(e0.elapsed_time(e1) for e0, e1 in roi_events[kind])
        n_l = max(len(roi_events[kind]), 1)
        per_step_ms = tot / max(args.steps, 1)
        roi[kind] = {"launches_per_step": len(roi_events[kind]) // max(args.steps, 1), "ms_per_step": per_step_ms,
                     "avg_us_per_launch": 1e3 * tot / n_l,
                     "compulsory_hbm_gbs": (plane_bytes + xa_bytes) / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0,
                     "frac_of_hbm_peak": ((plane_bytes + xa_bytes) / (per_step_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if per_step_ms > 0 else 0.0}
    roi["share_of_step"] = (roi["fwd"]["ms_per_step"] + roi["bwd"]["ms_per_step"]) / ms_per_step
    roi["bound"] = "shared-memory gather / read-add-write inside the SM (planes staged in smem); HBM carries only the compulsory bytes"
    roofline["roi_pool"] = roi
    ops.roi_pool_forward, ops.roi_pool_backward = orig_fwd, orig_bwd

    # ---- end to end through the plugin surface, host buffers, H2D/D2H inside the timed region ----
    heads.grad_hook = grad_hook if world > 1 else None
    params = [p for p in heads.parameters()]
    h2d = sum(t.numel() * t.element_size() for t in host[0]["feats"] + host[0]["rois"] + [host[0]["obj"]]) + host[0]["gt"].numel() * 8
    image_sizes = [SIZES[0], SIZES[0], SIZES[1], SIZES[1]]

    # Inputs of step i+1 are copied host->device on a side stream while step i computes (the reference's DataLoader
    # prefetches batches the same way); every step's copy is issued and completes inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    copy_events = []

    def stage_inputs(i):
        im = host[i % n_img]
        with torch.cuda.stream(copy_stream):
            c0 = torch.cuda.Event(enable_timing=True)
            c0.record(copy_stream)
            feats = [f.to(dev, non_blocking=True) for f in im["feats"]]
            rois = [r.to(dev, non_blocking=True) for r in im["rois"]]
            obj = im["obj"].to(dev, non_blocking=True)
            gt = im["gt"].to(dev, non_blocking=True)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(copy_stream)
            copy_events.append((c0, ev))
        for t in feats + rois + [obj, gt]:
            t.record_stream(main_stream)
        return feats, rois, obj, gt, ev

    staged = {}
    heads.loss_scale_check = "deferred"
    n_loss = 1 + 2 * REFINE_K
    loss_bufs = [torch.empty(n_loss, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    pending_loss, host_seen = [], []

    host_time = {"issue": 0.0, "wait": 0.0}

    def e2e_step(i, last=False):
        t_issue = time.perf_counter()
        feats, rois, obj, gt, ev = staged.pop(i) if i in staged else stage_inputs(i)
        if not last:
            staged[i + 1] = stage_inputs(i + 1)
        main_stream.wait_event(ev)
        feats = [f.requires_grad_(True) for f in feats]
        props = []
        for v in range(VIEWS):
            rr = rois[v // 2][(v % 2) * R_PROPOSALS:(v % 2 + 1) * R_PROPOSALS, 1:5]
            props.append([Instances(image_sizes[v], proposal_boxes=Boxes(rr),
                                    objectness_logits=obj[v * R_PROPOSALS:(v + 1) * R_PROPOSALS])])
        targets = [Instances(image_sizes[0], gt_classes=gt)]
        for p in params:
            p.grad = None
        _, losses = heads(None, [{"plain5": feats[0]}, {"plain5": feats[1]}], props, [targets, None, None, None])
        total = sum(losses.values())
        total.backward()
        wait_grads()
        # D2H of the step's result: queued behind the step into pinned memory, read on the host one step later so
        # that the host keeps issuing step i+1 while the device runs step i (every step's losses are read inside
        # the timed region; the last one before the closing synchronisation)
        dl = torch.stack([losses[k].detach() for k in sorted(losses)])
        hbuf = loss_bufs[i % 2]
        hbuf.copy_(dl, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        prev = pending_loss.pop() if pending_loss else None
        pending_loss.append((hbuf, ev))
        t_wait = time.perf_counter()
        host_time["issue"] += t_wait - t_issue
        if prev is not None:
            prev[1].synchronize()
            host_seen.append(float(prev[0].sum()))
        if last:
            ev.synchronize()
            host_seen.append(float(hbuf.sum()))
            pending_loss.clear()
            heads.check_deferred(wait=True)
        host_time["wait"] += time.perf_counter() - t_wait
        return hbuf

    for i in range(max(3, args.warmup)):
        e2e_step(i, last=(i == max(3, args.warmup) - 1))
    settle_blocks_e2e = settle(lambda j: e2e_step(j, last=(j == 2)))
    staged.clear()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    host_seen.clear()
    copy_events.clear()
    host_time["issue"] = host_time["wait"] = 0.0
    for i in range(args.steps):
        hl = e2e_step(i, last=(i == args.steps - 1))
    t1.record()
    sync_all()
    if len(host_seen) != args.steps or not all(v == v for v in host_seen):
        raise RuntimeError(f"e2e: read {len(host_seen)} finite step results on the host, expected {args.steps}")
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(e2e_ms.item()) / args.steps
    e2e = {"value": world * VIEWS / (e2e_ms_per_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(hl.numel() * 4), "ms_per_step": e2e_ms_per_step,
           "host_issue_ms_per_step": 1e3 * host_time["issue"] / args.steps,
           "host_wait_ms_per_step": 1e3 * host_time["wait"] / args.steps,
           "h2d_copy_ms_per_step": sum(a.elapsed_time(b) for a, b in copy_events) / max(len(copy_events), 1),
           "api": "OICRPlusHeads.forward(images, features, proposals, targets) + sum(losses).backward()",
           "h2d": "pinned host buffers, copied on a side stream one step ahead (double-buffered), inside the timed region",
           "d2h": "each step's loss vector copied to pinned memory behind an event and read on the host one step later"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        views, gt, cparams, r, gflops = cpu_sample_setup(8.0)
        cpu_reference_step(views, gt, cparams)
        t0c = time.perf_counter()
        reps = 2
        for _ in range(reps):
            cpu_reference_step(views, gt, cparams)
        dtc = (time.perf_counter() - t0c) / reps
        cpu_baseline = {"value": VIEWS * (r / R_PROPOSALS) / dtc, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{VIEWS} views x {r} proposals (scaled to 2000/view), fwd+bwd, fp32, torch "
                                  f"{torch.__version__} CPU + torchvision roi_pool, {reps} timed steps after 1 warm-up",
                        "host_fc6_gflops": gflops}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "proposals_per_s": value * R_PROPOSALS,
                "training_images_per_s": value / VIEWS,
                "config": {"workload": WORKLOAD, "per_gpu": "1 image (4 views) per step", "l2": "per-step working set "
                           "(bf16 pooled operand 401 MB + dgrad 401 MB + fc6 weights/grads 616 MB) >> 126 MB L2; 3 "
                           "distinct synthetic images are cycled", "parallelism": f"dp{world}",
                           "extra_untimed_warmup_steps": PRE_WARMUP + 3 * settle_blocks,
                           "extra_untimed_warmup_steps_e2e": 3 * settle_blocks_e2e,
                           "allreduce": "NCCL AVG per layer, async, overlapped with the remaining backward" if world > 1 else "none",
                           "fc_flops_per_step": 3 * 2.0 * VIEWS * R_PROPOSALS * (25088 * 4096 + 4096 * 4096)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "loss": loss_val}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shape", default="voc", choices=["voc", "coco"],
                    help="voc = BASELINE configs[1] (the bench line); coco = configs[3] (80 classes), a side measurement")
    args = ap.parse_args()
    if args.shape == "coco":
        global NUM_CLASSES, WORKLOAD
        NUM_CLASSES = 80
        WORKLOAD = WORKLOAD.replace("cfg2:", "cfg4:").replace("VOC07 shape", "COCO shape").replace("C=20", "C=80")
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
