#!/usr/bin/env python
"""Diagnostics for the e2e arm of bench.py: pinned H2D bandwidth at the step's input size, and the host time one
plugin-surface step takes to ISSUE (no synchronisation inside)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda", 0)
shapes = [(2, 512, 60, 80), (2, 512, 72, 96)]
host = [torch.randn(s).pin_memory() for s in shapes]
nbytes = sum(t.numel() * 4 for t in host)
s = torch.cuda.Stream()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(10):
            d = [t.to(dev, non_blocking=True) for t in host]
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {nbytes/1e6:.1f} MB pinned: {ms:.3f} ms  -> {nbytes/ms/1e6:.1f} GB/s")
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d = big.to(dev, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"H2D 256 MiB pinned single copy: {big.numel()/e0.elapsed_time(e1)/1e6:.1f} GB/s")

