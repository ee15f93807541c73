#!/usr/bin/env python
"""The fused statistics call (ertdiff_ensemble_summary) on a column window of an (N, 29) float32 array -- what one
rank of the column-sharded statistics runs per step -- timed as a whole and per kernel (torch.profiler), on ONE GPU.
    python scripts/summary_window_bench.py [--cases 2048x4,8192x4,151552x4,2048x29,8192x29,18944x29]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="2048x4,8192x4,151552x4,2048x29,8192x29,18944x29")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
QS = [2.5, 25.0, 50.0, 75.0, 97.5]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for case in a.cases.split(","):
    N, ncols = (int(v) for v in case.split("x"))
    # a chain-like output: every parameter spread over a few hundred units, a few narrow ones
    scale = torch.linspace(150.0, 600.0, 29, device=dev)
    scale[5], scale[17] = 0.5, 5.0
    x = torch.randn(N, 29, device=dev, generator=torch.Generator(dev).manual_seed(N)) * scale
    for _ in range(2):
        eb.ensemble_summary_packed(x, QS, 5000, col0=0, ncols=ncols)
    best = 1e30
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eb.ensemble_summary_packed(x, QS, 5000, col0=0, ncols=ncols)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.reps):
            eb.ensemble_summary_packed(x, QS, 5000, col0=0, ncols=ncols)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split("(")[0].replace("void ", "").replace("ertdiff::", "")[:60]
            agg[name][0] += 1
            agg[name][1] += ev.device_time
    parts = ", ".join(f"{k} {t / a.reps:.0f}" for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8])
    print(f"({N:6d} members, {ncols:2d} of 29 columns): {best * 1e3:8.1f} us per call | kernels (us): {parts}", flush=True)
