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
ap.add_argument("--chain", action="store_true", help="use the fields of a real T=1000 bf16 chain (bench weights) instead of synthetic columns")
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
    if a.chain:
        import bench
        spec = bench.Spec("w", N, "bf16")
        sd, cond = bench.synthetic_inputs(spec)
        model = eb.ConditionalDiffusionModel(29, 128)
        model.load_state_dict(sd)
        model.to(dev).eval()
        sched = [t.to(dev) for t in eb.get_diffusion_schedule(1000)]
        x = eb.run_chain(model, cond.to(dev).expand(N, 14, 4693), 1000, *sched, dev, seed=1234, offset=0, precision="bf16")
        sdv = x.std(dim=0)
        print(f"   chain fields: |x|max {x.abs().max().item():.0f}, column std min/median/max {sdv.min().item():.1f}/{sdv.median().item():.1f}/{sdv.max().item():.1f}")
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
