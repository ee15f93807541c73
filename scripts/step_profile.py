#!/usr/bin/env python
"""Kernel-level timeline of bench-like steps (torch.profiler / CUPTI): per-kernel GPU time inside a real,
back-to-back step (warm instruction caches, L2 flushed between steps).  python scripts/step_profile.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

dev = torch.device("cuda", 0)
P, H, C, L, T, B = 29, 128, 14, 4693, 1000, int(os.environ.get("MEMBERS", "256"))
prec = os.environ.get("PRECISION", "fp32")
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(P, H).to(dev).eval()
cond = torch.rand(1, C, L, device=dev).expand(B, C, L)
sched = [t.to(dev) for t in eb.get_diffusion_schedule(T)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step(i):
    x = eb.run_chain(model, cond, T, *sched, dev, seed=1, offset=4 * i, precision=prec)
    out = eb.ensemble_moments(x)
    out["pct"] = eb.ensemble_percentile(x, [2.5, 25.0, 50.0, 75.0, 97.5])
    out["mode"] = eb.ensemble_kde_mode(x, 5000)
    return out


for i in range(3):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(10):
        flush.zero_()
        step(i)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(c * t for _, c, t in rows if "FillFunctor" not in _)
for k, c, t in sorted(rows, key=lambda r: -r[1] * r[2]):
    print(f"{k[:80]:80s} n={c:3d} avg {t:8.1f} us")
print("sum per step (excl. flush):", tot / 10, "us")
