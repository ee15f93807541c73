#!/usr/bin/env python
"""Where one bench step goes (CUDA events around each public call, L2 flushed before each):
   python scripts/step_breakdown.py [--members 256] [--precision fp32]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=256)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
P, H, C, L = 29, 128, 14, 4693
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(P, H).to(dev).eval()
cond = torch.rand(1, C, L, device=dev)
sched = [t.to(dev) for t in eb.get_diffusion_schedule(a.T)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
x = eb.run_chain(model, cond.expand(a.members, C, L), a.T, *sched, dev, seed=1, precision=a.precision)


def timed(fn):
    ms = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)) * 1e3


parts = {
    "encode_condition (1 shared)": lambda: model.encode_condition(cond, precision=a.precision, return_bias=True),
    "run_chain (encoder + chain)": lambda: eb.run_chain(model, cond.expand(a.members, C, L), a.T, *sched, dev, seed=1, precision=a.precision),
    "ensemble_moments": lambda: eb.ensemble_moments(x),
    "ensemble_percentile x5": lambda: eb.ensemble_percentile(x, [2.5, 25.0, 50.0, 75.0, 97.5]),
    "ensemble_kde_mode": lambda: eb.ensemble_kde_mode(x, 5000),
    "empty (event pair only)": lambda: None,
}
for k, fn in parts.items():
    fn()
    print(f"{k:32s} {timed(fn):9.1f} us")
