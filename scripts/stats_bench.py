#!/usr/bin/env python
"""Statistics kernels on the reference's own L4 workload (SURVEY.md §8 d): an ensemble of simulated ERT maps
(N, 4693, 14) float64 (ECD.py:716, 747-762, 867-872) from `np.random.default_rng(3).lognormal`, N in {50, 1024,
8192}; plus the chain's own (N, 29) float32 output for N up to 151,552; plus the standalone posterior update.
Per call: CUDA-event time (best of reps, L2 flushed before every call), algorithmic GB/s against the measured HBM
peak, and exp-evaluations/s for the KDE mode (its useful work is N * 5000 kernel evaluations per pixel; the scan
skips grid points that provably cannot hold the maximum, so the *algorithmic* rate can exceed the MUFU rate).
    python scripts/stats_bench.py [--maps 50,1024,8192] [--fields 256,8192,18944,151552] [--reps 3] [--only kde]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--maps", default="50,1024,8192")
ap.add_argument("--fields", default="256,8192,18944,151552")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", default="")
ap.add_argument("--pixels", type=int, default=4693 * 14)
a = ap.parse_args()
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
G = 5000


def timed(fn):
    best = 1e30
    for _ in range(a.reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def line(what, ms, nbytes=None, evals=None):
    s = f"{what:58s} {ms:10.3f} ms"
    if nbytes is not None:
        gbs = nbytes / ms / 1e6
        s += f"  {gbs:8.1f} GB/s algorithmic = {100 * gbs / HBM:5.1f}% of {HBM:.0f}"
    if evals is not None:
        s += f"  {evals / ms / 1e9:8.2f} T exp-evaluations/s algorithmic"
    print(s, flush=True)


def run(tag, x):
    N, Q = x.shape
    esz = x.element_size()
    nb = N * Q * esz
    want = a.only.split(",") if a.only else ["moments", "pct", "kde"]
    if "moments" in want:
        ms, _ = timed(lambda: eb.ensemble_moments(x, std=False, var=False))
        line(f"{tag} mean", ms, nb)
        ms, _ = timed(lambda: eb.ensemble_moments(x))
        line(f"{tag} mean+std+var (two passes)", ms, 2 * nb)
    if "pct" in want:
        ms, _ = timed(lambda: eb.ensemble_percentile(x, [25.0, 50.0, 75.0]))
        line(f"{tag} percentiles 25/50/75", ms, nb)
    if "kde" in want and N >= 2:
        ms, _ = timed(lambda: eb.ensemble_kde_mode(x, G))
        line(f"{tag} KDE mode (5000-point grid)", ms, evals=float(N) * G * Q)


if a.maps:
    for N in [int(v) for v in a.maps.split(",")]:
        rng = np.random.default_rng(3)
        host = rng.lognormal(size=(N, a.pixels))                 # float64, as ECD.py:716
        x = torch.from_numpy(host).to(dev)
        del host
        run(f"maps ({N},{a.pixels}) f64", x)
        del x
        torch.cuda.empty_cache()
if a.fields:
    for N in [int(v) for v in a.fields.split(",")]:
        x = (torch.randn(N, 29, device=dev, generator=torch.Generator(dev).manual_seed(N)) *
             torch.logspace(-1, 3, 29, device=dev))              # parameters of very different spread, like a chain's output
        run(f"fields ({N},29) f32", x)
if not a.only or "update" in a.only:
    n = 1 << 26
    xs = [torch.randn(n, device=dev) for _ in range(3)]
    ms, _ = timed(lambda: eb.posterior_update(xs[0], xs[1], xs[2], 0.01, 1.01, 0.1))
    line("posterior update, 2^26 elements (x, eps, z in; x out)", ms, 16 * n)
