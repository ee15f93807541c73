#!/usr/bin/env python
"""Misfit-metric kernels on one GPU: time per call (CUDA events) and achieved bandwidth against the
algorithmic bytes (each simulated map read once + the observed map), next to numpy on the host.
Development aid:  python scripts/misfit_bench.py [--members 50,512]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", default="50,512,4096")
a = ap.parse_args()
dev = torch.device("cuda", 0)
L, C = 4693, 14
for N in [int(v) for v in a.members.split(",")]:
    for dt in (torch.float32, torch.float64):
        g = torch.Generator(device=dev).manual_seed(0)
        obs = torch.randn(L, C, device=dev, dtype=dt, generator=g)
        sims = obs[None] + 0.3 * torch.randn(N, L, C, device=dev, dtype=dt, generator=g)
        for _ in range(3):
            eb.misfit_metrics(sims, obs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            eb.misfit_metrics(sims, obs)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        nbytes = (N + 1) * L * C * sims.element_size()
        # the two kernels on their own, straight through the C ABI
        from ertdiff_b200 import _lib
        lib = _lib.load()
        w, tot, m = (torch.empty(N, C, device=dev, dtype=dt), torch.empty(N, device=dev, dtype=dt),
                     torch.empty(N, device=dev, dtype=dt))
        code = _lib.F32 if dt == torch.float32 else _lib.F64
        each = []
        for outs in ((w, tot, None), (None, None, m)):
            ptrs = [_lib.ptr(o) if o is not None else None for o in outs]
            e0.record()
            for _ in range(reps):
                _lib.check(lib.ertdiff_misfit_metrics(_lib.ptr(sims), _lib.ptr(obs), code, N, L, C, 0.1, 0.01, *ptrs,
                                                      _lib.stream_ptr(dev)), "misfit")
            e1.record()
            torch.cuda.synchronize()
            each.append(e0.elapsed_time(e1) / reps * 1e3)
        line = f"N {N:5d} {str(dt)[6:]:8s} {us:9.1f} us  {nbytes / us * 1e-3:8.1f} GB/s algorithmic  [wsse {each[0]:7.1f} us, mse {each[1]:7.1f} us]"
        if N <= 512:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
            from oracle import stats_oracle as so
            s_np, o_np = sims.cpu().numpy(), obs.cpu().numpy()
            t0 = time.perf_counter()
            so.misfit_metrics(s_np, o_np)
            line += f"   numpy (reference arithmetic, 1 core) {(time.perf_counter() - t0) * 1e3:8.1f} ms"
        print(line, flush=True)
