#!/usr/bin/env python
"""Chain-kernel sweep on one GPU: per (members, precision) the persistent kernel's duration (CUDA
events inside the library), ms per denoiser step and the FLOP rate.  Development aid, not a bench
line:  python scripts/chain_sweep.py [--T 1000] [--members 256,1024,...] [--distinct]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--members", default="256,1024,4096,8192,18944,37888")
ap.add_argument("--precisions", default="fp32,bf16")
ap.add_argument("--distinct", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--timing", action="store_true")
ap.add_argument("--flush", action="store_true", help="evict L2 (256 MiB memset) before every run, as bench.py does")
ap.add_argument("--mpc", default="", help="bf16 only: comma list of members-per-CTA overrides (32,64,128) to sweep")
a = ap.parse_args()
dev = torch.device("cuda", 0)
P, H, C, L = 29, 128, 14, 4693
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(P, H).to(dev).eval()
sched = [t.to(dev) for t in eb.get_diffusion_schedule(a.T)]
model.profile_chain(True)
for B in [int(v) for v in a.members.split(",")]:
    n = min(B, 512) if a.distinct else 1
    cond = torch.rand(n, C, L if not a.distinct else 256, device=dev)
    cond_b = cond.expand(B, C, cond.size(2)) if n == 1 else cond
    for prec, mpc in [(p_, m_) for p_ in a.precisions.split(",")
                      for m_ in (a.mpc.split(",") if (a.mpc and p_ == "bf16") else [""])]:
        if mpc:
            os.environ["ERTDIFF_UMMA_MPC"] = mpc
        else:
            os.environ.pop("ERTDIFF_UMMA_MPC", None)
        ms = []
        for i in range(a.reps + 1):
            if a.flush:
                if "flush_buf" not in globals():
                    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
                flush_buf.zero_()
            x = eb.run_chain(model, cond_b, a.T, *sched, dev, seed=7, offset=i, precision=prec,
                             n_members=B)
            ms.append(model.last_chain_ms())
        assert torch.isfinite(x).all()
        best = min(ms[1:])
        if prec in ("bf16", "bf16x3") and a.timing:
            model.umma_timing(True)
            eb.run_chain(model, cond_b, a.T, *sched, dev, seed=7, offset=0, precision=prec, n_members=B)
            tm = model.umma_timing(False)
            n = max(tm[15], 1)
            names = ["top", "waitD", "epi1", "waitZ", "waitE", "epi2+pub", "epi1:proxyfence", "epi1:arrive", "mma:waitX", "mma:gemm1", "mma:waitH+gemm2", "rng:waitEmpty", "rng:generate"]
            print("   cycles/step:", {k: round(tm[i] / n) for i, k in enumerate(names) if k}, flush=True)
        print(f"B {B:6d} {prec}{' mpc ' + mpc if mpc else ''} chain_ms {best:8.3f} us/step {best / a.T * 1e3:7.3f} "
              f"samples/s {B / best * 1e3:12.0f} TFLOP/s {B * a.T * 14848 / best / 1e9:7.2f} "
              f"|x|max {x.abs().max().item():.2f} umma_status {model.umma_status()}", flush=True)
