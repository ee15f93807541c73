#!/usr/bin/env python
"""Condition-encoder timing on one GPU (CUDA events, L2 flushed between runs): fp32 CUDA-core kernel
vs the tensor-core (tcgen05 + TMA) kernel.  python scripts/encoder_bench.py [--conds 256,1024]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--conds", default="1,64,256,1024,4096")
ap.add_argument("--L", type=int, default=4693)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(29, 128).to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
FLOP = 20751232
for n in [int(v) for v in a.conds.split(",")]:
    cond = torch.rand(n, 14, a.L, device=dev)
    for prec in ("fp32", "bf16"):
        best = 1e9
        for _ in range(a.reps + 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            emb = model.encode_condition(cond, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if prec == "bf16" and os.environ.get("ENC_TIMING"):
            model.umma_timing(True)
            model.encode_condition(cond, precision=prec)
            tm = model.umma_timing(False)
            names = ["waitTMA", "convert", "sync+issue1", "waitMMA1", "epi1", "sync+issue2", "waitMMA2", "epi2"]
            print("   cycles/tile:", {k: round(tm[i] / max(tm[15], 1)) for i, k in enumerate(names)}, flush=True)
        gb = n * 14 * a.L * 4 / 1e9
        print(f"conds {n:5d} {prec}: {best * 1e3:9.1f} us  {n * FLOP / best / 1e9:8.1f} TFLOP/s  "
              f"{gb / best * 1e3:7.1f} GB/s of condition reads  status {model.umma_status()}", flush=True)
