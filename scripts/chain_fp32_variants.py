#!/usr/bin/env python
"""fp32 chain kernel: members-per-CTA x hidden-units-per-thread variants and the arithmetic-free floor build,
per ensemble size.  Kernel duration from CUDA events inside the library (persistent loop), best of `reps`.
    python scripts/chain_fp32_variants.py [--T 1000] [--members 128,148,256,296,512,1024] [--hidden 128]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--members", default="128,148,256,296,512,1024")
ap.add_argument("--hidden", type=int, default=128)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--L", type=int, default=4693)
a = ap.parse_args()
dev = torch.device("cuda", 0)
P, H, C = 29, a.hidden, 14
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(P, H).to(dev).eval()
sched = [t.to(dev) for t in eb.get_diffusion_schedule(a.T)]
model.profile_chain(True)
cond = torch.rand(1, C, a.L, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"# H={H} T={a.T}: us per denoiser step (kernel duration / T), full | floor")
for B in [int(v) for v in a.members.split(",")]:
    cond_b = cond.expand(B, C, a.L)
    for mpb, upt in ((1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (4, 2)):
        os.environ["ERTDIFF_CHAIN_MPB"], os.environ["ERTDIFF_CHAIN_UPT"] = str(mpb), str(upt)
        res = {}
        for floor in (False, True):
            if floor and mpb > 2:
                continue
            model.chain_floor(floor)
            ms = []
            for i in range(a.reps + 1):
                flush.zero_()
                x = eb.run_chain(model, cond_b, a.T, *sched, dev, seed=7, offset=4 * i, n_members=B)
                ms.append(model.last_chain_ms())
            res[floor] = min(ms[1:])
        model.chain_floor(False)
        ctas = (B + mpb - 1) // mpb
        fl = f"{res[True] / a.T * 1e3:7.3f}" if True in res else "    n/a"
        print(f"B {B:5d} mpb {mpb} upt {upt} CTAs {ctas:5d} x {H // upt:3d} thr  full {res[False] / a.T * 1e3:7.3f}  floor {fl}  "
              f"samples/s {B / res[False] * 1e3:10.0f}", flush=True)
