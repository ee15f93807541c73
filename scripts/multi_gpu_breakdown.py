#!/usr/bin/env python
"""Per-kernel time of multi-GPU bench steps on rank 0 (torch.profiler / CUPTI: every kernel of the process,
NCCL's included), launched under torchrun like bench.py.  ncu cannot follow a multi-rank job; this is the
launch list of one N-GPU step that profiles/r02_multigpu_step_*.md records.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
          scripts/multi_gpu_breakdown.py [--members 256] [--precision fp32] [--steps 10]"""
import argparse
import collections
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ertdiff_b200 as eb  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=256)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--shard", default="auto", choices=["auto", "yes", "no"])
a = ap.parse_args()
SHARD = {"auto": None, "yes": True, "no": False}[a.shard]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
spec = bench.Spec("breakdown", a.members, a.precision, T=a.T)
sd, cond = bench.synthetic_inputs(spec)
model = eb.ConditionalDiffusionModel(bench.P, spec.hidden)
model.load_state_dict(sd)
model.to(dev).eval()
sched = [t.to(dev) for t in eb.get_diffusion_schedule(a.T)]
cond_b = cond.to(dev).expand(a.members, bench.C, spec.L)
total = a.members * world
peer_x = peer_s = None
if not os.environ.get("ERTDIFF_BENCH_NCCL"):
    peer_x = eb.parallel.PeerAllGather(a.members * bench.P * 4, dev)
    peer_s = eb.parallel.PeerAllGather(-(-bench.P // world) * (5 + len(bench.PERCENTILES)) * 8, dev)


def step(i):
    x = eb.run_chain(model, cond_b, a.T, *sched, dev, seed=1234, offset=4 * i, member_offset=rank * a.members,
                     precision=a.precision, check_status=False)
    x = eb.parallel.gather_members(x, total, peer=peer_x)
    return eb.parallel.ensemble_statistics_distributed(x, bench.PERCENTILES, bench.KDE_GRID, shard=SHARD, peer=peer_s)


for i in range(3):
    step(i)
dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(a.steps):
        step(i)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split("(")[0].replace("void ", "")[:88]
            agg[name][0] += 1
            agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print(f"# rank 0 of {world} GPUs, {a.members} members/GPU ({total} gathered), {a.precision}, T={a.T}: device time of "
          f"{a.steps} steps by kernel (torch.profiler); {tot / a.steps:.1f} us of kernels per step")
    print("| kernel | launches/step | avg us | us/step | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n / a.steps:.1f} | {t / n:.1f} | {t / a.steps:.1f} | {100 * t / tot:.1f}% |")
dist.barrier()
dist.destroy_process_group()
