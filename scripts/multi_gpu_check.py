#!/usr/bin/env python
"""Multi-GPU parity check, launched under torchrun (one process per GPU, NCCL):
every rank runs its contiguous slice of the ensemble, ONE all-gather returns the fields in member
order; the gathered fields must be bit-identical to a single-GPU run of the whole ensemble (device
RNG streams are keyed by the global member index), for both chain kernels and with an uneven split.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
          --master-port 29511 scripts/multi_gpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
P, H, C, L, T = 29, 128, 14, 4693, 200
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(P, H)
torch.manual_seed(1)
cond1 = torch.rand(1, C, L)
model.to(dev).eval()
sched = eb.get_diffusion_schedule(T)
peer_x = eb.parallel.PeerAllGather(600 * P * 4, dev)
peer_s = eb.parallel.PeerAllGather(32 * 16 * 8, dev)
ok = True
for prec in ("fp32", "bf16"):
    for B in (512, 515):                      # even and uneven splits
        a, b = eb.parallel.member_slice(B, rank, world)
        x_loc = eb.run_chain(model, cond1.to(dev).expand(b - a, C, L), T, *sched, dev, seed=99, offset=0,
                             member_offset=a, precision=prec)
        x_all = eb.parallel.gather_members(x_loc, B)
        if B % world == 0:                      # the library's own NVLink all-gather must deliver the same bytes as NCCL
            x_peer = peer_x.all_gather(x_loc)
            assert torch.equal(x_peer, x_all), "peer all-gather differs from NCCL"
        x_one = eb.run_chain(model, cond1.to(dev).expand(B, C, L), T, *sched, dev, seed=99, offset=0,
                             precision=prec)
        same = torch.equal(x_all, x_one)
        qs = (2.5, 50.0, 97.5)
        st_one = eb.ensemble_statistics(x_one, percentiles=(), n_grid=512)
        pct_one = eb.ensemble_percentile(x_one, list(qs))                        # list q: float64 index arithmetic
        st_sh = eb.parallel.sharded_statistics(x_all, qs, 512)                   # columns split over the ranks
        st_pp = eb.parallel.sharded_statistics(x_all, qs, 512, peer=peer_s)      # same, results through the peer kernel
        assert torch.equal(st_pp["packed"], st_sh["packed"]), "peer-gathered statistics differ"
        bad = [k for k in ("mean", "std", "var", "mode", "mode_index")
               if not torch.equal(st_sh[k].to(st_one[k].dtype), st_one[k])]
        if not torch.equal(st_sh["pct"], pct_one.double()):
            bad.append("pct")
        same_stats = not bad
        if bad:
            print(f"rank {rank} {prec} B={B}: sharded statistics differ in {bad}", flush=True)
        flag = torch.tensor([int(same and same_stats)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{prec} B={B} world={world}: gathered == single-GPU: {bool(flag.item())}", flush=True)
        ok = ok and bool(flag.item())
ok = ok and peer_x.status() == 0 and peer_s.status() == 0
dist.barrier()
peer_x.close(); peer_s.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
