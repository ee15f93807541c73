#!/usr/bin/env python
"""Benchmark of the ensemble posterior-sampling path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--members M]
                    [--T T] [--loop-mode persistent|graph] [--distinct-conditions]

One "step" = one pass of the hot path over one ensemble: condition encoder -> full DDPM reverse
chain of T steps for every member -> (N>1: all-gather) -> ensemble mean/std/var, percentiles
and KDE mode of the final fields.  Metric: posterior samples/sec (members completing the full
chain per second, whole job).  Default workload = BASELINE.json configs[1]: 256 members per
GPU, T = 1000, fp32, one synthetic condition of the reference's grid (14 x 4693) shared by all
members, random-init weights, device-side RNG.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU algorithm (the
oracle port, as written: condition encoder re-run every step) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, H, C, L = 29, 128, 14, 4693
FLOP_STEP_MEMBER = 14848            # 2*(29*128 + 128*29): x-block of mlp.0 + mlp.2 (SURVEY §8 d)
FLOP_ENCODER = 20751232 + 32768     # per distinct condition, once per chain
FLOP_TIME_ROW = 65536               # per step, shared by all members
FLOP_AS_WRITTEN = 20864384          # per member per step, reference-equivalent work
PERCENTILES = [2.5, 25.0, 50.0, 75.0, 97.5]
KDE_GRID = 5000
SHARD_STATS_ABOVE = 4096   # gathered members above which the statistics' columns are split over the ranks
                           # (below it the second all-gather and the packing cost more than they save)
MAX_STAT_MEMBERS = 25600   # one statistics call holds a column in shared memory (KDE: N*8 B <= 200 KB)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel):
    """DRAM bytes of one launch of `kernel` from the committed `ncu --set full` summary (or None)."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_kernels.json")
    try:
        return json.load(open(path))[kernel]["traffic_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def synthetic_inputs(members, distinct):
    """weights = the reference's default init under manual_seed(0) (the host
    mirror draws exactly what the reference's constructor draws), condition seed 1."""
    import ertdiff_b200 as eb
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in eb.ConditionalDiffusionModel(P, H).state_dict().items()}
    g = torch.Generator().manual_seed(1)
    n = members if distinct else 1
    cond = torch.rand(n, C, L, generator=g)
    return sd, cond


def cpu_reference_sample(sd, cond1, members, T, steps_sample, threads):
    """The reference algorithm as written (oracle port of ECD.py:102-119: the encoder runs
    every step) for `steps_sample` of the T steps; returns seconds."""
    from oracle import denoiser_oracle as do
    torch.set_num_threads(threads)
    betas, alphas, alpha_bar = do.diffusion_schedule(T)
    cond = cond1.expand(members, C, L) if cond1.size(0) == 1 else cond1
    noise = torch.randn(steps_sample, members, P, generator=torch.Generator().manual_seed(2))
    t0 = time.perf_counter()
    x = do.sample_chain(sd, cond, T, betas, alphas, alpha_bar, P, noise, num_steps=steps_sample)
    return time.perf_counter() - t0, x


def cpu_reference_stats(x):
    """The reference's statistics calls (ECD.py:747-762, 867-872) on the final fields; seconds."""
    from scipy import stats as sstats
    a = x.numpy()
    t0 = time.perf_counter()
    np.mean(a, axis=0); np.std(a, axis=0); np.var(a, axis=0)
    np.percentile(a, PERCENTILES, axis=0)
    grid = np.linspace(a.min(), a.max(), KDE_GRID)
    for j in range(a.shape[1]):
        np.argmax(sstats.gaussian_kde(a[:, j])(grid))
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    members, T = args.members, args.T
    sd, cond = synthetic_inputs(members, args.distinct_conditions)
    # each timed step = a bounded sample of the T-step chain; sized from a calibration run so that the
    # whole --steps K run stays within ~2.5 minutes of CPU time (at most --ref-sample-steps per sample)
    t_cal, _ = cpu_reference_sample(sd, cond, members, T, 2, threads)
    per_step_s = max(t_cal / 2, 1e-4)
    sample_steps = int(max(5, min(args.ref_sample_steps, 150.0 / (max(args.steps, 1) * per_step_s), T)))
    for _ in range(args.warmup):
        cpu_reference_sample(sd, cond, members, T, max(1, sample_steps // 8), threads)
    runs = [cpu_reference_sample(sd, cond, members, T, sample_steps, threads) for _ in range(args.steps)]
    secs = [r[0] for r in runs]
    stats_s = cpu_reference_stats(runs[-1][1])
    per_chain = float(np.mean(secs)) * (T / sample_steps) + stats_s   # chain cost is linear in steps
    value = members / per_chain
    line = {
        "impl": "reference", "metric": "posterior_samples_per_sec_full_chain", "value": value,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_chain * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, members),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{members} members x {sample_steps} of {T} steps of the as-written "
                                   f"chain (encoder re-run every step), scaled by {T}/{sample_steps}, + numpy/scipy statistics"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, members_per_gpu):
    return {"workload": f"BASELINE configs[1]: {members_per_gpu} members/GPU, T={args.T} full DDPM chain, "
                        f"{args.precision}, grid 14x{L}, {'distinct' if args.distinct_conditions else 'one shared'} "
                        "condition, + ensemble mean/std/var, 5 percentiles, KDE mode",
            "members_per_gpu": members_per_gpu, "T": args.T, "param_dim": P, "hidden_dim": H,
            "loop_mode": args.loop_mode, "rng": "device Philox4x32-10",
            "l2": "flushed between steps (256 MiB memset outside the per-step event pairs)",
            "parallelism": f"members sharded over {args.gpus} GPU(s), one all-gather of (B/G,29) f32; statistics on the "
                           f"gathered fields (columns split over the ranks above {SHARD_STATS_ABOVE} members)"}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import ertdiff_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    members, T = args.members, args.T                 # per GPU (weak scaling)
    total = members * world
    if total > MAX_STAT_MEMBERS:
        raise SystemExit(f"bench.py: {total} gathered members exceed the {MAX_STAT_MEMBERS} a single statistics "
                         "call supports (a column is held in shared memory); lower --members")
    sd, cond_host = synthetic_inputs(members, args.distinct_conditions)
    model = eb.ConditionalDiffusionModel(P, H)
    model.load_state_dict(sd)
    model.to(dev).eval()
    betas, alphas, alpha_bar = eb.get_diffusion_schedule(T)
    sched_dev = [t.to(dev) for t in (betas, alphas, alpha_bar)]
    cond_dev = cond_host.to(dev)
    cond_dev_b = cond_dev if args.distinct_conditions else cond_dev.expand(members, C, L)
    cond_pinned = cond_host.pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def stats(x):
        if world > 1 and total > SHARD_STATS_ABOVE:     # per-column statistics: columns split over the ranks
            return eb.parallel.sharded_statistics(x, PERCENTILES, KDE_GRID)
        out = eb.ensemble_moments(x)
        out["pct"] = eb.ensemble_percentile(x, PERCENTILES)
        out["mode"] = eb.ensemble_kde_mode(x, KDE_GRID)
        return out

    def step_device(i):
        """inputs resident in HBM"""
        x = eb.run_chain(model, cond_dev_b, T, *sched_dev, dev, seed=1234, offset=4 * i,
                         member_offset=rank * members, loop_mode=args.loop_mode, precision=args.precision)
        if world > 1:
            x = eb.parallel.gather_members(x, total)
        return x, stats(x)

    host_out = []

    def step_e2e(i):
        """public API with HOST buffers: H2D of the condition + schedule, D2H of fields + statistics"""
        c = cond_pinned.to(dev, non_blocking=True)
        c = c if args.distinct_conditions else c.expand(members, C, L)
        x = eb.sample_model(model, c, T, betas, alphas, alpha_bar, P, dev, seed=1234 + i,
                            loop_mode=args.loop_mode, precision=args.precision) if world == 1 else \
            eb.run_chain(model, c, T, betas, alphas, alpha_bar, dev, seed=1234, offset=4 * i,
                         member_offset=rank * members, loop_mode=args.loop_mode, precision=args.precision)
        if world > 1:
            x = eb.parallel.gather_members(x, total)
        st = stats(x)
        # D2H read of the step's results: fields + every statistic, straight into pinned host buffers
        # (no staging kernels), one stream synchronise at the end
        srcs = (x, st["mean"], st["std"], st["var"], st["pct"], st["mode"])
        if not host_out:
            host_out.extend(torch.empty(v.shape, dtype=v.dtype).pin_memory() for v in srcs)
        for dst, src in zip(host_out, srcs):
            dst.copy_(src, non_blocking=True)
        stream.synchronize()
        return host_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        wall0 = time.perf_counter()
        for i, (e0, e1) in enumerate(evs):
            flush.zero_()                       # L2 flush, outside the event pair
            e0.record(stream)
            fn(i)
            e1.record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    for i in range(args.warmup):
        step_device(i)
        step_e2e(i)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eb.launch_count(reset=True)
    model.profile_chain(True)
    ms_dev, _ = timed(step_device, args.steps)
    launches = eb.launch_count()
    chain_ms = []
    if args.loop_mode == "persistent":      # duration of the dominant kernel, CUDA events inside the library
        for i in range(args.steps):
            flush.zero_()
            step_device(i)
            chain_ms.append(model.last_chain_ms())
    model.profile_chain(False)
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # latency floor: the same persistent kernel structure cannot go below T dependent steps;
    # report an empty-dependency yardstick: ms per step of the chain kernel itself
    value = total * args.steps / (ms_dev * 1e-3)
    e2e_value = total * args.steps / (ms_e2e * 1e-3)
    n_cond = members if args.distinct_conditions else 1
    h2d = cond_host.numel() * 4          # the condition; the 3 x 4 KB schedule is content-cached on the device after step 1
    d2h = 4 * (total * P + 3 * P) + 8 * (len(PERCENTILES) * P + P)      # fields + moments fp32, percentiles + mode fp64

    if rank == 0:
        peaks = load_peaks()
        line = {
            "metric": "posterior_samples_per_sec_full_chain", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, members),
            "ms_per_denoiser_step": (float(np.mean(chain_ms)) / T) if chain_ms else ms_dev / args.steps / T,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "wall_ms_per_step": wall_e2e * 1e3 / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if chain_ms:
            k_ms = float(np.mean(chain_ms))
            flops = members * T * FLOP_STEP_MEMBER
            achieved = flops / (k_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            fp32_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12 if clocks else None
            fp32 = args.precision == "fp32"
            line["roofline"] = {
                "kernel": "k_chain (persistent reverse loop, fp32 FFMA2)" if fp32
                          else "k_chain_umma (persistent reverse loop, tcgen05 bf16, TMEM accumulators)",
                "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": ncu_traffic("chain_fp32" if fp32 else "chain_umma"),
                "peak_source": peaks["source"] + ", sustained bf16",
                "kernel_ms": k_ms, "flop_per_launch": flops,
                "share_of_step": k_ms / (ms_dev / args.steps),
                "note": ("structurally latency-bound: T dependent steps per member, 14,848 FLOP each, no HBM traffic "
                         "beyond the tables (device RNG); fp32 CUDA-core kernel, so also quoted against the fp32 FFMA peak"
                         if fp32 else
                         "T dependent steps per member; per step and 128-member tile the tensor pipe needs ~0.13 us, the "
                         "Philox/Box-Muller generator ~0.5 us of MUFU + integer-multiply pipe time (measured, "
                         "scripts/microbench/rng_bench.cu): the kernel is bound by issue slots / those pipes, not by the MMAs"),
                "fp32_ffma_peak_tflops": fp32_peak,
                "fp32_ffma_frac": (achieved / fp32_peak) if (fp32_peak and fp32) else None,
                "as_written_equivalent_tflops": members * T * FLOP_AS_WRITTEN / (k_ms * 1e-3) / 1e12,
                "traffic_note": "dram__bytes_read+write of one launch from the committed ncu capture "
                                "(profiles/r01_ncu_kernels.json; its workload is stated there), null if absent",
            }
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            t_cal, _ = cpu_reference_sample(sd, cond_host, members, T, 2, threads)            # warm-up + calibration
            sample_steps = int(max(5, min(args.ref_sample_steps, 12.0 / max(t_cal / 2, 1e-4), T)))   # ~10 s of CPU work
            secs, xs = cpu_reference_sample(sd, cond_host, members, T, sample_steps, threads)
            per_chain = secs * (T / sample_steps) + cpu_reference_stats(xs)
            line["cpu_baseline"] = {
                "value": members / per_chain, "unit": "samples/s", "cores": threads, "kind": "port",
                "sample": f"{members} members x {sample_steps} of {T} steps of the as-written chain "
                          f"(encoder re-run every step) in {secs:.2f} s, scaled by {T}/{sample_steps}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=256, help="ensemble members per GPU")
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--loop-mode", default="persistent", choices=["persistent", "graph", "stream"])
    ap.add_argument("--distinct-conditions", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"],
                    help="fp32 = CUDA-core FFMA chain (BASELINE config 2); bf16 = tcgen05 chain (configs 3/4)")
    ap.add_argument("--ref-sample-steps", type=int, default=250,
                    help="steps of the T-step chain the CPU arm actually runs per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                      # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
