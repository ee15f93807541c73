#!/usr/bin/env python
"""Copy the round-2 measurement logs brought back in gpurun_out/ into tracked summaries under profiles/
(gpurun_out/ is scratch).  Re-run after every GPU pass; files whose source log is absent are left untouched."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, out = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def emit(dst, title, sources, fence=True):
    parts = []
    for src, caption in sources:
        p = os.path.join(go, src)
        if not os.path.isfile(p):
            continue
        txt = open(p).read().rstrip()
        parts.append((f"## {caption}\n\n" if caption else "") + (f"```\n{txt}\n```" if fence else txt))
    if parts:
        open(os.path.join(out, dst), "w").write(f"# {title}\n\n" + "\n\n".join(parts) + "\n")
        print("wrote", dst)


emit("r02_parity_measured.md",
     "r02 measured parity errors, CUDA path vs the reference-made goldens (scripts/measure_parity.py, one B200)",
     [("parity.log", "per case: max |got - want|, the golden's scale, and the smallest atol that makes the element-wise "
                     "`|d| <= atol + rtol*|want|` hold at rtol = 1e-5 / 1e-4 / 1e-3 / 1e-2 (tests/ use ~10x these)")])
emit("r02_chain_fp32_variants.md",
     "r02 fp32 chain kernel: members per CTA (mpb) x hidden units per thread (upt), full kernel and arithmetic-free floor "
     "(scripts/chain_fp32_variants.py; us per denoiser step = kernel duration / T, T = 1000, best of 4, L2 flushed)",
     [("variants_h128.log", "hidden_dim 128, small ensembles"), ("variants_h128_b.log", "hidden_dim 128, larger ensembles"),
      ("variants_h256.log", "hidden_dim 256, L = 9386")])
emit("r02_chain_bf16_sweep.md", "r02 tensor-core chain (k_chain_umma), scripts/chain_sweep.py, T = 1000",
     [("sweep_bf16.log", None)])
emit("r02_chain_split_precision.md",
     "r02 the three chain kernels side by side (scripts/chain_sweep.py; kernel duration by CUDA events): fp32 CUDA-core, "
     "bf16 tensor-core, and the split-precision tensor-core build (precision=\"bf16x3\": bf16 hi + residual operands, "
     "three accumulating products per projection); accuracy of the three against the reference goldens is in "
     "r02_parity_measured.md",
     [("chain_sweep_h.log", "T = 200 (first block) and T = 1000 (last four lines)")])
emit("r02_chain_umma_rng_variants.md",
     "r02 tensor-core chain, noise-warp variants that were measured and NOT kept (scripts/chain_sweep.py, T = 1000; separately "
     "built libraries).  `a` = the shipped kernel (four noise warps, one item at a time: Philox rounds, then Box-Muller).  "
     "`main` = the noise warps skewed by one item, the next item's Philox rounds written between the current item's "
     "Box-Muller pairs so that the FMA/ALU and XU pipes would overlap inside the warp: ptxas still emits the MUFU stream "
     "as one cluster, and the extra live registers cost more than the overlap gives (bf16, 18,944 members: 1.237 -> 1.356 us "
     "per step); the code was reverted.  `b` = eight noise warps, two per scheduler (-DUC_RNG_WARPS_N=8; 544 threads, 96 "
     "registers): no gain for bf16, slower for bf16x3 (64 bytes of spills).  `c` = both.",
     [("rng_variants.log", None)])
emit("r02_kde_coarse_to_fine.md",
     "r02 the coarse-to-fine KDE scan against the full scan (ERTDIFF_KDE_COARSE=1 evaluates every grid point; the default "
     "largest coarse stride is 32): scripts/stats_bench.py --only kde, then the fused statistics call on real chain output "
     "(scripts/summary_window_bench.py --chain; whole call by CUDA events, per kernel by torch.profiler).  scripts/gpu_r2_j.sh",
     [("kde_coarse_ab.log", None)])
emit("r02_summary_window.md",
     "r02 the fused statistics call (ertdiff_ensemble_summary) on a column window of the fields of a real T = 1000 bf16 chain "
     "(scripts/summary_window_bench.py --chain; whole call by CUDA events, per kernel by torch.profiler; the side-stream "
     "kernels -- k_colstats_smallq<.,0>, the percentile kernels -- overlap the KDE kernels)",
     [("summary_window_chain.log", "real chain output (all 29 parameters spread over +-2000: every column covers most of the common grid)"),
      ("summary_window_small.log", "one-launch KDE (k_kde_small) against the staged kernels below 1024 members: ERTDIFF_KDE_SMALL_MAX sweep")])
emit("r02_multigpu_peer_vs_nccl.md",
     "r02 the default bench workload (256 members per GPU) with the two per-step collectives through the library's NVLink "
     "peer-memory kernel (k_peer_all_gather) and through NCCL (ERTDIFF_BENCH_NCCL=1); scripts/gpu_r2_peer.sh",
     [(f"bench_peer_n{n}.json", f"{n} GPUs, peer kernel") for n in (2, 4, 8)] + [(f"bench_nccl_n{n}.json", f"{n} GPUs, NCCL") for n in (2, 4, 8)])
emit("r02_stats_bench.md",
     "r02 statistics kernels on the reference's map-shaped workload and on the chain's own output (scripts/stats_bench.py; "
     "CUDA-event time of the whole library call, best of 3, L2 flushed; calls of a few microseconds are dominated by "
     "launch overhead -- the ncu durations in r02_ncu_kernels.json are the kernel-only figures)",
     [("stats_bench.log", None)])
for n in (2, 4, 8):
    emit(f"r02_multigpu_step_n{n}.md",
         f"r02 one bench step on {n} GPUs: per-kernel device time on rank 0 (scripts/multi_gpu_breakdown.py, torch.profiler / CUPTI; "
         "NCCL kernels included -- their duration contains the wait for the slowest rank)",
         [(f"breakdown_fp32_256_n{n}.md", None), (f"breakdown_bf16_1024_n{n}.md", None)], fence=False)
    emit(f"r02_multigpu_check_n{n}.md", f"r02 multi-GPU parity on {n} GPUs (scripts/multi_gpu_check.py): gathered fields and "
         "column-sharded statistics against a single-GPU run of the whole ensemble", [(f"multi_gpu_check_{n}.log", None)])
