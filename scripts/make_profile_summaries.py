#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/:
   profiles/<round>_launches_<cfg>.md   per-kernel shares of a bench step (launch list)
   profiles/<round>_ncu_kernels.json    key metrics of each `--set full` capture (bench.py reads `traffic` here)
   profiles/<round>_ncu_<kernel>.md     metrics + stall reasons + hottest SASS of each capture
usage: make_profile_summaries.py r01"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")

for cfg, title in (("fp32_b256", "bench.py default: 256 members, T=1000, fp32, shared condition"),
                   ("bf16_b8192", "bench.py --precision bf16 --members 8192"),
                   ("bf16_b18944", "bench.py --precision bf16 --members 18944"),
                   ("stats_maps1024", "scripts/stats_bench.py --maps 1024 --fields '' (maps (1024, 4693, 14) float64)")):
    src = os.path.join(go, f"launches_{cfg}.csv")
    if os.path.isfile(src):
        md = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"), src,
                             f"{rnd} launch list, {title} (`ncu --metrics gpu__time_duration.sum --clock-control none`; "
                             "cold-cache serialised launches: compare SHARES, not absolutes)"],
                            capture_output=True, text=True).stdout
        open(os.path.join(out, f"{rnd}_launches_{cfg}.md"), "w").write(md)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        # per-pipe instruction rates: XU = MUFU + conversions, the pipe the KDE scan and the chain's noise warps live on
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
jpath = os.path.join(out, f"{rnd}_ncu_kernels.json")
summary = json.load(open(jpath)) if os.path.isfile(jpath) else {}      # captures arrive one per gpurun call
for name, what in (("chain_fp32", "k_chain (two hidden units per thread), 256 members, T=1000 (chain_sweep.py)"),
                   ("moments", "k_moments<double>, maps (1024, 4693, 14) float64 (stats_bench.py)"),
                   ("percentiles", "k_percentiles_select<double> (exact radix selection), maps (1024, 4693, 14) float64 (stats_bench.py)"),
                   ("percentiles_n50", "k_percentiles_warp<double> (one warp per column), maps (50, 4693, 14) float64: the reference's own shape (stats_bench.py)"),
                   ("kde_scan", "k_kde_scan32<double>, maps (1024, 4693, 14) float64 (stats_bench.py)"),
                   ("kde_select", "k_kde_select64<double>, maps (1024, 4693, 14) float64 (stats_bench.py)"),
                   ("posterior_update", "k_posterior_update, 2^26 elements (stats_bench.py)"),
                   ("sort_runs", "k_sort_runs<float>, fields (151552, 29) float32 (stats_bench.py)"),
                   ("select_runs", "k_select_runs<float>, fields (151552, 29) float32 (stats_bench.py)"),
                   ("chain_umma", "k_chain_umma, 18,944 members, T=200 (chain_sweep.py)"),
                   ("chain_umma_split", "k_chain_umma<SPLIT> (precision bf16x3: bf16 hi + residual operands), 18,944 members, T=200 (chain_sweep.py)"),
                   ("chain_umma2", "k_chain_umma, two CTAs per SM build, 37,888 members, T=200 (chain_sweep.py)"),
                   ("encoder_umma", "k_encoder_umma, 1024 conditions of 14x4693 (encoder_bench.py)")):
    rep = os.path.join(go, f"prof_{name}.ncu-rep")
    if not os.path.isfile(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    m = {"what": what, "kernel": d["Kernel Name"][1]}
    for k in KEYS:
        if k in d:
            u, v = d[k]
            v = float(v.replace(",", ""))
            if k.startswith("dram__bytes"):
                v, u = v * SCALE.get(u, 1), "byte"
            if k.startswith("gpu__time"):
                v, u = v * SCALE.get(u, 1), "us"
            m[k] = {"value": v, "unit": u}
    m["traffic_bytes_per_launch"] = m["dram__bytes_read.sum"]["value"] + m["dram__bytes_write.sum"]["value"]
    summary[name] = m
    srccsv = os.path.join(go, f"{name}_source.csv")
    open(srccsv, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                                           capture_output=True, text=True).stdout)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_summary.py"), srccsv, "25"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(out, f"{rnd}_ncu_{name}.md"), "w") as f:
        f.write(f"# {rnd} `ncu --set full --clock-control none --import-source on`: {what}\n\n")
        for k, v in m.items():
            if isinstance(v, dict):
                f.write(f"- `{k}` = {v['value']:.6g} {v['unit']}\n")
        f.write(f"- DRAM traffic per launch = {m['traffic_bytes_per_launch']:.6g} bytes\n\n```\n{txt}```\n")
json.dump(summary, open(jpath, "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
