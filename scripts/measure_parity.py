#!/usr/bin/env python
"""Measured parity errors of the CUDA path against the reference-made goldens (tests/golden), printed case by
case: the tolerances in tests/ are set to ~10x these numbers.  Development aid (GPU box):
    python scripts/measure_parity.py > gpurun_out/parity.log"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ertdiff_b200 as eb  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda", 0)
P, C, L = 29, 14, 4693


def report(name, got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    d = np.abs(got - want)
    scale = np.abs(want).max()
    # smallest atol for which rtol = 1e-5 / 1e-4 / 1e-3 holds element-wise
    need = {r: float(np.maximum(d - r * np.abs(want), 0).max()) for r in (1e-5, 1e-4, 1e-3, 1e-2)}
    print(f"{name:44s} max|d| {d.max():.3e}  |want|max {scale:.3e}  d/scale {d.max() / scale:.3e}  "
          f"atol needed at rtol 1e-5/1e-4/1e-3/1e-2: {need[1e-5]:.2e} {need[1e-4]:.2e} {need[1e-3]:.2e} {need[1e-2]:.2e}",
          flush=True)


def model_from(npz, prefix="", hidden=128):
    sd = {k[len(prefix):]: torch.from_numpy(npz[k].copy()) for k in npz.files if k.startswith(prefix) and "." in k[len(prefix):]}
    m = eb.ConditionalDiffusionModel(P, hidden)
    m.load_state_dict(sd)
    return m.to(dev).eval()


m = model_from(np.load(os.path.join(G, "model_seed0.npz")))
c1 = np.load(os.path.join(G, "chain_cfg1.npz"))
cond1 = torch.from_numpy(c1["condition"]).to(dev)
noise = torch.from_numpy(c1["noise"]).to(dev)
b, a, ab = eb.get_diffusion_schedule(50)
for prec in ("fp32", "bf16", "bf16x3"):
    x, eps = eb.run_chain(m, cond1.expand(16, C, L), 50, b, a, ab, dev, noise=noise, return_eps=True, precision=prec)
    for t in (49, 25, 0):
        report(f"cfg1 {prec} eps t={t}", eps[t].cpu(), c1[f"eps_t{t}"])
    report(f"cfg1 {prec} x0 (T=50, B=16)", x.cpu(), c1["x0"])
f = np.load(os.path.join(G, "forward_cases.npz"))
for tag in ("L257", "L64", "L3", "L1", "L1000"):
    xx, tt, cc = (torch.from_numpy(f[f"{tag}_{k}"]).to(dev) for k in ("x", "t", "cond"))
    report(f"forward {tag} eps", m(xx, tt, cc).cpu(), f[tag + "_eps"])
    report(f"forward {tag} cond_emb", m.encode_condition(cc).cpu(), f[tag + "_cemb"])
    report(f"forward {tag} cond_emb bf16", m.encode_condition(cc, precision="bf16").cpu(), f[tag + "_cemb"])
g = np.load(os.path.join(G, "chain_long.npz"))
for name, (B, T, ns) in {"T1000_B4": (4, 1000, None), "T500_B3_steps120": (3, 500, 120)}.items():
    torch.manual_seed(2)
    nz = torch.randn(T if ns is None else ns, B, P)
    bb = eb.get_diffusion_schedule(T)
    for prec in ("fp32", "bf16", "bf16x3"):
        x = eb.sample_model(m, cond1.expand(B, C, L), T, *bb, P, dev, num_steps=ns, noise=nz.to(dev), precision=prec)
        report(f"long {name} {prec}", x.cpu(), g[name])
g2 = np.load(os.path.join(G, "chain_cfg2.npz"))
bb = eb.get_diffusion_schedule(1000)
for name, B in (("cfg2_B256_T1000", 256), ("B64_T1000", 64)):
    torch.manual_seed(2)
    nz = torch.randn(1000, B, P)
    assert np.array_equal(nz[:2].numpy(), g2[name + "_noise_head"])
    for prec in ("fp32", "bf16", "bf16x3"):
        for upt in ("1", "2") if prec == "fp32" else ("",):
            if upt:
                os.environ["ERTDIFF_CHAIN_UPT"] = upt
            x = eb.sample_model(m, cond1.expand(B, C, L), 1000, *bb, P, dev, noise=nz.to(dev), precision=prec)
            report(f"{name} {prec} upt={upt}", x.cpu(), g2[name])
            d = np.abs(x.cpu().numpy().astype(np.float64) - g2[name]).max(axis=1)
            s = np.abs(g2[name]).max(axis=1)
            print(f"    per-member max|d|/max|x|: median {np.median(d / s):.3e}  p99 {np.percentile(d / s, 99):.3e}  max {(d / s).max():.3e}")
    os.environ.pop("ERTDIFF_CHAIN_UPT", None)
# bf16 vs fp32 on the same device RNG stream, 1024 members, T = 1000; random-init and damped weights
for label, damp in (("random-init", 1.0), ("damped (mlp.2 x 0.25)", 0.25)):
    mm = model_from(np.load(os.path.join(G, "model_seed0.npz")))
    if damp != 1.0:
        sd = mm.state_dict()
        sd["mlp.2.weight"] = sd["mlp.2.weight"] * damp
        sd["mlp.2.bias"] = sd["mlp.2.bias"] * damp
        mm.load_state_dict(sd)
    xs = {p_: eb.run_chain(mm, cond1.expand(1024, C, L), 1000, *bb, dev, seed=5, offset=0, precision=p_) for p_ in ("fp32", "bf16", "bf16x3")}
    for p_ in ("bf16", "bf16x3"):
        report(f"{p_} vs fp32, 1024 x T1000, {label}", xs[p_].cpu(), xs["fp32"].cpu())
        d = (xs[p_] - xs["fp32"]).abs().max(dim=1).values / xs["fp32"].abs().max(dim=1).values
        print(f"    per-member max|d|/max|x|: median {d.median().item():.3e}  p99 {d.quantile(0.99).item():.3e}  max {d.max().item():.3e}")
# hidden_dim = 256, L = 9386
h = np.load(os.path.join(G, "model_h256_case.npz"))
m256 = model_from(h, "sd.", 256)
gh = np.load(os.path.join(G, "chain_h256.npz"))
gg = torch.Generator().manual_seed(int(gh["cond_seed"]))
cond = torch.rand(2, C, 2 * L, generator=gg)
nz = torch.randn(200, 8, P, generator=gg)
assert np.array_equal(nz[:2].numpy(), gh["noise_head"])
bb = eb.get_diffusion_schedule(200)
for prec in ("fp32", "bf16"):
    for upt in ("1", "2") if prec == "fp32" else ("",):
        if upt:
            os.environ["ERTDIFF_CHAIN_UPT"] = upt
        x = eb.sample_ensemble(m256, cond.to(dev), 200, *bb, P, dev, n_realizations=4, noise=nz.to(dev), precision=prec)
        report(f"H256 L9386 chain {prec} upt={upt}", x.reshape(8, P).cpu(), gh["x0"])
os.environ.pop("ERTDIFF_CHAIN_UPT", None)
print("umma status", m.umma_status(), m256.umma_status())
