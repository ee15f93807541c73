"""Generate ``tests/golden/*`` by running the REFERENCE ITSELF (unmodified, AST-loaded from
``/root/reference``) in the build container.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden          # from the repo root; needs /root/reference

The reference ships no golden vectors; these files are what pins both the
oracle restatement (``oracle/denoiser_oracle.py``) and the CUDA path on machines where the
reference is not present.  Seeds follow SURVEY.md §8(d): weights ``manual_seed(0)``,
condition ``manual_seed(1)``, noise ``manual_seed(2)``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
P, H, C, L = 29, 128, 14, 4693


def ref_model(ref, hidden=H, seed=0):
    torch.manual_seed(seed)
    m = ref.ConditionalDiffusionModel(P, hidden)
    m.eval()
    return m


def run_chain(model, cond, T, noise, num_steps=None, temperature=1.0):
    ref = load_reference(noise=noise)
    betas, alphas, alpha_bar = ref.get_diffusion_schedule(T)
    x = ref.sample_model(model, cond, T, betas, alphas, alpha_bar, P, "cpu",
                         num_steps=num_steps, temperature=temperature)
    n_expected = T if num_steps is None else num_steps
    assert ref.torch_proxy.draws == n_expected, (ref.torch_proxy.draws, n_expected)
    return x, (betas, alphas, alpha_bar)


def eps_trace(ref, model, cond, T, noise, at):
    """Replay the chain step by step with the reference's model and record pred_noise."""
    import math
    betas, alphas, alpha_bar = ref.get_diffusion_schedule(T)
    x = noise[0].clone()
    draw = 1
    out = {}
    with torch.no_grad():
        for t_ in reversed(range(T)):
            tt = torch.full((cond.size(0),), t_, dtype=torch.long)
            eps = model(x, tt, cond)
            if t_ in at:
                out[t_] = (x.clone(), eps.clone())
            coef = (1 - alphas[t_]) / (math.sqrt(1 - alpha_bar[t_]) + 1e-8)
            x = (1.0 / math.sqrt(alphas[t_])) * (x - coef * eps)
            if t_ > 0:
                x = x + math.sqrt(betas[t_]) * 1.0 * noise[draw]
                draw += 1
    return out, x


def make_misfit(ref):
    # ---- per-member misfit (ECD.py:764-785, 927-930): the reference's own WSSE_metric and sklearn's
    # mean_squared_error, inside the reference's inline loops as written there
    import contextlib
    import io
    from sklearn.metrics import mean_squared_error
    rng = np.random.default_rng(9)
    Lm, Cm = 211, 14
    observed = rng.normal(3.0, 1.0, size=(Lm, Cm))
    sim_data = observed[None] + rng.normal(scale=0.4, size=(5, Lm, Cm))
    misfit = {"sim_data": sim_data, "observed": observed}
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        sd_, ob_ = sim_data.astype(dt), observed.astype(dt)
        with contextlib.redirect_stdout(io.StringIO()):           # WSSE_metric prints every value
            wsse = np.array([[ref.WSSE_metric(0.1, 0.01, sd_[i][:, es], ob_[:, es])[0] for es in range(Cm)]
                             for i in range(sd_.shape[0])])
        mse = np.array([mean_squared_error(ob_.flatten(), sd_[i].flatten()) for i in range(sd_.shape[0])])
        misfit.update({f"wsse_{tag}": wsse, f"wsse_total_{tag}": wsse.sum(axis=1), f"mse_{tag}": mse})
    np.savez(os.path.join(OUT, "misfit.npz"), **misfit)


def load_parameter_limits():
    """``ParameterLimits().plims`` (Generate_ERT_utils.py:8-59), class loaded by AST like the rest."""
    import ast
    from oracle.reference_loader import REFERENCE_ROOT
    path = os.path.join(REFERENCE_ROOT, "Generate_ERT_utils.py")
    tree = ast.parse(open(path).read(), filename=path)
    picked = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ParameterLimits"]
    ns = {"np": np}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), ns)
    return ns["ParameterLimits"]().plims


def make_round2(ref):
    """Round-2 fixtures (each in its own file; the round-1 files stay byte-identical in git):
    the BASELINE config-2 chain at full size, a 64-member T=1000 chain (the bf16 path's reference anchor),
    a hidden_dim=256 / L=9386 chain (config 5's reference-expressible widening) and the f1 epilogue
    (the reference's own inverse_transform + check_param_bounds around sklearn's MinMaxScaler)."""
    import contextlib
    import io
    import re
    import time
    from sklearn.preprocessing import MinMaxScaler
    model = ref_model(ref)
    torch.manual_seed(1)
    cond1 = torch.rand(1, C, L)
    # ---- config 2: 256 members, T = 1000, one shared condition; noise regenerated from its seed ----------
    out = {}
    for name, B in (("cfg2_B256_T1000", 256), ("B64_T1000", 64)):
        torch.manual_seed(2)
        nz = torch.randn(1000, B, P)
        t0 = time.time()
        x, _ = run_chain(model, cond1.expand(B, C, L), 1000, nz)
        print(name, "reference chain", round(time.time() - t0, 1), "s", flush=True)
        out[name] = x.numpy()
        out[name + "_noise_head"] = nz[:2].numpy()
    np.savez(os.path.join(OUT, "chain_cfg2.npz"), **out)
    # ---- hidden_dim = 256, L = 9386 ("2x grid"), 8 members, T = 200 ---------------------------------------
    m256 = ref_model(ref, hidden=256, seed=5)
    prev = np.load(os.path.join(OUT, "model_h256_case.npz"))
    assert all(np.array_equal(prev["sd." + k], v.detach().numpy()) for k, v in m256.state_dict().items())
    g = torch.Generator().manual_seed(11)
    cond = torch.rand(2, C, 2 * L, generator=g)                  # two distinct conditions, 4 realisations each
    nz = torch.randn(200, 8, P, generator=g)
    refn = load_reference(noise=nz)
    b, a, ab = refn.get_diffusion_schedule(200)
    with torch.no_grad():
        x = refn.sample_model(m256, cond.repeat(4, 1, 1), 200, b, a, ab, P, "cpu")
    np.savez(os.path.join(OUT, "chain_h256.npz"), x0=x.numpy(), cond_seed=np.int64(11), noise_head=nz[:2].numpy(),
             cond_head=cond[:, :2, :8].numpy())
    # ---- f1: logits -> sigmoid -> MinMaxScaler.inverse_transform -> check_param_bounds (ECD.py:402-406) ----
    plims = load_parameter_limits()
    rng = np.random.default_rng(12)
    width = plims[:, 1] - plims[:, 0]
    sim_param = plims[:, 0] - 0.03 * width + rng.uniform(size=(500, P)) * 1.06 * width    # training range a little wider
    scaler = MinMaxScaler(feature_range=(0.0, 1.0)).fit(sim_param)
    u = torch.from_numpy(rng.normal(scale=2.5, size=(400, P)).astype(np.float32))
    gen = ref.inverse_transform(u, 0.0, 1.0)                     # torch branch, float32
    gen_np = scaler.inverse_transform(gen.cpu().numpy())
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kept = ref.check_param_bounds(gen_np, plims)
    first_bad = -np.ones(u.shape[0], dtype=np.int32)
    for m in re.finditer(r"Sample (\d+) Parameter (\d+):", buf.getvalue()):
        first_bad[int(m.group(1))] = int(m.group(2))
    valid = first_bad < 0
    assert kept is not None and np.array_equal(kept, gen_np[valid]) and 0 < valid.sum() < valid.size
    # the float64 form (numpy branch of inverse_transform), bounds only
    gen64 = scaler.inverse_transform(ref.inverse_transform(u.numpy().astype(np.float64), 0.0, 1.0))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kept64 = ref.check_param_bounds(gen64, plims)
    first_bad64 = -np.ones(u.shape[0], dtype=np.int32)
    for m in re.finditer(r"Sample (\d+) Parameter (\d+):", buf.getvalue()):
        first_bad64[int(m.group(1))] = int(m.group(2))
    np.savez(os.path.join(OUT, "transforms_f1.npz"), u=u.numpy(), sigmoid=gen.numpy(), phys=gen_np, valid=valid,
             first_bad=first_bad, scaler_min=scaler.min_, scaler_scale=scaler.scale_, limits=plims,
             phys64=gen64, first_bad64=first_bad64)
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in ("chain_cfg2.npz", "chain_h256.npz", "transforms_f1.npz")})


def main():
    if sys.argv[1:] == ["misfit"]:          # only this fixture (the others stay byte-identical in git)
        return make_misfit(load_reference())
    if sys.argv[1:] == ["round2"]:
        return make_round2(load_reference())
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    model = ref_model(ref)
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    np.savez(os.path.join(OUT, "model_seed0.npz"), **sd)

    # ---- config 1: 16 members, T=50, one shared condition ---------------------------------
    torch.manual_seed(1)
    cond1 = torch.rand(1, C, L)
    torch.manual_seed(2)
    noise = torch.randn(50, 16, P)
    cond = cond1.expand(16, C, L)
    x0, (betas, alphas, alpha_bar) = run_chain(model, cond, 50, noise)
    tr, x0_b = eps_trace(ref, model, cond, 50, noise, at=(49, 25, 0))
    assert torch.equal(x0, x0_b)
    x_tr, x0_trunc = None, None
    x0_trunc, _ = run_chain(model, cond, 50, noise[:20], num_steps=20, temperature=0.7)
    np.savez(os.path.join(OUT, "chain_cfg1.npz"),
             condition=cond1.numpy(), noise=noise.numpy(), x0=x0.numpy(),
             betas=betas.numpy(), alphas=alphas.numpy(), alpha_bar=alpha_bar.numpy(),
             x_t49=tr[49][0].numpy(), eps_t49=tr[49][1].numpy(),
             x_t25=tr[25][0].numpy(), eps_t25=tr[25][1].numpy(),
             x_t0=tr[0][0].numpy(), eps_t0=tr[0][1].numpy(),
             x0_steps20_temp07=x0_trunc.numpy())

    # ---- long chains: noise regenerated from the seed (torch CPU randn is deterministic) ---
    long = {}
    for name, (B, T, ns, temp) in {"T1000_B4": (4, 1000, None, 1.0),
                                   "T500_B3_steps120": (3, 500, 120, 1.0)}.items():
        torch.manual_seed(2)
        nz = torch.randn(T if ns is None else ns, B, P)
        x, _ = run_chain(model, cond1.expand(B, C, L), T, nz, num_steps=ns, temperature=temp)
        long[name] = x.numpy()
        long[name + "_noise_head"] = nz[:2].numpy()      # guards the regeneration
    np.savez(os.path.join(OUT, "chain_long.npz"), **long)

    # ---- standalone forward: per-row t, distinct conditions, ragged lengths ---------------
    fw = {}
    for tag, (B, Lx) in {"L4693": (3, 4693), "L257": (5, 257), "L64": (4, 64), "L3": (2, 3),
                         "L1": (2, 1), "L1000": (2, 1000)}.items():
        g = torch.Generator().manual_seed(100 + Lx)
        x = torch.randn(B, P, generator=g)
        t = torch.randint(0, 1000, (B,), generator=g)
        c = torch.rand(B, C, Lx, generator=g)
        with torch.no_grad():
            e = model(x, t, c)
            ce = model.condition_encoder(c)
        fw[f"{tag}_x"], fw[f"{tag}_t"], fw[f"{tag}_eps"] = x.numpy(), t.numpy(), e.numpy()
        fw[f"{tag}_cemb"] = ce.numpy()
        if Lx != 4693:
            fw[f"{tag}_cond"] = c.numpy()
        else:
            fw[f"{tag}_cond_seed"] = np.int64(100 + Lx)
    np.savez(os.path.join(OUT, "forward_cases.npz"), **fw)

    # ---- hidden_dim=256 model (BASELINE config 5's reference-expressible widening) --------
    m256 = ref_model(ref, hidden=256, seed=5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, P, generator=g)
    t = torch.randint(0, 1000, (6,), generator=g)
    c = torch.rand(6, C, 530, generator=g)
    with torch.no_grad():
        e = m256(x, t, c)
    d = {"sd." + k: v.detach().numpy() for k, v in m256.state_dict().items()}
    np.savez(os.path.join(OUT, "model_h256_case.npz"), x=x.numpy(), t=t.numpy(),
             cond=c.numpy(), eps=e.numpy(), **d)

    # ---- timestep embedding / schedule known answers -------------------------------------
    tt = torch.tensor([0, 1, 2, 17, 499, 999], dtype=torch.long)
    emb = ref.get_timestep_embedding(tt, 128)
    sch = {}
    for T in (50, 500, 1000):
        b, a, ab = ref.get_diffusion_schedule(T)
        sch[f"betas{T}"], sch[f"alphas{T}"], sch[f"alpha_bar{T}"] = b.numpy(), a.numpy(), ab.numpy()
    np.savez(os.path.join(OUT, "embedding_schedule.npz"), t=tt.numpy(), emb=emb.numpy(), **sch)

    # ---- statistics: numpy / scipy called exactly as the reference calls them ------------
    from scipy import stats as sstats
    rng = np.random.default_rng(3)
    sim = rng.lognormal(size=(50, 12, 3))                     # (N, H, W) float64 maps
    grid = np.linspace(np.min(sim), np.max(sim), 5000)
    mode = np.zeros(sim.shape[1:])
    midx = np.zeros(sim.shape[1:], dtype=np.int64)
    for i in range(sim.shape[1]):
        for j in range(sim.shape[2]):
            kv = sstats.gaussian_kde(sim[:, i, j])(grid)
            midx[i, j] = np.argmax(kv)
            mode[i, j] = grid[midx[i, j]]
    np.savez(os.path.join(OUT, "stats_maps.npz"), sim=sim,
             mean=np.mean(sim, axis=0), std=np.std(sim, axis=0), var=np.var(sim, axis=0),
             p25=np.percentile(sim, 25, axis=0), p50=np.percentile(sim, 50, axis=0),
             p75=np.percentile(sim, 75, axis=0), mode=mode, mode_index=midx,
             ci95=np.percentile(sim.astype(np.float32), [2.5, 97.5], axis=0))
    # ---- UQ calibration (ECD.py:1089-1137, 1191-1214): the reference's four helper functions run
    # unmodified; its inline percentile loop is replayed around them with numpy, as written there
    import warnings
    warnings.simplefilter("ignore", DeprecationWarning)          # np.trapz
    rng = np.random.default_rng(8)
    truth = rng.normal(size=(8, P)).astype(np.float32)
    gen = (truth[None] * 0.9 + rng.normal(scale=0.8, size=(50, 8, P))).astype(np.float32)
    prob_array = np.linspace(0.01, 0.99, 30)

    def curve(dist, true):
        avg = np.zeros(len(prob_array))
        for prob in enumerate(prob_array):
            p_low = (1 - prob[1]) / 2
            p_upp = (1 + prob[1]) / 2
            low_bound = np.percentile(dist, p_low * 100, axis=0)
            upp_bound = np.percentile(dist, p_upp * 100, axis=0)
            indicator_matrix = ((low_bound < true) & (true <= upp_bound)).astype(int)
            avg[prob[0]] = np.mean(indicator_matrix)
        return avg

    def metrics(avg):
        a_p = ref.avg_prop_indicator_function(avg, prob_array)
        acc = ref.accuracy_score(a_p, prob_array)
        return acc, ref.preccision_score(acc, avg, prob_array, a_p), ref.goodness_score(a_p, avg, prob_array)

    avg = curve(gen, truth)
    acc, prec, good = metrics(avg)
    pavg = np.stack([curve(gen[:, :, j], truth[:, j]) for j in range(P)])
    pm = np.array([metrics(pavg[j]) for j in range(P)])
    np.savez(os.path.join(OUT, "uq_calibration.npz"), generated=gen, true=truth, prob_array=prob_array,
             avg_proportion=avg, accuracy=acc, precision=prec, goodness=good, param_avg_proportion=pavg,
             param_accuracy=pm[:, 0], param_precision=pm[:, 1], param_goodness=pm[:, 2])
    make_misfit(ref)
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print(sizes, sum(sizes.values()))


if __name__ == "__main__":
    main()
