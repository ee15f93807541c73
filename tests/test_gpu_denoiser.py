"""GPU parity of the denoiser, the posterior update and the reverse chain against the oracle
and the golden vectors generated from the reference.  All calls go through the C ABI.

Tolerances (stated per SURVEY.md §8 d), all ELEMENT-WISE ``|got - want| <= atol + rtol*|want|`` and set to
about ten times the error measured on a B200 (``scripts/measure_parity.py``, log in
``profiles/r02_parity_measured.md``): contractions run in fp32 but in a different summation order than ATen's
CPU kernels, so a single forward's predicted noise agrees to a few 1e-7; over a chain the difference is
amplified with the fields themselves (the reverse process scales x by 1/sqrt(alpha_t) every step, 157x over
T = 1000 -- a pure-PyTorch reorder of the same T = 500 chain already moves the reference's own output by
1.5e-5), hence one (rtol, atol) pair per chain length.  The posterior update, the step coefficients and
everything the loop modes / sharding share are compared bit for bit."""
import numpy as np
import pytest
import torch

import ertdiff_b200 as eb
from oracle import denoiser_oracle as do
from conftest import assert_close

pytestmark = pytest.mark.gpu
P, C, L = 29, 14, 4693
# measured on a B200 (worst case over the golden cases)           -> tolerance (about 10x)
EPS_RTOL, EPS_ATOL = 1e-5, 5e-7          # one forward: max|d| 1.8e-7 on values <= 0.58; 4.4e-8 beyond 1e-5*|want|
X50_RTOL, X50_ATOL = 1e-5, 5e-6          # fields after <= 200 steps: max|d| 9.5e-7 on |x| <= 12, none beyond 1e-5*|want|
X1000_RTOL, X1000_ATOL = 1e-5, 1.5e-3    # fields after 1000 steps (|x| up to 1.5e3): max|d| 8.5e-4, 1.4e-4 beyond 1e-5*|want|


def test_forward_golden_cases(gpu_model, golden, cuda_dev):
    f = golden("forward_cases.npz")
    for tag in ("L257", "L64", "L3", "L1", "L1000"):
        x, t, c = (torch.from_numpy(f[f"{tag}_{k}"]).to(cuda_dev) for k in ("x", "t", "cond"))
        eps = gpu_model(x, t, c).cpu().numpy()
        assert_close(eps, f[tag + "_eps"], EPS_RTOL, EPS_ATOL, tag + " eps")
        ce = gpu_model.encode_condition(c).cpu().numpy()
        assert_close(ce, f[tag + "_cemb"], EPS_RTOL, EPS_ATOL, tag + " cond_emb")
    # full-length grid, condition regenerated from its seed
    g = torch.Generator().manual_seed(int(f["L4693_cond_seed"]))
    x = torch.randn(3, P, generator=g)
    t = torch.randint(0, 1000, (3,), generator=g)
    c = torch.rand(3, C, L, generator=g)
    assert np.array_equal(x.numpy(), f["L4693_x"]) and np.array_equal(t.numpy(), f["L4693_t"])
    eps = gpu_model(x.to(cuda_dev), t.to(cuda_dev), c.to(cuda_dev)).cpu().numpy()
    assert_close(eps, f["L4693_eps"], EPS_RTOL, EPS_ATOL, "L4693 eps")


def test_forward_shared_condition_and_per_row_t(gpu_model, ref_state_dict, cuda_dev):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(37, P, generator=g)
    t = torch.randint(0, 1000, (37,), generator=g)
    c1 = torch.rand(1, C, 777, generator=g)
    ref = do.denoiser_forward(ref_state_dict, x, t, c1.expand(37, C, 777))
    got = gpu_model(x.to(cuda_dev), t.to(cuda_dev), c1.to(cuda_dev).expand(37, C, 777))
    assert_close(got.cpu(), ref, EPS_RTOL, EPS_ATOL, "shared condition, per-row t")
    got2 = gpu_model(x.to(cuda_dev), t.to(cuda_dev), c1.expand(37, C, 777).contiguous().to(cuda_dev))
    assert torch.equal(got, got2)           # shared vs materialised condition: same bits


def test_hidden256_model(golden, cuda_dev):
    g = golden("model_h256_case.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m = eb.ConditionalDiffusionModel(29, 256)
    m.load_state_dict(sd)
    m.to(cuda_dev)
    eps = m(torch.from_numpy(g["x"]).to(cuda_dev), torch.from_numpy(g["t"]).to(cuda_dev),
            torch.from_numpy(g["cond"]).to(cuda_dev))
    assert_close(eps.cpu(), g["eps"], EPS_RTOL, EPS_ATOL, "hidden 256 forward")
    back = m.state_dict()
    assert all(np.array_equal(back[k].cpu().numpy(), sd[k].numpy()) for k in sd)


@pytest.mark.parametrize("T,temp", [(50, 1.0), (500, 1.0), (1000, 0.7), (1000, 1.0)])
def test_step_coefficients_bit_exact(cuda_dev, T, temp):
    b, a, ab = do.diffusion_schedule(T)
    tab = eb.step_coefficients(b, a, ab, T, temp, cuda_dev).cpu()
    for t_ in range(T):
        coef, c1, sigma = do.step_coefficients(b, a, ab, t_, temp)
        assert tab[t_, 0].item() == float(coef) and tab[t_, 1].item() == float(c1) \
            and tab[t_, 2].item() == float(sigma), t_


@pytest.mark.parametrize("n", [1, 3, 29, 16 * 29, 4097, 256 * 29 * 4 + 1])
def test_posterior_update_bit_exact(cuda_dev, n):
    g = torch.Generator().manual_seed(n)
    x, e, z = (torch.randn(n, generator=g) * 3 for _ in range(3))
    b, a, ab = do.diffusion_schedule(1000)
    for t_ in (999, 500, 1):
        coef, c1, sigma = do.step_coefficients(b, a, ab, t_, 1.0)
        ref = do.posterior_update(x, e, z, coef, c1, sigma)
        got = eb.posterior_update(x.to(cuda_dev), e.to(cuda_dev), z.to(cuda_dev), float(coef), float(c1), float(sigma))
        assert torch.equal(got.cpu(), ref)
    coef, c1, sigma = do.step_coefficients(b, a, ab, 0, 1.0)
    ref = do.posterior_update(x, e, None, coef, c1, sigma)
    got = eb.posterior_update(x.to(cuda_dev), e.to(cuda_dev), None, float(coef), float(c1), float(sigma))
    assert torch.equal(got.cpu(), ref)
    # unaligned views take the scalar path
    if n > 8:
        got = eb.posterior_update(x.to(cuda_dev)[1:], e.to(cuda_dev)[1:], None, float(coef), float(c1), float(sigma))
        assert torch.equal(got.cpu(), ref[1:])


def test_chain_config1_golden(gpu_model, golden, cuda_dev):
    c = golden("chain_cfg1.npz")
    cond = torch.from_numpy(c["condition"]).to(cuda_dev).expand(16, C, L)
    noise = torch.from_numpy(c["noise"]).to(cuda_dev)
    b, a, ab = eb.get_diffusion_schedule(50)
    x, eps = eb.run_chain(gpu_model, cond, 50, b, a, ab, cuda_dev, noise=noise, return_eps=True)
    for t in (49, 25, 0):
        assert_close(eps[t].cpu(), c[f"eps_t{t}"], X50_RTOL, X50_ATOL, f"eps at t={t} inside the chain")
    assert_close(x.cpu(), c["x0"], X50_RTOL, X50_ATOL, "config 1 fields")
    x2 = eb.sample_model(gpu_model, cond, 50, b, a, ab, P, cuda_dev, num_steps=20, temperature=0.7,
                         noise=noise[:20])
    assert_close(x2.cpu(), c["x0_steps20_temp07"], X50_RTOL, X50_ATOL, "truncated chain, temperature 0.7")


def test_single_step_eps_from_golden_state(gpu_model, golden, cuda_dev):
    # no accumulated drift: feed the reference's own x_t and compare that step's prediction
    c = golden("chain_cfg1.npz")
    cond = torch.from_numpy(c["condition"]).to(cuda_dev).expand(16, C, L)
    for t in (49, 25, 0):
        xt = torch.from_numpy(c[f"x_t{t}"]).to(cuda_dev)
        tt = torch.full((16,), t, dtype=torch.long, device=cuda_dev)
        eps = gpu_model(xt, tt, cond)
        assert_close(eps.cpu(), c[f"eps_t{t}"], EPS_RTOL, EPS_ATOL, f"single step t={t}")


def test_long_chains_golden(gpu_model, golden, cuda_dev):
    g = golden("chain_long.npz")
    c = golden("chain_cfg1.npz")
    cond1 = torch.from_numpy(c["condition"]).to(cuda_dev)
    for name, (B, T, ns) in {"T1000_B4": (4, 1000, None), "T500_B3_steps120": (3, 500, 120)}.items():
        torch.manual_seed(2)
        nz = torch.randn(T if ns is None else ns, B, P)
        assert np.array_equal(nz[:2].numpy(), g[name + "_noise_head"])
        b, a, ab = eb.get_diffusion_schedule(T)
        x = eb.sample_model(gpu_model, cond1.expand(B, C, L), T, b, a, ab, P, cuda_dev,
                            num_steps=ns, noise=nz.to(cuda_dev))
        tol = (X1000_RTOL, X1000_ATOL) if T == 1000 else (X50_RTOL, X50_ATOL)
        assert_close(x.cpu(), g[name], *tol, name)


def test_distinct_conditions_chain_vs_oracle(gpu_model, ref_state_dict, cuda_dev):
    g = torch.Generator().manual_seed(21)
    B, T = 6, 40
    cond = torch.rand(B, C, 900, generator=g)
    noise = torch.randn(T, B, P, generator=g)
    b, a, ab = do.diffusion_schedule(T)
    ref = do.sample_chain(ref_state_dict, cond, T, b, a, ab, P, noise)
    got = eb.sample_model(gpu_model, cond.to(cuda_dev), T, b, a, ab, P, cuda_dev, noise=noise.to(cuda_dev))
    assert_close(got.cpu(), ref, X50_RTOL, X50_ATOL, "distinct conditions")
    # ensemble driver: 3 realisations x 2 conditions, realisation-major (ECD.py:394-412)
    ens = eb.sample_ensemble(gpu_model, cond[:2].to(cuda_dev), T, b, a, ab, P, cuda_dev,
                             n_realizations=3, noise=noise.to(cuda_dev))
    ref_e = do.sample_chain(ref_state_dict, cond[:2].repeat(3, 1, 1), T, b, a, ab, P, noise)
    assert ens.shape == (3, 2, P)
    assert_close(ens.reshape(6, P).cpu(), ref_e, X50_RTOL, X50_ATOL, "ensemble driver")


@pytest.mark.parametrize("B", [1, 5, 256, 700, 1500, 5000])
def test_loop_modes_and_tilings_agree_bitwise(gpu_model, cuda_dev, B):
    # persistent / graph / stream run the same per-member arithmetic -> identical bits, for every
    # members-per-CTA tiling the launcher picks
    T = 24
    g = torch.Generator().manual_seed(B)
    cond = torch.rand(1, C, 300, generator=g).to(cuda_dev).expand(B, C, 300)
    noise = torch.randn(T, B, P, generator=g).to(cuda_dev)
    b, a, ab = eb.get_diffusion_schedule(T)
    outs = {m: eb.sample_model(gpu_model, cond, T, b, a, ab, P, cuda_dev, noise=noise, loop_mode=m)
            for m in ("persistent", "graph", "stream")}
    assert torch.equal(outs["persistent"], outs["graph"])
    assert torch.equal(outs["persistent"], outs["stream"])
    # a member's result does not depend on which tile it sits in
    sub = eb.sample_model(gpu_model, cond[:1], T, b, a, ab, P, cuda_dev, noise=noise[:, :1].contiguous())
    assert torch.equal(sub[0], outs["persistent"][0])


def test_device_rng_chain_equals_replay_of_its_own_draws(gpu_model, ref_state_dict, cuda_dev):
    B, T = 33, 37
    cond = torch.rand(1, C, 500, generator=torch.Generator().manual_seed(3))
    b, a, ab = eb.get_diffusion_schedule(T)
    x_rng = eb.sample_model(gpu_model, cond.to(cuda_dev).expand(B, C, 500), T, b, a, ab, P, cuda_dev, seed=77)
    draws = eb.philox_normal(77, 0, B, P, T, cuda_dev)
    x_rep = eb.sample_model(gpu_model, cond.to(cuda_dev).expand(B, C, 500), T, b, a, ab, P, cuda_dev, noise=draws)
    assert torch.equal(x_rng, x_rep)
    for mode in ("graph", "stream"):
        x_m = eb.sample_model(gpu_model, cond.to(cuda_dev).expand(B, C, 500), T, b, a, ab, P, cuda_dev,
                              seed=77, loop_mode=mode)
        assert torch.equal(x_rng, x_m), mode
    # and the oracle fed with the device's draws agrees within the chain tolerance
    ref = do.sample_chain(ref_state_dict, cond.expand(B, C, 500), T, b, a, ab, P, draws.cpu())
    assert_close(x_rng.cpu(), ref, X50_RTOL, X50_ATOL, "oracle fed with the device draws")


def test_device_rng_is_standard_normal_and_seeded(cuda_dev):
    z = eb.philox_normal(5, 0, 4096, 29, 64, cuda_dev).double().cpu().numpy().ravel()
    assert abs(z.mean()) < 2e-3 and abs(z.var() - 1) < 4e-3
    from scipy import stats
    assert stats.kstest(z[:200000], "norm").pvalue > 1e-3
    assert abs(stats.skew(z)) < 5e-3 and abs(stats.kurtosis(z)) < 1e-2
    z2 = eb.philox_normal(5, 0, 4096, 29, 64, cuda_dev).double().cpu().numpy().ravel()
    z3 = eb.philox_normal(6, 0, 4096, 29, 64, cuda_dev).double().cpu().numpy().ravel()
    assert np.array_equal(z, z2) and not np.array_equal(z, z3)
    assert abs(np.corrcoef(z, z3)[0, 1]) < 3e-3


def test_member_offset_makes_shards_identical_to_the_whole(gpu_model, cuda_dev):
    B, T = 96, 20
    cond = torch.rand(1, C, 300, generator=torch.Generator().manual_seed(4)).to(cuda_dev)
    b, a, ab = eb.get_diffusion_schedule(T)
    whole = eb.run_chain(gpu_model, cond.expand(B, C, 300), T, b, a, ab, cuda_dev, seed=9, offset=0)
    parts = [eb.run_chain(gpu_model, cond.expand(hi - lo, C, 300), T, b, a, ab, cuda_dev, seed=9, offset=0,
                          member_offset=lo) for lo, hi in ((0, 24), (24, 48), (48, 96))]
    assert torch.equal(whole, torch.cat(parts))
    noise = torch.randn(T, B, P, generator=torch.Generator().manual_seed(5)).to(cuda_dev)
    whole = eb.run_chain(gpu_model, cond.expand(B, C, 300), T, b, a, ab, cuda_dev, noise=noise)
    parts = [eb.run_chain(gpu_model, cond.expand(hi - lo, C, 300), T, b, a, ab, cuda_dev,
                          noise=noise[:, lo:hi]) for lo, hi in ((0, 40), (40, 96))]
    assert torch.equal(whole, torch.cat(parts))


def test_edge_cases(gpu_model, cuda_dev):
    b, a, ab = eb.get_diffusion_schedule(10)
    x = eb.sample_model(gpu_model, torch.rand(1, C, 5).to(cuda_dev), 10, b, a, ab, P, cuda_dev, num_steps=1, seed=1)
    assert x.shape == (1, P) and torch.isfinite(x).all()
    with pytest.raises(ValueError):
        eb.sample_model(gpu_model, torch.rand(2, C, 5).to(cuda_dev), 10, b, a, ab, 28, cuda_dev)
    with pytest.raises(Exception):
        eb.sample_model(gpu_model, torch.rand(2, C, 5).to(cuda_dev), 10, b, a, ab, P, cuda_dev, num_steps=11)
    out = gpu_model(torch.zeros(0, P, device=cuda_dev), torch.zeros(0, dtype=torch.long, device=cuda_dev),
                    torch.zeros(0, C, 9, device=cuda_dev))
    assert out.shape == (0, P)


def test_full_size_config2_golden(gpu_model, golden, cuda_dev):
    # BASELINE config 2 at full size (256 members, T = 1000, fp32) against the REFERENCE's own output
    # (tests/golden/chain_cfg2.npz, made by oracle/make_golden.py round2 from ECD.py:102-119 unmodified);
    # the noise is regenerated from its seed and guarded by its first rows
    g = golden("chain_cfg2.npz")
    cond1 = torch.from_numpy(golden("chain_cfg1.npz")["condition"]).to(cuda_dev)
    B, T = 256, 1000
    torch.manual_seed(2)
    nz = torch.randn(T, B, P)
    assert np.array_equal(nz[:2].numpy(), g["cfg2_B256_T1000_noise_head"])
    b, a, ab = eb.get_diffusion_schedule(T)
    x = eb.sample_model(gpu_model, cond1.expand(B, C, L), T, b, a, ab, P, cuda_dev, noise=nz.to(cuda_dev))
    assert_close(x.cpu(), g["cfg2_B256_T1000"], X1000_RTOL, X1000_ATOL, "config 2 fields")
    # every loop mode and the members run alone give the same bits
    xg = eb.sample_model(gpu_model, cond1.expand(B, C, L), T, b, a, ab, P, cuda_dev, noise=nz.to(cuda_dev), loop_mode="graph")
    assert torch.equal(x, xg)
    x8 = eb.run_chain(gpu_model, cond1.expand(8, C, L), T, b, a, ab, cuda_dev, noise=nz[:, 100:108].to(cuda_dev))
    assert torch.equal(x8, x[100:108])


def test_full_size_config2_properties(gpu_model, cuda_dev):
    # the same size with the device RNG: finite, reproducible, and identical to the same members run alone
    B, T = 256, 1000
    cond = torch.rand(1, C, L, generator=torch.Generator().manual_seed(1)).to(cuda_dev)
    b, a, ab = eb.get_diffusion_schedule(T)
    x1 = eb.sample_model(gpu_model, cond.expand(B, C, L), T, b, a, ab, P, cuda_dev, seed=3)
    x2 = eb.sample_model(gpu_model, cond.expand(B, C, L), T, b, a, ab, P, cuda_dev, seed=3)
    assert torch.isfinite(x1).all() and torch.equal(x1, x2)
    x8 = eb.run_chain(gpu_model, cond.expand(8, C, L), T, b, a, ab, cuda_dev, seed=3, offset=0, member_offset=100)
    assert torch.equal(x8, x1[100:108])


@pytest.mark.parametrize("upt,mpb", [(1, 1), (2, 1), (1, 2), (2, 2), (2, 4), (1, 8)])
def test_chain_kernel_variants_agree(gpu_model, golden, cuda_dev, monkeypatch, upt, mpb):
    # members per CTA x hidden units per thread: every build is checked against the reference golden; builds
    # with the same number of units per thread run the same per-member arithmetic (bit-identical)
    c = golden("chain_cfg1.npz")
    cond = torch.from_numpy(c["condition"]).to(cuda_dev).expand(16, C, L)
    noise = torch.from_numpy(c["noise"]).to(cuda_dev)
    b, a, ab = eb.get_diffusion_schedule(50)
    base = eb.run_chain(gpu_model, cond, 50, b, a, ab, cuda_dev, noise=noise)
    monkeypatch.setenv("ERTDIFF_CHAIN_UPT", str(upt))
    monkeypatch.setenv("ERTDIFF_CHAIN_MPB", str(mpb))
    x = eb.run_chain(gpu_model, cond, 50, b, a, ab, cuda_dev, noise=noise)
    assert_close(x.cpu(), c["x0"], X50_RTOL, X50_ATOL, f"upt={upt} mpb={mpb}")
    xr = eb.run_chain(gpu_model, cond, 50, b, a, ab, cuda_dev, seed=11, offset=0)
    monkeypatch.setenv("ERTDIFF_CHAIN_MPB", "1")
    xr1 = eb.run_chain(gpu_model, cond, 50, b, a, ab, cuda_dev, seed=11, offset=0)
    assert torch.equal(xr, xr1)                       # the tiling never changes a member's result
    assert_close(base.cpu(), x.cpu(), X50_RTOL, X50_ATOL, "default build vs this one")


def test_hidden256_long_grid_chain_golden(golden, cuda_dev):
    # BASELINE config 5's reference-expressible widening: hidden_dim = 256 (ECD.py:123), L = 9386 ("2x grid",
    # ECD.py:134-138 take any L), two distinct conditions x 4 realisations, T = 200 -- against the reference's
    # own sample_model output; fp32 CUDA-core chain in both thread layouts
    h = golden("model_h256_case.npz")
    sd = {k[3:]: torch.from_numpy(h[k]) for k in h.files if k.startswith("sd.")}
    m = eb.ConditionalDiffusionModel(29, 256)
    m.load_state_dict(sd)
    m.to(cuda_dev).eval()
    g = golden("chain_h256.npz")
    gen = torch.Generator().manual_seed(int(g["cond_seed"]))
    cond = torch.rand(2, C, 2 * L, generator=gen)
    nz = torch.randn(200, 8, P, generator=gen)
    assert np.array_equal(nz[:2].numpy(), g["noise_head"]) and np.array_equal(cond[:, :2, :8].numpy(), g["cond_head"])
    b, a, ab = eb.get_diffusion_schedule(200)
    x = eb.sample_ensemble(m, cond.to(cuda_dev), 200, b, a, ab, P, cuda_dev, n_realizations=4, noise=nz.to(cuda_dev))
    assert_close(x.reshape(8, P).cpu(), g["x0"], X50_RTOL, X50_ATOL, "hidden 256, L 9386")
    # device RNG at config-5 scale per GPU (512 members): finite and tiling-independent
    xr = eb.run_chain(m, cond[:1].to(cuda_dev).expand(512, C, 2 * L), 200, b, a, ab, cuda_dev, seed=2, offset=0)
    x4 = eb.run_chain(m, cond[:1].to(cuda_dev).expand(4, C, 2 * L), 200, b, a, ab, cuda_dev, seed=2, offset=0, member_offset=300)
    assert torch.isfinite(xr).all() and torch.equal(x4, xr[300:304])
