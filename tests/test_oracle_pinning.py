"""The oracle restatement against the golden vectors generated from the reference itself
(oracle/make_golden.py) and, where /root/reference exists, against the reference live."""
import numpy as np
import pytest
import torch

from oracle import denoiser_oracle as do
from oracle.reference_loader import load_reference, reference_available

P, C, L = 29, 14, 4693


def test_schedule_and_embedding_match_golden(golden):
    g = golden("embedding_schedule.npz")
    for T in (50, 500, 1000):
        b, a, ab = do.diffusion_schedule(T)
        assert np.array_equal(b.numpy(), g[f"betas{T}"])
        assert np.array_equal(a.numpy(), g[f"alphas{T}"])
        assert np.array_equal(ab.numpy(), g[f"alpha_bar{T}"])
    emb = do.timestep_embedding(torch.from_numpy(g["t"]), 128)
    assert np.array_equal(emb.numpy(), g["emb"])
    # known answer: t = 0 -> 64 zeros then 64 ones
    assert np.array_equal(emb[0].numpy(), np.r_[np.zeros(64), np.ones(64)].astype(np.float32))


def test_forward_cases_bit_exact(golden, ref_state_dict):
    f = golden("forward_cases.npz")
    for tag in ("L257", "L64", "L3", "L1", "L1000"):
        e = do.denoiser_forward(ref_state_dict, torch.from_numpy(f[tag + "_x"]),
                                torch.from_numpy(f[tag + "_t"]), torch.from_numpy(f[tag + "_cond"]))
        assert np.array_equal(e.numpy(), f[tag + "_eps"]), tag
        ce = do.encode_condition(ref_state_dict, torch.from_numpy(f[tag + "_cond"]))
        assert np.array_equal(ce.numpy(), f[tag + "_cemb"]), tag


def test_chain_config1_bit_exact(golden, ref_state_dict):
    c = golden("chain_cfg1.npz")
    cond = torch.from_numpy(c["condition"]).expand(16, C, L)
    noise = torch.from_numpy(c["noise"])
    b, a, ab = do.diffusion_schedule(50)
    x, tr = do.sample_chain(ref_state_dict, cond, 50, b, a, ab, P, noise, trace_eps_at=(49, 25, 0))
    assert np.array_equal(x.numpy(), c["x0"])
    for t in (49, 25, 0):
        assert np.array_equal(tr[t].numpy(), c[f"eps_t{t}"])
    x = do.sample_chain(ref_state_dict, cond, 50, b, a, ab, P, noise[:20], num_steps=20, temperature=0.7)
    assert np.array_equal(x.numpy(), c["x0_steps20_temp07"])


def test_init_state_dict_shapes():
    sd = do.init_state_dict(29, 128, seed=0)
    spec = do.state_dict_spec(29, 128)
    assert list(sd) == list(spec)
    assert all(tuple(sd[k].shape) == spec[k] for k in spec)
    assert sum(v.numel() for v in sd.values()) == 72765


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_restatement_equals_reference_live():
    ref = load_reference()
    torch.manual_seed(11)
    m = ref.ConditionalDiffusionModel(P, 128).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x, t, c = torch.randn(3, P, generator=g), torch.randint(0, 1000, (3,), generator=g), torch.rand(3, C, 211, generator=g)
    with torch.no_grad():
        assert torch.equal(m(x, t, c), do.denoiser_forward(sd, x, t, c))
    noise = torch.randn(30, 3, P, generator=g)
    ref2 = load_reference(noise=noise)
    b, a, ab = ref2.get_diffusion_schedule(40)
    xr = ref2.sample_model(m, c, 40, b, a, ab, P, "cpu", num_steps=30, temperature=1.3)
    xo = do.sample_chain(sd, c, 40, b, a, ab, P, noise, num_steps=30, temperature=1.3)
    assert torch.equal(xr, xo)
    assert ref2.torch_proxy.draws == 30


def test_f1_restatement_equals_the_reference_fixture(golden):
    # the oracle-side restatement of the f1 epilogue (what the GPU test's tolerances are stated against) reproduces
    # the reference-made fixture bit for bit: torch sigmoid, sklearn's `X -= min_; X /= scale_`, the bounds loop
    import torch
    g = golden("transforms_f1.npz")
    s = do.logistic_unconstrain_inverse(torch.from_numpy(g["u"]), 0.0, 1.0).numpy()
    assert np.array_equal(s, g["sigmoid"])
    x = s.copy()
    x -= g["scaler_min"]
    x /= g["scaler_scale"]
    assert x.dtype == np.float32 and np.array_equal(x, g["phys"])
    lim = g["limits"]
    first = np.array([next((j for j in range(x.shape[1]) if x[i, j] < lim[j, 0] or x[i, j] > lim[j, 1]), -1)
                      for i in range(x.shape[0])], dtype=np.int32)
    assert np.array_equal(first, g["first_bad"]) and np.array_equal(first < 0, g["valid"])


def test_hoisted_chain_stays_within_reorder_noise_of_the_as_written_chain(ref_state_dict, golden):
    # the hoisted CPU chain (bench.py's second CPU baseline) against the as-written oracle and the reference golden
    c = golden("chain_cfg1.npz")
    cond = torch.from_numpy(c["condition"]).expand(16, C, L)
    noise = torch.from_numpy(c["noise"])
    b, a, ab = do.diffusion_schedule(50)
    x = do.sample_chain_hoisted(ref_state_dict, cond, 50, b, a, ab, P, noise)
    assert np.abs(x.numpy() - c["x0"]).max() <= 2e-5 + 1e-4 * np.abs(c["x0"]).max()
    g = torch.Generator().manual_seed(3)
    cond = torch.rand(3, C, 200, generator=g)
    nz = torch.randn(12, 6, P, generator=g)
    b, a, ab = do.diffusion_schedule(40)
    x1 = do.sample_chain(ref_state_dict, cond.repeat(2, 1, 1), 40, b, a, ab, P, nz, num_steps=12, temperature=0.8)
    x2 = do.sample_chain_hoisted(ref_state_dict, cond.repeat(2, 1, 1), 40, b, a, ab, P, nz, num_steps=12, temperature=0.8)
    assert (x1 - x2).abs().max().item() <= 2e-5
