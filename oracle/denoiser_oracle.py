"""torch-CPU restatement of the reference denoiser and DDPM reverse chain.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker, never the product.

Every function cites the reference lines it restates
(``/root/reference/ERT_Conditional_Diffusion.py`` = ``ECD.py``).  The restatement is
functional (it works on the reference's 12-tensor ``state_dict``) rather than an
``nn.Module``, and is pinned bit-for-bit against the reference itself by
``tests/test_oracle_pinning.py`` (in the build container) and against
``tests/golden`` (everywhere).

Pinning status: pinned to reference outputs generated in the build container by
``oracle/make_golden.py`` (the reference has no tests or golden vectors of its own).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# the reference's state_dict layout (ECD.py:133-153), name -> shape for (P, H)
def state_dict_spec(param_dim: int = 29, hidden_dim: int = 128, in_channels: int = 14):
    P, H, C = param_dim, hidden_dim, in_channels
    return OrderedDict([
        ("condition_encoder.0.weight", (32, C, 3)), ("condition_encoder.0.bias", (32,)),
        ("condition_encoder.2.weight", (64, 32, 3)), ("condition_encoder.2.bias", (64,)),
        ("condition_encoder.6.weight", (H, 64)), ("condition_encoder.6.bias", (H,)),
        ("time_embed.0.weight", (H, H)), ("time_embed.0.bias", (H,)),
        ("mlp.0.weight", (H, P + 2 * H)), ("mlp.0.bias", (H,)),
        ("mlp.2.weight", (P, H)), ("mlp.2.bias", (P,)),
    ])


def init_state_dict(param_dim=29, hidden_dim=128, seed=0, in_channels=14):
    """Random weights with torch's default Conv1d/Linear initialisation law
    (uniform(+-1/sqrt(fan_in)) for weight and bias), drawn from a private generator.

    NOT bit-identical to ``torch.manual_seed(seed); ConditionalDiffusionModel(...)``
    (different draw order) -- golden fixtures use the reference class itself; this is
    for synthetic benchmark weights on machines without the reference.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    spec = state_dict_spec(param_dim, hidden_dim, in_channels)
    fan_in = {}
    for name, shape in spec.items():
        if name.endswith("weight"):
            fi = 1
            for s in shape[1:]:
                fi *= s
            fan_in[name[:-len("weight")]] = fi
    for name, shape in spec.items():
        bound = 1.0 / math.sqrt(fan_in[name.rsplit(".", 1)[0] + "."])
        sd[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    return sd


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """ECD.py:80-88.  ``[sin(t f_i) | cos(t f_i)]``, ``f_i = exp(-i ln(1e4)/(half-1))``."""
    half = dim // 2
    scale = math.log(10000.0) / (half - 1)                     # python double
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -scale)
    arg = t.float().unsqueeze(1) * freq.unsqueeze(0)
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(t.size(0), 1, device=t.device)], dim=1)
    return emb


def diffusion_schedule(T: int, beta_start: float = 1e-4, beta_end: float = 0.02, device="cpu"):
    """ECD.py:90-94."""
    betas = torch.linspace(beta_start, beta_end, T, device=device)
    alphas = 1 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def encode_condition(sd, condition: torch.Tensor) -> torch.Tensor:
    """ECD.py:133-142: conv(s2,p1)-relu-conv(s2,p1)-relu-global mean-linear-relu."""
    h = F.relu(F.conv1d(condition, sd["condition_encoder.0.weight"],
                        sd["condition_encoder.0.bias"], stride=2, padding=1))
    h = F.relu(F.conv1d(h, sd["condition_encoder.2.weight"],
                        sd["condition_encoder.2.bias"], stride=2, padding=1))
    h = F.adaptive_avg_pool1d(h, 1).flatten(1)
    return F.relu(F.linear(h, sd["condition_encoder.6.weight"], sd["condition_encoder.6.bias"]))


def denoiser_forward(sd, x: torch.Tensor, t: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
    """ECD.py:155-164 as written (nothing hoisted): returns predicted noise (B, P)."""
    H = sd["time_embed.0.weight"].shape[1]
    t_emb = F.relu(F.linear(timestep_embedding(t, H), sd["time_embed.0.weight"],
                            sd["time_embed.0.bias"]))
    c_emb = encode_condition(sd, condition)
    h = torch.cat([x, t_emb, c_emb], dim=1)
    h = F.relu(F.linear(h, sd["mlp.0.weight"], sd["mlp.0.bias"]))
    return F.linear(h, sd["mlp.2.weight"], sd["mlp.2.bias"])


def step_coefficients(betas, alphas, alpha_bar, t_: int, temperature: float = 1.0):
    """The three per-step scalars of ECD.py:111-118 with the reference's exact rounding:
    ``1-alpha`` and ``1-alpha_bar`` are f32 tensor ops, ``math.sqrt`` goes through a python
    double, the divisor and results are rounded back to f32."""
    one_minus_alpha = (1 - alphas[t_])                       # f32 0-d tensor
    denom = math.sqrt(1 - alpha_bar[t_]) + 1e-8              # double
    coef = one_minus_alpha / denom                           # f32 tensor / python scalar -> f32
    c1 = torch.tensor(1.0 / math.sqrt(alphas[t_]), dtype=torch.float32)
    sigma = torch.tensor(math.sqrt(betas[t_]) * temperature, dtype=torch.float32)
    return coef, c1, sigma


def posterior_update(x, eps, z, coef, c1, sigma):
    """ECD.py:114-118: ``x <- c1*(x - coef*eps) [+ sigma*z]``, each op rounded to f32, no FMA."""
    x = c1 * (x - coef * eps)
    if z is not None:
        x = x + sigma * z
    return x


@torch.no_grad()
def sample_chain(sd, condition, T, betas, alphas, alpha_bar, param_dim, noise,
                 num_steps=None, temperature=1.0, trace_eps_at=()):
    """ECD.py:102-119 with replayed noise: ``noise`` is ``(num_steps, B, P)``, row 0 = x_T,
    row k (k>=1) the k-th in-loop draw.  Returns ``x_0`` and, optionally, the predicted noise
    at the timesteps listed in ``trace_eps_at``."""
    if num_steps is None:
        num_steps = T
    x = noise[0].clone()
    draw = 1
    B = condition.size(0)
    trace = {}
    for t_ in reversed(range(num_steps)):
        t_tensor = torch.full((B,), t_, dtype=torch.long)
        eps = denoiser_forward(sd, x, t_tensor, condition)
        if t_ in trace_eps_at:
            trace[t_] = eps.clone()
        coef, c1, sigma = step_coefficients(betas, alphas, alpha_bar, t_, temperature)
        z = None
        if t_ > 0:
            z = noise[draw]
            draw += 1
        x = posterior_update(x, eps, z, coef, c1, sigma)
    return (x, trace) if trace_eps_at else x


@torch.no_grad()
def sample_chain_hoisted(sd, condition, T, betas, alphas, alpha_bar, param_dim, noise,
                         num_steps=None, temperature=1.0):
    """The same chain with the loop-invariant work hoisted out of the step loop -- the algebra the CUDA
    path uses, in plain torch on the CPU: the condition encoder runs once per DISTINCT condition
    (a stride-0 ``expand`` is encoded once), the time embedding once per step for all members, and a step is
    ``eps = W2 @ relu(W0x @ x + c_t + c_b) + b2``.  Not bit-identical to ``sample_chain`` (different summation
    split; ``tests/test_oracle_pinning.py`` bounds the difference): it exists as the CPU baseline that separates
    the algorithmic saving from the hardware speed-up in ``bench.py``."""
    if num_steps is None:
        num_steps = T
    P = param_dim
    H = sd["time_embed.0.weight"].shape[1]
    W0 = sd["mlp.0.weight"]
    W0x, W0t, W0c = W0[:, :P].contiguous(), W0[:, P:P + H].contiguous(), W0[:, P + H:].contiguous()
    shared = condition.size(0) > 1 and condition.stride(0) == 0
    cemb = encode_condition(sd, condition[:1] if shared else condition)
    cb = F.linear(cemb, W0c, sd["mlp.0.bias"])                     # (n_cond, H); broadcasts when shared
    ts = torch.arange(num_steps)
    temb = F.relu(F.linear(timestep_embedding(ts, H), sd["time_embed.0.weight"], sd["time_embed.0.bias"]))
    ct = F.linear(temb, W0t)                                       # (num_steps, H)
    x = noise[0].clone()
    draw = 1
    for t_ in reversed(range(num_steps)):
        h = F.relu(F.linear(x, W0x) + ct[t_] + cb)
        eps = F.linear(h, sd["mlp.2.weight"], sd["mlp.2.bias"])
        coef, c1, sigma = step_coefficients(betas, alphas, alpha_bar, t_, temperature)
        z = None
        if t_ > 0:
            z = noise[draw]
            draw += 1
        x = posterior_update(x, eps, z, coef, c1, sigma)
    return x


def logistic_unconstrain_inverse(u, a, b):
    """ECD.py:42-53 (tensor branch): ``a + (b-a)*sigmoid(u)``."""
    return a + (b - a) * torch.sigmoid(u)
