"""Schedule, timestep embedding and the DDPM reverse chain: drop-ins for
``get_diffusion_schedule`` (ECD.py:90-94), ``get_timestep_embedding`` (ECD.py:80-88) and
``sample_model`` (ECD.py:102-119), plus the ensemble driver of ECD.py:394-412 as one batched
chain (``sample_ensemble``).

The chain runs entirely inside ``libertdiff_b200.so``: condition encoder once per distinct
condition, the per-step time table, then one persistent kernel (or a CUDA graph of step
kernels) for all ``num_steps`` steps.  Noise is either injected (``noise=``, parity runs) or
drawn on the device by a Philox counter RNG seeded from torch's CUDA generator.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .model import ConditionalDiffusionModel, IN_CHANNELS


def get_diffusion_schedule(T, beta_start=1e-4, beta_end=0.02, device="cpu"):
    """ECD.py:90-94.  Tiny, host-side, done once; the chain consumes the caller's tensors
    as they are (the sampler uses ``1 - alphas[t]``, not ``betas[t]``)."""
    betas = torch.linspace(beta_start, beta_end, T, device=device)
    alphas = 1 - betas
    alpha_bar = torch.cumprod(alphas, dim=0)
    return betas, alphas, alpha_bar


def get_timestep_embedding(timesteps, embedding_dim):
    """ECD.py:80-88: ``[sin(t f_i) | cos(t f_i)]`` with ``f_i = exp(-i ln(1e4)/(half-1))``.
    Provided for API completeness (the kernels evaluate it on the device)."""
    half = embedding_dim // 2
    scale = math.log(10000.0) / (half - 1)
    freq = torch.exp(torch.arange(half, device=timesteps.device, dtype=torch.float32) * -scale)
    arg = timesteps.float().unsqueeze(1) * freq.unsqueeze(0)
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(timesteps.size(0), 1, device=timesteps.device)], dim=1)
    return emb


def _next_philox_stream(device):
    """(seed, offset) for one chain, taken from torch's CUDA generator so that
    ``torch.manual_seed`` makes sampling reproducible; the generator offset is advanced."""
    gen = torch.cuda.default_generators[device.index if device.index is not None
                                        else torch.cuda.current_device()]
    seed = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF
    offset = int(gen.get_offset())
    gen.set_offset(offset + 4)          # offsets must stay multiples of 4
    return seed, offset


_SCHEDULE_CACHE = {}          # (device, numel) -> (host copies, device copies); tiny tensors


def _schedule_on(device, *tensors):
    """The caller's (T,) schedule tensors on `device`.  They are usually built once on the host
    (``get_diffusion_schedule(T)`` defaults to CPU, ECD.py:90) and passed to every
    ``sample_model`` call: the device copies are cached and re-used while the CONTENT is
    unchanged (checked bit for bit against a host snapshot -- 3 x 4 KB)."""
    if all(t.device == device and t.dtype == torch.float32 and t.is_contiguous() for t in tensors):
        return [t.detach() for t in tensors]
    if any(t.device.type != "cpu" for t in tensors):
        return [t.detach().to(device=device, dtype=torch.float32).contiguous() for t in tensors]
    key = (device, tuple(t.numel() for t in tensors))
    hit = _SCHEDULE_CACHE.get(key)
    if hit is not None and all(h.dtype == t.dtype and torch.equal(h, t) for h, t in zip(hit[0], tensors)):
        return hit[1]
    host = [t.detach().clone() for t in tensors]
    packed = torch.stack([h.to(torch.float32) for h in host]).to(device)     # one H2D copy
    dev = [packed[i] for i in range(len(host))]
    if len(_SCHEDULE_CACHE) > 16:
        _SCHEDULE_CACHE.clear()
    _SCHEDULE_CACHE[key] = (host, dev)
    return dev


@torch.no_grad()
def run_chain(model, condition, T, betas, alphas, alpha_bar, device, num_steps=None,
              temperature=1.0, *, n_members=None, noise=None, seed=None, offset=0,
              member_offset=0, loop_mode="persistent", precision="fp32", return_eps=False,
              check_status=True):
    """The engine behind ``sample_model`` / ``sample_ensemble``.

    condition: ``(n_cond, 14, L)``; member ``i`` of the ``n_members`` (default ``n_cond``)
    uses condition ``i % n_cond``.  ``noise``: optional ``(num_steps, n_members, P)`` tensor, row 0
    = ``x_T``, row k = k-th in-loop draw (the order ``torch.randn`` is called in ECD.py:107,116).
    Returns ``x_0 (n_members, P)`` on ``device`` (and the per-step predicted noise
    ``(num_steps, n_members, P)`` indexed by t when ``return_eps``).

    ``precision``: ``"fp32"`` (CUDA-core FFMA, pinned to the reference), ``"bf16"`` (tensor cores, bf16 operands,
    fp32 accumulation; hidden_dim 128 or 256) or ``"bf16x3"`` (tensor cores, every operand as two bf16 terms and
    three accumulating products per projection: fp32-class fields at tensor-core speed; hidden_dim 128; the
    condition encoder stays in fp32).  Tensor-core modes: a tile whose MMA never completes (a hardware or
    driver fault; the waits are bounded) poisons its members with NaN and raises the handle's status
    word.  With ``check_status`` (default) that word is read after the launch -- one 4-byte D2H copy,
    which waits for the chain -- and ``ErtdiffError`` is raised; a caller that pipelines several
    chains passes ``check_status=False`` and polls ``model.umma_status()`` itself.
    """
    if not isinstance(model, ConditionalDiffusionModel):
        raise TypeError("model must be an ertdiff_b200 ConditionalDiffusionModel")
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.ErtdiffError(f"device={device}: the sampler runs on CUDA only (no CPU path)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if model.device != device:
        raise _lib.ErtdiffError(f"model is on {model.device}, sampling requested on {device}")
    if num_steps is None:
        num_steps = T
    num_steps, T = int(num_steps), int(T)
    P = model.param_dim
    h = model.handle()
    lib = _lib.load()

    condition = condition.to(device=device, dtype=torch.float32)
    cond, stride = ConditionalDiffusionModel._condition_layout(condition)
    n_rows = condition.size(0)
    n_cond = 1 if stride == 0 else n_rows
    B = int(n_members) if n_members is not None else n_rows
    if B % n_cond != 0 and n_cond != 1:
        raise ValueError("n_members must be a multiple of the number of conditions")
    L = cond.size(2)
    betas_d, alphas_d, abar_d = _schedule_on(device, betas, alphas, alpha_bar)
    if min(betas_d.numel(), alphas_d.numel(), abar_d.numel()) < num_steps:
        raise ValueError("schedule tensors are shorter than num_steps")

    x_out = torch.empty(B, P, device=device, dtype=torch.float32)
    if B == 0 or num_steps == 0:
        return x_out
    eps = torch.empty(num_steps, B, P, device=device, dtype=torch.float32) if return_eps else None

    args = _lib.ChainArgs()
    args.B, args.n_cond, args.T, args.num_steps = B, n_cond, T, num_steps
    args.temperature = float(temperature)
    args.d_betas, args.d_alphas, args.d_alpha_bar = (betas_d.data_ptr(), alphas_d.data_ptr(),
                                                     abar_d.data_ptr())
    args.d_cond_bias = None
    keep = None
    if noise is not None:
        noise = noise.to(device=device, dtype=torch.float32)
        if noise.dim() != 3 or noise.size(0) < num_steps or noise.size(2) != P or noise.size(1) < B:
            raise ValueError(f"noise must be (>= {num_steps}, >= {B}, {P}), got {tuple(noise.shape)}")
        # a member-slice of a larger (steps, B_total, P) tensor is read in place
        sliced = (noise.stride(2) == 1 and noise.stride(1) == P and noise.stride(0) % P == 0
                  and noise.stride(0) >= noise.size(1) * P)
        if not sliced:
            noise = noise.contiguous()
        keep = noise
        args.d_x_T = noise.data_ptr()
        args.d_noise = noise[1:].data_ptr() if num_steps > 1 else None
        args.noise_member_stride_B = noise.stride(0) // P if noise.size(0) > 1 else noise.size(1)
        args.seed = args.offset = 0
    else:
        args.d_x_T = None
        args.d_noise = None
        args.noise_member_stride_B = B
        if seed is None:
            seed, offset = _next_philox_stream(device)
        args.seed, args.offset = int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset)
    args.member_offset = int(member_offset)
    args.loop_mode = _lib.LOOP_MODES[loop_mode]
    args.precision = _lib.PRECISIONS[precision]
    args.d_x_out = x_out.data_ptr()
    args.d_eps_trace = eps.data_ptr() if eps is not None else None
    with torch.cuda.device(device):
        _lib.check(lib.ertdiff_sample_model(h, _lib.ptr(cond), L, stride if n_cond > 1 else 0,
                                            C.byref(args), _lib.stream_ptr(device)),
                   "sample_model")
    del keep
    if check_status and precision in ("bf16", "bf16x3") and model.umma_status() != 0:
        raise _lib.ErtdiffError("tensor-core chain: an MMA completion wait timed out; the affected members are NaN")
    return (x_out, eps) if return_eps else x_out


def sample_model(model, condition, T, betas, alphas, alpha_bar, param_dim, device,
                 num_steps=None, temperature=1.0, *, noise=None, seed=None,
                 loop_mode="persistent", precision="fp32", check_status=True):
    """Drop-in for the reference's ``sample_model`` (ECD.py:102-119): same positional
    signature, returns ``x_0 (B, param_dim)`` float32 on ``device``.

    Keyword-only extras: ``noise`` replays a ``(num_steps, B, param_dim)`` tensor instead of
    drawing (row 0 = ``x_T``); ``seed`` fixes the device RNG stream; ``loop_mode`` is
    ``"persistent"`` (one kernel for the whole chain), ``"graph"`` (CUDA graph of step kernels)
    or ``"stream"``.
    """
    if int(param_dim) != model.param_dim:
        raise ValueError(f"param_dim={param_dim} but the model was built with {model.param_dim}")
    return run_chain(model, condition, T, betas, alphas, alpha_bar, device, num_steps,
                     temperature, noise=noise, seed=seed, loop_mode=loop_mode,
                     precision=precision, check_status=check_status)


def sample_ensemble(model, condition, T, betas, alphas, alpha_bar, param_dim, device,
                    n_realizations=50, num_steps=None, temperature=1.0, *, noise=None,
                    seed=None, loop_mode="persistent", precision="fp32", check_status=True):
    """The ensemble driver of ECD.py:394-412 / 1037-1079 as ONE batched chain.

    The reference loops ``for realization in range(50): sample_model(...)`` over the same
    ``n_cond`` conditions and stacks to ``(50, n_cond, P)``.  Here the ``n_realizations *
    n_cond`` members run together (realisation-major, member ``r*n_cond + c``), the condition
    encoder runs once per distinct condition, and the result comes back already stacked:
    ``(n_realizations, n_cond, param_dim)``.  ``noise``: optional
    ``(num_steps, n_realizations*n_cond, P)``.
    """
    if int(param_dim) != model.param_dim:
        raise ValueError(f"param_dim={param_dim} but the model was built with {model.param_dim}")
    n_cond = condition.size(0)
    x = run_chain(model, condition, T, betas, alphas, alpha_bar, device, num_steps, temperature,
                  n_members=n_realizations * n_cond, noise=noise, seed=seed,
                  loop_mode=loop_mode, precision=precision, check_status=check_status)
    return x.view(n_realizations, n_cond, model.param_dim)


def step_coefficients(betas, alphas, alpha_bar, num_steps, temperature=1.0, device="cuda"):
    """The (num_steps, 3) table ``[coef, 1/sqrt(alpha_t), sqrt(beta_t)*temperature]`` of
    ECD.py:111-118 as the chain kernel uses it (exposed for tests)."""
    device = torch.device(device)
    b, a, ab = _schedule_on(device, betas, alphas, alpha_bar)
    out = torch.empty(num_steps, 4, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(_lib.load().ertdiff_step_coefficients(
            _lib.ptr(b), _lib.ptr(a), _lib.ptr(ab), int(num_steps), float(temperature),
            _lib.ptr(out), _lib.stream_ptr(device)), "step_coefficients")
    return out[:, :3]


def posterior_update(x, eps, z, coef, c1, sigma):
    """ECD.py:114-118 as one vectorised kernel: ``c1*(x - coef*eps) [+ sigma*z]``."""
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ertdiff_posterior_update(
            _lib.ptr(x), _lib.ptr(eps.contiguous()), _lib.ptr(z.contiguous()) if z is not None else None,
            float(coef), float(c1), float(sigma), x.numel(), _lib.ptr(out),
            _lib.stream_ptr(x.device)), "posterior_update")
    return out


def philox_normal(seed, offset, n_members, param_dim, draws, device="cuda", member_offset=0):
    """The device noise source of the chain, materialised: ``(draws, n_members, P)``."""
    device = torch.device(device)
    out = torch.empty(draws, n_members, param_dim, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(_lib.load().ertdiff_philox_normal(
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset), int(member_offset), n_members, param_dim,
            draws, _lib.ptr(out), _lib.stream_ptr(device)), "philox_normal")
    return out


def debug_umma_gemm(a, b):
    """Self-test of the tcgen05 path: ``a (128,K) @ b (N,K)^T`` with bf16 operands on the tensor
    cores (``ertdiff_debug_umma_gemm``)."""
    a = a.contiguous().float()
    b = b.contiguous().float()
    out = torch.empty(128, b.size(0), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().ertdiff_debug_umma_gemm(_lib.ptr(a), _lib.ptr(b), b.size(0), a.size(1),
                                                       _lib.ptr(out), _lib.stream_ptr(a.device)),
                   "debug_umma_gemm")
    return out
