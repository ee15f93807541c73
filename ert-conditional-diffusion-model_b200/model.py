"""Host-side mirror of the reference denoiser, backed by the CUDA library.

``ConditionalDiffusionModel`` has the reference's constructor, call signature, attribute
``param_dim`` and 12-key ``state_dict`` (ERT_Conditional_Diffusion.py:122-164), so the
reference's sampling cells (``load_best_model`` -> ``model.eval()`` -> ``sample_model``,
ECD.py:369-399) run unchanged against it.  It owns no arithmetic: ``forward`` calls
``ertdiff_forward`` in ``libertdiff_b200.so``; parameters are ordinary ``nn.Parameter`` s that
are mirrored into the library's device handle whenever they change.

Inference only (the north-star path): outputs carry no autograd graph.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib

IN_CHANNELS = 14   # ECD.py:134 hard-codes the 14 ERT surveys as conv input channels


class _Affine(nn.Module):
    """A weight/bias pair initialised like torch's Conv1d/Linear ``reset_parameters`` (same
    RNG draws in the same order, so ``torch.manual_seed(s)`` gives the reference's weights)."""

    def __init__(self, *weight_shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*weight_shape))
        self.bias = nn.Parameter(torch.empty(weight_shape[0]))
        fan_in = 1
        for s in weight_shape[1:]:
            fan_in *= s
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
        nn.init.uniform_(self.bias, -bound, bound)


class _Slots(nn.Module):
    """Children registered under the integer names the reference's ``nn.Sequential`` s give
    their parametrised layers, so ``state_dict`` keys match (``condition_encoder.0.weight`` ...)."""

    def __init__(self, slots):
        super().__init__()
        for idx, mod in slots:
            self.add_module(str(idx), mod)

    def __getitem__(self, idx):
        return getattr(self, str(idx))


class ConditionalDiffusionModel(nn.Module):
    """ECD.py:122-164.  ``model(x, t, condition) -> predicted noise``.

    x ``(B, param_dim)`` f32, t ``(B,)`` int64 (per-row timesteps), condition ``(B, 14, L)`` f32.
    All on the CUDA device the model lives on.
    """

    def __init__(self, param_dim, hidden_dim=128):
        super().__init__()
        self.param_dim = int(param_dim)
        self.hidden_dim = int(hidden_dim)
        P, H = self.param_dim, self.hidden_dim
        # creation order == the reference's, so seeded initialisation is identical
        self.condition_encoder = _Slots([(0, _Affine(32, IN_CHANNELS, 3)),   # Conv1d k3 s2 p1
                                         (2, _Affine(64, 32, 3)),            # Conv1d k3 s2 p1
                                         (6, _Affine(H, 64))])               # Linear
        self.time_embed = _Slots([(0, _Affine(H, H))])
        self.mlp = _Slots([(0, _Affine(H, P + 2 * H)), (2, _Affine(P, H))])
        self._handle = None
        self._handle_device = None
        self._stamp = None

    # ------------------------------------------------------------------ library handle
    def _ordered_params(self):
        ce, te, mlp = self.condition_encoder, self.time_embed, self.mlp
        return [ce[0].weight, ce[0].bias, ce[2].weight, ce[2].bias, ce[6].weight, ce[6].bias,
                te[0].weight, te[0].bias, mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias]

    def _frequency_table(self):
        """ECD.py:81-83, evaluated with the very torch ops the reference uses."""
        half = self.hidden_dim // 2
        emb = math.log(10000.0) / (half - 1)
        return torch.exp(torch.arange(half, dtype=torch.float32) * -emb).contiguous()

    def handle(self):
        """The library's device handle, (re)loaded if parameters moved or changed."""
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.ErtdiffError(
                "ConditionalDiffusionModel is on %s: this implementation runs on CUDA only "
                "(call model.to('cuda')); there is no CPU path" % dev)
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise _lib.ErtdiffError("all parameters must be float32 on one CUDA device")
        lib = _lib.load()
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != index:
            self._release()
            h = C.c_void_p()
            _lib.check(lib.ertdiff_model_create(C.byref(h), index, self.param_dim,
                                                self.hidden_dim), "model_create")
            self._handle, self._handle_device, self._stamp = h, index, None
        stamp = tuple((p.data_ptr(), p._version) for p in params)
        if stamp != self._stamp:
            tensors = [p.detach().contiguous() for p in params]
            arr = (C.c_void_p * 12)(*[t.data_ptr() for t in tensors])
            freq = self._frequency_table()
            with torch.cuda.device(index):
                _lib.check(lib.ertdiff_model_load(self._handle, arr, 1, C.c_void_p(freq.data_ptr()),
                                                  _lib.stream_ptr(index)), "model_load")
            self._stamp = stamp
        return self._handle

    def _release(self):
        if self._handle is not None:
            try:
                _lib.load().ertdiff_model_destroy(self._handle)
            except Exception:
                pass
            # object.__setattr__: nn.Module.__setattr__ touches module globals that are already
            # gone when this runs from __del__ at interpreter shutdown
            object.__setattr__(self, "_handle", None)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __getstate__(self):
        # the device handle is not copyable/picklable; a copy re-creates its own on first use
        d = self.__dict__.copy()
        d["_handle"], d["_handle_device"], d["_stamp"] = None, None, None
        return d

    def profile_chain(self, enable=True):
        """Bracket the persistent chain kernel with CUDA events (bench.py's roofline figure)."""
        _lib.check(_lib.load().ertdiff_model_profile(self.handle(), int(bool(enable))), "model_profile")

    def umma_status(self):
        """0 unless a tile of the tensor-core chain timed out (its output is NaN)."""
        st = C.c_int()
        _lib.check(_lib.load().ertdiff_model_umma_status(self.handle(), C.byref(st)), "umma_status")
        return int(st.value)

    def umma_timing(self, enable=True):
        """Development aid (``ertdiff_debug_umma_timing``): returns the 16 cycle sums recorded by
        CTA 0 of the last tensor-core chain launch, then switches the recording on/off."""
        out = (C.c_int64 * 16)()
        _lib.check(_lib.load().ertdiff_debug_umma_timing(self.handle(), int(bool(enable)), out), "umma_timing")
        return [int(v) for v in out]

    def chain_floor(self, enable=True):
        """Measurement aid (``ertdiff_debug_chain_floor``): fp32 persistent chains run the kernel with the
        matrix-vector arithmetic removed -- the latency floor of its structure.  Outputs are meaningless."""
        _lib.check(_lib.load().ertdiff_debug_chain_floor(self.handle(), int(bool(enable))), "chain_floor")

    def graph_stats(self):
        """(graphs instantiated, in-place graph updates) of ``loop_mode="graph"`` on this handle."""
        out = (C.c_int64 * 2)()
        _lib.check(_lib.load().ertdiff_debug_graph_stats(self.handle(), out), "graph_stats")
        return int(out[0]), int(out[1])

    def last_chain_ms(self):
        ms = C.c_float()
        _lib.check(_lib.load().ertdiff_model_last_chain_ms(self.handle(), C.byref(ms)), "last_chain_ms")
        return float(ms.value)

    @property
    def device(self):
        return self.mlp[2].weight.device

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _condition_layout(condition):
        """(contiguous-per-member tensor, member stride in elements; 0 = one shared condition)."""
        if condition.dim() != 3 or condition.size(1) != IN_CHANNELS:
            raise ValueError(f"condition must be (B, {IN_CHANNELS}, L), got {tuple(condition.shape)}")
        L = condition.size(2)
        if condition.size(0) > 1 and condition.stride(0) == 0:
            one = condition[:1].contiguous()          # an expand()ed shared condition
            return one, 0
        c = condition.contiguous()
        return c, IN_CHANNELS * L

    @torch.no_grad()
    def forward(self, x, t, condition):
        h = self.handle()
        dev = self.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        condition = condition.to(device=dev, dtype=torch.float32)
        B = x.size(0)
        if x.dim() != 2 or x.size(1) != self.param_dim:
            raise ValueError(f"x must be (B, {self.param_dim}), got {tuple(x.shape)}")
        if t.shape != (B,) or condition.size(0) != B:
            raise ValueError("x, t and condition disagree on the batch size")
        cond, stride = self._condition_layout(condition)
        out = torch.empty_like(x)
        if B == 0:
            return out
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ertdiff_forward(
                h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(cond), B, cond.size(2), stride,
                _lib.ptr(out), _lib.stream_ptr(dev)), "forward")
        return out

    @torch.no_grad()
    def encode_condition(self, condition, precision="fp32", return_bias=False):
        """``condition_encoder(condition)`` (ECD.py:133-142, 161): ``(n, 14, L) -> (n, H)``.
        ``precision="bf16"`` runs both convolutions on the tensor cores (bf16 operands, fp32
        accumulation).  ``return_bias`` also returns ``mlp.0.weight[:, P+H:] @ emb + mlp.0.bias``,
        the per-condition constant the chain consumes."""
        h = self.handle()
        dev = self.device
        condition = condition.to(device=dev, dtype=torch.float32).contiguous()
        n, L = condition.size(0), condition.size(2)
        emb = torch.empty(n, self.hidden_dim, device=dev, dtype=torch.float32)
        bias = torch.empty(n, self.hidden_dim, device=dev, dtype=torch.float32) if return_bias else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ertdiff_encode_condition_prec(
                h, _lib.ptr(condition), n, L, IN_CHANNELS * L, _lib.ptr(emb),
                _lib.ptr(bias) if bias is not None else None, _lib.PRECISIONS[precision],
                _lib.stream_ptr(dev)), "encode_condition")
        return (emb, bias) if return_bias else emb
