"""Load the reference's own functions without importing (or copying) its script.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``import ERT_Conditional_Diffusion`` cannot work: the file imports plotting
libraries that are not installed, loads data files that are not shipped, trains
for 500 epochs and shells out to PFLOTRAN at module level.  What
does work is to parse the file and execute only the top-level ``def`` / ``class``
statements of the hot path, in a namespace that supplies the handful of modules
they use.  Nothing of the reference is copied into this repository; the source is
read where it lies, at run time, and only in the container that has
``/root/reference``.

Noise replay: ``sample_model`` (ERT_Conditional_Diffusion.py:102-119) draws its
noise with ``torch.randn`` / ``torch.randn_like``.  ``load_reference(noise=...)``
puts a proxy under the name ``torch`` in that namespace which forwards every
attribute to the real module except those two, which hand out successive rows of
a pre-generated ``(num_steps, B, P)`` tensor (row 0 is ``x_T``).
"""
from __future__ import annotations

import ast
import math
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import Dataset

REFERENCE_ROOT = os.environ.get("ERTDIFF_REFERENCE_ROOT", "/root/reference")
REFERENCE_FILE = os.path.join(REFERENCE_ROOT, "ERT_Conditional_Diffusion.py")

# top-level definitions on (or next to) the hot path, with their line ranges for the reader
WANTED = (
    "transform_to_unconstrained",   # :26-40
    "inverse_transform",            # :42-53
    "get_timestep_embedding",       # :80-88
    "get_diffusion_schedule",       # :90-94
    "q_sample",                     # :96-99
    "sample_model",                 # :102-119
    "ConditionalDiffusionModel",    # :122-164
    "mode_kde_calculation",         # :166-181
    "check_param_bounds",           # :183-218
    "load_best_model",              # :369-377
    "avg_prop_indicator_function",  # :1089-1096
    "accuracy_score",               # :1098-1100
    "preccision_score",             # :1102-1109
    "goodness_score",               # :1111-1115
    "WSSE_metric",                  # :764-770
)


def reference_available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


class _ReplayTorch:
    """Forwards to ``torch`` except ``randn`` / ``randn_like`` which replay ``noise`` rows."""

    def __init__(self, noise: torch.Tensor):
        self._noise = noise
        self._next = 0

    def _draw(self, shape, device):
        row = self._noise[self._next]
        self._next += 1
        assert tuple(row.shape) == tuple(shape), (row.shape, shape)
        return row.clone().to(device)

    def randn(self, *shape, device=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return self._draw(shape, device or "cpu")

    def randn_like(self, x, **kw):
        return self._draw(x.shape, x.device)

    @property
    def draws(self):
        return self._next

    def __getattr__(self, name):
        return getattr(torch, name)


def load_reference(noise: torch.Tensor | None = None) -> SimpleNamespace:
    """Return a namespace holding the reference's own callables (see ``WANTED``).

    With ``noise`` given, ``sample_model`` replays it instead of drawing from the RNG;
    ``ns.torch_proxy.draws`` then tells how many rows were consumed.
    """
    if not reference_available():
        raise FileNotFoundError(
            f"{REFERENCE_FILE} not found: the reference exists only in the build "
            "container; on other machines use tests/golden instead")
    with open(REFERENCE_FILE, "r") as fh:
        tree = ast.parse(fh.read(), filename=REFERENCE_FILE)
    picked = [n for n in tree.body
              if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in WANTED]
    missing = set(WANTED) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"reference definitions not found: {sorted(missing)}")
    import scipy.stats as stats
    torch_like = _ReplayTorch(noise) if noise is not None else torch
    ns = {"torch": torch_like, "nn": nn, "math": math, "np": np, "stats": stats,
          "Dataset": Dataset, "__name__": "reference_hot_path"}
    module = ast.Module(body=picked, type_ignores=[])
    exec(compile(module, REFERENCE_FILE, "exec"), ns)
    out = SimpleNamespace(**{k: ns[k] for k in WANTED})
    out.torch_proxy = torch_like
    return out
