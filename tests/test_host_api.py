"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, the host mirror has the reference's surface, and compute fails loudly without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import ertdiff_b200 as eb
from ertdiff_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ertdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ertdiff_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(eb.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ertdiff_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes prototypes and header disagree"
    assert _lib.load().ertdiff_abi_version() == 1


def test_chain_args_layout_matches_header():
    # field order and natural alignment of ertdiff_chain_args
    names = [f[0] for f in _lib.ChainArgs._fields_]
    src = open(os.path.join(ROOT, "include", "ertdiff_b200.h")).read()
    body = src[src.index("typedef struct ertdiff_chain_args {"):src.index("} ertdiff_chain_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"\b(\w+);", body)
    assert declared == names
    assert ctypes.sizeof(_lib.ChainArgs) == 136


def test_enum_values_match_header():
    # the host mirror's names for the loop modes and precisions carry the header's enum values
    src = open(os.path.join(ROOT, "include", "ertdiff_b200.h")).read()
    enum = {k: int(v) for k, v in re.findall(r"\b(ERTDIFF_(?:PREC|LOOP)_[A-Z0-9]+)\s*=\s*(\d+)", src)}
    assert {k: enum["ERTDIFF_PREC_" + k.upper()] for k in _lib.PRECISIONS} == _lib.PRECISIONS
    assert sorted(_lib.PRECISIONS) == ["bf16", "bf16x3", "fp32"]
    assert {k: enum["ERTDIFF_LOOP_" + k.upper()] for k in _lib.LOOP_MODES} == _lib.LOOP_MODES


def test_model_surface_and_seeded_init_equal_reference(golden):
    torch.manual_seed(0)
    m = eb.ConditionalDiffusionModel(29, 128)
    g = golden("model_seed0.npz")
    sd = m.state_dict()
    assert list(sd.keys()) == list(g.files)
    for k in g.files:       # same RNG draws as the reference's constructor under the same seed
        assert np.array_equal(sd[k].numpy(), g[k]), k
    assert m.param_dim == 29 and m.eval() is m
    assert sum(p.numel() for p in m.parameters()) == 72765
    m2 = eb.ConditionalDiffusionModel(29, 128)
    m2.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_checkpoint_roundtrip(tmp_path):
    m = eb.ConditionalDiffusionModel(29, 64)
    path = tmp_path / "best_model.pt"
    eb.save_checkpoint(path, m, epoch=7, best_val_loss=0.25)
    m2 = eb.ConditionalDiffusionModel(29, 64)
    ck = eb.load_best_model(path, m2, map_location="cpu")        # weights_only=True by default
    assert set(ck) == {"epoch", "model_state_dict", "best_val_loss", "train_history", "val_history", "param_dim"}
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    torch.save(m.state_dict(), tmp_path / "bare.pt")
    eb.load_best_model(tmp_path / "bare.pt", m2, map_location="cpu")
    # with an optimizer the reference's 7-key dict is written and its loader pattern (ECD.py:369-377) works
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    eb.save_checkpoint(path, m, epoch=3, best_val_loss=0.5, optimizer=opt, train_history=[1.0], val_history=[2.0])
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-4)
    ck = eb.load_best_model(path, m2, optimizer=opt2, map_location="cpu")
    assert set(ck) == set(eb.checkpoint.CHECKPOINT_KEYS) and ck["epoch"] == 3
    # a checkpoint written without optimizer state loads into a caller that passes an optimizer
    eb.save_checkpoint(path, m)
    eb.load_best_model(path, m2, optimizer=opt2, map_location="cpu")


def test_no_cpu_fallback():
    m = eb.ConditionalDiffusionModel(29, 128)
    with pytest.raises(eb.ErtdiffError, match="CUDA only"):
        m(torch.zeros(2, 29), torch.zeros(2, dtype=torch.long), torch.zeros(2, 14, 33))
    b, a, ab = eb.get_diffusion_schedule(10)
    with pytest.raises(eb.ErtdiffError, match="CUDA only"):
        eb.sample_model(m, torch.zeros(2, 14, 33), 10, b, a, ab, 29, "cpu")
    # CPU tensors without a device= are refused outright; numpy inputs need a GPU to land on
    with pytest.raises(eb.ErtdiffError, match="CUDA only"):
        eb.misfit_metrics(torch.zeros(2, 5, 3), torch.zeros(5, 3))
    with pytest.raises(eb.ErtdiffError, match="CUDA only"):
        eb.wasserstein_distance(torch.zeros(7), torch.zeros(5))
    if not torch.cuda.is_available():
        for call in (lambda: eb.ensemble_mean(np.zeros((4, 3))),
                     lambda: eb.misfit_metrics(np.zeros((2, 5, 3)), np.zeros((5, 3))),
                     lambda: eb.wasserstein_distance(np.zeros(7), np.ones(5))):
            with pytest.raises(Exception):
                call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ert-conditional-diffusion-model_b200")
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.+oracle)|oracle/|oracle\.", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} refers to the oracle package"
    import sys
    assert not any(m == "oracle" or m.startswith("oracle.") for m in sys.modules
                   if "ertdiff" in m)


def test_schedule_and_embedding_host_helpers(golden):
    g = golden("embedding_schedule.npz")
    b, a, ab = eb.get_diffusion_schedule(500)
    assert np.array_equal(b.numpy(), g["betas500"]) and np.array_equal(ab.numpy(), g["alpha_bar500"])
    emb = eb.get_timestep_embedding(torch.from_numpy(g["t"]), 128)
    assert np.array_equal(emb.numpy(), g["emb"])


def test_member_slices_cover_and_preserve_order():
    from ertdiff_b200.parallel import member_slice
    for n, w, mult in [(8192, 8, 1), (1000, 3, 1), (50 * 32, 4, 32), (7, 8, 1), (96, 2, 32)]:
        spans = [member_slice(n, r, w, mult) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert all((b - a) % mult == 0 for a, b in spans)
