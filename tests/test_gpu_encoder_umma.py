"""The tensor-core condition encoder (`precision="bf16"`): conv1 / conv2 as tcgen05 implicit GEMMs.

Two references, as for the bf16 chain:
  * an emulation of the kernel's rounding points (input, both weight tensors and the conv1
    activations rounded to bf16; accumulation, biases, ReLU, pooling in fp32/fp64) -- tight
    tolerance, validates the data flow (phase blocks, shifted row windows, carried rows, masks);
  * the fp32 oracle (ECD.py:133-142) -- the precision cost of bf16 operands, stated below.
"""
import numpy as np
import pytest
import torch

import ertdiff_b200 as eb
from oracle import denoiser_oracle as do

pytestmark = pytest.mark.gpu
C = 14


from bf16_emulation import emulate_encoder as emulate


# L2 = 1174 at the reference grid: two chunks, the second ends inside a tile; the small grids
# exercise a single partial tile, an exact tile boundary and odd lengths (ragged L1 / L2)
@pytest.mark.parametrize("n,L", [(3, 4693), (1, 37), (2, 512), (5, 1029), (2, 2051), (1, 6000)])
def test_tensor_core_encoder_matches_its_emulation(gpu_model, ref_state_dict, cuda_dev, n, L):
    g = torch.Generator().manual_seed(100 * n + L)
    cond = torch.rand(n, C, L, generator=g)
    emb = gpu_model.encode_condition(cond.to(cuda_dev), precision="bf16").cpu()
    assert gpu_model.umma_status() == 0
    ref = emulate(ref_state_dict, cond)
    scale = ref.abs().max().item()
    assert (emb - ref).abs().max().item() <= 2e-5 * scale + 1e-6, (emb - ref).abs().max().item() / scale


def test_tensor_core_encoder_vs_fp32_oracle_tolerance(gpu_model, ref_state_dict, cuda_dev):
    # bf16 operands: 8 mantissa bits per product, averaged over 1174 pooled positions
    cond = torch.rand(4, C, 4693, generator=torch.Generator().manual_seed(5))
    emb16, bias16 = gpu_model.encode_condition(cond.to(cuda_dev), precision="bf16", return_bias=True)
    emb32, bias32 = gpu_model.encode_condition(cond.to(cuda_dev), return_bias=True)
    ref = do.encode_condition(ref_state_dict, cond)
    scale = ref.abs().max().item()
    assert (emb32.cpu() - ref).abs().max().item() <= 1e-5 * scale + 1e-6
    assert (emb16.cpu() - ref).abs().max().item() <= 2e-3 * scale
    assert (bias16 - bias32).abs().max().item() <= 2e-3 * bias32.abs().max().item()


def test_tensor_core_encoder_strided_and_unaligned_conditions(gpu_model, cuda_dev):
    # conditions that start at addresses which are only 4-byte aligned (odd L), and a slice
    base = torch.rand(7, C, 1029, generator=torch.Generator().manual_seed(9)).to(cuda_dev)
    full = gpu_model.encode_condition(base, precision="bf16")
    part = gpu_model.encode_condition(base[3:6], precision="bf16")
    assert torch.equal(part, full[3:6])
