"""Emulations of the bf16 tensor-core kernels' rounding points, built from torch ops (test helper)."""
import torch
import torch.nn.functional as F


def bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


def emulate_encoder(sd, cond):
    """condition_encoder (ECD.py:133-142) with the tensor-core encoder's rounding points: input, both
    conv weights and the conv1 activations rounded to bf16; accumulation, biases, ReLU, pooling and
    the Linear layer in fp32/fp64."""
    w1, b1 = sd["condition_encoder.0.weight"], sd["condition_encoder.0.bias"]
    w2, b2 = sd["condition_encoder.2.weight"], sd["condition_encoder.2.bias"]
    h1 = F.relu(F.conv1d(bf(cond), bf(w1), None, stride=2, padding=1).float() + b1[None, :, None])
    h2 = F.relu(F.conv1d(bf(h1), bf(w2), None, stride=2, padding=1).float() + b2[None, :, None])
    pooled = h2.double().mean(dim=2).float()
    return F.relu(F.linear(pooled, sd["condition_encoder.6.weight"], sd["condition_encoder.6.bias"]))


