"""Small host-side helpers that callers of the reference import from ``diffusion.nn``
(reference src/diffusion/nn.py).  Only what the sampling / training drivers touch is mirrored."""
import math

import torch as th


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding (nn.py:103-121).  The CUDA path computes this in ``k_sinusoid``; this torch
    version exists for callers that use it directly and to derive the frequency table bit-exactly."""
    half = dim // 2
    freqs = sinusoid_freqs(dim, max_period).to(device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = th.cat([th.cos(args), th.sin(args)], dim=-1)
    if dim % 2:
        emb = th.cat([emb, th.zeros_like(emb[:, :1])], dim=-1)
    return emb


def sinusoid_freqs(dim, max_period=10000):
    half = dim // 2
    return th.exp(-math.log(max_period) * th.arange(start=0, end=half, dtype=th.float32) / half)


def mean_flat(tensor):
    return tensor.mean(dim=list(range(1, len(tensor.shape))))


def update_ema(target_params, source_params, rate=0.99):
    for targ, src in zip(target_params, source_params):
        targ.detach().mul_(rate).add_(src, alpha=1 - rate)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module
