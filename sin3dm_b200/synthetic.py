"""Deterministic synthetic checkpoints for benchmarks and smoke runs (no network, no pretrained weights: BASELINE.json asks for
"random-init weights of that architecture").  A freshly constructed reference UNet outputs exactly zero (zero_module on every
block's second conv and on the output conv, unet_triplane.py:243-246, 444), so every tensor is redrawn:
conv / linear weights ~ U(-a, a) with a = sqrt(3 / fan_in), biases ~ N(0, 0.05), norm gains ~ 1 + N(0, 0.1), norm biases ~ N(0, 0.1),
from one CPU generator in state_dict order (the same recipe the test oracle uses, asserted equal in tests/test_host_logic.py)."""
import math

import torch


def synthetic_state_dict_like(module, seed=1234):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        if ".norm_" in key:
            v = torch.randn(shape, generator=g) * 0.1
            if key.endswith("weight"):
                v = v + 1.0
        elif key.endswith("bias"):
            v = torch.randn(shape, generator=g) * 0.05
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        sd[key] = v.float().contiguous()
    return sd
