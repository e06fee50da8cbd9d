"""Helpers shared by the -m gpu parity tests: build the CUDA model from oracle weights, error metrics."""
import numpy as np
import torch

import sin3dm_b200 as s3
from oracle import unet_ref as ur


def make_cuda_model(spec: ur.UNetSpec, sd, precision=3, conv_impl="tc"):
    cls = s3.TriplaneUNetModelSmall if spec.rollout else s3.TriplaneUNetModelSmallRaw
    m = cls(spec.in_channels, spec.model_channels, spec.out_channels, spec.num_res_blocks, 0, tuple(spec.channel_mult),
            use_scale_shift_norm=spec.use_scale_shift_norm)
    m.load_state_dict(sd)
    m.s3d_precision = precision
    m.s3d_conv_impl = 1 if conv_impl == "ffma" else 0
    return m.cuda().eval()


def plane_errors(got, want, H, W, D):
    """(rel_l2, max_abs / max|want|) over the three decomposed planes (the dead corner is excluded)."""
    g = [p.double() for p in ur.split_planes(torch.as_tensor(got).cpu(), H, W, D)]
    w = [p.double() for p in ur.split_planes(torch.as_tensor(want).cpu(), H, W, D)]
    num = sum(((a - b) ** 2).sum() for a, b in zip(g, w)).sqrt()
    den = sum((b ** 2).sum() for b in w).sqrt()
    mx = max((a - b).abs().max() for a, b in zip(g, w))
    ref = max(b.abs().max() for b in w)
    return float(num / den), float(mx / ref)


def nhwc(p):
    """oracle plane [B,C,R,Cc] -> NHWC like the kernels' activations."""
    return p.permute(0, 2, 3, 1).contiguous()
