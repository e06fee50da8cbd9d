"""Composed-triplane layout helpers (reference src/utils/triplane_util.py:7-25).

The composed tensor is [..., H+D, W+D]: xy in the top-left, xz top-right, yz (transposed) bottom-left,
and an all-zero D x D corner.  The CUDA kernels read / write this layout directly
(``k_in_conv`` / ``k_out_head``); these torch views are for callers and tests.
"""
import torch


def decompose_featmaps(composed_map, sizes):
    H, W, D = sizes
    return (composed_map[..., :H, :W], composed_map[..., :H, W:], composed_map[..., H:, :W].transpose(-1, -2))


def compose_featmaps(feat_xy, feat_xz, feat_yz):
    H, W = feat_xy.shape[-2:]
    D = feat_xz.shape[-1]
    out = feat_xy.new_zeros(*feat_xy.shape[:-2], H + D, W + D)
    out[..., :H, :W] = feat_xy
    out[..., :H, W:] = feat_xz
    out[..., H:, :W] = feat_yz.transpose(-1, -2)
    return out, (H, W, D)


def pad_composed_featmaps(composed_map, sizes, pad_sizes):
    """reference :28-35.  pad_sizes = [[padH1, padH2], [padW1, padW2], [padD1, padD2]] (zero padding of every plane)."""
    import torch.nn.functional as F
    feat_xy, feat_xz, feat_yz = decompose_featmaps(composed_map, sizes)
    feat_xy = F.pad(feat_xy, list(pad_sizes[1]) + list(pad_sizes[0]))
    feat_xz = F.pad(feat_xz, list(pad_sizes[2]) + list(pad_sizes[0]))
    feat_yz = F.pad(feat_yz, list(pad_sizes[2]) + list(pad_sizes[1]))
    return compose_featmaps(feat_xy, feat_xz, feat_yz)


def save_triplane_data(path, feat_xy, feat_xz, feat_yz):
    """reference :38-41: the latent of one sample as a compressed ``feat.npz`` (host zlib; arrays [C, ., .])."""
    import os
    import numpy as np
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    to_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    np.savez_compressed(path, feat_xy=to_np(feat_xy), feat_xz=to_np(feat_xz), feat_yz=to_np(feat_yz))


def load_triplane_data(path, device="cuda:0", compose=True):
    """reference :44-61: -> (composed [C, H+D, W+D], (H, W, D)), or the three planes when ``compose`` is False."""
    import numpy as np
    data = np.load(path)
    planes = [torch.from_numpy(data[k][:]).float().to(device) for k in ("feat_xy", "feat_xz", "feat_yz")]
    if not compose:
        return tuple(planes)
    return compose_featmaps(*planes)
