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
