"""CPU restatement of the BACKWARD of the rollout TriplaneConv in its folded form (test infrastructure; groundwork for the
training row).

Forward (reference src/diffusion/unet_triplane.py:31-60, oracle.unet_ref.tri_conv): every plane's 3x3 conv sees its own C
channels plus 2C channels that are axis means of the other two planes broadcast along one image axis.  The forward kernels never
materialise those 2C channels: their conv collapses to a 1-D conv of the mean vector (border classes first / interior / last).
The adjoint has the same shape, and this file states it with plain torch ops so that the kernels of the next round have an oracle
that already has their structure:

  * gradient wrt the plane's own channels: the ordinary 3x3 transposed conv of dY;
  * gradient wrt a broadcast mean vector: a 1-D transposed conv of THREE axis sums of dY — the full sum and the sums without the
    first / last line across (the zero padding makes the outer taps miss one border line) — one per tap across;
  * weight gradient of the broadcast channels: the 1-D correlation of the mean vector with the same three axis sums;
  * the mean's own adjoint: divide by the averaged length and broadcast back into the source plane.

``tri_conv_backward_folded`` is checked against torch.autograd through ``oracle.unet_ref.tri_conv`` (tests/test_oracle_backward.py).
The file also states, in the form a kernel would use, the backward of GroupNorm (+ FiLM) + SiLU (two reductions per group, then one
apply pass) and the adjoints of the 2x2 average pool and of the bilinear resamplings (gather form), each checked against autograd.
"""
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

PLANES = ("xy", "xz", "yz")
# plane -> ((source plane, axis of the source that is averaged, "row" | "col": which image axis of THIS plane indexes the vector), ...)
# in the order the reference concatenates them (unet_triplane.py:37-46)
SOURCES = {
    0: ((2, -1, "col"), (1, -1, "row")),      # xy [H,W]: mean_D(yz)[W] along columns, mean_D(xz)[H] along rows
    1: ((0, -1, "row"), (2, -2, "col")),      # xz [H,D]: mean_W(xy)[H] along rows,    mean_W(yz)[D] along columns
    2: ((0, -2, "row"), (1, -2, "col")),      # yz [W,D]: mean_H(xy)[W] along rows,    mean_H(xz)[D] along columns
}


def _axis_sums(dy: torch.Tensor, kind: str):
    """The three sums of dY across the broadcast axis, indexed by the tap across (0, 1, 2): without the first line, all, without
    the last line.  kind == "row": the vector is indexed by row, sums run over columns."""
    if kind == "row":
        s, first, last = dy.sum(-1), dy[..., 0], dy[..., -1]
    else:
        s, first, last = dy.sum(-2), dy[..., 0, :], dy[..., -1, :]
    return s - first, s, s - last


def tri_conv_backward_folded(sd: Dict[str, torch.Tensor], prefix: str, planes: Sequence[torch.Tensor],
                             dys: Sequence[torch.Tensor]) -> Tuple[list, Dict[str, torch.Tensor]]:
    """Gradients of the rollout TriplaneConv (3x3, padding 1).  planes / dys: (xy, xz, yz) inputs [B,C,.,.] and output gradients
    [B,Cout,.,.].  -> ([d_xy, d_xz, d_yz], {"<prefix>.conv_<plane>.weight" / ".bias": grad})."""
    C = planes[0].shape[1]
    dplanes = [torch.zeros_like(p) for p in planes]
    grads = {}
    for pi, name in enumerate(PLANES):
        w = sd[f"{prefix}.conv_{name}.weight"]                       # [Cout, 3C, 3, 3]
        dy, a = dys[pi], planes[pi]
        gw = torch.zeros_like(w)
        grads[f"{prefix}.conv_{name}.bias"] = dy.sum(dim=(0, 2, 3))
        # own channels: ordinary dgrad / wgrad
        w0 = w[:, :C]
        dplanes[pi] += F.conv_transpose2d(dy, w0, padding=1)
        gw[:, :C] = torch.nn.grad.conv2d_weight(a, w0.shape, dy, padding=1)
        # broadcast channels: 1-D forms
        for si, (src, axis, kind) in enumerate(SOURCES[pi]):
            ws = w[:, (si + 1) * C:(si + 2) * C]                     # [Cout, C, kh, kw]
            n_avg = planes[src].shape[axis]
            vec = planes[src].mean(dim=axis)                         # [B, C, L]
            sums = _axis_sums(dy, kind)                              # tap across -> [B, Cout, L]
            dvec = torch.zeros_like(vec)
            for across in range(3):
                w1d = ws[:, :, :, across] if kind == "row" else ws[:, :, across, :]      # [Cout, C, along]
                dvec += F.conv_transpose1d(sums[across], w1d, padding=1)
                gw1d = torch.nn.grad.conv1d_weight(vec, w1d.shape, sums[across], padding=1)
                if kind == "row":
                    gw[:, (si + 1) * C:(si + 2) * C, :, across] = gw1d
                else:
                    gw[:, (si + 1) * C:(si + 2) * C, across, :] = gw1d
            dplanes[src] += (dvec / n_avg).unsqueeze(axis).expand_as(planes[src])
        grads[f"{prefix}.conv_{name}.weight"] = gw
    return dplanes, grads


# --------------------------------------------------------------------------- GroupNorm (+ FiLM) + SiLU
def gn_film_silu_backward(x, gamma, beta, dy, groups=32, eps=1e-5, scale=None, shift=None):
    """Backward of y = silu(GN(x) * (1 + scale) + shift) (unet_triplane.py:63-95, 285-297; scale / shift [B, C] or None) in the
    two-pass form a kernel would use: pass 1 forms, per (sample, channel), S1 = sum dxhat and S2 = sum dxhat * xhat over the plane
    (plus the parameter / FiLM sums), pass 2 applies dx = rstd * (dxhat - mean_g(S1) - xhat * mean_g(S2)).
    -> dict(dx, dgamma, dbeta, dscale, dshift)."""
    B, C, R, Cc = x.shape
    cpg = C // groups
    xg = x.reshape(B, groups, cpg * R * Cc)
    mu = xg.mean(-1, keepdim=True)
    var = xg.var(-1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    xhat = ((xg - mu) * rstd).reshape(B, C, R, Cc)
    n = xhat * gamma.view(1, C, 1, 1) + beta.view(1, C, 1, 1)
    if scale is not None:
        f = n * (1 + scale.view(B, C, 1, 1)) + shift.view(B, C, 1, 1)
    else:
        f = n
    sig = torch.sigmoid(f)
    df = dy * (sig * (1 + f * (1 - sig)))                            # SiLU'
    out = {}
    if scale is not None:
        out["dscale"] = (df * n).sum(dim=(2, 3))                     # [B, C]
        out["dshift"] = df.sum(dim=(2, 3))
        dn = df * (1 + scale.view(B, C, 1, 1))
    else:
        out["dscale"] = out["dshift"] = None
        dn = df
    out["dgamma"] = (dn * xhat).sum(dim=(0, 2, 3))
    out["dbeta"] = dn.sum(dim=(0, 2, 3))
    dxhat = dn * gamma.view(1, C, 1, 1)
    # pass 1: per-channel sums, then per-group means
    s1 = dxhat.sum(dim=(2, 3)).reshape(B, groups, cpg).sum(-1) / (cpg * R * Cc)            # [B, G]
    s2 = (dxhat * xhat).sum(dim=(2, 3)).reshape(B, groups, cpg).sum(-1) / (cpg * R * Cc)
    # pass 2
    expand = lambda v: v.repeat_interleave(cpg, dim=1).view(B, C, 1, 1)
    out["dx"] = expand(rstd.view(B, groups)) * (dxhat - expand(s1) - xhat * expand(s2))
    return out


# --------------------------------------------------------------------------- resampling adjoints (gather form)
def avgpool2_backward(dy, in_rows, in_cols):
    """Adjoint of F.avg_pool2d(x, 2) (unet_triplane.py:127-145; floor on odd sizes: the last odd line gets no gradient)."""
    dx = dy.new_zeros(*dy.shape[:2], in_rows, in_cols)
    R, Cc = dy.shape[-2:]
    dx[..., :2 * R, :2 * Cc] = dy.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1) * 0.25
    return dx


def _bilinear_weights(out_size, in_size, scale):
    """ATen area_pixel_compute_source_index(align_corners=False) as a dense [out, in] interpolation matrix."""
    m = torch.zeros(out_size, in_size, dtype=torch.float64)
    for o in range(out_size):
        s = max(scale * (o + 0.5) - 0.5, 0.0)
        i0 = min(int(s), in_size - 1)
        i1 = min(i0 + 1, in_size - 1)
        l1 = s - i0
        m[o, i0] += 1 - l1
        m[o, i1] += l1
    return m


def bilinear_resize_backward(dy, in_rows, in_cols, scale_factor=None):
    """Adjoint of F.interpolate(x, mode="bilinear", align_corners=False) to dy's size — either scale_factor=2
    (unet_triplane.py:106-124) or an explicit size (the resize-to-skip step, :494-499).  Gather form: every input pixel sums its
    contributors, i.e. dx = Wr^T dy Wc with the two 1-D interpolation matrices."""
    R, Cc = dy.shape[-2:]
    sr = (1.0 / scale_factor) if scale_factor else in_rows / R
    sc = (1.0 / scale_factor) if scale_factor else in_cols / Cc
    wr = _bilinear_weights(R, in_rows, sr).to(dy.dtype)
    wc = _bilinear_weights(Cc, in_cols, sc).to(dy.dtype)
    return torch.einsum("oi,bcop,pj->bcij", wr, dy, wc)
