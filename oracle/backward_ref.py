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
