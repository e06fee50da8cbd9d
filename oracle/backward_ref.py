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


# --------------------------------------------------------------------------- whole-UNet backward from the pieces above
def unet_param_grads(sd, spec, x, t, H, W, D, dout_fn):
    """Parameter gradients of the triplane UNet computed WITHOUT autograd: a forward pass that keeps what each adjoint needs, then
    the adjoints above in reverse order (the order a training step's backward kernels would run in).
    ``dout_fn(out_planes) -> (d_xy, d_xz, d_yz)`` supplies dL/d(output planes).  -> (output planes, {param name: grad})."""
    from oracle import unet_ref as ur
    grads: Dict[str, torch.Tensor] = {}
    tape = []                                                          # closures, run in reverse

    def acc(name, g):
        grads[name] = grads[name] + g if name in grads else g

    def linear(xv, wn, bn):
        w, b = sd[wn], sd[bn]
        y = F.linear(xv, w, b)
        return y, (lambda dy: (acc(wn, dy.t() @ xv), acc(bn, dy.sum(0)), dy @ w)[-1])

    def silu_vec(xv):
        sig = torch.sigmoid(xv)
        return xv * sig, (lambda dy: dy * (sig * (1 + xv * (1 - sig))))

    def conv1x1(prefix, planes):
        outs = []
        for a, n in zip(planes, PLANES):
            outs.append(F.conv2d(a, sd[f"{prefix}.conv_{n}.weight"], sd[f"{prefix}.conv_{n}.bias"]))

        def back(dys):
            dxs = []
            for a, n, dy in zip(planes, PLANES, dys):
                w = sd[f"{prefix}.conv_{n}.weight"]
                acc(f"{prefix}.conv_{n}.weight", torch.nn.grad.conv2d_weight(a, w.shape, dy))
                acc(f"{prefix}.conv_{n}.bias", dy.sum(dim=(0, 2, 3)))
                dxs.append(F.conv_transpose2d(dy, w))
            return dxs
        return tuple(outs), back

    def conv3x3(prefix, planes):
        outs = ur.tri_conv(sd, prefix, planes, 1, spec.rollout)

        def back(dys):
            if spec.rollout:
                dxs, g = tri_conv_backward_folded(sd, prefix, planes, dys)
            else:
                dxs, g = [], {}
                for a, n, dy in zip(planes, PLANES, dys):
                    w = sd[f"{prefix}.conv_{n}.weight"]
                    g[f"{prefix}.conv_{n}.weight"] = torch.nn.grad.conv2d_weight(a, w.shape, dy, padding=1)
                    g[f"{prefix}.conv_{n}.bias"] = dy.sum(dim=(0, 2, 3))
                    dxs.append(F.conv_transpose2d(dy, w, padding=1))
            for k, v in g.items():
                acc(k, v)
            return dxs
        return outs, back

    def norm_silu(prefix, planes, scale=None, shift=None):
        """silu(GN(x) [* (1 + scale) + shift]) per plane.  back -> (dplanes, dscale, dshift)"""
        outs = []
        for a, n in zip(planes, PLANES):
            h = F.group_norm(a, ur.GN_GROUPS, sd[f"{prefix}.norm_{n}.weight"], sd[f"{prefix}.norm_{n}.bias"], ur.GN_EPS)
            if scale is not None:
                h = h * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
            outs.append(ur.silu(h))

        def back(dys):
            dxs, dsc, dsh = [], None, None
            for a, n, dy in zip(planes, PLANES, dys):
                o = gn_film_silu_backward(a, sd[f"{prefix}.norm_{n}.weight"], sd[f"{prefix}.norm_{n}.bias"], dy, ur.GN_GROUPS,
                                          ur.GN_EPS, scale, shift)
                acc(f"{prefix}.norm_{n}.weight", o["dgamma"])
                acc(f"{prefix}.norm_{n}.bias", o["dbeta"])
                dxs.append(o["dx"])
                if scale is not None:
                    dsc = o["dscale"] if dsc is None else dsc + o["dscale"]
                    dsh = o["dshift"] if dsh is None else dsh + o["dshift"]
            return dxs, dsc, dsh
        return tuple(outs), back

    # ---- forward with a tape of adjoint closures; every closure maps output gradient(s) to input gradient(s)
    emb0 = ur.sinusoid(t, spec.model_channels)
    e1, b_l0 = linear(emb0, "time_embed.0.weight", "time_embed.0.bias")
    e2, b_s0 = silu_vec(e1)
    emb, b_l1 = linear(e2, "time_embed.2.weight", "time_embed.2.bias")
    demb = [torch.zeros_like(emb)]

    def res_block(prefix, planes):
        h1, b_n1 = norm_silu(f"{prefix}.in_layers.0", planes)
        h2, b_c1 = conv3x3(f"{prefix}.in_layers.2", h1)
        se, b_se = silu_vec(emb)
        e, b_le = linear(se, f"{prefix}.emb_layers.1.weight", f"{prefix}.emb_layers.1.bias")
        if spec.use_scale_shift_norm:
            scale, shift = e.chunk(2, dim=1)
            h3, b_n2 = norm_silu(f"{prefix}.out_layers.0", h2, scale, shift)
        else:
            h2e = tuple(a + e[:, :, None, None] for a in h2)
            h3, b_n2 = norm_silu(f"{prefix}.out_layers.0", h2e)
        h4, b_c2 = conv3x3(f"{prefix}.out_layers.2", h3)
        has_skip = f"{prefix}.skip_connection.conv_xy.weight" in sd
        if has_skip:
            s, b_sk = conv1x1(f"{prefix}.skip_connection", planes)
        else:
            s, b_sk = planes, None
        out = tuple(a + b for a, b in zip(h4, s))

        def back(douts):
            dh3 = b_c2(douts)
            dh2, dsc, dsh = b_n2(dh3)
            if spec.use_scale_shift_norm:
                de = torch.cat([dsc, dsh], dim=1)
            else:
                de = sum(d.sum(dim=(2, 3)) for d in dh2)
            demb[0] = demb[0] + b_se(b_le(de))
            dh1 = b_c1(dh2)
            dx, _, _ = b_n1(dh1)
            dskip = b_sk(douts) if has_skip else douts
            return [a + b for a, b in zip(dx, dskip)]
        return out, back

    p = tuple(a.contiguous() for a in ur.split_planes(x, H, W, D))
    p, b_in = conv1x1("in_conv.0", p)
    downs, ups = ur.block_plan(spec)
    stack = []                                    # (planes, index into skip_grads)
    skip_grads = []
    seq = []                                      # top-level tape: closures over plane-gradient lists
    for ops in downs:
        for op in ops:
            if op[0] == "down":
                shp = [a.shape[-2:] for a in p]
                p = ur.avgpool2(p)
                seq.append(lambda dys, shp=shp: [avgpool2_backward(dy, r, c) for dy, (r, c) in zip(dys, shp)])
            else:
                p, bk = res_block(op[1], p)
                seq.append(bk)
        skip_grads.append(None)
        stack.append((p, len(skip_grads) - 1, len(seq)))          # gradient of this skip joins the chain after closure #len(seq)-1
    join_at = {}                                   # position in seq (exclusive) -> skip index whose gradient is added there
    for j, ops in enumerate(ups):
        if j == 0:
            p, si, pos = stack.pop()
            # the deepest skip IS the chain: nothing to add
        else:
            s, si, pos = stack.pop()
            join_at[pos] = si
            if spec.rollout:
                shp = [a.shape[-2:] for a in p]
                tgt = [b.shape[-2:] for b in s]
                p = tuple(a if a.shape[2:] == b.shape[2:] else F.interpolate(a, size=b.shape[2:], mode="bilinear", align_corners=False)
                          for a, b in zip(p, s))
                seq.append(lambda dys, shp=shp, tgt=tgt: [dy if tuple(sh) == tuple(tg) else bilinear_resize_backward(dy, sh[0], sh[1])
                                                         for dy, sh, tg in zip(dys, shp, tgt)])
            cu = p[0].shape[1]
            p = tuple(torch.cat([a, b], dim=1) for a, b in zip(p, s))

            def split_cat(dys, cu=cu, si=si):
                skip_grads[si] = [dy[:, cu:] for dy in dys]
                return [dy[:, :cu] for dy in dys]
            seq.append(split_cat)
        for op in ops:
            if op[0] == "up":
                shp = [a.shape[-2:] for a in p]
                p = ur.up2(p)
                seq.append(lambda dys, shp=shp: [bilinear_resize_backward(dy, r, c, scale_factor=2) for dy, (r, c) in zip(dys, shp)])
            else:
                p, bk = res_block(op[1], p)
                seq.append(bk)
    hN, b_on = norm_silu("out.0", p)
    outp, b_oc = conv1x1("out.2", hN)

    # ---- backward
    d = list(dout_fn(outp))
    d = b_oc(d)
    d, _, _ = b_on(d)
    for pos in range(len(seq), 0, -1):
        if pos in join_at:                         # the skip saved after closure #pos-1 also fed a decoder concat
            d = [a + b for a, b in zip(d, skip_grads[join_at[pos]])]
        d = seq[pos - 1](d)
    b_in(d)
    b_l0(b_s0(b_l1(demb[0])))
    return outp, grads
