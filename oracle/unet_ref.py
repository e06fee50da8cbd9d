"""CPU oracle for the triplane UNet (test infrastructure — see oracle/__init__.py).

Functional fp32 restatement, over a flat ``state_dict``, of

  * ``TriplaneUNetModelSmall.forward``      reference src/diffusion/unet_triplane.py:465-510
  * ``TriplaneUNetModelSmallRaw.forward``   reference src/diffusion/unet_triplane.py:665-702
  * ``TriplaneResBlock._forward``           reference src/diffusion/unet_triplane.py:269-311
  * ``TriplaneConv.forward`` (rollout)      reference src/diffusion/unet_triplane.py:31-60
  * ``GroupNorm32`` / ``timestep_embedding`` reference src/diffusion/nn.py:17-19, 103-121
  * ``compose/decompose_featmaps``          reference src/utils/triplane_util.py:7-25

It is written against plain ``torch.nn.functional`` on CPU tensors and takes the reference's own
checkpoint keys, so a reference ``.pt`` drops in.  No module objects, no autograd.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

PLANES = ("xy", "xz", "yz")
GN_GROUPS = 32      # nn.py:93-100  normalization() == GroupNorm32(32, C)
GN_EPS = 1e-5       # torch.nn.GroupNorm default, never overridden by the reference


@dataclass
class UNetSpec:
    """Constructor arguments of the reference UNet (unet_triplane.py:346-357)."""
    in_channels: int = 12
    model_channels: int = 64
    out_channels: int = 12
    num_res_blocks: int = 1
    channel_mult: Sequence[int] = (1, 2)
    use_scale_shift_norm: bool = True
    rollout: bool = True            # True: ...Small, False: ...SmallRaw

    @property
    def emb_dim(self) -> int:
        return 4 * self.model_channels


# --------------------------------------------------------------------------- layout
def split_planes(x: torch.Tensor, H: int, W: int, D: int):
    """[B,C,H+D,W+D] -> xy[B,C,H,W], xz[B,C,H,D], yz[B,C,W,D]   (triplane_util.py:20-25)."""
    return x[..., :H, :W], x[..., :H, W:], x[..., H:, :W].transpose(-1, -2)


def join_planes(xy, xz, yz):
    """Inverse of split_planes; the D x D corner is zero (triplane_util.py:7-17)."""
    D = xz.shape[-1]
    corner = xy.new_zeros(*xy.shape[:-2], D, D)
    top = torch.cat([xy, xz], dim=-1)
    bot = torch.cat([yz.transpose(-1, -2), corner], dim=-1)
    return torch.cat([top, bot], dim=-2)


# --------------------------------------------------------------------------- layers
def sinusoid(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)    # as nn.py:114-116
    ang = t[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    if dim % 2:
        e = torch.cat([e, torch.zeros_like(e[:, :1])], dim=-1)
    return e


def silu(x):
    return x * torch.sigmoid(x)     # nn.py:12-14


def rollout_inputs(p):
    """Axis-mean 'rollout' concat feeding every 3x3 conv (unet_triplane.py:37-46).

    xy[H,W], xz[H,D], yz[W,D]; every plane gets the other two planes' means over the axis it
    does not share, broadcast along the axis it does not have.
    """
    xy, xz, yz = p
    m = lambda a, d: a.mean(dim=d, keepdim=True)
    xy_in = torch.cat([xy, m(yz, -1).transpose(-1, -2).expand_as(xy), m(xz, -1).expand_as(xy)], 1)
    xz_in = torch.cat([xz, m(xy, -1).expand_as(xz), m(yz, -2).expand_as(xz)], 1)
    yz_in = torch.cat([yz, m(xy, -2).transpose(-1, -2).expand_as(yz), m(xz, -2).expand_as(yz)], 1)
    return xy_in, xz_in, yz_in


def tri_conv(sd, prefix, p, pad, rollout):
    if rollout:
        p = rollout_inputs(p)
    return tuple(F.conv2d(a, sd[f"{prefix}.conv_{n}.weight"], sd[f"{prefix}.conv_{n}.bias"], padding=pad)
                 for a, n in zip(p, PLANES))


def tri_norm(sd, prefix, p):
    return tuple(F.group_norm(a.float(), GN_GROUPS, sd[f"{prefix}.norm_{n}.weight"],
                              sd[f"{prefix}.norm_{n}.bias"], GN_EPS)
                 for a, n in zip(p, PLANES))


def res_block(sd, prefix, p, emb, spec: UNetSpec, trace=None):
    """unet_triplane.py:269-311 (no up/down inside the block)."""
    h = tri_norm(sd, f"{prefix}.in_layers.0", p)
    h = tuple(silu(a) for a in h)
    h = tri_conv(sd, f"{prefix}.in_layers.2", h, 1, spec.rollout)
    e = F.linear(silu(emb), sd[f"{prefix}.emb_layers.1.weight"], sd[f"{prefix}.emb_layers.1.bias"])
    e = e[:, :, None, None]
    if spec.use_scale_shift_norm:
        scale, shift = e.chunk(2, dim=1)
        h = tri_norm(sd, f"{prefix}.out_layers.0", h)
        h = tuple(a * (1 + scale) + shift for a in h)
    else:
        h = tuple(a + e for a in h)
        h = tri_norm(sd, f"{prefix}.out_layers.0", h)
    h = tuple(silu(a) for a in h)
    h = tri_conv(sd, f"{prefix}.out_layers.2", h, 1, spec.rollout)
    if f"{prefix}.skip_connection.conv_xy.weight" in sd:
        s = tri_conv(sd, f"{prefix}.skip_connection", p, 0, False)
    else:
        s = p
    return tuple(a + b for a, b in zip(h, s))


def avgpool2(p):
    return tuple(F.avg_pool2d(a, 2, 2) for a in p)                   # :127-145


def up2(p):
    return tuple(F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=False) for a in p)   # :106-124


def block_plan(spec: UNetSpec):
    """Names and channel counts of every block, mirroring the constructor loops (:377-434)."""
    mc = spec.model_channels
    ch = int(spec.channel_mult[0] * mc)
    chans = [ch]
    downs = []      # list of levels; each: list of ("down",) | ("res", name, cin, cout)
    for level, mult in enumerate(spec.channel_mult):
        ops = []
        idx = 0
        if level != 0:
            ops.append(("down",))
            idx = 1
        for _ in range(spec.num_res_blocks):
            cout = int(mult * mc)
            ops.append(("res", f"input_blocks.{level}.{idx}", ch, cout))
            idx += 1
            ch = cout       # NB the reference only updates ch after the loop; equal for nrb==1
        ch = int(mult * mc)
        chans.append(ch)
        downs.append(ops)
    ups = []
    n = len(spec.channel_mult)
    for j, (level, mult) in enumerate(list(enumerate(spec.channel_mult))[::-1]):
        ops = []
        for i in range(spec.num_res_blocks):
            ich = chans.pop()
            if level == n - 1 and i == 0:
                ich = 0
            cout = int(mc * mult)
            ops.append(("res", f"output_blocks.{j}.{i}", ch + ich, cout))
        ch = int(mc * mult)
        if level > 0:
            ops.append(("up",))
        ups.append(ops)
    return downs, ups


def unet_forward(sd: Dict[str, torch.Tensor], spec: UNetSpec, x, t, H, W, D, trace: dict = None):
    """x [B,C,H+D,W+D] fp32, t [B] (int64 or float) -> [B,C_out,H+D,W+D]."""
    if spec.num_res_blocks != 1:
        # The reference constructor keeps `ch` stale inside the res-block loop and pops one skip per
        # output res block, so num_res_blocks>1 builds a model whose forward cannot run.
        raise NotImplementedError("reference UNet only runs with num_res_blocks == 1")
    emb = sinusoid(t, spec.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    p = split_planes(x.float(), H, W, D)
    p = tri_conv(sd, "in_conv.0", p, 0, False)
    if trace is not None:
        trace["emb"] = emb
        trace["in_conv"] = p
    downs, ups = block_plan(spec)
    stack = []
    for ops in downs:
        for op in ops:
            if op[0] == "down":
                p = avgpool2(p)
            else:
                p = res_block(sd, op[1], p, emb, spec)
                if trace is not None:
                    trace[op[1]] = p
        stack.append(p)
    for j, ops in enumerate(ups):
        if j == 0:
            p = stack.pop()
        else:
            s = stack.pop()
            if spec.rollout:   # only ...Small resizes to the skip's size (:494-499)
                p = tuple(a if a.shape[2:] == b.shape[2:] else
                          F.interpolate(a, size=b.shape[2:], mode="bilinear", align_corners=False)
                          for a, b in zip(p, s))
            p = tuple(torch.cat([a, b], dim=1) for a, b in zip(p, s))
        for op in ops:
            if op[0] == "up":
                p = up2(p)
            else:
                p = res_block(sd, op[1], p, emb, spec)
                if trace is not None:
                    trace[op[1]] = p
    p = tri_norm(sd, "out.0", p)
    p = tuple(silu(a) for a in p)
    p = tri_conv(sd, "out.2", p, 0, False)
    return join_planes(*p)


# --------------------------------------------------------------------------- synthetic weights
def param_shapes(spec: UNetSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list == reference ``state_dict()`` order (138 tensors for defaults)."""
    mc, ed = spec.model_channels, spec.emb_dim
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def lin(name, i, o):
        out.extend([(f"{name}.weight", (o, i)), (f"{name}.bias", (o,))])

    def conv(name, i, o, k):
        for n in PLANES:
            out.extend([(f"{name}.conv_{n}.weight", (o, i, k, k)), (f"{name}.conv_{n}.bias", (o,))])

    def norm(name, c):
        for n in PLANES:
            out.extend([(f"{name}.norm_{n}.weight", (c,)), (f"{name}.norm_{n}.bias", (c,))])

    r = 3 if spec.rollout else 1

    def res(name, cin, cout):
        norm(f"{name}.in_layers.0", cin)
        conv(f"{name}.in_layers.2", cin * r, cout, 3)
        lin(f"{name}.emb_layers.1", ed, 2 * cout if spec.use_scale_shift_norm else cout)
        norm(f"{name}.out_layers.0", cout)
        conv(f"{name}.out_layers.2", cout * r, cout, 3)
        if cin != cout:
            conv(f"{name}.skip_connection", cin, cout, 1)

    lin("time_embed.0", mc, ed)
    lin("time_embed.2", ed, ed)
    conv("in_conv.0", spec.in_channels, int(spec.channel_mult[0] * mc), 1)
    downs, ups = block_plan(spec)
    for ops in downs + ups:
        for op in ops:
            if op[0] == "res":
                res(op[1], op[2], op[3])
    c0 = int(spec.channel_mult[0] * mc)
    norm("out.0", c0)
    conv("out.2", c0, spec.out_channels, 1)
    return out


def synthetic_state_dict(spec: UNetSpec, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Deterministic non-degenerate weights (nothing left at the reference's zero-init).

    Conv/linear weights ~ U(-a, a) with a = sqrt(3/fan_in) (unit gain), biases ~ N(0, 0.05),
    norm gains ~ 1 + N(0, 0.1), norm biases ~ N(0, 0.1).  Drawn from one CPU generator in key
    order, so any process with the same torch build reproduces them bit-for-bit.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in param_shapes(spec):
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
            a = math.sqrt(3.0 / fan_in)
            v = (torch.rand(shape, generator=g) * 2 - 1) * a
        sd[key] = v.float().contiguous()
    return sd
