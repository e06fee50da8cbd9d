"""CPU oracle for the triplane decoder (test infrastructure — see oracle/__init__.py).

Functional fp32 restatement, over a flat ``state_dict`` with the reference's own checkpoint keys, of

  * ``AutoEncoderGroupSkip.decode``           reference src/encoding/networks.py:192-220
  * ``sample_feature_plane2D``                reference src/encoding/networks.py:182-190
  * ``TriplaneGroupResnetBlock.forward``      reference src/encoding/blocks.py:189-256 (input_norm=False, input_act=False,
                                              as constructed at networks.py:152-160)
  * ``compose/decompose_triplane_channelwise`` reference src/encoding/blocks.py:164-186
  * ``DecoderMLPSkipConcat.forward``          reference src/encoding/blocks.py:65-91
  * ``ShapeAutoEncoder.decode_batch / decode_grid``  reference src/encoding/model.py:319-349
  * ``sample_grid_points_aabb``               reference src/encoding/utils3d.py:13-25
  * ``AutoEncoderGroupSkip.encode``           reference src/encoding/networks.py:164-180
  * ``AutoEncoderGroupPBR``                   reference src/encoding/networks.py:227-331 (``net_kind="pbr"``: geo block ks 5, two
                                              texture blocks ks 3 — the second with input_norm / input_act — and the rgb / mr /
                                              normal heads on the shared texture planes)

Written against plain ``torch.nn.functional`` on CPU tensors.  No module objects, no autograd.
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

PLANES = ("xy", "xz", "yz")
IN_EPS = 1e-6          # blocks.py:213-215  nn.InstanceNorm2d(..., eps=1e-6, affine=True)
COORDS = ((0, 1), (0, 2), (1, 2))     # networks.py:201: plane i is sampled at point coordinates COORDS[i]


@dataclass
class DecoderSpec:
    """Constructor arguments of the reference auto-encoder that matter for decode (networks.py:134-162;
    defaults = utils/parser_util.py:19-26: sdftex, fdim_geo 4, fdim_tex 8, fdim_up 64, hidden 256, 4 hidden layers)."""
    geo_feat_channels: int = 4
    tex_feat_channels: int = 8
    feat_channel_up: int = 64
    mlp_hidden_channels: int = 256
    mlp_hidden_layers: int = 4
    use_tex: bool = True
    tex_channels: int = 3
    ks: int = 5
    mlp_kind: str = "skip"      # "skip": AutoEncoderGroupSkip / DecoderMLPSkipConcat; "base": AutoEncoderGroupV3 / DecoderMLP (networks.py:21-131)
    net_kind: str = "skip"      # "pbr": AutoEncoderGroupPBR (networks.py:227-331; tex_channels 8, `ks` unused: 5 for geo, 3 for tex)

    @property
    def out_channels(self) -> int:
        return 1 + (self.tex_channels if self.use_tex else 0)


def _branches(spec: DecoderSpec):
    b = [("geo", spec.geo_feat_channels, 1)]
    if spec.use_tex:
        b.append(("tex", spec.tex_feat_channels, spec.tex_channels))
    return b


# AutoEncoderGroupPBR (networks.py:241-257): (block prefix, in channels (None = feat_channel_up), ks, input norm + act) per branch,
# and (head prefix, branch, outputs) in output order
def pbr_blocks(spec: DecoderSpec):
    out = {"geo": [("geo_convs.", spec.geo_feat_channels, 5, False)]}
    if spec.use_tex:
        out["tex"] = [("tex_convs.0.", spec.tex_feat_channels, 3, False), ("tex_convs.1.", spec.feat_channel_up, 3, True)]
    return out


def pbr_heads(spec: DecoderSpec):
    h = [("geo_decoder.", "geo", 1)]
    if spec.use_tex:
        h += [("rgb_decoder.", "tex", 3), ("mr_decoder.", "tex", 2), ("normal_decoder.", "tex", 3)]
    return h


def mlp_layer_names(spec: DecoderSpec):
    """(sequential name, index) of every Linear of DecoderMLPSkipConcat, in evaluation order (blocks.py:65-91); for the plain
    DecoderMLP (blocks.py:46-62) everything is one Sequential `layers` and the second list is empty."""
    nh = spec.mlp_hidden_layers
    if spec.mlp_kind == "base":
        return [("layers", 2 * i) for i in range(nh + 2)], []
    first = [("first_layers", 2 * i) for i in range(1 + nh // 2)]
    second = [("second_layers", 2 * i) for i in range(1 + max(nh // 2 - 1, 0) + 1)]
    return first, second


def param_shapes(spec: DecoderSpec):
    """state_dict keys and shapes of the reference module, in state_dict() order (decode side + the encoder convs a
    reference checkpoint also carries)."""
    up, hid = spec.feat_channel_up, spec.mlp_hidden_channels
    out = [("aabb", (6,)), ("geo_encoder.weight", (spec.geo_feat_channels, 1, 4, 4, 4)),
           ("geo_encoder.bias", (spec.geo_feat_channels,))]
    if spec.use_tex:
        out += [("tex_encoder.weight", (spec.tex_feat_channels, spec.tex_channels + 1, 4, 4, 4)),
                ("tex_encoder.bias", (spec.tex_feat_channels,))]
    def mlp_shapes(q, oc):
        o = []
        first, second = mlp_layer_names(spec)
        for j, (seq, i) in enumerate(first):
            cin = up if j == 0 else hid
            cout = oc if (spec.mlp_kind == "base" and j == len(first) - 1) else hid
            o += [(q + f"{seq}.{i}.weight", (cout, cin)), (q + f"{seq}.{i}.bias", (cout,))]
        for j, (seq, i) in enumerate(second):
            cin = up + hid if j == 0 else hid
            cout = oc if j == len(second) - 1 else hid
            o += [(q + f"{seq}.{i}.weight", (cout, cin)), (q + f"{seq}.{i}.bias", (cout,))]
        return o

    if spec.net_kind == "pbr":
        blocks, heads = pbr_blocks(spec), pbr_heads(spec)
        for br in blocks:
            for p, c, ks, in_norm in blocks[br]:
                conv = p + ("in_layers.1" if in_norm else "in_layers.0")      # Sequential(SiLU, Conv2d) with input_act (blocks.py:199-204)
                out += [(conv + ".weight", (3 * up, c, ks, ks)), (conv + ".bias", (3 * up,))]
                for pl in PLANES:
                    out += [(p + f"norm_{pl}.weight", (up,)), (p + f"norm_{pl}.bias", (up,))]
                out += [(p + "out_layers.1.weight", (3 * up, up, ks, ks)), (p + "out_layers.1.bias", (3 * up,))]
                if c != up:
                    out += [(p + "shortcut.weight", (3 * up, c, 1, 1)), (p + "shortcut.bias", (3 * up,))]
            for q, hb, oc in heads:
                if hb == br:
                    out += mlp_shapes(q, oc)
        return out
    for name, c, oc in _branches(spec):
        p = f"{name}_convs."
        out += [(p + "in_layers.0.weight", (3 * up, c, spec.ks, spec.ks)), (p + "in_layers.0.bias", (3 * up,))]
        for pl in PLANES:
            out += [(p + f"norm_{pl}.weight", (up,)), (p + f"norm_{pl}.bias", (up,))]
        out += [(p + "out_layers.1.weight", (3 * up, up, spec.ks, spec.ks)), (p + "out_layers.1.bias", (3 * up,))]
        if c != up:
            out += [(p + "shortcut.weight", (3 * up, c, 1, 1)), (p + "shortcut.bias", (3 * up,))]
        q = f"{name}_decoder."
        first, second = mlp_layer_names(spec)
        for j, (seq, i) in enumerate(first):
            cin = up if j == 0 else hid
            cout = oc if (spec.mlp_kind == "base" and j == len(first) - 1) else hid
            out += [(q + f"{seq}.{i}.weight", (cout, cin)), (q + f"{seq}.{i}.bias", (cout,))]
        for j, (seq, i) in enumerate(second):
            cin = up + hid if j == 0 else hid
            cout = oc if j == len(second) - 1 else hid
            out += [(q + f"{seq}.{i}.weight", (cout, cin)), (q + f"{seq}.{i}.bias", (cout,))]
    return out


def synthetic_state_dict(spec: DecoderSpec, seed: int) -> Dict[str, torch.Tensor]:
    """Random weights of the checkpoint layout.  The reference zero-initialises out_layers.1 (blocks.py:222-224); a
    trained checkpoint does not have zeros there, so every tensor is random here (norm weights around 1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(spec):
        if k == "aabb":
            sd[k] = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
        elif "norm_" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = torch.randn(shp, generator=g) * (1.5 / fan_in ** 0.5)
    return sd


# --------------------------------------------------------------------------- encoder
def encode(sd, spec: DecoderSpec, vol):
    """AutoEncoderGroupSkip.encode (networks.py:164-180): vol [1, 1 (+tex_channels), X, Y, Z] -> [xy, xz, yz] latent planes.
    Conv3d(k 4, stride 2, pad 1) per branch, mean over one volume axis, InstanceNorm2d(eps 1e-5, no affine), tanh(x / 2)."""
    feat = F.conv3d(vol[:, :1], sd["geo_encoder.weight"], sd["geo_encoder.bias"], stride=2, padding=1)
    if spec.use_tex:
        feat = torch.cat([feat, F.conv3d(vol, sd["tex_encoder.weight"], sd["tex_encoder.bias"], stride=2, padding=1)], dim=1)
    return [(F.instance_norm(feat.mean(dim=d), eps=1e-5) * 0.5).tanh() for d in (4, 3, 2)]


# --------------------------------------------------------------------------- feature planes
def silu(x):
    return x * torch.sigmoid(x)     # blocks.py:94-96


def compose_channelwise(maps):
    """blocks.py:164-177: zero-pad the three planes to [max(H,W), max(W,D)] and stack them on the channel axis."""
    xy, xz, yz = maps
    H, W = xy.shape[-2:]
    D = xz.shape[-1]
    nh, nw = max(H, W), max(W, D)
    xy = F.pad(xy, (0, nw - W, 0, nh - H))
    xz = F.pad(xz, (0, nw - D, 0, nh - H))
    yz = F.pad(yz, (0, nw - D, 0, nh - W))
    return torch.cat([xy, xz, yz], dim=1), (H, W, D)


def decompose_channelwise(h, sizes):
    """blocks.py:180-186."""
    H, W, D = sizes
    C = h.shape[1] // 3
    return h[:, :C, :H, :W], h[:, C:2 * C, :H, :D], h[:, 2 * C:, :W, :D]


def group_resnet_block(sd, prefix: str, maps, ks: int, in_norm: bool = False):
    """TriplaneGroupResnetBlock.forward (blocks.py:232-256): [input_norm + input_act: the block's own per-plane InstanceNorms
    on the input, then SiLU ->] grouped conv ks x ks -> per-plane InstanceNorm -> SiLU -> grouped conv ks x ks, + grouped 1x1
    shortcut (identity on the — normed — input when the channel counts agree)."""
    pad = (ks - 1) // 2
    if in_norm:
        maps = [F.instance_norm(a, weight=sd[prefix + f"norm_{pl}.weight"], bias=sd[prefix + f"norm_{pl}.bias"], eps=IN_EPS)
                for a, pl in zip(maps, PLANES)]
        x, sizes = compose_channelwise(maps)
        h = F.conv2d(silu(x), sd[prefix + "in_layers.1.weight"], sd[prefix + "in_layers.1.bias"], padding=pad, groups=3)
    else:
        x, sizes = compose_channelwise(maps)
        h = F.conv2d(x, sd[prefix + "in_layers.0.weight"], sd[prefix + "in_layers.0.bias"], padding=pad, groups=3)
    hs = decompose_channelwise(h, sizes)
    hs = [F.instance_norm(a, weight=sd[prefix + f"norm_{pl}.weight"], bias=sd[prefix + f"norm_{pl}.bias"], eps=IN_EPS)
          for a, pl in zip(hs, PLANES)]
    h, _ = compose_channelwise(hs)
    h = F.conv2d(silu(h), sd[prefix + "out_layers.1.weight"], sd[prefix + "out_layers.1.bias"], padding=pad, groups=3)
    if (prefix + "shortcut.weight") in sd:
        h = h + F.conv2d(x, sd[prefix + "shortcut.weight"], sd[prefix + "shortcut.bias"], groups=3)
    else:
        h = h + x
    return decompose_channelwise(h, sizes)


def feature_planes(sd, spec: DecoderSpec, feat_maps):
    """The up-convolved geo / tex planes decode() samples (networks.py:203-213).  They depend on the latent only, so a
    caller may compute them once per latent; the reference recomputes them for every chunk of points."""
    g = spec.geo_feat_channels
    if spec.net_kind == "pbr":          # networks.py:307-316
        out = {}
        for br, blocks in pbr_blocks(spec).items():
            maps = [fm[:, :g] if br == "geo" else fm[:, g:] for fm in feat_maps]
            for p, _, ks, in_norm in blocks:
                maps = group_resnet_block(sd, p, maps, ks, in_norm)
            out[br] = maps
        return out
    out = {"geo": group_resnet_block(sd, "geo_convs.", [fm[:, :g] for fm in feat_maps], spec.ks)}
    if spec.use_tex:
        out["tex"] = group_resnet_block(sd, "tex_convs.", [fm[:, g:] for fm in feat_maps], spec.ks)
    return out


# --------------------------------------------------------------------------- point decode
def sample_plane(feat_map, xy):
    """networks.py:182-190: bilinear grid_sample, border padding, align_corners=False, coordinates flipped."""
    n = xy.shape[0]
    return F.grid_sample(feat_map, xy.view(1, 1, n, 2).flip(-1), align_corners=False,
                         padding_mode="border")[0, :, 0, :].transpose(0, 1)


def mlp_skip_concat(sd, prefix: str, spec: DecoderSpec, x):
    """DecoderMLPSkipConcat.forward (blocks.py:84-91), posenc == 0."""
    first, second = mlp_layer_names(spec)
    h = x
    if spec.mlp_kind == "base":             # DecoderMLP.forward (blocks.py:59-62): Linear + ReLU ..., last Linear bare
        for j, (seq, i) in enumerate(first):
            h = F.linear(h, sd[prefix + f"{seq}.{i}.weight"], sd[prefix + f"{seq}.{i}.bias"])
            if j < len(first) - 1:
                h = F.relu(h)
        return h
    for seq, i in first:
        h = F.relu(F.linear(h, sd[prefix + f"{seq}.{i}.weight"], sd[prefix + f"{seq}.{i}.bias"]))
    h = torch.cat([x, h], dim=-1)
    for j, (seq, i) in enumerate(second):
        h = F.linear(h, sd[prefix + f"{seq}.{i}.weight"], sd[prefix + f"{seq}.{i}.bias"])
        if j < len(second) - 1:
            h = F.relu(h)
    return h


def decode(sd, spec: DecoderSpec, pts, feat_maps, aabb=None, planes=None):
    """AutoEncoderGroupSkip.decode (networks.py:192-220): pts [N,3] -> [N, 1 (+tex)]."""
    if aabb is None:
        aabb = sd["aabb"]
    x = 2 * (pts - aabb[:3]) / (aabb[3:] - aabb[:3]) - 1
    if planes is None:
        planes = feature_planes(sd, spec, feat_maps)
    outs = []
    if spec.net_kind == "pbr":          # networks.py:318-331: four heads, no sigmoid
        feat = {}
        for br in planes:
            feat[br] = sum(sample_plane(planes[br][i], x[..., list(COORDS[i])]) for i in range(3))
        return torch.cat([mlp_skip_concat(sd, q, spec, feat[hb]) for q, hb, _ in pbr_heads(spec)], dim=1)
    for name, _, _ in _branches(spec):
        h = 0
        for i in range(3):
            h = h + sample_plane(planes[name][i], x[..., list(COORDS[i])])
        h = mlp_skip_concat(sd, f"{name}_decoder.", spec, h)
        outs.append(h.sigmoid() if name == "tex" else h)
    return torch.cat(outs, dim=1)


def decode_batch(sd, spec: DecoderSpec, feat_maps, points, batch_size=2 ** 14, aabb=None, hoist=True):
    """model.py:319-333: chunked decode, colour channels clamped to [0,1].  `hoist` computes the feature planes once
    (same values: they do not depend on the points)."""
    planes = feature_planes(sd, spec, feat_maps) if hoist else None
    preds = [decode(sd, spec, points[i:i + batch_size], feat_maps, aabb=aabb, planes=planes)
             for i in range(0, points.shape[0], batch_size)]
    preds = torch.cat(preds, dim=0)
    preds[..., 1:] = preds[..., 1:].clamp_(0, 1)
    return preds


def grid_axes(aabb, resolution: int):
    """The three coordinate vectors of sample_grid_points_aabb (utils3d.py:13-25)."""
    aabb_min, aabb_max = torch.split(aabb, 3, dim=-1)
    size = aabb_max - aabb_min
    res = (resolution * size / size.max()).long()
    return [torch.linspace(0.5, res[i] - 0.5, int(res[i])) / res[i] * size[i] + aabb_min[i] for i in range(3)]


def grid_points(aabb, resolution: int):
    xs, ys, zs = grid_axes(aabb, resolution)
    return torch.stack(torch.meshgrid(xs, ys, zs, indexing="ij"), dim=-1)


def decode_grid(sd, spec: DecoderSpec, feat_maps, reso: int, batch_size=2 ** 14, aabb=None):
    """model.py:335-349."""
    if aabb is None:
        aabb = sd["aabb"]
    coords = grid_points(aabb, reso)
    nx, ny, nz, _ = coords.shape
    return decode_batch(sd, spec, feat_maps, coords.view(-1, 3), batch_size=batch_size, aabb=aabb).view(nx, ny, nz, -1)
