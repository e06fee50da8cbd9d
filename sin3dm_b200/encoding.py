"""Host mirror of the reference triplane decoder (decode side of the auto-encoder).

  reference: ``AutoEncoderGroupSkip``               src/encoding/networks.py:134-223
             ``ShapeAutoEncoder.decode_batch/_grid`` src/encoding/model.py:319-349
             ``sample_grid_points_aabb``             src/encoding/utils3d.py:13-25

``AutoEncoderGroupSkip`` keeps the reference constructor signature, parameter names (so a reference ``.pt`` loads with
``load_state_dict``) and ``decode`` / ``reset_aabb``; the arithmetic runs in ``libsin3dm_b200.so`` (``s3d_decoder_*``,
include/sin3dm_b200.h), including the encoder half (``encode``: Conv3d + axis means + InstanceNorm + tanh, SURVEY §8(f)
rank 3, ``s3d_decoder_encode``).  Inference only (no autograd).  There is no CPU / torch fallback.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib


class _Holder(nn.Module):
    """Parameter-only stand-in for a reference sub-module (never called)."""


def _conv_params(cout, cin, ks, dims=2):
    m = _Holder()
    m.weight = nn.Parameter(torch.empty(cout, cin, *([ks] * dims)))
    m.bias = nn.Parameter(torch.empty(cout))
    nn.init.kaiming_uniform_(m.weight, a=5 ** 0.5)
    nn.init.uniform_(m.bias, -0.1, 0.1)
    return m


def _affine_params(c):
    m = _Holder()
    m.weight = nn.Parameter(torch.ones(c))
    m.bias = nn.Parameter(torch.zeros(c))
    return m


def _group_resnet_params(cin, cout, ks):
    """Parameters of TriplaneGroupResnetBlock(cin, cout, ks, input_norm=False, input_act=False) (blocks.py:189-230)."""
    m = _Holder()
    m.in_layers = nn.Sequential(_conv_params(3 * cout, cin, ks))
    m.norm_xy, m.norm_xz, m.norm_yz = _affine_params(cout), _affine_params(cout), _affine_params(cout)
    out_conv = _conv_params(3 * cout, cout, ks)
    with torch.no_grad():                      # zero_module (blocks.py:222-224)
        out_conv.weight.zero_()
        out_conv.bias.zero_()
    m.out_layers = nn.Sequential(_Holder(), out_conv)
    m.shortcut = _conv_params(3 * cout, cin, 1)
    return m


def _group_resnet_params_in(c, ks):
    """Parameters of TriplaneGroupResnetBlock(c, c, ks, input_norm=True, input_act=True) (blocks.py:199-204): the conv sits at
    in_layers.1 behind the SiLU, the shortcut is the identity."""
    m = _Holder()
    m.in_layers = nn.Sequential(_Holder(), _conv_params(3 * c, c, ks))
    m.norm_xy, m.norm_xz, m.norm_yz = _affine_params(c), _affine_params(c), _affine_params(c)
    out_conv = _conv_params(3 * c, c, ks)
    with torch.no_grad():
        out_conv.weight.zero_()
        out_conv.bias.zero_()
    m.out_layers = nn.Sequential(_Holder(), out_conv)
    return m


def _mlp_params(cin, cout, hidden, n_hidden):
    """Parameters of DecoderMLPSkipConcat (blocks.py:65-83); ReLUs sit at the odd indices."""
    m = _Holder()
    first = [nn.Linear(cin, hidden), _Holder()]
    for _ in range(n_hidden // 2):
        first += [nn.Linear(hidden, hidden), _Holder()]
    second = [nn.Linear(cin + hidden, hidden), _Holder()]
    for _ in range(n_hidden // 2 - 1):
        second += [nn.Linear(hidden, hidden), _Holder()]
    second.append(nn.Linear(hidden, cout))
    m.first_layers, m.second_layers = nn.Sequential(*first), nn.Sequential(*second)
    return m


def sample_grid_points_axes(aabb, resolution):
    """The three coordinate vectors whose 'ij' meshgrid is sample_grid_points_aabb(aabb, resolution)
    (utils3d.py:13-25); same torch expressions, evaluated on aabb's device."""
    aabb_min, aabb_max = torch.split(aabb, 3, dim=-1)
    size = aabb_max - aabb_min
    res = (resolution * size / size.max()).long()
    return [torch.linspace(0.5, res[i] - 0.5, int(res[i]), device=aabb.device) / res[i] * size[i] + aabb_min[i]
            for i in range(3)]


def _mlp_params_plain(cin, cout, hidden, n_hidden):
    """Parameters of the plain DecoderMLP (blocks.py:46-62): one Sequential `layers`, ReLUs at the odd indices."""
    m = _Holder()
    layers = [nn.Linear(cin, hidden), _Holder()]
    for _ in range(n_hidden):
        layers += [nn.Linear(hidden, hidden), _Holder()]
    layers.append(nn.Linear(hidden, cout))
    m.layers = nn.Sequential(*layers)
    return m


class AutoEncoderGroupSkip(nn.Module):
    """Decode side of the reference auto-encoder (networks.py:134-223) on the sm_100a kernels."""
    _mlp_kind = 0                      # s3d_decoder_config.mlp_kind
    _net_kind = 0                      # s3d_decoder_config.net_kind
    _mlp_factory = staticmethod(_mlp_params)

    def __init__(self, geo_feat_channels, tex_feat_channels, feat_channel_up, mlp_hidden_channels, mlp_hidden_layers,
                 use_tex=True, tex_channels=3, posenc=0):
        super().__init__()
        if posenc != 0:
            raise NotImplementedError("posenc > 0 is never constructed by the reference (networks.py:14)")
        self.use_tex = use_tex
        self.geo_feat_dim, self.tex_feat_dim = geo_feat_channels, tex_feat_channels
        self.feat_channel_up, self.mlp_hidden_channels = feat_channel_up, mlp_hidden_channels
        self.mlp_hidden_layers, self.tex_channels = mlp_hidden_layers, tex_channels
        self.geo_encoder = _conv_params(geo_feat_channels, 1, 4, dims=3)
        if use_tex:
            self.tex_encoder = _conv_params(tex_feat_channels, tex_channels + 1, 4, dims=3)
        self._build_decode_side(geo_feat_channels, tex_feat_channels, feat_channel_up, mlp_hidden_channels, mlp_hidden_layers)
        self.register_buffer("aabb", torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32))
        # kernel options: fp16x3 split (fp32-grade) unless S3D_PRECISION=1; tcgen05 MLP unless S3D_MLP_IMPL=ffma
        self.s3d_precision = int(os.environ.get("S3D_PRECISION", "3"))
        self.s3d_mlp_impl = 1 if os.environ.get("S3D_MLP_IMPL", "tc") == "ffma" else 0
        self._handle = self._handle_key = self._weights_key = self._planes_key = None

    def _build_decode_side(self, geo, tex, up, hidden, n_hidden):
        """Sub-modules in the reference's registration order (state_dict order)."""
        self.geo_convs = _group_resnet_params(geo, up, 5)
        self.geo_decoder = self._mlp_factory(up, 1, hidden, n_hidden)
        if self.use_tex:
            self.tex_convs = _group_resnet_params(tex, up, 5)
            self.tex_decoder = self._mlp_factory(up, self.tex_channels, hidden, n_hidden)

    # ------------------------------------------------------------------ reference surface
    def geo_parameters(self):
        return list(self.geo_encoder.parameters()) + list(self.geo_convs.parameters()) + list(self.geo_decoder.parameters())

    def tex_parameters(self):
        return list(self.tex_encoder.parameters()) + list(self.tex_convs.parameters()) + list(self.tex_decoder.parameters())

    def reset_aabb(self, aabb):
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.aabb = aabb.to(self.geo_encoder.weight.device)

    def encode(self, vol):
        """vol [1, 1 (+ tex_channels), X, Y, Z] -> [xy [1,C,H,W], xz [1,C,H,D], yz [1,C,W,D]]  (networks.py:164-180).
        Inference only: strided Conv3d + axis means + InstanceNorm2d + tanh(x/2) in two launches (s3d_decoder_encode)."""
        h = self.handle()
        dev = self.geo_encoder.weight.device
        cv = 1 + (self.tex_channels if self.use_tex else 0)
        if vol.dim() != 5 or vol.shape[0] != 1 or vol.shape[1] != cv:
            raise ValueError(f"vol must have shape [1, {cv}, X, Y, Z]; got {tuple(vol.shape)}")
        v = vol.detach().to(dev, torch.float32).contiguous()
        X, Y, Z = v.shape[2:]
        H, W, D = ((n - 2) // 2 + 1 for n in (X, Y, Z))
        c = self.geo_feat_dim + (self.tex_feat_dim if self.use_tex else 0)
        xy, xz, yz = (torch.empty(1, c, a, b, device=dev, dtype=torch.float32) for a, b in ((H, W), (H, D), (W, D)))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().s3d_decoder_encode(h, C.c_void_p(v.data_ptr()), X, Y, Z, C.c_void_p(xy.data_ptr()),
                                                     C.c_void_p(xz.data_ptr()), C.c_void_p(yz.data_ptr()),
                                                     _lib.current_stream_ptr()))
        return [xy, xz, yz]

    def forward(self, vol, x, aabb=None):
        return self.decode(x, self.encode(vol), aabb=aabb)

    # ------------------------------------------------------------------ handle management
    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.lib().s3d_decoder_destroy(self._handle)
            self._handle = self._handle_key = self._weights_key = self._planes_key = None

    def handle(self):
        dev = self.geo_encoder.weight.device
        if dev.type != "cuda":
            raise _lib.S3DError("sin3dm_b200 runs on CUDA (sm_100a) only: move the model with .to('cuda') first "
                                "(no CPU fallback)")
        L = _lib.lib()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        hkey = (idx, self.s3d_precision, self.s3d_mlp_impl)
        if self._handle is None or self._handle_key != hkey:
            self._drop_handle()
            cfg = _lib.DecoderConfig(self.geo_feat_dim, self.tex_feat_dim if self.use_tex else 0, self.feat_channel_up,
                                     self.mlp_hidden_channels, self.mlp_hidden_layers, int(bool(self.use_tex)),
                                     self.tex_channels, 5, self.s3d_precision, self.s3d_mlp_impl, self._mlp_kind, self._net_kind)
            h = C.c_void_p()
            _lib.check(L.s3d_decoder_create(C.byref(cfg), idx, C.byref(h)))
            self._handle, self._handle_key = h, hkey
            names = []
            for i in range(L.s3d_decoder_num_tensors(h)):
                nm, nd, shp = C.c_char_p(), C.c_int(), (C.c_int64 * 5)()
                _lib.check(L.s3d_decoder_tensor_info(h, i, C.byref(nm), C.byref(nd), shp))
                names.append((nm.value.decode(), tuple(shp[k] for k in range(nd.value))))
            mine = [(k, tuple(v.shape)) for k, v in self.state_dict().items()]
            if names != mine:
                raise _lib.S3DError("state_dict layout mismatch between host mirror and C library")
        wkey = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._weights_key != wkey:
            with torch.no_grad():
                for k, v in self.state_dict().items():
                    t = v.detach().to("cpu", torch.float32).contiguous()
                    shp = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.s3d_decoder_load_tensor(self._handle, k.encode(), C.c_void_p(t.data_ptr()), shp, t.dim()))
            _lib.check(L.s3d_decoder_finalize(self._handle))
            self._weights_key, self._planes_key = wkey, None
        return self._handle

    def _bind_planes(self, feat_maps):
        """Runs the two TriplaneGroupResnetBlocks for this latent unless they are already resident."""
        h = self.handle()
        xy, xz, yz = (f.detach().to(torch.float32).contiguous() for f in feat_maps)
        c = self.geo_feat_dim + (self.tex_feat_dim if self.use_tex else 0)
        H, W = xy.shape[-2:]
        D = xz.shape[-1]
        if tuple(xy.shape) != (1, c, H, W) or tuple(xz.shape) != (1, c, H, D) or tuple(yz.shape) != (1, c, W, D):
            raise ValueError(f"feat_maps must be [1,{c},H,W], [1,{c},H,D], [1,{c},W,D]; got "
                             f"{tuple(xy.shape)}, {tuple(xz.shape)}, {tuple(yz.shape)}")
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (xy, xz, yz))
        if self._planes_key != key:
            with torch.cuda.device(xy.device):
                _lib.check(_lib.lib().s3d_decoder_set_planes(h, C.c_void_p(xy.data_ptr()), C.c_void_p(xz.data_ptr()),
                                                             C.c_void_p(yz.data_ptr()), H, W, D, _lib.current_stream_ptr()))
            self._planes_key, self._planes_keep = key, (xy, xz, yz)
        return h

    def _aabb6(self, aabb):
        a = (self.aabb if aabb is None else aabb).detach().to("cpu", torch.float32).reshape(6)
        return (C.c_float * 6)(*a.tolist())

    @property
    def out_channels(self):
        return 1 + (self.tex_channels if self.use_tex else 0)

    # ------------------------------------------------------------------ decode
    def decode(self, x, feat_maps, aabb=None, clamp_tex=False):
        """x [N,3], feat_maps [xy, xz, yz] -> [N, 1 (+ tex_channels)]  (networks.py:192-220).  Inference only."""
        h = self._bind_planes(feat_maps)
        dev = self.geo_encoder.weight.device
        pts = x.detach().to(dev, torch.float32).contiguous()
        if pts.dim() != 2 or pts.shape[1] != 3:
            raise ValueError("x must have shape [N, 3]")
        out = torch.empty(pts.shape[0], self.out_channels, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().s3d_decoder_decode(h, C.c_void_p(pts.data_ptr()), pts.shape[0], self._aabb6(aabb),
                                                     int(clamp_tex), C.c_void_p(out.data_ptr()), _lib.current_stream_ptr()))
        return out

    def decode_axes(self, xs, ys, zs, feat_maps, aabb=None, clamp_tex=True):
        """Decode the 'ij' meshgrid of three coordinate vectors without materialising it -> [nx, ny, nz, C]."""
        h = self._bind_planes(feat_maps)
        dev = self.geo_encoder.weight.device
        xs, ys, zs = (v.detach().to(dev, torch.float32).contiguous() for v in (xs, ys, zs))
        out = torch.empty(xs.numel(), ys.numel(), zs.numel(), self.out_channels, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().s3d_decoder_decode_grid(h, C.c_void_p(xs.data_ptr()), C.c_void_p(ys.data_ptr()),
                                                          C.c_void_p(zs.data_ptr()), xs.numel(), ys.numel(), zs.numel(),
                                                          self._aabb6(aabb), int(clamp_tex), C.c_void_p(out.data_ptr()),
                                                          _lib.current_stream_ptr()))
        return out

    def feature_planes(self):
        """Bring-up aid: the up-convolved planes of the bound latent, [rows, cols, 64 (+64)] fp32 each (tests only)."""
        L, h = _lib.lib(), self.handle()
        out = []
        for p, t in enumerate(self._planes_keep):
            buf = torch.empty(t.shape[-2], t.shape[-1], self.feat_channel_up * (2 if self.use_tex else 1))
            _lib.check(L.s3d_decoder_planes_read(h, p, C.c_void_p(buf.data_ptr()), buf.numel()))
            out.append(buf)
        return out


class AutoEncoderGroupV3(AutoEncoderGroupSkip):
    """``enc_net_type == "base"`` (reference networks.py:21-131): the same encoder and feature-plane blocks with plain DecoderMLP
    heads (blocks.py:46-62) instead of the skip-concat ones; same kernels, the fourth layer just has no concat chunk."""
    _mlp_kind = 1
    _mlp_factory = staticmethod(_mlp_params_plain)


class AutoEncoderGroupPBR(AutoEncoderGroupSkip):
    """``enc_net_type == "pbr"`` (reference networks.py:227-331, data_type "sdfpbr"): geometry block ks 5; two texture blocks ks 3,
    the second with input InstanceNorm + SiLU and an identity shortcut; four skip-concat heads — sdf(1), rgb(3),
    metallic-roughness(2), normal(3) — on the shared texture planes, no sigmoid.  decode -> [N, 9]."""
    _net_kind = 1

    def __init__(self, geo_feat_channels, tex_feat_channels, feat_channel_up, mlp_hidden_channels, mlp_hidden_layers,
                 use_tex=True, tex_channels=3, posenc=0):
        if use_tex and tex_channels != 8:
            raise ValueError("AutoEncoderGroupPBR decodes rgb(3) + mr(2) + normal(3): tex_channels must be 8 (data_type 'sdfpbr')")
        super().__init__(geo_feat_channels, tex_feat_channels, feat_channel_up, mlp_hidden_channels, mlp_hidden_layers,
                         use_tex=use_tex, tex_channels=tex_channels, posenc=posenc)

    def _build_decode_side(self, geo, tex, up, hidden, n_hidden):
        self.geo_convs = _group_resnet_params(geo, up, 5)
        self.geo_decoder = _mlp_params(up, 1, hidden, n_hidden)
        if self.use_tex:
            self.tex_convs = nn.Sequential(_group_resnet_params(tex, up, 3), _group_resnet_params_in(up, 3))
            self.rgb_decoder = _mlp_params(up, 3, hidden, n_hidden)
            self.mr_decoder = _mlp_params(up, 2, hidden, n_hidden)
            self.normal_decoder = _mlp_params(up, 3, hidden, n_hidden)

    def tex_parameters(self):
        return list(self.tex_encoder.parameters()) + list(self.tex_convs.parameters()) + list(self.rgb_decoder.parameters()) + \
            list(self.mr_decoder.parameters()) + list(self.normal_decoder.parameters())


def get_networks(cfg):
    """networks.py:7-18."""
    use_tex = cfg.data_type != "sdf"
    tex_channels = 8 if cfg.data_type == "sdfpbr" else 3
    cls = {"base": AutoEncoderGroupV3, "skip": AutoEncoderGroupSkip, "pbr": AutoEncoderGroupPBR}.get(cfg.enc_net_type)
    if cls is None:
        raise ValueError("Unknown / unsupported net type: {}".format(cfg.enc_net_type))
    return cls(cfg.fdim_geo, cfg.fdim_tex, cfg.fdim_up, cfg.hidden_dim, cfg.n_hidden_layers, use_tex=use_tex, tex_channels=tex_channels)


class TriplaneDecoder:
    """``decode_batch`` / ``decode_grid`` of the reference's ShapeAutoEncoder (model.py:319-349) around a net."""

    def __init__(self, net: AutoEncoderGroupSkip, aabb=None):
        self.net = net
        self.aabb = net.aabb if aabb is None else torch.as_tensor(aabb, dtype=torch.float32)

    @torch.no_grad()
    def decode_batch(self, triplane_feat, points, batch_size=2 ** 14, aabb=None):
        """model.py:319-333.  ``batch_size`` is accepted for compatibility: the reference chunks to bound activation
        memory; here activations never leave the SM, so all points go out in one launch."""
        return self.net.decode(points, triplane_feat, aabb=aabb, clamp_tex=True)

    @torch.no_grad()
    def decode_grid(self, triplane_feat, reso, batch_size=2 ** 14, aabb=None):
        """model.py:335-349 -> [nx, ny, nz, C]; the grid points are formed inside the kernel."""
        if aabb is None:
            aabb = self.aabb
        aabb = torch.as_tensor(aabb, dtype=torch.float32)
        xs, ys, zs = sample_grid_points_axes(aabb.cpu(), reso)
        return self.net.decode_axes(xs, ys, zs, triplane_feat, aabb=aabb, clamp_tex=True)
