"""Drop-in ``TriplaneUNetModelSmall`` / ``TriplaneUNetModelSmallRaw`` backed by libsin3dm_b200.

Mirrors reference src/diffusion/unet_triplane.py:315-510 (…Small) and :513-702 (…SmallRaw):
same constructor arguments, same ``forward(x, timesteps, H=, W=, D=, y=)``, same ``state_dict`` keys
(138 tensors for the defaults) and the same construction-time initialisation (stock ``nn.Conv2d`` /
``nn.Linear`` / ``nn.GroupNorm`` initialisers drawn in the reference's order, second conv of every
block and the output conv zeroed, unet_triplane.py:243-246, 444).

The torch modules below only OWN the parameters.  All arithmetic of ``forward`` runs in the
hand-written sm_100a kernels through the C ABI (include/sin3dm_b200.h); there is no torch fallback.
"""
import ctypes as C
import os

import torch as th
import torch.nn as nn

from . import _lib
from .nn import sinusoid_freqs, zero_module

_PLANES = ("xy", "xz", "yz")
# conv operand terms of the tcgen05 kernels (include/sin3dm_b200.h: s3d_unet_config.precision); S3D_PRECISION overrides
DEFAULT_PRECISION = 3


class _TriConvParams(nn.Module):
    """Parameter holder with the reference's TriplaneConv names (unet_triplane.py:21-29)."""

    def __init__(self, channels, out_channels, kernel_size, padding, is_rollout=True):
        super().__init__()
        cin = channels * 3 if is_rollout else channels
        for n in _PLANES:
            setattr(self, f"conv_{n}", nn.Conv2d(cin, out_channels, kernel_size, padding=padding))


class _TriNormParams(nn.Module):
    """TriplaneNorm names (unet_triplane.py:63-68): three GroupNorm(32, C)."""

    def __init__(self, channels):
        super().__init__()
        for n in _PLANES:
            setattr(self, f"norm_{n}", nn.GroupNorm(32, channels))


class _Empty(nn.Module):
    """Stands in for the parameter-free modules (SiLU, up/down-sample) so Sequential indices match."""


class _ResBlockParams(nn.Module):
    def __init__(self, channels, emb_channels, out_channels, use_scale_shift_norm, is_rollout):
        super().__init__()
        self.in_layers = nn.Sequential(_TriNormParams(channels), _Empty(),
                                       _TriConvParams(channels, out_channels, 3, 1, is_rollout))
        self.emb_layers = nn.Sequential(
            _Empty(), nn.Linear(emb_channels, 2 * out_channels if use_scale_shift_norm else out_channels))
        self.out_layers = nn.Sequential(_TriNormParams(out_channels), _Empty(),
                                        zero_module(_TriConvParams(out_channels, out_channels, 3, 1, is_rollout)))
        if out_channels != channels:
            self.skip_connection = _TriConvParams(channels, out_channels, 1, 0, is_rollout=False)


class _UNetFn(th.autograd.Function):
    """forward = s3d_unet_forward_film on a training plan, backward = s3d_unet_backward (include/sin3dm_b200.h): gradients of every
    conv / norm parameter come back in one flat buffer, the gradient of the conditioning rows goes on through the embedding MLP
    (which torch differentiates: a [B, 256] GEMV chain).  Replaces autograd through TriplaneUNetModelSmall in
    TrainLoop.forward_backward (reference src/diffusion/train_util.py:198-235)."""

    @staticmethod
    def forward(ctx, module, x, film, H, W, D, *params):
        L, h = _lib.lib(), module.handle()
        _lib.check(L.s3d_unet_set_training(h, 1))
        B = x.shape[0]
        out = th.empty(B, module.out_channels, H + D, W + D, device=x.device, dtype=th.float32)
        film = film.detach().to(th.float32).contiguous()
        with th.cuda.device(x.device):
            _lib.check(L.s3d_unet_forward_film(h, C.c_void_p(x.data_ptr()), C.c_void_p(film.data_ptr()), None,
                                               C.c_void_p(out.data_ptr()), B, H, W, D, _lib.current_stream_ptr()))
        ctx.module, ctx.keep, ctx.n_params = module, (x, film), len(params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        L, h = _lib.lib(), module._handle
        x, film = ctx.keep
        g = grad_out.to(th.float32).contiguous()
        flat = th.empty(L.s3d_unet_grad_numel(h), device=g.device, dtype=th.float32)
        dfilm = th.empty_like(film)
        with th.cuda.device(g.device):
            _lib.check(L.s3d_unet_backward(h, C.c_void_p(g.data_ptr()), C.c_void_p(flat.data_ptr()), C.c_void_p(dfilm.data_ptr()),
                                           _lib.current_stream_ptr()))
        slots = module._kernel_param_slots()
        # Parameters that live in FusedAdamWEMA's flat buffers (same 4-float-aligned state_dict layout as the library's gradient
        # buffer): ONE add into the flat gradient instead of 130 AccumulateGrad kernels (the embedding MLP's slots are zero here
        # and are filled by torch afterwards).
        params = [p for _, p in module._kernel_params()]
        tag = getattr(params[0], "_s3d_flat", None)
        if tag is not None and tag[0].numel() == flat.numel() and tag[0].device == flat.device and all(
                getattr(p, "_s3d_flat", (None, -1))[0] is tag[0] and p._s3d_flat[1] == o and p.grad is not None
                and p.grad.data_ptr() == tag[0].data_ptr() + 4 * o for p, (o, _, _) in zip(params, slots)):
            tag[0].add_(flat)
            return (None, None, dfilm, None, None, None, *([None] * len(params)))
        grads = [flat[o:o + n].view(shape) for o, n, shape in slots]
        return (None, None, dfilm, None, None, None, *grads)


class _S3DUNet(nn.Module):
    _rollout = True

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks=1, dropout=0, channel_mult=(1, 2),
                 use_checkpoint=False, use_fp16=False, use_scale_shift_norm=False):
        super().__init__()
        if isinstance(channel_mult, str):
            channel_mult = tuple(int(v) for v in channel_mult.split(","))
        if num_res_blocks != 1:
            raise NotImplementedError(
                "num_res_blocks != 1: the reference constructor (unet_triplane.py:389-405) keeps a stale channel "
                "count inside the block loop and cannot build a runnable model either")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.dropout, self.channel_mult = num_res_blocks, dropout, tuple(channel_mult)
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.dtype = th.float16 if use_fp16 else th.float32     # accepted for compatibility; kernels pick precision

        ro = self._rollout
        emb = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, emb), _Empty(), nn.Linear(emb, emb))
        ch = int(channel_mult[0] * model_channels)
        self.in_conv = nn.Sequential(_TriConvParams(in_channels, ch, 1, 0, is_rollout=False))
        chans = [ch]
        self.input_blocks = nn.ModuleList()
        for level, mult in enumerate(channel_mult):
            layers = [_Empty()] if level != 0 else []
            cout = int(mult * model_channels)
            layers.append(_ResBlockParams(ch, emb, cout, use_scale_shift_norm, ro))
            ch = cout
            self.input_blocks.append(nn.Sequential(*layers))
            chans.append(ch)
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            ich = chans.pop()
            if level == len(channel_mult) - 1:
                ich = 0
            cout = int(model_channels * mult)
            layers = [_ResBlockParams(ch + ich, emb, cout, use_scale_shift_norm, ro)]
            ch = cout
            if level > 0:
                layers.append(_Empty())
            self.output_blocks.append(nn.Sequential(*layers))
        c0 = int(channel_mult[0] * model_channels)
        self.out = nn.Sequential(_TriNormParams(ch), _Empty(),
                                 zero_module(_TriConvParams(c0, out_channels, 1, 0, is_rollout=False)))

        # kernel options: fp16x3 split (fp32-grade) unless S3D_PRECISION=1; tcgen05 conv unless S3D_CONV_IMPL=ffma
        self.s3d_precision = int(os.environ.get("S3D_PRECISION", str(DEFAULT_PRECISION)))
        self.s3d_conv_impl = 1 if os.environ.get("S3D_CONV_IMPL", "tc") == "ffma" else 0
        self._handle = None
        self._handle_key = None
        self._weights_key = None

    # ------------------------------------------------------------------ compatibility no-ops
    def convert_to_fp16(self):
        """unet_triplane.py:451-456 retypes the torso convs; the kernels choose their own operand format."""

    def convert_to_fp32(self):
        """unet_triplane.py:458-463."""

    # ------------------------------------------------------------------ handle management
    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.lib().s3d_unet_destroy(self._handle)
            self._handle = None
            self._handle_key = None
            self._weights_key = None

    def _device(self):
        return next(self.parameters()).device

    def _fast_params(self):
        """Current parameters without nn.Module's recursive traversal (the module tree is fixed after construction; the Parameter
        objects are re-read from every leaf, so replaced parameters are seen).  handle() runs once per sampling call: this keeps
        its revalidation at ~20 us instead of ~200."""
        leaves = self.__dict__.get("_param_leaves")
        if leaves is None:
            leaves = self.__dict__["_param_leaves"] = [m for m in self.modules() if m._parameters]
        return [p for m in leaves for p in m._parameters.values() if p is not None]

    def handle(self):
        """C handle bound to the parameters' device, with the current parameter values packed."""
        params = self._fast_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.S3DError("sin3dm_b200 runs on CUDA (sm_100a) only: move the model with .to('cuda') first "
                                "(no CPU fallback)")
        L = _lib.lib()
        idx = dev.index if dev.index is not None else th.cuda.current_device()
        hkey = (idx, self.s3d_precision, self.s3d_conv_impl)
        if self._handle is None or self._handle_key != hkey:
            self._drop_handle()
            cfg = _lib.UNetConfig()
            cfg.in_channels, cfg.model_channels, cfg.out_channels = self.in_channels, self.model_channels, self.out_channels
            cfg.num_res_blocks, cfg.n_levels = self.num_res_blocks, len(self.channel_mult)
            for i, m in enumerate(self.channel_mult):
                cfg.channel_mult[i] = int(m)
            cfg.use_scale_shift_norm = int(bool(self.use_scale_shift_norm))
            cfg.rollout = int(self._rollout)
            cfg.precision, cfg.conv_impl = self.s3d_precision, self.s3d_conv_impl
            h = C.c_void_p()
            _lib.check(L.s3d_unet_create(C.byref(cfg), idx, C.byref(h)))
            self._handle, self._handle_key = h, hkey
            self._slots = None
            # the C side derives the expected checkpoint layout itself: cross-check with our state_dict
            names = []
            for i in range(L.s3d_unet_num_tensors(h)):
                nm, nd, shp = C.c_char_p(), C.c_int(), (C.c_int64 * 4)()
                _lib.check(L.s3d_unet_tensor_info(h, i, C.byref(nm), C.byref(nd), shp))
                names.append((nm.value.decode(), tuple(shp[k] for k in range(nd.value))))
            mine = [(k, tuple(v.shape)) for k, v in self.state_dict().items()]
            if names != mine:
                raise _lib.S3DError("state_dict layout mismatch between host mirror and C library")
        wkey = tuple((p.data_ptr(), p._version) for p in params)
        if self._weights_key != wkey and self._weights_key is not None and not os.environ.get("S3D_HOST_REPACK"):
            # the weights changed (optimizer step) and the operand buffers exist: re-pack on the device, no host round trip
            sd = [v.detach() for v in self.state_dict().values()]
            if all(v.is_cuda and v.dtype == th.float32 and v.is_contiguous() for v in sd):
                ptrs = (C.c_void_p * len(sd))(*[v.data_ptr() for v in sd])
                with th.cuda.device(dev):
                    _lib.check(L.s3d_unet_refresh_dev(self._handle, ptrs, len(sd), _lib.current_stream_ptr()))
                self._weights_key = wkey
                self.__dict__.pop("_film_cache", None)
        if self._weights_key != wkey:
            with th.no_grad():
                for k, v in self.state_dict().items():
                    t = v.detach().to("cpu", th.float32).contiguous()
                    shp = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.s3d_unet_load_tensor(self._handle, k.encode(), C.c_void_p(t.data_ptr()), shp, t.dim()))
                fr = sinusoid_freqs(self.model_channels).contiguous()
                shp = (C.c_int64 * 1)(fr.numel())
                _lib.check(L.s3d_unet_load_tensor(self._handle, b"__freqs", C.c_void_p(fr.data_ptr()), shp, 1))
            _lib.check(L.s3d_unet_finalize(self._handle))
            self._weights_key = wkey
        return self._handle

    @property
    def film_dim(self):
        if self._handle is not None:
            return _lib.lib().s3d_unet_film_dim(self._handle)
        return _lib.lib().s3d_unet_film_dim(self.handle())

    def film_table(self, timesteps, cache=False, handle=None, cache_key=None):
        """[n, film_dim] conditioning rows for ``timesteps`` (float32 values).  ``cache=True`` keeps the table per (weights,
        timesteps) so that repeated sampling loops see the same device buffer (the sampling graph is cached by address);
        ``handle`` / ``cache_key``: a handle() result and a precomputed key of ``timesteps`` the caller already has."""
        h = handle if handle is not None else self.handle()
        key = None
        if cache:
            key = (self._weights_key, cache_key if cache_key is not None else timesteps.detach().to("cpu", th.float32).numpy().tobytes())
            hit = self.__dict__.setdefault("_film_cache", {}).get(key)
            if hit is not None:
                return hit
        t = timesteps.to(self._device(), th.float32).contiguous()
        out = th.empty(t.numel(), self.film_dim, device=t.device, dtype=th.float32)
        with th.cuda.device(t.device):
            _lib.check(_lib.lib().s3d_unet_film(h, C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(out.data_ptr()),
                                                 _lib.current_stream_ptr()))
        if cache:
            if len(self._film_cache) >= 4:
                self._film_cache.clear()
            self._film_cache[key] = out
        return out

    # ------------------------------------------------------------------ training path
    def _is_embedding_param(self, name):
        return name.startswith("time_embed.") or ".emb_layers." in name

    def _kernel_params(self):
        """(name, parameter) of everything the backward kernels differentiate (all but the embedding MLP), state_dict order."""
        return [(n, p) for n, p in self.named_parameters() if not self._is_embedding_param(n)]

    def _kernel_param_slots(self):
        """(offset, numel, shape) inside the flat gradient buffer of s3d_unet_backward, aligned with _kernel_params()."""
        if getattr(self, "_slots", None) is None:
            L, h = _lib.lib(), self._handle
            index = {k: i for i, k in enumerate(self.state_dict().keys())}
            self._slots = [(L.s3d_unet_grad_offset(h, index[n]), p.numel(), tuple(p.shape)) for n, p in self._kernel_params()]
        return self._slots

    def _film_rows_torch(self, timesteps):
        """The conditioning rows [B, film_dim] through the torch-owned embedding MLP (differentiable): sinusoid -> time_embed ->
        every block's emb_layers, concatenated in block order (nn.py:103-121, unet_triplane.py:232-238, 371-375)."""
        fr = sinusoid_freqs(self.model_channels).to(timesteps.device)
        ang = timesteps.float()[:, None] * fr[None]
        e = th.cat([th.cos(ang), th.sin(ang)], dim=-1)
        emb = self.time_embed[2](th.nn.functional.silu(self.time_embed[0](e)))
        se = th.nn.functional.silu(emb)
        blocks = [m for seq in list(self.input_blocks) + list(self.output_blocks) for m in seq if isinstance(m, _ResBlockParams)]
        return th.cat([b.emb_layers[1](se) for b in blocks], dim=1)

    # ------------------------------------------------------------------ forward
    def forward(self, x, timesteps, H=None, W=None, D=None, y=None):
        """[N, C, H+D, W+D] -> same shape (unet_triplane.py:465-510).  With autograd enabled and parameters that require a gradient
        the call is differentiable w.r.t. the parameters (training); otherwise it is the plain inference forward."""
        assert H is not None and W is not None and D is not None
        train = th.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if th.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("sin3dm_b200 produces parameter gradients only (no gradient w.r.t. the UNet input)")
        h = self.handle()
        xin = x if y is None else th.cat([x, y], dim=1)
        xin = xin.to(th.float32).contiguous()
        B, Cin, Hc, Wc = xin.shape
        if Cin != self.in_channels or Hc != H + D or Wc != W + D:
            raise ValueError(f"input shape {tuple(xin.shape)} does not match in_channels={self.in_channels}, "
                             f"(H+D, W+D)=({H + D}, {W + D})")
        if timesteps.shape != (B,):
            raise ValueError("timesteps must have shape [N]")
        if train:
            film = self._film_rows_torch(timesteps.to(xin.device))
            return _UNetFn.apply(self, xin, film, int(H), int(W), int(D), *[p for _, p in self._kernel_params()])
        _lib.check(_lib.lib().s3d_unet_set_training(h, 0))
        t = timesteps.to(xin.device, th.float32).contiguous()
        out = th.empty(B, self.out_channels, Hc, Wc, device=xin.device, dtype=th.float32)
        with th.cuda.device(xin.device):
            _lib.check(_lib.lib().s3d_unet_forward(h, C.c_void_p(xin.data_ptr()), C.c_void_p(t.data_ptr()),
                                                    C.c_void_p(out.data_ptr()), B, H, W, D, _lib.current_stream_ptr()))
        return out

    # ------------------------------------------------------------------ debugging aid used by tests/
    def debug_activations(self):
        """{name: (xy, xz, yz)} NHWC fp32 activations of the last forward (tests only)."""
        L, h = _lib.lib(), self.handle()
        out = {}
        for i in range(L.s3d_unet_debug_count(h)):
            nm, ch, bt = C.c_char_p(), C.c_int(), C.c_int()
            rows, cols = (C.c_int * 3)(), (C.c_int * 3)()
            _lib.check(L.s3d_unet_debug_info(h, i, C.byref(nm), C.byref(ch), rows, cols, C.byref(bt)))
            planes = []
            for p in range(3):
                buf = th.empty(bt.value, rows[p], cols[p], ch.value, dtype=th.float32)
                _lib.check(L.s3d_unet_debug_read(h, i, p, C.c_void_p(buf.data_ptr()), buf.numel()))
                planes.append(buf)
            out[nm.value.decode()] = tuple(planes)
        return out


class TriplaneUNetModelSmall(_S3DUNet):
    """Rollout ("triplane-aware") UNet — reference unet_triplane.py:315-510."""
    _rollout = True


class TriplaneUNetModelSmallRaw(_S3DUNet):
    """Same network without the axis-mean rollout — reference unet_triplane.py:513-702."""
    _rollout = False
