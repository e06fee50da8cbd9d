"""Drop-in ``GaussianDiffusion`` whose sampling math runs in the sm_100a scheduler kernels.

Mirrors the public surface of reference src/diffusion/gaussian_diffusion.py (class at :102-931):
enums, fp64 numpy coefficient tables with the reference's attribute names, ``q_sample``,
``p_mean_variance``, ``p_sample(_loop)(_progressive)``, ``ddim_sample(_loop)(_progressive)``,
``ddim_reverse_sample`` and ``training_losses`` (MSE branch; the reference's KL branches raise, :792-793).

Execution
  * ``*_sample_loop`` with one of this package's UNets, no ``cond_fn`` / ``denoised_fn`` and
    ``progress=False`` runs entirely on the device: ``s3d_sample_loop`` replays one captured CUDA graph
    (UNet forward + fused scheduler step) per iteration with the step index held in device memory.
  * every other combination steps from Python but still uses the fused ``s3d_sched_step`` kernel for
    the per-step arithmetic (one launch instead of ~10 element-wise ops + 2-6 table uploads,
    gaussian_diffusion.py:294-315, 431-439, 581-599, 944).
Per-step noise: the reference draws ``th.randn_like`` every step (:431, :591).  Here the scheduler kernel
generates it in-register with Philox4x32-10 keyed by (seed, global sample index, step), so results do not
depend on batch split or GPU count; pass ``step_noise=`` ([n_steps, *shape] tensor or ``f(i) -> tensor``)
to inject explicit noise (parity tests).
"""
import ctypes as C
import enum
import math

import numpy as np
import torch as th

from . import _lib
from .triplane_util import decompose_featmaps


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:19-43."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:46-63."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _f32(arr):
    """table -> fp32 exactly like _extract_into_tensor's ``.float()`` (gaussian_diffusion.py:944)."""
    return th.from_numpy(np.asarray(arr, dtype=np.float64)).float()


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)
        if model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learned variances: the triplane UNets have no variance head "
                                      "(learn_sigma=False everywhere in the reference, parser_util.py:75-85)")
        if model_mean_type == ModelMeanType.PREVIOUS_X:
            raise NotImplementedError("PREVIOUS_X is unreachable through the reference factory (script_util.py:46-48)")
        self._coef_cache = {}

    # ------------------------------------------------------------------ tables for the kernels
    def _model_variance_tables(self):
        if self.model_var_type == ModelVarType.FIXED_LARGE:            # :279-285
            var = np.append(self.posterior_variance[1], self.betas[1:])
            return var, np.log(var)
        return self.posterior_variance, self.posterior_log_variance_clipped

    def coef_table(self, device, eta=0.0):
        """[T, 12] fp32 device table (column meaning: include/sin3dm_b200.h), every entry produced with the
        same fp32 torch expression the reference evaluates per element."""
        key = (str(device), float(eta))
        if key not in self._coef_cache:
            _, logvar = self._model_variance_tables()
            ab, abp, abn = _f32(self.alphas_cumprod), _f32(self.alphas_cumprod_prev), _f32(self.alphas_cumprod_next)
            sigma = eta * th.sqrt((1 - abp) / (1 - ab)) * th.sqrt(1 - ab / abp)          # :585-589
            cols = [
                _f32(self.sqrt_recip_alphas_cumprod), _f32(self.sqrt_recipm1_alphas_cumprod),
                _f32(self.posterior_mean_coef1), _f32(self.posterior_mean_coef2),
                th.exp(0.5 * _f32(logvar)),                                              # :439
                th.sqrt(abp), th.sqrt(1 - abp - sigma ** 2), sigma,                      # :593-594
                th.sqrt(abn), th.sqrt(1 - abn),                                          # :634-635
                _f32(self.sqrt_alphas_cumprod), _f32(self.sqrt_one_minus_alphas_cumprod),
            ]
            self._coef_cache[key] = th.stack(cols, dim=1).contiguous().to(device)
        return self._coef_cache[key]

    def _mean_code(self):
        return _lib.START_X if self.model_mean_type == ModelMeanType.START_X else _lib.EPSILON

    # ------------------------------------------------------------------ q(x_t | x_0)
    def q_mean_variance(self, x_start, t):
        mean = _extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = _extract_into_tensor(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = _extract_into_tensor(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def q_sample(self, x_start, t, noise=None):
        """:189-207, one fused kernel."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        if not x_start.is_cuda:
            raise _lib.S3DError("sin3dm_b200 runs on CUDA only")
        x0 = x_start.float().contiguous()
        nz = noise.float().contiguous()
        out = th.empty_like(x0)
        ti = t.to(x0.device, th.int32).contiguous()
        coef = self.coef_table(x0.device)
        with th.cuda.device(x0.device):
            _lib.check(_lib.lib().s3d_q_sample(_ptr(x0), _ptr(nz), _ptr(out), _ptr(coef), _ptr(ti), x0.shape[0],
                                                x0[0].numel(), _lib.current_stream_ptr()))
        return out

    def q_posterior_mean_variance(self, x_start, x_t, t):
        mean = (_extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + _extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        var = _extract_into_tensor(self.posterior_variance, t, x_t.shape)
        logvar = _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    # ------------------------------------------------------------------ model-side helpers
    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _model_timesteps(self, t):
        """Values handed to the network for step indices ``t`` (overridden by SpacedDiffusion)."""
        return self._scale_timesteps(t)

    def _call_model(self, model, x, t, model_kwargs):
        return model(x, self._model_timesteps(t), **(model_kwargs or {}))

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return ((_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart)
                / _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """:233-327.  Returns the four tensors the reference returns (torch element-wise ops: this entry point
        is only used by callers that want the distribution itself; sampling goes through the fused kernel)."""
        B = x.shape[0]
        assert t.shape == (B,)
        out = self._call_model(model, x, t, model_kwargs)
        var, logvar = self._model_variance_tables()
        model_variance = _extract_into_tensor(var, t, x.shape)
        model_log_variance = _extract_into_tensor(logvar, t, x.shape)

        def proc(v):
            if denoised_fn is not None:
                v = denoised_fn(v)
            return v.clamp(-1, 1) if clip_denoised else v

        if self.model_mean_type == ModelMeanType.START_X:
            pred = proc(out)
        else:
            pred = proc(self._predict_xstart_from_eps(x, t, out))
        mean, _, _ = self.q_posterior_mean_variance(pred, x, t)
        return {"mean": mean, "variance": model_variance, "log_variance": model_log_variance, "pred_xstart": pred}

    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        gradient = cond_fn(x, self._model_timesteps(t), **(model_kwargs or {}))
        return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()

    def condition_score(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        alpha_bar = _extract_into_tensor(self.alphas_cumprod, t, x.shape)
        eps = self._predict_eps_from_xstart(x, t, p_mean_var["pred_xstart"])
        eps = eps - (1 - alpha_bar).sqrt() * cond_fn(x, self._model_timesteps(t), **(model_kwargs or {}))
        out = dict(p_mean_var)
        out["pred_xstart"] = self._predict_xstart_from_eps(x, t, eps)
        out["mean"], _, _ = self.q_posterior_mean_variance(out["pred_xstart"], x, t)
        return out

    # ------------------------------------------------------------------ fused scheduler step
    def _sched(self, kind, model_out, x, t, noise, clip, eta=0.0, y0=None, mask=None, is_mask_t0=False, seed=0,
               sample_base=0):
        if not x.is_cuda:
            raise _lib.S3DError("sin3dm_b200 runs on CUDA only (no CPU fallback)")
        x = x.float().contiguous()
        mo = model_out.float().contiguous()
        B = x.shape[0]
        sample, x0 = th.empty_like(x), th.empty_like(x)
        a = _lib.SchedArgs()
        a.kind, a.mean_type, a.clip_denoised, a.is_mask_t0 = kind, self._mean_code(), int(bool(clip)), int(bool(is_mask_t0))
        a.B, a.C, a.n_per_sample = B, x.shape[1], x[0].numel()
        keep = [x, mo, sample, x0]
        a.model_out, a.x, a.sample, a.pred_xstart = mo.data_ptr(), x.data_ptr(), sample.data_ptr(), x0.data_ptr()
        if noise is not None:
            noise = noise.to(x.device, th.float32).contiguous()
            assert noise.shape == x.shape
            keep.append(noise)
            a.noise = noise.data_ptr()
        if y0 is not None and mask is not None:
            y0c, mc = y0.to(x.device, th.float32).contiguous(), mask.to(x.device, th.float32).contiguous()
            assert y0c.shape == x.shape and mc.shape == x.shape
            keep += [y0c, mc]
            a.y0, a.mask = y0c.data_ptr(), mc.data_ptr()
        coef = self.coef_table(x.device, eta)
        ti = t.to(x.device, th.int32).contiguous()
        keep += [coef, ti]
        a.coef_dev, a.t_idx_dev = coef.data_ptr(), ti.data_ptr()
        a.seed, a.sample_base = int(seed) & (2 ** 64 - 1), int(sample_base)
        with th.cuda.device(x.device):
            _lib.check(_lib.lib().s3d_sched_step(C.byref(a), _lib.current_stream_ptr()))
        return {"sample": sample, "pred_xstart": x0}

    @staticmethod
    def _fresh_seed():
        return int(th.randint(0, 2 ** 62, (1,)).item())      # follows torch.manual_seed

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 noise=None, _seed=None, _sample_base=0):
        """:396-440."""
        if denoised_fn is None and cond_fn is None:
            out = self._call_model(model, x, t, model_kwargs)
            return self._sched(_lib.DDPM, out, x, t, noise, clip_denoised,
                               seed=self._fresh_seed() if (noise is None and _seed is None) else (_seed or 0),
                               sample_base=_sample_base)
        o = self.p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)
        nz = noise if noise is not None else th.randn_like(x)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        if cond_fn is not None:
            o["mean"] = self.condition_mean(cond_fn, o, x, t, model_kwargs=model_kwargs)
        return {"sample": o["mean"] + nonzero * th.exp(0.5 * o["log_variance"]) * nz, "pred_xstart": o["pred_xstart"]}

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0,
                    y0=None, mask=None, is_mask_t0=False, noise=None, _seed=None, _sample_base=0):
        """:538-600."""
        if denoised_fn is None and cond_fn is None:
            out = self._call_model(model, x, t, model_kwargs)
            return self._sched(_lib.DDIM, out, x, t, noise, clip_denoised, eta, y0, mask, is_mask_t0,
                               seed=self._fresh_seed() if (noise is None and _seed is None) else (_seed or 0),
                               sample_base=_sample_base)
        o = self.p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)
        if cond_fn is not None:
            o = self.condition_score(cond_fn, o, x, t, model_kwargs=model_kwargs)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        if y0 is not None and mask is not None:
            mix = mask * y0 + (1 - mask) * o["pred_xstart"]
            o["pred_xstart"] = mix if is_mask_t0 else mix * nonzero + o["pred_xstart"] * (1 - nonzero)
        eps = self._predict_eps_from_xstart(x, t, o["pred_xstart"])
        ab = _extract_into_tensor(self.alphas_cumprod, t, x.shape)
        abp = _extract_into_tensor(self.alphas_cumprod_prev, t, x.shape)
        sigma = eta * th.sqrt((1 - abp) / (1 - ab)) * th.sqrt(1 - ab / abp)
        nz = noise if noise is not None else th.randn_like(x)
        mean = o["pred_xstart"] * th.sqrt(abp) + th.sqrt(1 - abp - sigma ** 2) * eps
        return {"sample": mean + nonzero * sigma * nz, "pred_xstart": o["pred_xstart"]}

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        """:602-638."""
        assert eta == 0.0, "Reverse ODE only for deterministic path"
        if denoised_fn is None:
            out = self._call_model(model, x, t, model_kwargs)
            return self._sched(_lib.DDIM_REVERSE, out, x, t, None, clip_denoised)
        o = self.p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)
        eps = self._predict_eps_from_xstart(x, t, o["pred_xstart"])
        abn = _extract_into_tensor(self.alphas_cumprod_next, t, x.shape)
        return {"sample": o["pred_xstart"] * th.sqrt(abn) + th.sqrt(1 - abn) * eps, "pred_xstart": o["pred_xstart"]}

    # ------------------------------------------------------------------ loops
    def _device_loop_ok(self, model, denoised_fn, cond_fn, model_kwargs, progress):
        from .unet_triplane import _S3DUNet
        return (isinstance(model, _S3DUNet) and denoised_fn is None and cond_fn is None and not progress
                and model_kwargs is not None and set(model_kwargs) == {"H", "W", "D"})

    def _run_device_loop(self, kind, model, shape, noise, clip_denoised, model_kwargs, device, eta, y0, mask, is_mask_t0,
                         step_noise, seed, sample_base, use_graph=True):
        h = model.handle()
        mdev = model._fast_params()[0].device
        dev = th.device(device) if device is not None else mdev
        if dev.type == "cuda" and dev.index is None:
            dev = th.device("cuda", th.cuda.current_device())
        if dev != mdev:
            raise ValueError(f"device={dev} but the model lives on {mdev}")
        H, W, D = (int(model_kwargs[k]) for k in ("H", "W", "D"))
        shape = tuple(int(v) for v in shape)
        # the C side sizes every access as out_channels * (H+D) * (W+D) per sample: refuse anything else here
        if len(shape) != 4 or shape[1] != model.in_channels or model.in_channels != model.out_channels \
                or shape[2:] != (H + D, W + D):
            raise ValueError(f"shape {shape} does not match the model (in/out channels {model.in_channels}/{model.out_channels}) "
                             f"and (H+D, W+D) = ({H + D}, {W + D})")
        B = shape[0]
        T = self.num_timesteps
        # Persistent buffers: the captured CUDA graph is cached by buffer address, so x lives in a per-shape workspace (the result
        # is returned as a copy) and the conditioning table is cached per (weights, timesteps).
        ws = self.__dict__.setdefault("_loop_ws", {})
        img = ws.get((str(dev), shape))
        if img is None:
            if len(ws) >= 4:
                ws.clear()
            img = ws[(str(dev), shape)] = th.empty(shape, device=dev, dtype=th.float32)
        if noise is not None:
            if tuple(noise.shape) != shape:
                raise ValueError(f"noise has shape {tuple(noise.shape)}, expected {shape}")
            img.copy_(noise.to(dev, th.float32))
        else:
            img.normal_()
        ts = self.__dict__.get("_loop_ts")          # the network's timesteps of the whole chain + their cache key, per instance
        if ts is None or ts[0] != T:
            tsv = self._model_timesteps(th.arange(T)).float()
            ts = self.__dict__["_loop_ts"] = (T, tsv, tsv.numpy().tobytes())
        film = model.film_table(ts[1], cache=True, handle=h, cache_key=ts[2])
        coef = self.coef_table(dev, eta)
        a = _lib.LoopArgs()
        a.kind, a.mean_type, a.clip_denoised, a.is_mask_t0 = kind, self._mean_code(), int(bool(clip_denoised)), int(bool(is_mask_t0))
        a.n_steps, a.t_start, a.B, a.H, a.W, a.D = T, T - 1, B, H, W, D
        a.n_per_sample = img[0].numel()
        keep = [img, film, coef]
        a.x_dev, a.coef_dev, a.film_dev = img.data_ptr(), coef.data_ptr(), film.data_ptr()
        if step_noise is not None:
            if callable(step_noise):
                step_noise = th.stack([step_noise(i).to(dev, th.float32) for i in range(T)])
            sn = step_noise.to(dev, th.float32).contiguous()
            if tuple(sn.shape) != (T, *shape):
                raise ValueError("step_noise must be [n_steps, *shape] indexed by step index")
            keep.append(sn)
            a.step_noise_dev = sn.data_ptr()
        if (y0 is None) != (mask is None):
            raise ValueError("y0 and mask go together")
        if y0 is not None:
            if tuple(y0.shape) != shape or tuple(mask.shape) != shape:      # the reference asserts the same (:568)
                raise ValueError(f"y0 / mask must have the sample's shape {shape} (got {tuple(y0.shape)}, {tuple(mask.shape)})")
            y0c, mc = y0.to(dev, th.float32).contiguous(), mask.to(dev, th.float32).contiguous()
            keep += [y0c, mc]
            a.y0_dev, a.mask_dev = y0c.data_ptr(), mc.data_ptr()
        a.seed = (self._fresh_seed() if seed is None else int(seed)) & (2 ** 64 - 1)
        a.sample_base = int(sample_base)
        a.use_graph = int(bool(use_graph))
        _lib.check(_lib.lib().s3d_unet_set_training(h, 0))
        with th.cuda.device(dev):
            _lib.check(_lib.lib().s3d_sample_loop(h, C.byref(a), _lib.current_stream_ptr()))
            out = img.clone()
        self._keepalive = keep      # buffers referenced by the cached graph stay alive until the next loop
        return out

    def _progressive(self, step_fn, model, shape, noise, device, progress, step_noise, seed, sample_base, **kw):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise.to(device) if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        seed = self._fresh_seed() if seed is None else seed
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            nz = None
            if step_noise is not None:
                nz = step_noise(i) if callable(step_noise) else step_noise[i]
            with th.no_grad():
                out = step_fn(model, img, t, noise=nz, _seed=seed, _sample_base=sample_base, **kw)
                yield out
                img = out["sample"]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, step_noise=None, seed=None, sample_base=0):
        """:442-486."""
        if self._device_loop_ok(model, denoised_fn, cond_fn, model_kwargs, progress):
            return self._run_device_loop(_lib.DDPM, model, shape, noise, clip_denoised, model_kwargs, device, 0.0, None,
                                         None, False, step_noise, seed, sample_base)
        final = None
        for s in self.p_sample_loop_progressive(model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                                device, progress, step_noise, seed, sample_base):
            final = s
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, step_noise=None, seed=None,
                                  sample_base=0):
        """:488-536."""
        yield from self._progressive(self.p_sample, model, shape, noise, device, progress, step_noise, seed, sample_base,
                                     clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                     model_kwargs=model_kwargs)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, y0=None, mask=None, is_mask_t0=False,
                         step_noise=None, seed=None, sample_base=0):
        """:640-678."""
        if self._device_loop_ok(model, denoised_fn, cond_fn, model_kwargs, progress):
            return self._run_device_loop(_lib.DDIM, model, shape, noise, clip_denoised, model_kwargs, device, eta, y0,
                                         mask, is_mask_t0, step_noise, seed, sample_base)
        final = None
        for s in self.ddim_sample_loop_progressive(model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                                   device, progress, eta, y0, mask, is_mask_t0, step_noise, seed,
                                                   sample_base):
            final = s
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, y0=None, mask=None,
                                     is_mask_t0=False, step_noise=None, seed=None, sample_base=0):
        """:680-734."""
        yield from self._progressive(self.ddim_sample, model, shape, noise, device, progress, step_noise, seed,
                                     sample_base, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                     model_kwargs=model_kwargs, eta=eta, y0=y0, mask=mask, is_mask_t0=is_mask_t0)

    # ------------------------------------------------------------------ variational bound (bits per dimension)
    def _vb_device(self, model, x_start, x_t, t, clip_denoised, model_kwargs, noise=None):
        """model -> fused k_vb_terms.  Returns (out [B, 3] fp32: vb term, xstart mse, eps mse; pred_xstart)."""
        if not x_t.is_cuda:
            raise _lib.S3DError("sin3dm_b200 runs on CUDA only (no CPU fallback)")
        xs, xt = x_start.float().contiguous(), x_t.float().contiguous()
        mo = self._call_model(model, xt, t, model_kwargs).float().contiguous()
        assert mo.shape == xs.shape == xt.shape
        B, n = xs.shape[0], xs[0].numel()
        dev = xs.device
        key = ("logvar", str(dev))
        if key not in self._coef_cache:
            _, logvar = self._model_variance_tables()
            self._coef_cache[key] = th.stack([_f32(self.posterior_log_variance_clipped), _f32(logvar)], dim=1).contiguous().to(dev)
        a = _lib.VbArgs()
        a.mean_type, a.clip_denoised, a.B, a.n_per_sample = self._mean_code(), int(bool(clip_denoised)), B, n
        pred, out = th.empty_like(xs), th.empty(B, 3, device=dev, dtype=th.float32)
        ws = th.empty(_lib.lib().s3d_vb_workspace_bytes(B, n), device=dev, dtype=th.uint8)
        ti = t.to(dev, th.int32).contiguous()
        coef = self.coef_table(dev)
        nz = noise.to(dev, th.float32).contiguous() if noise is not None else None
        a.x_start, a.x_t, a.model_out, a.pred_xstart = xs.data_ptr(), xt.data_ptr(), mo.data_ptr(), pred.data_ptr()
        a.noise = nz.data_ptr() if nz is not None else None
        a.coef_dev, a.logvar_dev, a.t_idx_dev = coef.data_ptr(), self._coef_cache[key].data_ptr(), ti.data_ptr()
        a.workspace, a.out = ws.data_ptr(), out.data_ptr()
        with th.cuda.device(dev):
            _lib.check(_lib.lib().s3d_vb_terms(C.byref(a), _lib.current_stream_ptr()))
        return out, pred

    def _vb_terms_bpd(self, model, x_start, x_t, t, clip_denoised=True, model_kwargs=None):
        """:736-769.  {'output': [N] KL (decoder NLL where t == 0) in bits per dimension, 'pred_xstart'}."""
        out, pred = self._vb_device(model, x_start, x_t, t, clip_denoised, model_kwargs)
        return {"output": out[:, 0].clone(), "pred_xstart": pred}

    def _prior_bpd(self, x_start):
        """:858-874: KL(q(x_T | x_0) || N(0, I)) in bits per dimension (depends on the schedule only; tiny element-wise host
        expression on the device tensor, evaluated once per calc_bpd_loop)."""
        B = x_start.shape[0]
        t = th.tensor([self.num_timesteps - 1] * B, device=x_start.device)
        mean, _, logvar = self.q_mean_variance(x_start, t)
        kl = 0.5 * (-1.0 - logvar + th.exp(logvar) + mean ** 2)
        return kl.mean(dim=list(range(1, kl.dim()))) / np.log(2.0)

    def calc_bpd_loop(self, model, x_start, clip_denoised=True, model_kwargs=None):
        """:876-931.  Per step: q_sample (fused) -> model -> k_vb_terms (vb term + both MSEs in one pass)."""
        dev = x_start.device
        B = x_start.shape[0]
        vb, xstart_mse, mse = [], [], []
        for i in range(self.num_timesteps - 1, -1, -1):
            t_batch = th.tensor([i] * B, device=dev)
            noise = th.randn_like(x_start)
            x_t = self.q_sample(x_start=x_start, t=t_batch, noise=noise)
            with th.no_grad():
                out, _ = self._vb_device(model, x_start, x_t, t_batch, clip_denoised, model_kwargs, noise=noise)
            vb.append(out[:, 0])
            xstart_mse.append(out[:, 1])
            mse.append(out[:, 2])
        vb, xstart_mse, mse = th.stack(vb, dim=1), th.stack(xstart_mse, dim=1), th.stack(mse, dim=1)
        prior_bpd = self._prior_bpd(x_start)
        return {"total_bpd": vb.sum(dim=1) + prior_bpd, "prior_bpd": prior_bpd, "vb": vb, "xstart_mse": xstart_mse,
                "mse": mse}

    # ------------------------------------------------------------------ training objective (forward values)
    def training_losses(self, model, x_start, t, model_kwargs=None, noise=None):
        """:771-856, MSE branch: q_sample (fused kernel) -> model -> per-plane mean squared error (k_plane_mse: one pass over the
        composed tensors, fixed-order fp64 reduction)."""
        if model_kwargs is None:
            model_kwargs = {}
        if noise is None:
            noise = th.randn_like(x_start)
        if self.loss_type in (LossType.KL, LossType.RESCALED_KL):
            raise NotImplementedError      # same as the reference (:792-793)
        x_t = self.q_sample(x_start, t, noise=noise)
        model_output = self._call_model(model, x_t, t, model_kwargs)
        target = {ModelMeanType.START_X: x_start, ModelMeanType.EPSILON: noise}[self.model_mean_type]
        assert model_output.shape == target.shape == x_start.shape
        H, W, D = (int(model_kwargs[k]) for k in ("H", "W", "D"))
        if model_output.requires_grad:
            # training: the loss has to stay on the autograd tape (TrainLoop.forward_backward calls loss.backward(),
            # train_util.py:198-235); the three plane means are tiny element-wise reductions on top of the UNet's output, whose
            # backward is s3d_unet_backward
            terms = {}
            for name, a, b in zip(("xy", "xz", "yz"), decompose_featmaps(target, (H, W, D)), decompose_featmaps(model_output, (H, W, D))):
                terms[f"mse_{name}"] = ((a - b) ** 2).mean(dim=(1, 2, 3))
            terms["loss"] = terms["mse_xy"] + terms["mse_xz"] + terms["mse_yz"]
            return terms
        tgt, outp = target.float().contiguous(), model_output.float().contiguous()
        B, Cc = tgt.shape[0], tgt.shape[1]
        assert tuple(tgt.shape[2:]) == (H + D, W + D)
        mse = th.empty(B, 3, device=tgt.device, dtype=th.float32)
        ws = th.empty(_lib.lib().s3d_vb_workspace_bytes(B, tgt[0].numel()), device=tgt.device, dtype=th.uint8)
        with th.cuda.device(tgt.device):
            _lib.check(_lib.lib().s3d_plane_mse(_ptr(tgt), _ptr(outp), B, Cc, H, W, D, _ptr(ws), _ptr(mse), _lib.current_stream_ptr()))
        terms = {"mse_xy": mse[:, 0], "mse_xz": mse[:, 1], "mse_yz": mse[:, 2]}
        terms["loss"] = terms["mse_xy"] + terms["mse_xz"] + terms["mse_yz"]
        return terms
