"""Fused AdamW + EMA over one flat parameter buffer (optimizer part).

  reference: ``TrainLoop`` builds ``AdamW(master_params, lr=, weight_decay=)`` and, after every ``opt.step()``, runs
  ``update_ema(ema_params, master_params, rate)`` for each EMA rate (src/diffusion/train_util.py:82-95, 160-167, 237-246;
  src/diffusion/nn.py:53-63) — about ten element-wise launches per tensor and step.  Here the parameters are re-pointed to views
  of one contiguous fp32 buffer (as are the gradients), and one kernel (``s3d_adamw_ema_step``) does the whole update in a single
  HBM pass.  The backward pass that fills the gradients is not part of this library yet: any autograd-capable
  module can be driven with it, e.g. the reference UNet.
"""
import ctypes as C

import torch

from . import _lib


class FusedAdamWEMA:
    def __init__(self, params, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, ema_rates=()):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.S3DError("sin3dm_b200 runs on CUDA only (no CPU fallback)")
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ValueError("all parameters must be fp32 tensors on one CUDA device")
        if isinstance(ema_rates, str):                      # the reference's flag form: "0.9999,0.999" (train_util.py:52-56)
            ema_rates = [float(x) for x in ema_rates.split(",")]
        self.ema_rates = [float(r) for r in ([ema_rates] if isinstance(ema_rates, float) else ema_rates)]
        if len(self.ema_rates) > 4:
            raise ValueError("at most 4 EMA rates")
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), tuple(betas), float(eps)
        self.step_count = 0
        sizes = [p.numel() for p in self.params]
        # every tensor starts on a 16-byte boundary inside the flat buffers (float4 accesses)
        self.offsets, n = [], 0
        for s in sizes:
            self.offsets.append(n)
            n += (s + 3) // 4 * 4
        self.n = n
        self.flat = torch.zeros(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                                   # the module now lives in the flat buffer
                p.grad = self.grad[o:o + p.numel()].view_as(p)  # autograd accumulates straight into the flat gradient
                p._s3d_flat = (self.grad, o)                    # lets the UNet's backward add its whole flat gradient in one pass
        self.ema = [self.flat.clone() for _ in self.ema_rates]  # copies of the initial parameters (train_util.py:93-96)

    def zero_grad(self):
        self.grad.zero_()

    def ema_params(self, k):
        """The k-th EMA copy as a list of tensors shaped like the parameters (views of its flat buffer)."""
        return [self.ema[k][o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]

    @torch.no_grad()
    def step(self, lr=None):
        """One AdamW update followed by every EMA update; ``lr`` overrides the base rate (TrainLoop._anneal_lr)."""
        for p, o in zip(self.params, self.offsets):             # a caller may have replaced .grad (e.g. zero_grad(set_to_none))
            if p.grad is None:
                # torch skips parameters without a gradient; here the flat kernel visits them, so give them a zero gradient
                # (AdamW still decays them and advances their moments: a parameter that never gets a gradient is not expected)
                self.grad[o:o + p.numel()].zero_()
                p.grad = self.grad[o:o + p.numel()].view_as(p)
            elif p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                self.grad[o:o + p.numel()].view_as(p).copy_(p.grad)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
        self.step_count += 1
        a = _lib.AdamWArgs()
        a.param, a.grad, a.exp_avg, a.exp_avg_sq = (t.data_ptr() for t in (self.flat, self.grad, self.exp_avg, self.exp_avg_sq))
        for k, (e, r) in enumerate(zip(self.ema, self.ema_rates)):
            a.ema[k], a.ema_rate[k] = e.data_ptr(), r
        a.n_ema, a.n = len(self.ema), self.n
        a.lr = self.lr if lr is None else float(lr)
        a.beta1, a.beta2, a.eps, a.weight_decay, a.step = self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count
        with torch.cuda.device(self.flat.device):
            _lib.check(_lib.lib().s3d_adamw_ema_step(C.byref(a), _lib.current_stream_ptr()))
        # the kernel wrote the parameter storage behind torch's back: bump the version counters so that weight caches keyed on
        # them (handle()) re-pack
        bump = getattr(torch.autograd.graph, "increment_version", None)
        for p in self.params:
            if bump is not None:
                bump(p)
            else:
                p.add_(0)

    def state_dict(self):
        """step count, moments and EMA copies (the reference saves opt*.pt / ema_*.pt, train_util.py:258-281)."""
        return dict(step_count=self.step_count, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    ema=[e.clone() for e in self.ema], ema_rates=list(self.ema_rates), lr=self.lr, weight_decay=self.weight_decay)

    def load_state_dict(self, sd):
        if sd["exp_avg"].numel() != self.n or len(sd["ema"]) != len(self.ema):
            raise ValueError("optimizer state does not match this parameter set")
        self.step_count = int(sd["step_count"])
        with torch.no_grad():
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            for e, s in zip(self.ema, sd["ema"]):
                e.copy_(s)
