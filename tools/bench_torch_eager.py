#!/usr/bin/env python
"""The "library kernels to beat" (SURVEY §8d): the reference's arithmetic run by torch eager on the GPU (cuDNN / cuBLAS / ATen
kernels) — the oracle's functional restatement moved to cuda:0, cfg2 shape, UNet forward + DDPM step per iteration.

    python tools/bench_torch_eager.py [--steps 20]        # prints one JSON line (steps/s); NOT part of the product path"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import diffusion_ref as dr
    from oracle import unet_ref as ur
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tf32", type=int, default=0)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = bool(a.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    dev = torch.device("cuda", 0)
    spec = ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12)
    sd = {k: v.to(dev) for k, v in ur.synthetic_state_dict(spec, 1234).items()}
    H, W, D = 92, 128, 92
    x = torch.randn(1, 12, H + D, W + D, device=dev)
    o = dr.RefDiffusion(1000, "")
    o._x = lambda arr, t, like: torch.from_numpy(__import__("numpy").asarray(arr))[t.cpu()].float().to(like.device).view(-1, *([1] * (like.dim() - 1)))
    o.model_t = lambda t: t
    model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(i, img):
        t = torch.full((1,), i, dtype=torch.long, device=dev)
        return o.p_sample(model, img, t, torch.randn_like(img))["sample"]
    with torch.no_grad():
        img = x
        for i in range(3):
            img = step(999 - i, img)
        torch.cuda.synchronize()
        e0.record()
        for i in range(a.steps):
            img = step(996 - i, img)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps(dict(what="torch eager (oracle restatement on cuda:0, library kernels), cfg2 DDPM step", tf32=bool(a.tf32),
                          steps=a.steps, ms_per_step=ms, steps_per_s=1e3 / ms, torch=torch.__version__)), flush=True)


if __name__ == "__main__":
    main()
