"""Full-length BASELINE chains from the REAL reference (authoring container only; test infrastructure).

    python -m oracle.make_golden_full [case ...]        # needs /root/reference/src (read-only import)

Runs the unmodified ``SpacedDiffusion.p_sample_loop`` / ``ddim_sample_loop`` (reference src/diffusion/gaussian_diffusion.py:442-536,
640-734) with ``TriplaneUNetModelSmall`` on the CPU for the configurations of BASELINE.json — cfg2 DDPM-1000 (B=1), cfg3 DDIM-100
(B=8, D=138), cfg5-shape 20-step DDPM (B=8) — with CPU-drawn x_T and per-step noise replayed through a patched ``th.randn_like``
(SURVEY §4.2), and freezes the final latent as ``tests/golden/full_<case>.npz`` (``oracle.cases.pack_full``).  The GPU test
(tests/test_gpu_full_chains.py) redraws the same noise from the same seeds and compares the CUDA-graph loop with these files.
"""
import os
import sys
import time

import numpy as np
import torch

from oracle import unet_ref as ur
from oracle.cases import FULL_CASES, make_step_noise, pack_full
from oracle.make_golden import OUT, build_ref_diffusion, build_ref_unet, ref_modules


def main(names):
    gd, respace, ut = ref_modules()
    torch.set_num_threads(os.cpu_count())
    for name in names:
        case = FULL_CASES[name]
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd)
        d = build_ref_diffusion(gd, respace, case)
        H, W, D = case["HWD"]
        shape = [case["B"], spec.in_channels, H + D, W + D]
        x_T, noises = make_step_noise(case, d.num_timesteps)
        it = iter(range(d.num_timesteps - 1, -1, -1))
        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: noises[next(it)]
        t0 = time.time()
        try:
            with torch.no_grad():
                fn = d.ddim_sample_loop if case["ddim"] else d.p_sample_loop
                want = fn(m, shape, noise=x_T, model_kwargs=dict(H=H, W=W, D=D), clip_denoised=True)
        finally:
            torch.randn_like = orig
        np.savez_compressed(os.path.join(OUT, f"full_{name}.npz"), **pack_full(want))
        print(f"full {name}: {d.num_timesteps} reference steps in {time.time() - t0:.0f} s, |x|max {want.abs().max().item():.3f}, "
              f"plane std {want[0, :, :H, :W].std().item():.3f}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(FULL_CASES))
