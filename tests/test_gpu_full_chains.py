"""-m gpu: the BASELINE.json configurations END TO END against the real reference.

Fixtures tests/golden/full_*.npz hold the final latent the unmodified reference produced on the CPU (oracle/make_golden_full.py:
SpacedDiffusion.p_sample_loop / ddim_sample_loop, reference src/diffusion/gaussian_diffusion.py:442-536, 640-734) for
  cfg2  DDPM-1000, C=12, (92,128,92), B=1          cfg3  DDIM-100, D=138, B=8          cfg5 shape, 20-step DDPM, B=8
with x_T and the per-step noise drawn on the CPU from the seeds in oracle/cases.py.  Here the same noise is redrawn, replayed
through `step_noise=` and the whole chain runs in the CUDA-graph loop.  Tolerance: north_star's 1e-3 rel-fp32 (rel-L2 over the
three planes of sample 0 and over the strided elements of the other samples)."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_ref as ur
from oracle.cases import FULL_CASES, full_errors, make_step_noise
from sin3dm_b200.script_util import create_gaussian_diffusion
from sin3dm_b200.unet_triplane import DEFAULT_PRECISION
from tests.gpu_util import make_cuda_model

pytestmark = pytest.mark.gpu
TOL = 1e-3


def run_full_chain(name, precision):
    case = FULL_CASES[name]
    spec = ur.UNetSpec(**case["spec"])
    m = make_cuda_model(spec, ur.synthetic_state_dict(spec, case["wseed"]), precision)
    d = create_gaussian_diffusion(steps=case["T"], predict_xstart=True, timestep_respacing=case["respacing"])
    H, W, D = case["HWD"]
    if case["ddim"]:                      # eta = 0: the per-step noise is multiplied by sigma = 0 on both sides
        g = torch.Generator().manual_seed(case["nseed"])
        x_T = torch.randn(case["B"], spec.in_channels, H + D, W + D, generator=g)
        sn = None
    else:
        x_T, noises = make_step_noise(case, d.num_timesteps)
        sn = torch.stack([noises[i] for i in range(d.num_timesteps)]).cuda()
        del noises
    kw = dict(noise=x_T, clip_denoised=True, model_kwargs=dict(H=H, W=W, D=D), step_noise=sn)
    with torch.no_grad():
        fn = d.ddim_sample_loop if case["ddim"] else d.p_sample_loop
        out = fn(m, list(x_T.shape), **kw).cpu()
    del sn
    torch.cuda.empty_cache()
    return out


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_full_chain_matches_reference(golden_dir, name):
    case = FULL_CASES[name]
    fx = np.load(os.path.join(golden_dir, f"full_{name}.npz"))
    got = run_full_chain(name, DEFAULT_PRECISION)
    rel0, rel_rest = full_errors(got, fx, *case["HWD"])
    print(f"full chain {name} precision {DEFAULT_PRECISION}: rel-L2 sample0 {rel0:.3e}, others {rel_rest:.3e}")
    assert rel0 < TOL and rel_rest < TOL, (name, rel0, rel_rest)
