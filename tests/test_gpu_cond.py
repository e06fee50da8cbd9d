"""-m gpu: guided-sampling hooks of GaussianDiffusion against fixtures from the real reference.

cond_fn -> condition_mean (p_sample, reference gaussian_diffusion.py:357-370) / condition_score (ddim_sample, :372-394),
denoised_fn (p_mean_variance, :294-315) and q_mean_variance (:172-187).  These combinations step from Python (the UNet forward is
the CUDA path, the hook arithmetic is the host mirror's torch expressions on device tensors)."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_ref as ur
from oracle.cases import COND_CASES, cond_fn, denoised_fn, make_step_noise
from sin3dm_b200.script_util import create_gaussian_diffusion
from tests.gpu_util import make_cuda_model, plane_errors

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("name", list(COND_CASES))
def test_cond_hooks_match_reference(golden_dir, name):
    case = COND_CASES[name]
    g = np.load(os.path.join(golden_dir, f"cond_{name}.npz"))
    spec = ur.UNetSpec(**case["spec"])
    m = make_cuda_model(spec, ur.synthetic_state_dict(spec, case["wseed"]))
    d = create_gaussian_diffusion(steps=case["T"], predict_xstart=case.get("mean_type", "start_x") == "start_x",
                                  rescale_timesteps=case.get("rescale_timesteps", False), timestep_respacing=case["respacing"])
    H, W, D = case["HWD"]
    x_T, noises = make_step_noise(case, d.num_timesteps)
    sn = torch.stack([noises[i] for i in range(d.num_timesteps)]).cuda()
    kw = dict(noise=x_T.cuda(), clip_denoised=True, model_kwargs=dict(H=H, W=W, D=D), step_noise=sn,
              cond_fn=cond_fn if case["cond"] else None, denoised_fn=denoised_fn if case["denoise"] else None)
    with torch.no_grad():
        if case["ddim"]:
            got = d.ddim_sample_loop(m, list(x_T.shape), eta=case.get("eta", 0.0), **kw)
        else:
            got = d.p_sample_loop(m, list(x_T.shape), **kw)
    rel, mx = plane_errors(got.cpu(), g["sample"], H, W, D)
    assert rel < TOL and mx < TOL, (name, rel, mx)
    # q_mean_variance: table look-ups + one multiply, bit-exact
    t = torch.from_numpy(g["t"]).cuda()
    qm, qv, qlv = d.q_mean_variance(x_T.cuda(), t)
    assert np.array_equal(qm.cpu().numpy(), g["q_mean"])
    assert np.array_equal(qv[:, 0, 0, 0].cpu().numpy(), g["q_var"]) and np.array_equal(qlv[:, 0, 0, 0].cpu().numpy(), g["q_logvar"])
