"""Variational-bound path (_vb_terms_bpd / _prior_bpd / calc_bpd_loop, gaussian_diffusion.py:736-931).

CPU: the oracle restatement against the fixtures the real reference produced (oracle/make_golden_bpd.py).
GPU: k_vb_terms against the oracle on the same model output, and the whole calc_bpd_loop against the fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import BPD_CASES, make_bpd_inputs

KEYS = ("total_bpd", "prior_bpd", "vb", "xstart_mse", "mse")


def _oracle(case):
    return dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"),
                           case.get("var_type", "fixed_large"))


@pytest.mark.parametrize("name", list(BPD_CASES))
def test_oracle_bpd_matches_reference_golden(golden_dir, name):
    case = BPD_CASES[name]
    g = np.load(os.path.join(golden_dir, f"bpd_{name}.npz"))
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    H, W, D = case["HWD"]
    o = _oracle(case)
    x0, noises = make_bpd_inputs(case, o.num_timesteps)
    model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
    got = o.calc_bpd_loop(model, x0, lambda i: noises[i], clip=case.get("clip", True))
    for k in KEYS:
        assert np.allclose(got[k].numpy(), g[k], rtol=2e-5, atol=1e-6), k
    assert got["vb"].shape == (x0.shape[0], o.num_timesteps)


@pytest.mark.gpu
@pytest.mark.parametrize("mean_type", ["start_x", "epsilon"])
@pytest.mark.parametrize("var_type", ["fixed_large", "fixed_small"])
@pytest.mark.parametrize("clip", [True, False])
def test_vb_kernel_matches_oracle(mean_type, var_type, clip):
    from sin3dm_b200.script_util import create_gaussian_diffusion
    d = create_gaussian_diffusion(steps=1000, predict_xstart=(mean_type == "start_x"), sigma_small=(var_type == "fixed_small"),
                                  timestep_respacing="20")
    o = dr.RefDiffusion(1000, "20", "linear", mean_type, var_type)
    g = torch.Generator().manual_seed(9)
    shape = (4, 5, 37, 29)                        # n per sample not a multiple of the block size
    x0 = torch.rand(shape, generator=g) * 2 - 1
    x0.view(-1)[::11] = 1.0
    x0.view(-1)[3::13] = -1.0
    nz = torch.randn(shape, generator=g)
    t = torch.tensor([0, 1, 7, 19])               # decoder NLL, and KL at small / large variance, in one batch
    x_t = o.q_sample(x0, t, nz)
    # a model output close to the truth for some samples, far for others (deep tails of the discretised likelihood)
    mo = (x0 if mean_type == "start_x" else nz) + torch.randn(shape, generator=g) * torch.tensor([0.01, 0.3, 1.0, 0.05]).view(4, 1, 1, 1)
    want = o.vb_terms_bpd(lambda xx, tt: mo, x0, x_t, t, clip)
    got = d._vb_terms_bpd(lambda xx, tt, **k: mo.cuda(), x0.cuda(), x_t.cuda(), t.cuda(), clip_denoised=clip)
    assert torch.equal(got["pred_xstart"].cpu(), want["pred_xstart"])          # same fp32 op order: bit-exact
    rel = ((got["output"].cpu() - want["output"]).abs() / want["output"].abs().clamp(min=1e-6)).max().item()
    print("vb term rel err", rel)
    assert rel < 2e-5, (got["output"], want["output"])                          # fp64 block sums vs torch's fp32 mean
    # the two MSEs calc_bpd_loop adds
    out, _ = d._vb_device(lambda xx, tt, **k: mo.cuda(), x0.cuda(), x_t.cuda(), t.cuda(), clip, None, noise=nz.cuda())
    flat = lambda v: v.mean(dim=(1, 2, 3))
    assert torch.allclose(out[:, 1].cpu(), flat((want["pred_xstart"] - x0) ** 2), rtol=2e-5, atol=1e-9)
    assert torch.allclose(out[:, 2].cpu(), flat((o.eps_from_x0(x_t, t, want["pred_xstart"]) - nz) ** 2), rtol=2e-5, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(BPD_CASES))
def test_calc_bpd_loop_matches_reference_golden(golden_dir, name):
    from sin3dm_b200.script_util import create_gaussian_diffusion
    from tests.gpu_util import make_cuda_model
    case = BPD_CASES[name]
    g = np.load(os.path.join(golden_dir, f"bpd_{name}.npz"))
    spec = ur.UNetSpec(**case["spec"])
    m = make_cuda_model(spec, ur.synthetic_state_dict(spec, case["wseed"]))
    d = create_gaussian_diffusion(steps=case["T"], predict_xstart=case.get("mean_type", "start_x") == "start_x",
                                  sigma_small=case.get("var_type", "fixed_large") == "fixed_small",
                                  timestep_respacing=case["respacing"])
    H, W, D = case["HWD"]
    x0, noises = make_bpd_inputs(case, d.num_timesteps)
    it = iter(range(d.num_timesteps - 1, -1, -1))
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: noises[next(it)].to(x.device)
    try:
        with torch.no_grad():
            got = d.calc_bpd_loop(m, x0.cuda(), clip_denoised=case.get("clip", True), model_kwargs=dict(H=H, W=W, D=D))
    finally:
        torch.randn_like = orig
    for k in KEYS:
        a, b = got[k].cpu().numpy(), g[k]
        rel = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        print(name, k, "rel", rel)
        assert rel < 1e-3, (k, a, b)                  # north-star tolerance (UNet output feeds an exp(-logvar)-scaled term)
    # stand-alone _vb_terms_bpd at t == 0 / last step
    t = torch.from_numpy(g["vt_t"]).cuda()
    x_t = d.q_sample(x0.cuda(), t, noise=noises[0].cuda())
    with torch.no_grad():
        vt = d._vb_terms_bpd(m, x0.cuda(), x_t, t, clip_denoised=case.get("clip", True), model_kwargs=dict(H=H, W=W, D=D))
    assert np.allclose(vt["output"].cpu().numpy(), g["vt_output"], rtol=1e-3)
    assert np.abs(vt["pred_xstart"].cpu().numpy() - g["vt_pred_xstart"]).max() < 1e-3 * max(1.0, np.abs(g["vt_pred_xstart"]).max())
