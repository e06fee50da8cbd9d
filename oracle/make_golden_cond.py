"""Fixtures for the guided-sampling hooks of GaussianDiffusion from the REAL reference (authoring container only).

    python -m oracle.make_golden_cond

cond_fn -> condition_mean (DDPM, gaussian_diffusion.py:357-370) / condition_score (DDIM, :372-394), denoised_fn (:294-315) and
q_mean_variance (:172-187): the unmodified reference loops run with oracle.cases.cond_fn / denoised_fn, the oracle restatement is
asserted against them on the spot, and the final samples are frozen as tests/golden/cond_*.npz."""
import os

import numpy as np
import torch

from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import COND_CASES, cond_fn, denoised_fn, make_step_noise
from oracle.make_golden import OUT, build_ref_diffusion, build_ref_unet, ref_modules


def main():
    gd, respace, ut = ref_modules()
    for name, case in COND_CASES.items():
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd)
        d = build_ref_diffusion(gd, respace, case)
        H, W, D = case["HWD"]
        x_T, noises = make_step_noise(case, d.num_timesteps)
        shape = list(x_T.shape)
        kw = dict(model_kwargs=dict(H=H, W=W, D=D), noise=x_T, clip_denoised=True,
                  cond_fn=cond_fn if case["cond"] else None, denoised_fn=denoised_fn if case["denoise"] else None)
        it = iter(range(d.num_timesteps - 1, -1, -1))
        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: noises[next(it)]
        try:
            with torch.no_grad():
                want = d.ddim_sample_loop(m, shape, eta=case.get("eta", 0.0), **kw) if case["ddim"] else d.p_sample_loop(m, shape, **kw)
        finally:
            torch.randn_like = orig
        o = dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"), "fixed_large",
                            case.get("rescale_timesteps", False))
        model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
        okw = dict(clip=True, cond_fn=cond_fn if case["cond"] else None, denoised_fn=denoised_fn if case["denoise"] else None)
        if case["ddim"]:
            okw["eta"] = case.get("eta", 0.0)
        got = o.sample_loop(model, x_T, lambda i: noises[i], ddim=case["ddim"], **okw)
        err = (want - got).abs().max().item()
        assert err <= 1e-5 * max(1.0, want.abs().max().item()), (name, err)
        # q_mean_variance at a few timesteps
        t = torch.tensor([0, d.num_timesteps - 1][: case["B"]])
        qm, qv, qlv = d.q_mean_variance(x_T, t)
        om, ov, olv = o.q_mean_variance(x_T, t)
        assert torch.equal(qm, om) and torch.equal(qv.expand_as(x_T), ov) and torch.equal(qlv.expand_as(x_T), olv), name
        np.savez_compressed(os.path.join(OUT, f"cond_{name}.npz"), sample=want.numpy(), q_mean=qm.numpy(),
                            q_var=qv.expand_as(x_T)[:, 0, 0, 0].numpy(), q_logvar=qlv.expand_as(x_T)[:, 0, 0, 0].numpy(), t=t.numpy())
        print("cond", name, "oracle-vs-ref max abs", err, "absmax", want.abs().max().item())


if __name__ == "__main__":
    main()
