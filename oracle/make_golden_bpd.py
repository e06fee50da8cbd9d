"""Generate tests/golden/bpd_*.npz from the REAL reference (authoring container only).

    python -m oracle.make_golden_bpd        # needs /root/reference/src (read-only import)

Runs the unmodified ``SpacedDiffusion.calc_bpd_loop`` / ``_vb_terms_bpd`` / ``_prior_bpd``
(src/diffusion/gaussian_diffusion.py:736-931) on the reference UNet with synthetic weights and replayed CPU noise, asserts the
oracle restatement (oracle/diffusion_ref.py) against it and freezes the outputs.
"""
import os

import numpy as np
import torch

from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import BPD_CASES, make_bpd_inputs
from oracle.make_golden import OUT, build_ref_diffusion, build_ref_unet, ref_modules


def main():
    gd, respace, ut = ref_modules()
    torch.set_num_threads(os.cpu_count())
    for name, case in BPD_CASES.items():
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd)
        d = build_ref_diffusion(gd, respace, case)
        H, W, D = case["HWD"]
        x0, noises = make_bpd_inputs(case, d.num_timesteps)
        clip = case.get("clip", True)
        it = iter(range(d.num_timesteps - 1, -1, -1))
        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: noises[next(it)]
        try:
            with torch.no_grad():
                want = d.calc_bpd_loop(m, x0, clip_denoised=clip, model_kwargs=dict(H=H, W=W, D=D))
        finally:
            torch.randn_like = orig
        o = dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"),
                            case.get("var_type", "fixed_large"))
        model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
        got = o.calc_bpd_loop(model, x0, lambda i: noises[i], clip=clip)
        for k in ("total_bpd", "prior_bpd", "vb", "xstart_mse", "mse"):
            assert torch.allclose(want[k], got[k], rtol=2e-5, atol=1e-6), (name, k, want[k], got[k])
        # one stand-alone _vb_terms_bpd call at t = 0 and at the last step (mixed batch when B > 1)
        B = x0.shape[0]
        t = torch.tensor(([0, d.num_timesteps - 1] * B)[:B])
        x_t = d.q_sample(x0, t, noise=noises[0])
        with torch.no_grad():
            vt = d._vb_terms_bpd(m, x0, x_t, t, clip_denoised=clip, model_kwargs=dict(H=H, W=W, D=D))
        ov = o.vb_terms_bpd(model, x0, x_t, t, clip)
        assert torch.allclose(vt["output"], ov["output"], rtol=2e-5, atol=1e-6)
        assert torch.allclose(vt["pred_xstart"], ov["pred_xstart"], rtol=1e-5, atol=1e-5)
        np.savez_compressed(os.path.join(OUT, f"bpd_{name}.npz"), vt_t=t.numpy(), vt_output=vt["output"].numpy(),
                            vt_pred_xstart=vt["pred_xstart"].numpy(), **{k: v.numpy() for k, v in want.items()})
        print("bpd", name, {k: want[k].flatten()[:3].tolist() for k in ("total_bpd", "prior_bpd", "vb")})


if __name__ == "__main__":
    main()
