"""Generate tests/golden/*.npz|json from the REAL reference (authoring container only).

    python -m oracle.make_golden            # needs /root/reference/src (read-only import)

Every fixture is produced by the unmodified reference classes
(``diffusion.unet_triplane.TriplaneUNetModelSmall[Raw]``, ``diffusion.respace.SpacedDiffusion``,
``diffusion.respace.space_timesteps``) fed with ``oracle.unet_ref.synthetic_state_dict`` weights and
CPU-generated noise, and the oracle restatement is asserted against it on the spot.  The GPU box has no
/root/reference; there ``tests/`` check oracle and CUDA path against these committed files.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import UNET_CASES, SAMPLER_CASES, RESPACE_CASES, TABLE_CASES, make_inputs, make_step_noise


def ref_modules():
    sys.path.insert(0, REF)
    from diffusion import gaussian_diffusion as gd
    from diffusion import respace, unet_triplane
    return gd, respace, unet_triplane


def build_ref_unet(ut, spec: ur.UNetSpec, sd):
    cls = ut.TriplaneUNetModelSmall if spec.rollout else ut.TriplaneUNetModelSmallRaw
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(in_channels=spec.in_channels, model_channels=spec.model_channels, out_channels=spec.out_channels,
                num_res_blocks=spec.num_res_blocks, dropout=0, channel_mult=tuple(spec.channel_mult),
                use_scale_shift_norm=spec.use_scale_shift_norm)
    assert [k for k, _ in ur.param_shapes(spec)] == list(m.state_dict().keys()), "state_dict key order differs"
    m.load_state_dict(sd)
    return m.eval()


def build_ref_diffusion(gd, respace, case):
    betas = gd.get_named_beta_schedule(case.get("schedule", "linear"), case["T"])
    resp = case["respacing"] if case["respacing"] else [case["T"]]
    return respace.SpacedDiffusion(
        use_timesteps=respace.space_timesteps(case["T"], resp), betas=betas,
        model_mean_type=gd.ModelMeanType.START_X if case.get("mean_type", "start_x") == "start_x" else gd.ModelMeanType.EPSILON,
        model_var_type=gd.ModelVarType.FIXED_LARGE if case.get("var_type", "fixed_large") == "fixed_large" else gd.ModelVarType.FIXED_SMALL,
        loss_type=gd.LossType.MSE, rescale_timesteps=case.get("rescale_timesteps", False))


def main():
    os.makedirs(OUT, exist_ok=True)
    gd, respace, ut = ref_modules()
    torch.set_num_threads(os.cpu_count())

    # ---- 1. respacing index sets (bit-exact integer path)
    resp = {}
    for T, spec in RESPACE_CASES:
        key = f"{T}|{spec if isinstance(spec, str) else ','.join(map(str, spec))}"
        try:
            want = sorted(respace.space_timesteps(T, spec))
        except ValueError as e:
            want = "ValueError"
        try:
            got = dr.kept_steps(T, spec)
        except ValueError:
            got = "ValueError"
        assert want == got, (key, want, got)
        resp[key] = want
    json.dump(resp, open(os.path.join(OUT, "respace.json"), "w"))
    print("respace cases", len(resp))

    # ---- 2. fp64 coefficient tables (bit-exact)
    tabs = {}
    for name, case in TABLE_CASES.items():
        d = build_ref_diffusion(gd, respace, case)
        o = dr.RefDiffusion(case["T"], case["respacing"], case.get("schedule", "linear"))
        assert d.timestep_map == o.timestep_map
        for k in dr.tables(np.ones(2) * 0.5).keys():
            a, b = getattr(d, k), o.tab[k]
            assert a.dtype == np.float64 and np.array_equal(a, b), (name, k)
            tabs[f"{name}/{k}"] = a
        tabs[f"{name}/timestep_map"] = np.array(d.timestep_map, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tabs)
    print("table cases", len(TABLE_CASES))

    # ---- 3. UNet forward
    for name, case in UNET_CASES.items():
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd)
        x, t = make_inputs(case)
        H, W, D = case["HWD"]
        with torch.no_grad():
            want = m(x, t, H=H, W=W, D=D)
        trace = {}
        got = ur.unet_forward(sd, spec, x, t, H, W, D, trace=trace)
        err = (want - got).abs().max().item()
        assert err <= 2e-5 * max(1.0, want.abs().max().item()), (name, err)
        save = dict(x=x.numpy(), t=t.numpy(), out=want.numpy())
        # a few intermediate activations from the oracle (already tied to the reference through `out`)
        for k, v in trace.items():
            if k == "emb":
                save["trace/emb"] = v.numpy()
            elif k in ("in_conv", "input_blocks.0.0"):
                for n, a in zip(ur.PLANES, v):
                    save[f"trace/{k}/{n}"] = a.numpy()
        np.savez_compressed(os.path.join(OUT, f"unet_{name}.npz"), **save)
        print("unet", name, "oracle-vs-ref max abs", err, "out absmax", want.abs().max().item())

    # ---- 4. sampler loops
    for name, case in SAMPLER_CASES.items():
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd)
        d = build_ref_diffusion(gd, respace, case)
        H, W, D = case["HWD"]
        B, C = case["B"], spec.in_channels
        shape = [B, C, H + D, W + D]
        x_T, noises = make_step_noise(case, d.num_timesteps)
        kw = dict(model_kwargs=dict(H=H, W=W, D=D), noise=x_T, clip_denoised=case.get("clip", True))
        extra = {}
        if case.get("mask"):
            g = torch.Generator().manual_seed(77)
            extra["y0"] = torch.rand(shape, generator=g) * 2 - 1
            extra["mask"] = (torch.rand(shape, generator=g) > 0.5).float()
            extra["is_mask_t0"] = case.get("is_mask_t0", False)
        # replay per-step noise into the reference by patching randn_like
        it = iter(range(d.num_timesteps - 1, -1, -1))
        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: noises[next(it)]
        try:
            with torch.no_grad():
                if case["ddim"]:
                    want = d.ddim_sample_loop(m, shape, eta=case.get("eta", 0.0), **extra, **kw)
                else:
                    want = d.p_sample_loop(m, shape, **kw)
        finally:
            torch.randn_like = orig
        o = dr.RefDiffusion(case["T"], case["respacing"], case.get("schedule", "linear"),
                            case.get("mean_type", "start_x"), case.get("var_type", "fixed_large"),
                            case.get("rescale_timesteps", False))
        model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
        okw = dict(clip=case.get("clip", True))
        if case["ddim"]:
            okw.update(eta=case.get("eta", 0.0), **extra)
        got = o.sample_loop(model, x_T, lambda i: noises[i], ddim=case["ddim"], **okw)
        pw, pg = ur.split_planes(want, H, W, D), ur.split_planes(got, H, W, D)
        err = max((a - b).abs().max().item() for a, b in zip(pw, pg))
        ref = max(a.abs().max().item() for a in pw)
        assert err <= 1e-4 * max(ref, 1.0), (name, err, ref)
        np.savez_compressed(os.path.join(OUT, f"sampler_{name}.npz"), sample=want.numpy())
        print("sampler", name, "oracle-vs-ref max abs", err, "absmax", ref)

    # ---- 5. q_sample / training_losses (forward values only)
    case = SAMPLER_CASES["ddpm20_small"]
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    m = build_ref_unet(ut, spec, sd)
    d = build_ref_diffusion(gd, respace, case)
    H, W, D = case["HWD"]
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(case["B"], spec.in_channels, H + D, W + D, generator=g) * 2 - 1
    nz = torch.randn(x0.shape, generator=g)
    t = torch.tensor([0, d.num_timesteps - 1][: case["B"]])
    with torch.no_grad():
        terms = d.training_losses(m, x0, t, model_kwargs=dict(H=H, W=W, D=D), noise=nz)
        qs = d.q_sample(x0, t, noise=nz)
    o = dr.RefDiffusion(case["T"], case["respacing"])
    oterms = o.training_losses(lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D), x0, t, nz, H, W, D)
    for k in ("mse_xy", "mse_xz", "mse_yz", "loss"):
        assert torch.allclose(terms[k], oterms[k], rtol=1e-5, atol=1e-6), k
    assert torch.equal(qs, o.q_sample(x0, t, nz))
    np.savez_compressed(os.path.join(OUT, "train_terms.npz"), x0=x0.numpy(), noise=nz.numpy(), t=t.numpy(),
                        q_sample=qs.numpy(), **{k: v.numpy() for k, v in terms.items()})
    print("training_losses ok", {k: v.tolist() for k, v in terms.items()})


if __name__ == "__main__":
    main()
