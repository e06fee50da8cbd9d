"""The CPU oracle against the fixtures produced from the real reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import diffusion_ref as dr
from oracle import philox_ref
from oracle import unet_ref as ur
from oracle.cases import RESPACE_CASES, SAMPLER_CASES, TABLE_CASES, UNET_CASES, make_inputs, make_step_noise


def test_respace_bit_exact(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "respace.json")))
    assert len(want) == len(RESPACE_CASES)
    for T, spec in RESPACE_CASES:
        key = f"{T}|{spec if isinstance(spec, str) else ','.join(map(str, spec))}"
        try:
            got = dr.kept_steps(T, spec)
        except ValueError:
            got = "ValueError"
        assert got == want[key], key


def test_tables_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    for name, case in TABLE_CASES.items():
        o = dr.RefDiffusion(case["T"], case["respacing"], case.get("schedule", "linear"))
        assert np.array_equal(np.array(o.timestep_map), g[f"{name}/timestep_map"])
        for k, v in o.tab.items():
            assert v.dtype == np.float64
            assert np.array_equal(v, g[f"{name}/{k}"]), (name, k)


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_forward(golden_dir, name):
    case = UNET_CASES[name]
    g = np.load(os.path.join(golden_dir, f"unet_{name}.npz"))
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    x, t = make_inputs(case)
    assert np.array_equal(x.numpy(), g["x"]) and np.array_equal(t.numpy(), g["t"])
    H, W, D = case["HWD"]
    out = ur.unet_forward(sd, spec, x, t, H, W, D).numpy()
    # same torch build => bit-exact; leave head-room for a different CPU kernel selection
    assert np.abs(out - g["out"]).max() <= 2e-5 * max(1.0, np.abs(g["out"]).max())
    # dead corner is exactly zero
    assert np.all(out[..., H:, W:] == 0)


@pytest.mark.parametrize("name", list(SAMPLER_CASES))
def test_sampler(golden_dir, name):
    case = SAMPLER_CASES[name]
    g = np.load(os.path.join(golden_dir, f"sampler_{name}.npz"))["sample"]
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    H, W, D = case["HWD"]
    o = dr.RefDiffusion(case["T"], case["respacing"], case.get("schedule", "linear"),
                        case.get("mean_type", "start_x"), case.get("var_type", "fixed_large"),
                        case.get("rescale_timesteps", False))
    x_T, noises = make_step_noise(case, o.num_timesteps)
    kw = dict(clip=case.get("clip", True))
    if case["ddim"]:
        kw["eta"] = case.get("eta", 0.0)
        if case.get("mask"):
            gen = torch.Generator().manual_seed(77)
            kw["y0"] = torch.rand(x_T.shape, generator=gen) * 2 - 1
            kw["mask"] = (torch.rand(x_T.shape, generator=gen) > 0.5).float()
            kw["is_mask_t0"] = case.get("is_mask_t0", False)
    got = o.sample_loop(lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D), x_T, lambda i: noises[i],
                        ddim=case["ddim"], **kw).numpy()
    pg = [p.numpy() for p in ur.split_planes(torch.from_numpy(got), H, W, D)]
    pw = [p.numpy() for p in ur.split_planes(torch.from_numpy(g), H, W, D)]
    for a, b in zip(pg, pw):
        assert np.abs(a - b).max() <= 1e-4


def test_training_terms(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_terms.npz"))
    case = SAMPLER_CASES["ddpm20_small"]
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    H, W, D = case["HWD"]
    o = dr.RefDiffusion(case["T"], case["respacing"])
    x0, nz, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["t"])
    assert np.array_equal(o.q_sample(x0, t, nz).numpy(), g["q_sample"])
    terms = o.training_losses(lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D), x0, t, nz, H, W, D)
    for k in ("mse_xy", "mse_xz", "mse_yz", "loss"):
        assert np.allclose(terms[k].numpy(), g[k], rtol=1e-5, atol=1e-6)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox_ref.philox4x32_10(np.array([c], dtype=np.uint32), np.array(k, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_philox_normals_moments():
    z = philox_ref.normals(seed=12345, sample_idx=3, step=7, C=12, hw=20000)
    assert z.shape == (12, 20000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    z2 = philox_ref.normals(seed=12345, sample_idx=3, step=7, C=10, hw=20000)      # same quad count: same stream
    assert np.array_equal(z[:10], z2)
    z3 = philox_ref.normals(seed=12345, sample_idx=3, step=7, C=8, hw=20000)       # 2 quads per pixel instead of 3
    assert not np.array_equal(z[:8], z3)
    assert not np.array_equal(z, philox_ref.normals(seed=12345, sample_idx=4, step=7, C=12, hw=20000))


# ---------------------------------------------------------------------------- guided sampling hooks
def _cond_oracle(case):
    from oracle.cases import cond_fn, denoised_fn
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    H, W, D = case["HWD"]
    o = dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"), "fixed_large",
                        case.get("rescale_timesteps", False))
    x_T, noises = make_step_noise(case, o.num_timesteps)
    kw = dict(clip=True, cond_fn=cond_fn if case["cond"] else None, denoised_fn=denoised_fn if case["denoise"] else None)
    if case["ddim"]:
        kw["eta"] = case.get("eta", 0.0)
    return o, x_T, o.sample_loop(lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D), x_T, lambda i: noises[i],
                                 ddim=case["ddim"], **kw)


def test_cond_hooks_match_reference(golden_dir):
    """condition_mean / condition_score / denoised_fn loops and q_mean_variance (gaussian_diffusion.py:172-187, 357-394)."""
    from oracle.cases import COND_CASES
    for name, case in COND_CASES.items():
        g = np.load(os.path.join(golden_dir, f"cond_{name}.npz"))
        o, x_T, got = _cond_oracle(case)
        assert np.abs(got.numpy() - g["sample"]).max() <= 1e-4, name
        qm, qv, qlv = o.q_mean_variance(x_T, torch.from_numpy(g["t"]))
        assert np.array_equal(qm.numpy(), g["q_mean"])
        assert np.array_equal(qv[:, 0, 0, 0].numpy(), g["q_var"]) and np.array_equal(qlv[:, 0, 0, 0].numpy(), g["q_logvar"])


def test_full_chain_fixtures_are_wellformed(golden_dir):
    """The full-length BASELINE chains run on the GPU only (tests/test_gpu_full_chains.py); here: the committed fixtures have the
    shapes oracle.cases.pack_full promises and a clamped final latent (x_0 prediction of the last step is clipped to [-1, 1])."""
    from oracle.cases import FULL_CASES, FULL_STRIDE
    for name, case in FULL_CASES.items():
        g = np.load(os.path.join(golden_dir, f"full_{name}.npz"))
        H, W, D = case["HWD"]
        C, B = case["spec"]["in_channels"], case["B"]
        assert g["sample0"].shape == (C, H + D, W + D)
        assert np.abs(g["sample0"][:, :H, :W]).max() <= 1.0 + 1e-6
        if B > 1:
            n = C * (H + D) * (W + D)
            assert g["rest_strided"].shape == (B - 1, (n + FULL_STRIDE - 1) // FULL_STRIDE)
