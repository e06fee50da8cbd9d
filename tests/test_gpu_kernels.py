"""-m gpu: scheduler / RNG kernels against the oracle (bit-exact where the arithmetic allows)."""
import numpy as np
import pytest
import torch

import sin3dm_b200 as s3
from oracle import diffusion_ref as dr
from oracle import philox_ref
from sin3dm_b200 import _lib
from sin3dm_b200.script_util import create_gaussian_diffusion

pytestmark = pytest.mark.gpu


def _mk(mean_type="start_x", var_type="fixed_large", respacing="20", T=1000):
    d = create_gaussian_diffusion(steps=T, predict_xstart=(mean_type == "start_x"), sigma_small=(var_type == "fixed_small"),
                                  timestep_respacing=respacing)
    o = dr.RefDiffusion(T, respacing, "linear", mean_type, var_type)
    return d, o


@pytest.mark.parametrize("mean_type", ["start_x", "epsilon"])
@pytest.mark.parametrize("var_type", ["fixed_large", "fixed_small"])
@pytest.mark.parametrize("clip", [True, False])
def test_ddpm_step_bit_exact(mean_type, var_type, clip):
    d, o = _mk(mean_type, var_type)
    g = torch.Generator().manual_seed(3)
    shape = (3, 5, 9, 7)                     # n not a multiple of 4: exercises the scalar tail
    x, out, nz = (torch.randn(shape, generator=g) for _ in range(3))
    t = torch.tensor([0, 7, 19])
    want = o.p_sample(lambda xx, tt: out, x, t, nz, clip=clip)
    got = d.p_sample(lambda xx, tt, **k: out.cuda(), x.cuda(), t.cuda(), clip_denoised=clip, noise=nz.cuda())
    assert torch.equal(got["sample"].cpu(), want["sample"])
    assert torch.equal(got["pred_xstart"].cpu(), want["pred_xstart"])


@pytest.mark.parametrize("mean_type", ["start_x", "epsilon"])
@pytest.mark.parametrize("eta", [0.0, 0.6])
@pytest.mark.parametrize("maskmode", [None, "t0", "nz"])
def test_ddim_step_bit_exact(mean_type, eta, maskmode):
    d, o = _mk(mean_type)
    g = torch.Generator().manual_seed(4)
    shape = (3, 4, 8, 8)
    x, out, nz = (torch.randn(shape, generator=g) for _ in range(3))
    t = torch.tensor([0, 11, 19])
    kw = {}
    if maskmode:
        kw = dict(y0=torch.rand(shape, generator=g) * 2 - 1, mask=(torch.rand(shape, generator=g) > 0.5).float(),
                  is_mask_t0=(maskmode == "t0"))
    want = o.ddim_sample(lambda xx, tt: out, x, t, nz, eta=eta, **kw)
    ckw = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
    got = d.ddim_sample(lambda xx, tt, **k: out.cuda(), x.cuda(), t.cuda(), eta=eta, noise=nz.cuda(), **ckw)
    assert torch.equal(got["sample"].cpu(), want["sample"])
    assert torch.equal(got["pred_xstart"].cpu(), want["pred_xstart"])


def test_ddim_reverse_and_q_sample_bit_exact():
    d, o = _mk()
    g = torch.Generator().manual_seed(5)
    shape = (2, 4, 8, 8)
    x, out, nz = (torch.randn(shape, generator=g) for _ in range(3))
    t = torch.tensor([2, 19])
    want = o.ddim_reverse_sample(lambda xx, tt: out, x, t)
    got = d.ddim_reverse_sample(lambda xx, tt, **k: out.cuda(), x.cuda(), t.cuda())
    assert torch.equal(got["sample"].cpu(), want["sample"])
    assert torch.equal(d.q_sample(x.cuda(), t.cuda(), nz.cuda()).cpu(), o.q_sample(x, t, nz))


def test_philox_matches_oracle():
    import ctypes as C
    Cc, hw, B = 10, 4099, 3
    out = torch.empty(B, Cc, hw, device="cuda")
    _lib.check(_lib.lib().s3d_philox_normal(C.c_void_p(out.data_ptr()), B, Cc, hw, 0x1234567890ABCDEF, 5, 17,
                                            _lib.current_stream_ptr()))
    got = out.cpu().numpy()
    for b in range(B):
        want = philox_ref.normals(0x1234567890ABCDEF, 5 + b, 17, Cc, hw)
        assert np.allclose(got[b], want, rtol=0, atol=2e-6), np.abs(got[b] - want).max()
    assert abs(got.mean()) < 0.05 and abs(got.std() - 1) < 0.05


def test_in_kernel_noise_equals_philox_fill():
    """p_sample with noise=None must use exactly the generator s3d_philox_normal exposes."""
    import ctypes as C
    d, _ = _mk()
    shape = (2, 4, 6, 10)
    g = torch.Generator().manual_seed(6)
    x, out = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    t = torch.tensor([5, 5]).cuda()
    a = d.p_sample(lambda xx, tt, **k: out, x, t, _seed=99, _sample_base=3)
    nz = torch.empty(shape, device="cuda")
    _lib.check(_lib.lib().s3d_philox_normal(C.c_void_p(nz.data_ptr()), 2, shape[1], shape[2] * shape[3], 99, 3, 5,
                                            _lib.current_stream_ptr()))
    b = d.p_sample(lambda xx, tt, **k: out, x, t, noise=nz)
    assert torch.equal(a["sample"], b["sample"])


def test_plane_mse_matches_decomposed_mean():
    """k_plane_mse vs mean_flat((target - output)^2) over decompose_featmaps' three planes (gaussian_diffusion.py:822-851,
    triplane_util.py:19-25), ragged sizes, batch 3; the D x D corner must not count."""
    import ctypes as C
    g = torch.Generator().manual_seed(11)
    B, Cc, H, W, D = 3, 5, 13, 18, 7
    tgt = torch.randn(B, Cc, H + D, W + D, generator=g)
    out = torch.randn(B, Cc, H + D, W + D, generator=g)
    out[..., H:, W:] += 100.0                      # garbage in the unused corner
    a, b = tgt.cuda(), out.cuda()
    L = _lib.lib()
    ws = torch.empty(L.s3d_vb_workspace_bytes(B, tgt[0].numel()), dtype=torch.uint8, device="cuda")
    mse = torch.empty(B, 3, device="cuda")
    _lib.check(L.s3d_plane_mse(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), B, Cc, H, W, D, C.c_void_p(ws.data_ptr()),
                               C.c_void_p(mse.data_ptr()), _lib.current_stream_ptr()))
    want = torch.stack([((tp - op) ** 2).mean(dim=(1, 2, 3)) for tp, op in zip(s3.decompose_featmaps(tgt, (H, W, D)),
                                                                                 s3.decompose_featmaps(out, (H, W, D)))], dim=1)
    assert torch.allclose(mse.cpu(), want, rtol=2e-6, atol=0), (mse.cpu(), want)
