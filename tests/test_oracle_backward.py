"""Groundwork for the training row: the folded-form backward of the rollout TriplaneConv (oracle/backward_ref.py) against
torch.autograd through the oracle forward (which is pinned to the real reference)."""
import pytest
import torch

from oracle import backward_ref as br
from oracle import unet_ref as ur


@pytest.mark.parametrize("C,Cout,HWD,B", [(4, 6, (7, 9, 5), 2), (3, 3, (1, 4, 2), 1), (5, 2, (6, 1, 3), 3)])
def test_folded_backward_equals_autograd(C, Cout, HWD, B):
    g = torch.Generator().manual_seed(C * 100 + Cout)
    H, W, D = HWD
    sd = {}
    for n in br.PLANES:
        sd[f"c.conv_{n}.weight"] = torch.randn(Cout, 3 * C, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
        sd[f"c.conv_{n}.bias"] = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True)
    planes = [torch.randn(B, C, r, c, generator=g, dtype=torch.float64, requires_grad=True) for r, c in ((H, W), (H, D), (W, D))]
    outs = ur.tri_conv(sd, "c", planes, 1, True)
    dys = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in outs]
    loss = sum((o * d).sum() for o, d in zip(outs, dys))
    loss.backward()
    dplanes, grads = br.tri_conv_backward_folded({k: v.detach() for k, v in sd.items()}, "c", [p.detach() for p in planes], dys)
    for p, dp in zip(planes, dplanes):
        assert torch.allclose(p.grad, dp, rtol=1e-10, atol=1e-10)
    for k, v in sd.items():
        assert torch.allclose(v.grad, grads[k], rtol=1e-10, atol=1e-10), k


@pytest.mark.parametrize("film", [False, True])
def test_gn_film_silu_backward_equals_autograd(film):
    g = torch.Generator().manual_seed(7)
    B, C, R, Cc = 2, 64, 5, 7
    x = torch.randn(B, C, R, Cc, generator=g, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(C, generator=g, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(C, generator=g, dtype=torch.float64, requires_grad=True)
    scale = torch.randn(B, C, generator=g, dtype=torch.float64, requires_grad=True) if film else None
    shift = torch.randn(B, C, generator=g, dtype=torch.float64, requires_grad=True) if film else None
    n = torch.nn.functional.group_norm(x, 32, gamma, beta, 1e-5)
    f = n * (1 + scale.view(B, C, 1, 1)) + shift.view(B, C, 1, 1) if film else n
    y = ur.silu(f)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * dy).sum().backward()
    o = br.gn_film_silu_backward(x.detach(), gamma.detach(), beta.detach(), dy, scale=scale.detach() if film else None,
                                 shift=shift.detach() if film else None)
    assert torch.allclose(x.grad, o["dx"], rtol=1e-9, atol=1e-10)
    assert torch.allclose(gamma.grad, o["dgamma"], rtol=1e-9, atol=1e-10) and torch.allclose(beta.grad, o["dbeta"], rtol=1e-9, atol=1e-10)
    if film:
        assert torch.allclose(scale.grad, o["dscale"], rtol=1e-9, atol=1e-10) and torch.allclose(shift.grad, o["dshift"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("rows,cols", [(6, 8), (7, 5), (1, 3)])
def test_resampling_adjoints_equal_autograd(rows, cols):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(rows * 10 + cols)
    x = torch.randn(2, 3, rows, cols, generator=g, dtype=torch.float64, requires_grad=True)
    # avg-pool 2x2 (floor)
    if rows >= 2 and cols >= 2:
        y = F.avg_pool2d(x, 2)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (y * dy).sum().backward()
        assert torch.allclose(x.grad, br.avgpool2_backward(dy, rows, cols), atol=1e-12)
        x.grad = None
    # bilinear x2
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * dy).sum().backward()
    assert torch.allclose(x.grad, br.bilinear_resize_backward(dy, rows, cols, scale_factor=2), atol=1e-12)
    x.grad = None
    # bilinear to an explicit (odd) size: the resize-to-skip step
    y = F.interpolate(x, size=(2 * rows + 1, 2 * cols + 1), mode="bilinear", align_corners=False)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * dy).sum().backward()
    assert torch.allclose(x.grad, br.bilinear_resize_backward(dy, rows, cols), atol=1e-12)


@pytest.mark.parametrize("name", ["startx", "eps_odd"])
def test_manual_unet_backward_matches_reference_gradients(golden_dir, name):
    """The whole UNet backward assembled from the kernel-form adjoints (no autograd) against the gradient fixtures the real
    reference produced (oracle/make_golden_grads.py)."""
    import os
    import numpy as np
    from oracle import diffusion_ref as dr
    from oracle.cases import GRAD_CASES, make_grad_inputs
    from oracle.make_golden_grads import sample_idx
    case = GRAD_CASES[name]
    g = np.load(os.path.join(golden_dir, f"grads_{name}.npz"))
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    H, W, D = case["HWD"]
    x0, nz, t = make_grad_inputs(case)
    o = dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"))
    x_t = o.q_sample(x0, t, nz)
    target = x0 if case.get("mean_type", "start_x") == "start_x" else nz
    tgt_planes = ur.split_planes(target, H, W, D)
    B = x0.shape[0]

    def dout(out_planes):          # loss = mean_b sum_planes mean_plane (target - out)^2   (gaussian_diffusion.py:822-851)
        return [2 * (op - tp) / (op[0].numel() * B) for op, tp in zip(out_planes, tgt_planes)]
    with torch.no_grad():
        _, grads = br.unet_param_grads(sd, spec, x_t, o.model_t(t), H, W, D, dout)
    assert len(grads) == 138
    for k, gr in grads.items():
        flat = gr.reshape(-1).numpy()
        n_ref = float(g[f"norm/{k}"])
        assert abs(np.linalg.norm(flat.astype(np.float64)) - n_ref) <= 2e-4 * max(n_ref, 1e-12), (k, np.linalg.norm(flat), n_ref)
        want = g[f"sample/{k}"]
        tol = 2e-4 * max(np.abs(want).max(), n_ref / np.sqrt(flat.size), 1e-12)
        assert np.abs(flat[sample_idx(flat.size)] - want).max() <= tol, k
