"""-m gpu: UNet forward through the C ABI against the golden fixtures (real-reference outputs) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_ref as ur
from oracle.cases import UNET_CASES, make_inputs
from tests.gpu_util import make_cuda_model, nhwc, plane_errors

pytestmark = pytest.mark.gpu

TOL = 1e-3      # BASELINE.json north_star: "within 1e-3 rel-fp32"


def _run(case, precision, impl):
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    x, t = make_inputs(case)
    H, W, D = case["HWD"]
    m = make_cuda_model(spec, sd, precision, impl)
    with torch.no_grad():
        out = m(x.cuda(), t.cuda(), H=H, W=W, D=D)
    return m, sd, spec, x, t, out.cpu()


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("name", list(UNET_CASES))
def test_forward_matches_reference_golden(golden_dir, name, impl):
    case = UNET_CASES[name]
    g = np.load(os.path.join(golden_dir, f"unet_{name}.npz"))
    H, W, D = case["HWD"]
    m, sd, spec, x, t, out = _run(case, 3, impl)
    rel, mx = plane_errors(out, g["out"], H, W, D)
    # localise a failure: compare intermediate activations with the oracle's trace
    if not (rel < TOL and mx < TOL):
        trace = {}
        ur.unet_forward(sd, spec, x, t, H, W, D, trace=trace)
        acts = m.debug_activations()
        for k, v in acts.items():
            key = k[:-4] if k.endswith(".out") else k
            if key in trace:
                errs = [float((a - nhwc(b)).abs().max() / (b.abs().max() + 1e-9)) for a, b in zip(v, trace[key])]
                print(f"  layer {k}: max-rel per plane {errs}")
    assert rel < TOL and mx < TOL, (name, impl, rel, mx)
    assert torch.all(out[..., H:, W:] == 0)       # dead corner


@pytest.mark.parametrize("name", ["small_odd", "raw"])
def test_single_fp16_mode_is_close_but_coarser(golden_dir, name):
    case = UNET_CASES[name]
    g = np.load(os.path.join(golden_dir, f"unet_{name}.npz"))
    H, W, D = case["HWD"]
    _, _, _, _, _, out1 = _run(case, 1, "tc")
    _, _, _, _, _, out3 = _run(case, 3, "tc")
    r1, _ = plane_errors(out1, g["out"], H, W, D)
    r3, _ = plane_errors(out3, g["out"], H, W, D)
    assert r1 < 5e-3 and r3 < r1


def test_tc_equals_ffma_at_benchmark_shape():
    """cfg2 shape (C=12, 92x128x92): tcgen05 path vs CUDA-core path on identical operands, plus batch invariance
    and run-to-run bit reproducibility."""
    spec = ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12)
    sd = ur.synthetic_state_dict(spec, 77)
    H, W, D = 92, 128, 92
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 12, H + D, W + D, generator=g).cuda()
    t = torch.tensor([321.0]).cuda()
    mt, mf = make_cuda_model(spec, sd, 3, "tc"), make_cuda_model(spec, sd, 3, "ffma")
    with torch.no_grad():
        a, b = mt(x, t, H=H, W=W, D=D), mf(x, t, H=H, W=W, D=D)
        rel, mx = plane_errors(a, b, H, W, D)
        assert rel < 2e-5 and mx < 2e-5, (rel, mx)
        a2 = mt(x, t, H=H, W=W, D=D)
        assert torch.equal(a, a2)
        x2 = torch.cat([x, torch.randn(1, 12, H + D, W + D, generator=g).cuda()])
        t2 = torch.tensor([321.0, 5.0]).cuda()
        c = mt(x2, t2, H=H, W=W, D=D)
        assert torch.equal(c[:1], a)
    assert torch.all(a[..., H:, W:] == 0)


def test_reference_checkpoint_keys_roundtrip(tmp_path):
    """state_dict written by this class loads into it again through torch.save / load (sample.py:15-16)."""
    import sin3dm_b200 as s3
    m = s3.TriplaneUNetModelSmall(12, 64, 12, 1, 0, (1, 2), use_scale_shift_norm=True)
    p = tmp_path / "ema.pt"
    torch.save(m.state_dict(), p)
    m2 = s3.TriplaneUNetModelSmall(12, 64, 12, 1, 0, "1,2", use_scale_shift_norm=True)
    m2.load_state_dict(torch.load(p, map_location="cpu"))
    m2.cuda().eval()
    with torch.no_grad():
        out = m2(torch.randn(1, 12, 24, 24).cuda(), torch.tensor([3]).cuda(), H=16, W=16, D=8)
    assert torch.all(out == 0)      # zero-initialised out conv: a fresh model predicts exactly 0


@pytest.mark.parametrize("B", [8, 20])
def test_large_batch_several_roll_tiles_per_cta(B):
    """More rollout 1-D tiles than CTAs (batch >= 5 at the half-resolution level): a CTA then runs several roll tiles
    before its first conv tile, and the two producers of the A ring (epilogue warps / TMA warp) must stay in step.
    Sample b of the batch must equal the same sample run alone, bit for bit, and the oracle within tolerance."""
    spec = ur.UNetSpec(in_channels=8, model_channels=64, out_channels=8)
    sd = ur.synthetic_state_dict(spec, 91)
    H, W, D = 12, 16, 10
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 8, H + D, W + D, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    m = make_cuda_model(spec, sd, 3, "tc")
    with torch.no_grad():
        out = m(x.cuda(), t.cuda(), H=H, W=W, D=D)
        one = m(x[B - 1:].cuda(), t[B - 1:].cuda(), H=H, W=W, D=D)
    torch.cuda.synchronize()
    assert torch.equal(out[B - 1:], one)
    want = ur.unet_forward(sd, spec, x[:2], t[:2], H, W, D)
    rel, mx = plane_errors(out[:2].cpu(), want, H, W, D)
    assert rel < TOL and mx < TOL, (rel, mx)


@pytest.mark.parametrize("name,HWD,B", [("cfg2", (92, 128, 92), 1), ("cfg3_resized", (92, 128, 138), 2)])
def test_full_size_forward_matches_oracle(name, HWD, B):
    """BASELINE.json configs[1] / configs[2] shapes (C=12, default triplane; --resize 1 1 1.5 -> D=138): one UNet forward of the
    tcgen05 path against the CPU oracle at FULL size (the oracle is pinned to the real reference by the golden fixtures)."""
    spec = ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12)
    sd = ur.synthetic_state_dict(spec, 78)
    H, W, D = HWD
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 12, H + D, W + D, generator=g)
    t = torch.tensor([977, 12][:B])
    m = make_cuda_model(spec, sd, 3, "tc")
    with torch.no_grad():
        got = m(x.cuda(), t.cuda(), H=H, W=W, D=D).cpu()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    want = ur.unet_forward(sd, spec, x, t, H, W, D)
    rel, mx = plane_errors(got, want, H, W, D)
    print(name, "rel_l2", rel, "max", mx)
    assert rel < TOL and mx < TOL, (rel, mx)
    assert torch.all(got[..., H:, W:] == 0)


def test_full_size_ddim_loop_matches_oracle():
    """cfg3 shape, DDIM with 3 respaced steps, batch 2, through the CUDA-graph loop vs the oracle's loop (same x_T; eta = 0 so no
    step noise enters)."""
    from oracle import diffusion_ref as dr
    from sin3dm_b200.script_util import create_gaussian_diffusion
    spec = ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12)
    sd = ur.synthetic_state_dict(spec, 79)
    H, W, D = 92, 128, 138
    g = torch.Generator().manual_seed(4)
    x_T = torch.randn(2, 12, H + D, W + D, generator=g)
    m = make_cuda_model(spec, sd, 3, "tc")
    d = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="3")
    with torch.no_grad():
        got = d.ddim_sample_loop(m, list(x_T.shape), noise=x_T, model_kwargs=dict(H=H, W=W, D=D)).cpu()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    o = dr.RefDiffusion(1000, "3")
    want = o.sample_loop(lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D), x_T, lambda i: torch.zeros_like(x_T), ddim=True)
    rel, mx = plane_errors(got, want, H, W, D)
    print("cfg3 ddim-3 rel_l2", rel, "max", mx)
    assert rel < TOL and mx < TOL, (rel, mx)


@pytest.mark.parametrize("name", ["small_odd", "three_level"])
def test_fused_pool_equals_standalone_pool(name, monkeypatch):
    """Downsample2x computed in the conv epilogue (default) vs the stand-alone k_avgpool2 launch (S3D_FUSE_POOL=0): same values up
    to the order of the four additions, odd plane sizes (floor pooling) included."""
    case = UNET_CASES[name]
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    x, t = make_inputs(case)
    H, W, D = case["HWD"]
    from sin3dm_b200 import _lib
    outs, launches = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("S3D_FUSE_POOL", flag)          # read when the handle is created
        m = make_cuda_model(spec, sd, 3, "tc")
        with torch.no_grad():
            outs.append(m(x.cuda(), t.cuda(), H=H, W=W, D=D).cpu())
        launches.append(_lib.lib().s3d_unet_last_launches(m.handle()))
    rel, mx = plane_errors(outs[0], outs[1], H, W, D)
    assert rel < 2e-5 and mx < 2e-5, (rel, mx)
    assert launches[1] > launches[0], launches          # one k_avgpool2 launch per Downsample2x comes back
