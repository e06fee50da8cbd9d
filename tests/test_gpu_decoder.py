"""-m gpu: the triplane decoder through the C ABI against the golden fixtures (real-reference outputs) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_ref as de
from oracle.cases import DECODER_CASES, make_decoder_inputs
from sin3dm_b200.encoding import AutoEncoderGroupPBR, AutoEncoderGroupSkip, AutoEncoderGroupV3, TriplaneDecoder

pytestmark = pytest.mark.gpu

TOL = 1e-3          # BASELINE.json north_star: "within 1e-3 rel-fp32"
TOL_SPLIT = 2e-5    # what the fp16 hi/lo split and the fp32 CUDA-core kernel actually deliver


def make_net(spec, sd, precision=3, impl="tc"):
    cls = AutoEncoderGroupPBR if spec.net_kind == "pbr" else (AutoEncoderGroupV3 if spec.mlp_kind == "base" else AutoEncoderGroupSkip)
    net = cls(spec.geo_feat_channels, spec.tex_feat_channels, spec.feat_channel_up,
                               spec.mlp_hidden_channels, spec.mlp_hidden_layers, use_tex=spec.use_tex,
                               tex_channels=spec.tex_channels)
    net.load_state_dict(sd)
    net.s3d_precision = precision
    net.s3d_mlp_impl = 1 if impl == "ffma" else 0
    return net.cuda().eval()


def errors(got, want):
    g, w = torch.as_tensor(got).double().cpu(), torch.as_tensor(want).double()
    return float((g - w).norm() / w.norm()), float((g - w).abs().max() / w.abs().max())


def run_case(name, precision, impl):
    case = DECODER_CASES[name]
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    maps, pts, aabb = make_decoder_inputs(case)
    net = make_net(spec, sd, precision, impl)
    dec = TriplaneDecoder(net)
    cmaps = [m.cuda() for m in maps]
    if "grid" in case:
        out = dec.decode_grid(cmaps, case["grid"], aabb=aabb)
        shape = tuple(out.shape[:3])
        out = out.reshape(-1, spec.out_channels)
    else:
        out = dec.decode_batch(cmaps, pts.cuda(), aabb=aabb)
        shape = None
    return net, spec, out.cpu(), shape


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("name", list(DECODER_CASES))
def test_decode_matches_reference_golden(golden_dir, name, impl):
    g = np.load(os.path.join(golden_dir, f"decoder_{name}.npz"))
    net, spec, out, shape = run_case(name, 3, impl)
    if shape is not None:
        assert list(shape) == list(g["grid_shape"])
    # the up-convolved planes first: localises a failure
    planes = net.feature_planes()
    for bi, br in enumerate(["geo", "tex"] if spec.use_tex else ["geo"]):
        for pi, pl in enumerate(de.PLANES):
            want = torch.from_numpy(g[f"planes/{br}/{pl}"])[0].permute(1, 2, 0)
            rel, mx = errors(planes[pi][..., bi * 64:(bi + 1) * 64], want)
            assert rel < TOL_SPLIT and mx < 1e-4, (name, br, pl, rel, mx)
    rel, mx = errors(out, g["out"])
    print(f"decoder {name} {impl}: rel_l2={rel:.2e} max={mx:.2e}")
    assert rel < TOL_SPLIT and mx < TOL_SPLIT * 5, (name, impl, rel, mx)
    tex = out[:, 1:].numpy()
    assert tex.min(initial=0.0) >= 0.0 and tex.max(initial=0.0) <= 1.0


@pytest.mark.parametrize("name", ["default", "sdf_only"])
def test_single_fp16_mode_is_close_but_coarser(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"decoder_{name}.npz"))
    _, _, out1, _ = run_case(name, 1, "tc")
    _, _, out3, _ = run_case(name, 3, "tc")
    r1, _ = errors(out1, g["out"])
    r3, _ = errors(out3, g["out"])
    print(f"decoder {name}: single fp16 rel_l2={r1:.2e}, split rel_l2={r3:.2e}")
    assert r1 < 5e-3 and r3 < r1


def test_benchmark_shape_tc_equals_ffma_and_is_tiling_invariant():
    """cfg2 latent (92,128,92), decode_grid at resolution 96 (~480k points, many tiles per CTA, ragged last tile):
    tensor-core kernel vs the fp32 CUDA-core kernel; a point's value does not depend on which tile it lands in; a
    second run is bit-identical."""
    spec = de.DecoderSpec()
    sd = de.synthetic_state_dict(spec, 7)
    H, W, D = 92, 128, 92
    g = torch.Generator().manual_seed(3)
    maps = [torch.tanh(torch.randn(1, 12, a, b, generator=g)).cuda() for a, b in ((H, W), (H, D), (W, D))]
    aabb = torch.tensor([-0.72, -1.0, -0.72, 0.72, 1.0, 0.72])
    nt, nf = make_net(spec, sd, 3, "tc"), make_net(spec, sd, 3, "ffma")
    a = TriplaneDecoder(nt).decode_grid(maps, 96, aabb=aabb)
    b = TriplaneDecoder(nf).decode_grid(maps, 96, aabb=aabb)
    assert a.shape == b.shape and a.shape[-1] == 4 and a.numel() // 4 > 400_000
    rel, mx = errors(a, b.cpu())
    assert rel < TOL_SPLIT and mx < TOL_SPLIT * 5, (rel, mx)
    assert torch.equal(a, TriplaneDecoder(nt).decode_grid(maps, 96, aabb=aabb))
    # same points, explicit list, shifted by 37 so every point sits in a different tile row
    pts = de.grid_points(aabb, 96).view(-1, 3)
    sub = pts[37:37 + 5000].cuda()
    c = TriplaneDecoder(nt).decode_batch(maps, sub, aabb=aabb)
    assert torch.equal(c, a.view(-1, 4)[37:37 + 5000])
    # oracle on a slice (CPU, seconds)
    want = de.decode_batch(sd, spec, [m.cpu() for m in maps], pts[:4096], aabb=aabb)
    rel, mx = errors(a.view(-1, 4)[:4096], want)
    assert rel < TOL_SPLIT and mx < TOL_SPLIT * 5, (rel, mx)


def test_empty_and_single_point():
    spec = de.DecoderSpec()
    net = make_net(spec, de.synthetic_state_dict(spec, 9))
    maps = [torch.zeros(1, 12, 8, 8).cuda()] * 3
    assert net.decode(torch.zeros(0, 3).cuda(), maps).shape == (0, 4)
    one = net.decode(torch.zeros(1, 3).cuda(), maps)
    assert one.shape == (1, 4) and torch.isfinite(one).all()


# ------------------------------------------------------------------------------------------------ encoder half
def test_pbr_unclamped_heads_match_oracle():
    """AutoEncoderGroupPBR.decode (networks.py:296-331) without decode_batch's clamp: the rgb / mr / normal heads are raw
    (no sigmoid), so the unclamped values are what pins them; both MLP kernels."""
    case = DECODER_CASES["pbr"]
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    maps, pts, aabb = make_decoder_inputs(case)
    want = de.decode(sd, spec, pts, maps, aabb=aabb)
    assert want.shape == (case["n"], 9) and float(want[:, 1:].abs().max()) > 1.0      # the clamp would have hidden these
    for impl in ("ffma", "tc"):
        net = make_net(spec, sd, 3, impl)
        got = net.decode(pts.cuda(), [m.cuda() for m in maps], aabb=aabb)
        rel, mx = errors(got, want)
        print(f"pbr unclamped {impl}: rel_l2={rel:.2e} max={mx:.2e}")
        assert rel < TOL_SPLIT and mx < TOL_SPLIT * 5, (impl, rel, mx)
        for c0, c1 in ((0, 1), (1, 4), (4, 6), (6, 9)):                                # every head on its own columns
            r, _ = errors(got[:, c0:c1], want[:, c0:c1])
            assert r < TOL_SPLIT * 2, (impl, c0, r)


@pytest.mark.parametrize("name", ["default", "odd", "sdf_only", "aligned", "custom", "pbr"])
def test_encode_matches_reference_golden(golden_dir, name):
    """AutoEncoderGroupSkip.encode (networks.py:164-180) through s3d_decoder_encode vs the real reference's planes."""
    from oracle.cases import ENCODER_CASES, make_encoder_inputs
    case = ENCODER_CASES[name]
    g = np.load(os.path.join(golden_dir, f"encoder_{name}.npz"))
    spec = de.DecoderSpec(**case["spec"])
    net = make_net(spec, de.synthetic_state_dict(spec, case["wseed"]))
    vol = make_encoder_inputs(case)
    got = net.encode(vol.cuda())
    for pl, a in zip(de.PLANES, got):
        w = g[pl]
        assert tuple(a.shape) == w.shape
        err = np.abs(a.cpu().numpy() - w).max()
        print(name, pl, "max abs", err)
        assert err < 2e-5, (name, pl, err)              # outputs are tanh values in (-1, 1): absolute == relative to full scale
    # fixed-point axis sums: a second run is bit-identical
    again = net.encode(vol.cuda())
    assert all(torch.equal(a, b) for a, b in zip(got, again))


def test_encode_then_decode_round_trip_runs_through_forward():
    """forward(vol, x) = decode(x, encode(vol)) (networks.py:222-224) against the oracle chain at the default-latent scale."""
    from oracle.cases import ENCODER_CASES, make_encoder_inputs
    case = ENCODER_CASES["default"]
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    net = make_net(spec, sd)
    vol = make_encoder_inputs(case)
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(500, 3, generator=g) * 2 - 1
    want = de.decode(sd, spec, pts, de.encode(sd, spec, vol))
    got = net(vol.cuda(), pts.cuda()).cpu()
    rel, mx = errors(got, want)
    assert rel < TOL_SPLIT * 5 and mx < TOL, (rel, mx)


def test_encode_full_size_properties():
    """BASELINE-size volume (184 x 256 x 184 -> the cfg2 latent 92 x 128 x 92): shapes, range, finiteness, and the
    InstanceNorm invariant: 2*atanh(plane) has zero mean and variance var/(var+eps) per channel."""
    spec = de.DecoderSpec()
    net = make_net(spec, de.synthetic_state_dict(spec, 7))
    g = torch.Generator(device="cuda").manual_seed(1)
    vol = torch.rand(1, 4, 184, 256, 184, device="cuda", generator=g) * 2 - 1
    planes = net.encode(vol)
    assert [tuple(p.shape) for p in planes] == [(1, 12, 92, 128), (1, 12, 92, 92), (1, 12, 128, 92)]
    for p in planes:
        assert torch.isfinite(p).all() and p.abs().max() < 1
        y = 2 * torch.atanh(p.double())
        assert y.mean(dim=(2, 3)).abs().max() < 1e-4
        v = y.var(dim=(2, 3), unbiased=False)             # = var / (var + 1e-5): just below 1
        assert v.max() <= 1 + 1e-6 and v.min() > 0.99
