"""CPU-only checks of the host mirror: respacing index path, fp64 tables, fp32 coefficient table,
state_dict layout, and that the C-ABI library loads and exports every declared symbol."""
import json
import os
import re

import numpy as np
import pytest
import torch

import sin3dm_b200 as s3
from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import RESPACE_CASES, TABLE_CASES
from sin3dm_b200 import _lib
from sin3dm_b200.script_util import create_gaussian_diffusion


def test_space_timesteps_bit_exact(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "respace.json")))
    for T, spec in RESPACE_CASES:
        key = f"{T}|{spec if isinstance(spec, str) else ','.join(map(str, spec))}"
        try:
            got = sorted(s3.space_timesteps(T, spec))
        except ValueError:
            got = "ValueError"
        assert got == want[key], key


def test_tables_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    for name, case in TABLE_CASES.items():
        d = create_gaussian_diffusion(steps=case["T"], noise_schedule=case.get("schedule", "linear"),
                                      predict_xstart=True, timestep_respacing=case["respacing"])
        assert np.array_equal(np.array(d.timestep_map), g[f"{name}/timestep_map"])
        for k in dr.tables(np.ones(2) * 0.5):
            assert np.array_equal(getattr(d, k), g[f"{name}/{k}"]), (name, k)


@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_coef_table_matches_reference_expressions(eta):
    d = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="25")
    o = dr.RefDiffusion(1000, "25")
    cf = d.coef_table("cpu", eta)
    assert cf.shape == (25, _lib.S3D_NCOEF) and cf.dtype == torch.float32
    t = torch.arange(25)
    one = torch.zeros(25, 1)
    f = lambda a: o._x(a, t, one)[:, 0]
    ab, abp = f(o.tab["alphas_cumprod"]), f(o.tab["alphas_cumprod_prev"])
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    assert torch.equal(cf[:, 0], f(o.tab["sqrt_recip_alphas_cumprod"]))
    assert torch.equal(cf[:, 2], f(o.tab["posterior_mean_coef1"]))
    assert torch.equal(cf[:, 4], torch.exp(0.5 * f(o.logvar)))
    assert torch.equal(cf[:, 5], torch.sqrt(abp))
    assert torch.equal(cf[:, 6], torch.sqrt(1 - abp - sigma ** 2))
    assert torch.equal(cf[:, 7], sigma)


def test_model_timesteps_map_and_rescale():
    d = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="10", rescale_timesteps=True)
    o = dr.RefDiffusion(1000, "10", rescale_timesteps=True)
    t = torch.tensor([0, 3, 9])
    assert torch.equal(d._model_timesteps(t), o.model_t(t))
    d2 = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="ddim50")
    assert d2._model_timesteps(t).dtype == torch.long
    assert d2._model_timesteps(t).tolist() == [0, 60, 180]


@pytest.mark.parametrize("rollout", [True, False])
@pytest.mark.parametrize("mult", [(1, 2), (1, 2, 2), (1,)])
def test_state_dict_layout(rollout, mult):
    cls = s3.TriplaneUNetModelSmall if rollout else s3.TriplaneUNetModelSmallRaw
    m = cls(12, 64, 12, 1, 0, mult, use_scale_shift_norm=True)
    spec = ur.UNetSpec(12, 64, 12, 1, mult, True, rollout)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in ur.param_shapes(spec)]
    # zero-init trap: second conv of every block and the out conv start at zero
    zeros = [k for k, v in m.state_dict().items() if (v == 0).all()]
    assert any("out_layers.2" in k for k in zeros) and any(k.startswith("out.2") for k in zeros)


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        s3.TriplaneUNetModelSmall(8, 64, 8, num_res_blocks=2)
    from sin3dm_b200 import gaussian_diffusion as gd
    with pytest.raises(NotImplementedError):
        create_gaussian_diffusion(learn_sigma=True)
    with pytest.raises(ValueError):
        s3.space_timesteps(50, "60")


def test_cpu_tensors_fail_loudly():
    m = s3.TriplaneUNetModelSmall(8, 64, 8, use_scale_shift_norm=True)
    with pytest.raises(_lib.S3DError), torch.no_grad():
        m(torch.zeros(1, 8, 16, 16), torch.zeros(1), H=8, W=8, D=8)


def test_abi_exports_every_declared_symbol():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "sin3dm_b200.h")).read()
    declared = set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib._SIGNATURES), declared ^ set(_lib._SIGNATURES)
    L = _lib.lib()            # raises if the .so is missing or a symbol cannot be bound
    assert L.s3d_abi_version() == 4
    for name in declared:
        assert hasattr(L, name)
    # struct sizes agree with the header layout (no compute call: there is no GPU here)
    import ctypes as C
    assert C.sizeof(_lib.UNetConfig) == 4 * (5 + 8 + 4)
    assert C.sizeof(_lib.DecoderConfig) == 4 * 12


@pytest.mark.parametrize("use_tex", [True, False])
def test_decoder_state_dict_layout(use_tex):
    """The decoder mirror carries the reference checkpoint keys (networks.py:134-162) in state_dict() order."""
    from oracle import decoder_ref as de
    from sin3dm_b200.encoding import AutoEncoderGroupSkip
    spec = de.DecoderSpec(use_tex=use_tex, tex_feat_channels=8 if use_tex else 0)
    m = AutoEncoderGroupSkip(4, spec.tex_feat_channels, 64, 256, 4, use_tex=use_tex, tex_channels=3)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in de.param_shapes(spec)]
    m.load_state_dict(de.synthetic_state_dict(spec, 1))
    # zero_module on the second conv of each block (blocks.py:222-224)
    m2 = AutoEncoderGroupSkip(4, spec.tex_feat_channels, 64, 256, 4, use_tex=use_tex)
    assert (m2.geo_convs.out_layers[1].weight == 0).all()
    with pytest.raises(_lib.S3DError), torch.no_grad():
        m.decode(torch.zeros(4, 3), [torch.zeros(1, spec.geo_feat_channels + spec.tex_feat_channels, 8, 8)] * 3)
    with pytest.raises(_lib.S3DError), torch.no_grad():        # encode runs on the GPU only, like decode
        m.encode(torch.zeros(1, 4 if use_tex else 1, 8, 8, 8))


@pytest.mark.parametrize("use_tex", [True, False])
def test_pbr_decoder_state_dict_layout(use_tex):
    """AutoEncoderGroupPBR (networks.py:227-262): checkpoint keys and order (tex_convs.0 / tex_convs.1, rgb / mr / normal heads)."""
    from oracle import decoder_ref as de
    from sin3dm_b200.encoding import AutoEncoderGroupPBR, get_networks
    spec = de.DecoderSpec(net_kind="pbr", use_tex=use_tex, tex_feat_channels=8 if use_tex else 0, tex_channels=8)
    m = AutoEncoderGroupPBR(4, spec.tex_feat_channels, 64, 256, 4, use_tex=use_tex, tex_channels=8)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in de.param_shapes(spec)]
    m.load_state_dict(de.synthetic_state_dict(spec, 1))
    assert m.out_channels == (9 if use_tex else 1)
    if use_tex:
        assert "tex_convs.1.in_layers.1.weight" in m.state_dict() and "tex_convs.1.shortcut.weight" not in m.state_dict()
        with pytest.raises(ValueError):
            AutoEncoderGroupPBR(4, 8, 64, 256, 4, use_tex=True, tex_channels=3)

    class Cfg:
        data_type, enc_net_type, fdim_geo, fdim_tex, fdim_up, hidden_dim, n_hidden_layers = "sdfpbr", "pbr", 4, 8, 64, 256, 4
    assert isinstance(get_networks(Cfg), AutoEncoderGroupPBR)


def test_grid_axes_match_oracle():
    from oracle import decoder_ref as de
    from sin3dm_b200.encoding import sample_grid_points_axes
    for aabb, reso in (([-0.72, -1.0, -0.55, 0.72, 1.0, 0.55], 20), ([-1, -1, -1, 1, 1, 1], 7), ([-0.3, -1, -1, 0.3, 1, 0.9], 33)):
        a = torch.tensor(aabb, dtype=torch.float32)
        for u, v in zip(sample_grid_points_axes(a, reso), de.grid_axes(a, reso)):
            assert torch.equal(u, v)


def test_triplane_io_round_trip_and_pad(tmp_path):
    """save/load_triplane_data and pad_composed_featmaps (reference src/utils/triplane_util.py:28-61) against the reference's
    definitions restated inline: npz keys feat_xy/xz/yz, compose on load, zero padding per plane."""
    import numpy as np
    import sin3dm_b200 as s3
    g = torch.Generator().manual_seed(0)
    H, W, D, C = 5, 7, 4, 3
    xy, xz, yz = torch.randn(C, H, W, generator=g), torch.randn(C, H, D, generator=g), torch.randn(C, W, D, generator=g)
    p = str(tmp_path / "sub" / "feat.npz")
    s3.save_triplane_data(p, xy, xz, yz)
    assert sorted(np.load(p).files) == ["feat_xy", "feat_xz", "feat_yz"]
    comp, sizes = s3.load_triplane_data(p, device="cpu")
    assert sizes == (H, W, D) and comp.shape == (C, H + D, W + D)
    a, b, c = s3.decompose_featmaps(comp, sizes)
    assert torch.equal(a, xy) and torch.equal(b, xz) and torch.equal(c, yz)
    assert (comp[..., H:, W:] == 0).all()
    planes = s3.load_triplane_data(p, device="cpu", compose=False)
    assert torch.equal(planes[2], yz)
    padded, ns = s3.pad_composed_featmaps(comp, sizes, [[1, 2], [0, 3], [2, 0]])
    assert ns == (H + 3, W + 3, D + 2)
    pa, pb, pc = s3.decompose_featmaps(padded, ns)
    assert torch.equal(pa[..., 1:1 + H, 0:W], xy) and int((pa != 0).sum()) == int((xy != 0).sum())
    assert torch.equal(pb[..., 1:1 + H, 2:2 + D], xz) and torch.equal(pc[..., 0:W, 2:2 + D], yz)


def test_schedule_samplers_match_reference_values():
    """sin3dm_b200.resample against values produced by the reference's resample.py under the same numpy seeds (UniformSampler
    draw, LossSecondMomentResampler weights before / after warm-up and after the history shifts, and a weighted draw)."""
    import numpy as np
    from sin3dm_b200.resample import LossSecondMomentResampler, UniformSampler, create_named_schedule_sampler

    class D:
        def __init__(self, n):
            self.num_timesteps = n
    np.random.seed(0)
    s = create_named_schedule_sampler("uniform", D(1000))
    assert isinstance(s, UniformSampler)
    i, w = s.sample(8, "cpu")
    assert i.dtype == torch.int64 and i.tolist() == [548, 715, 602, 544, 423, 645, 437, 891] and w.tolist() == [1.0] * 8
    np.random.seed(1)
    r = LossSecondMomentResampler(D(5), history_per_term=2, uniform_prob=0.01)
    assert r.weights().tolist() == [1.0] * 5
    r.update_with_all_losses([0, 1, 2, 3, 4, 0, 1, 2, 3, 4], [1., 2., 3., 4., 5., 2., 1., 0.5, 4., 3.])
    assert np.allclose(r.weights(), [0.11850279589800725, 0.11850279589800725, 0.160460934259197, 0.29673135105233966,
                                     0.30580212289244885], rtol=1e-14)
    r.update_with_local_losses(torch.tensor([0, 0]), torch.tensor([10., 20.]))           # single process: no collective
    assert np.allclose(r.weights(), [0.5677902587017709, 0.058579025870177104, 0.07895579517861935, 0.1451348716346799,
                                     0.14954004861475256], rtol=1e-14)
    i, w = r.sample(6, "cpu")
    assert i.tolist() == [0, 3, 0, 0, 0, 0]
    assert np.allclose(w.numpy(), [0.35224273800849915, 1.3780286312103271] + [0.35224273800849915] * 4, rtol=1e-6)
    with pytest.raises(NotImplementedError):
        create_named_schedule_sampler("nope", D(3))


def test_fused_optimizer_refuses_cpu_parameters():
    from sin3dm_b200.optim import FusedAdamWEMA
    with pytest.raises(_lib.S3DError):
        FusedAdamWEMA([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)


def test_synthetic_weights_match_the_oracle_recipe():
    """bench.py's GPU arm draws its checkpoint with sin3dm_b200.synthetic (no oracle import on the product arm); the CPU arm and
    the parity tests use oracle.unet_ref.synthetic_state_dict.  Same generator, same order, same values."""
    import torch
    import sin3dm_b200 as s3
    from oracle import unet_ref as ur
    from sin3dm_b200.synthetic import synthetic_state_dict_like
    m = s3.TriplaneUNetModelSmall(12, 64, 12, 1, 0, (1, 2), use_scale_shift_norm=True)
    a = synthetic_state_dict_like(m, 1234)
    b = ur.synthetic_state_dict(ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12), 1234)
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
