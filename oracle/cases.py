"""Shared, seeded definitions of the parity cases (test infrastructure).

Used by ``oracle/make_golden.py`` (to produce fixtures from the real reference), by ``tests/`` (to
rebuild the same inputs and compare oracle / CUDA path with the fixtures) and by ``bench.py``.
"""
import torch

SMALL = dict(in_channels=8, model_channels=64, out_channels=8)

# (T, section spec) — respace.py:7-60 index path, must be bit-exact
RESPACE_CASES = [
    (1000, "100"), (1000, "10"), (1000, "250"), (1000, "1000"), (1000, "ddim100"), (1000, "ddim50"),
    (1000, "ddim25"), (1000, "10,10,20"), (1000, "1"), (1000, "2"), (1000, "3,1,7"), (300, [10, 15, 20]),
    (100, "25"), (100, "ddim7"), (1000, "ddim3"), (50, "60"), (4000, "100"), (1000, "999"), (1000, "333"),
]

TABLE_CASES = {
    "ddpm1000": dict(T=1000, respacing=""),
    "resp100": dict(T=1000, respacing="100"),
    "resp10": dict(T=1000, respacing="10"),
    "ddim100": dict(T=1000, respacing="ddim100"),
    "cos200_25": dict(T=200, respacing="25", schedule="cosine"),
}

UNET_CASES = {
    # odd sizes: exercises floor avg-pool + bilinear resize-to-skip (unet_triplane.py:494-499)
    "small_odd": dict(spec=dict(**SMALL), wseed=11, HWD=(13, 18, 11), B=2, t=[3, 977], xseed=21),
    "small_even": dict(spec=dict(**SMALL), wseed=12, HWD=(16, 24, 8), B=1, t=[500], xseed=22),
    "raw": dict(spec=dict(in_channels=4, model_channels=64, out_channels=4, rollout=False), wseed=13,
                HWD=(12, 16, 8), B=2, t=[0, 999], xseed=23),
    "add_emb": dict(spec=dict(**SMALL, use_scale_shift_norm=False), wseed=14, HWD=(8, 12, 10), B=1, t=[42],
                    xseed=24),
    "three_level": dict(spec=dict(**SMALL, channel_mult=(1, 2, 2)), wseed=15, HWD=(16, 20, 12), B=1, t=[700],
                        xseed=25),
    "c12_float_t": dict(spec=dict(in_channels=12, model_channels=64, out_channels=12), wseed=16, HWD=(10, 16, 6),
                        B=2, t=[12.5, 999.0], xseed=26),
}

SAMPLER_CASES = {
    # BASELINE.json configs[0]: DDIM 10-step, 3x8chx32x32 triplane
    "cfg1_ddim10": dict(spec=dict(**SMALL), wseed=31, HWD=(32, 32, 32), B=1, T=1000, respacing="10", ddim=True,
                        nseed=41),
    "ddpm20_small": dict(spec=dict(**SMALL), wseed=32, HWD=(12, 16, 10), B=2, T=1000, respacing="20", ddim=False,
                         nseed=42),
    "ddpm_eps": dict(spec=dict(**SMALL), wseed=33, HWD=(8, 12, 8), B=1, T=1000, respacing="15", ddim=False,
                     mean_type="epsilon", nseed=43),
    "ddim_eta_mask": dict(spec=dict(**SMALL), wseed=34, HWD=(8, 12, 8), B=2, T=1000, respacing="ddim10", ddim=True,
                          eta=0.7, mask=True, nseed=44),
    "ddim_eps_small_var_t0mask": dict(spec=dict(**SMALL), wseed=35, HWD=(8, 8, 8), B=1, T=1000, respacing="8",
                                      ddim=True, mean_type="epsilon", var_type="fixed_small", mask=True,
                                      is_mask_t0=True, rescale_timesteps=True, nseed=45),
    "ddpm_noclip_small_var": dict(spec=dict(**SMALL), wseed=36, HWD=(8, 12, 8), B=1, T=100, respacing="12",
                                  ddim=False, var_type="fixed_small", clip=False, nseed=46),
}


def make_inputs(case):
    H, W, D = case["HWD"]
    g = torch.Generator().manual_seed(case["xseed"])
    x = torch.randn(case["B"], case["spec"]["in_channels"], H + D, W + D, generator=g)
    tv = case["t"]
    t = torch.tensor(tv, dtype=torch.float32 if any(isinstance(v, float) for v in tv) else torch.long)
    return x, t


def make_step_noise(case, n_steps):
    """-> (x_T, {step_index: noise}) drawn on CPU in a fixed order (x_T first, then step T-1 .. 0)."""
    H, W, D = case["HWD"]
    shape = (case["B"], case["spec"]["in_channels"], H + D, W + D)
    g = torch.Generator().manual_seed(case["nseed"])
    x_T = torch.randn(shape, generator=g)
    noises = {i: torch.randn(shape, generator=g) for i in range(n_steps - 1, -1, -1)}
    return x_T, noises


# ---------------------------------------------------------------------------- triplane decoder
DECODER_CASES = {
    # default auto-encoder (parser_util.py:19-26), ragged plane sizes, points inside and a little outside the box
    "default": dict(spec=dict(), wseed=51, HWD=(20, 28, 18), n=777, aabb=[-0.7, -1.0, -0.6, 0.7, 1.0, 0.66], seed=61,
                    spill=1.15),
    # H > W > D
    "tall": dict(spec=dict(), wseed=52, HWD=(26, 14, 9), n=300, aabb=[-1.0, -0.5, -0.3, 1.0, 0.55, 0.35], seed=62,
                 spill=1.0),
    # data_type == "sdf": geometry branch only
    "sdf_only": dict(spec=dict(use_tex=False, tex_feat_channels=0), wseed=53, HWD=(12, 16, 12), n=257,
                     aabb=[-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], seed=63, spill=1.3),
    # enc_net_type == "base": AutoEncoderGroupV3 with the plain DecoderMLP heads (networks.py:21-131, blocks.py:46-62)
    "base_v3": dict(spec=dict(mlp_kind="base"), wseed=55, HWD=(18, 12, 22), n=500, aabb=[-0.8, -0.6, -1.0, 0.8, 0.6, 1.0], seed=65,
                    spill=1.1),
    # enc_net_type == "pbr": AutoEncoderGroupPBR (networks.py:227-331), data_type "sdfpbr" -> 8 texture channels
    "pbr": dict(spec=dict(net_kind="pbr", tex_channels=8), wseed=56, HWD=(14, 20, 17), n=400, aabb=[-0.9, -1.0, -0.7, 0.9, 1.0, 0.7], seed=66,
                spill=1.1),
    # decode_grid (model.py:335-349) over sample_grid_points_aabb (utils3d.py:13-25)
    "grid": dict(spec=dict(), wseed=54, HWD=(16, 22, 12), grid=20, aabb=[-0.72, -1.0, -0.55, 0.72, 1.0, 0.55], seed=64),
}


def make_decoder_inputs(case):
    """-> (feat_maps [xy, xz, yz] each [1, c, ., .] in (-1, 1) like the encoder's tanh output, points [n, 3] or None,
    aabb [6])."""
    from oracle.decoder_ref import DecoderSpec
    spec = DecoderSpec(**case["spec"])
    c = spec.geo_feat_channels + (spec.tex_feat_channels if spec.use_tex else 0)
    H, W, D = case["HWD"]
    g = torch.Generator().manual_seed(case["seed"])
    maps = [torch.tanh(torch.randn(1, c, a, b, generator=g)) for a, b in ((H, W), (H, D), (W, D))]
    aabb = torch.tensor(case["aabb"], dtype=torch.float32)
    pts = None
    if "n" in case:
        u = torch.rand(case["n"], 3, generator=g) * 2 - 1
        centre, half = (aabb[3:] + aabb[:3]) / 2, (aabb[3:] - aabb[:3]) / 2
        pts = centre + u * half * case["spill"]
    return maps, pts, aabb


# ---------------------------------------------------------------------------- variational bound
BPD_CASES = {
    "startx_large": dict(spec=dict(**SMALL), wseed=71, HWD=(8, 12, 8), B=2, T=1000, respacing="6", nseed=81),
    "eps_small_noclip": dict(spec=dict(**SMALL), wseed=72, HWD=(8, 8, 10), B=1, T=1000, respacing="5", mean_type="epsilon",
                             var_type="fixed_small", clip=False, nseed=82),
}


def make_bpd_inputs(case, n_steps):
    """-> (x_start in [-1, 1], {step: q_sample noise}); x_start holds a few exact +-1 so both edge branches of the
    discretised likelihood (losses.py:71-75) are exercised."""
    H, W, D = case["HWD"]
    shape = (case["B"], case["spec"]["in_channels"], H + D, W + D)
    g = torch.Generator().manual_seed(case["nseed"])
    x0 = torch.rand(shape, generator=g) * 2 - 1
    x0.view(-1)[::17] = 1.0
    x0.view(-1)[5::23] = -1.0
    noises = {i: torch.randn(shape, generator=g) for i in range(n_steps - 1, -1, -1)}
    return x0, noises


# ---------------------------------------------------------------------------- encoder
ENCODER_CASES = {
    # even / odd volume sizes (odd: the stride-2 conv drops the last voxel), not multiples of the kernel's tile
    "default": dict(spec=dict(), wseed=91, XYZ=(22, 36, 70), seed=95),
    "odd": dict(spec=dict(), wseed=92, XYZ=(15, 9, 33), seed=96),
    "sdf_only": dict(spec=dict(use_tex=False, tex_feat_channels=0), wseed=93, XYZ=(12, 20, 16), seed=97),
    # Z % 4 == 0 with colour: the TMA-staged kernel, several z tiles with a ragged last one
    "aligned": dict(spec=dict(), wseed=94, XYZ=(10, 14, 148), seed=98),
    # non-default channel counts (fdim_geo 3, fdim_tex 5, two colour channels): the generic direct-convolution kernel
    "custom": dict(spec=dict(geo_feat_channels=3, tex_feat_channels=5, tex_channels=2), wseed=99, XYZ=(14, 11, 150), seed=100),
    # AutoEncoderGroupPBR.encode (networks.py:270-287): sdf + 8 material channels in
    "pbr": dict(spec=dict(net_kind="pbr", tex_channels=8), wseed=101, XYZ=(18, 12, 26), seed=102),
}


def make_encoder_inputs(case):
    """-> vol [1, 1 (+3), X, Y, Z]: a smooth signed-distance-like channel plus colour channels in [0, 1]."""
    from oracle.decoder_ref import DecoderSpec
    spec = DecoderSpec(**case["spec"])
    X, Y, Z = case["XYZ"]
    g = torch.Generator().manual_seed(case["seed"])
    ax = [torch.linspace(-1, 1, n) for n in (X, Y, Z)]
    gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
    sdf = (gx ** 2 + 0.7 * gy ** 2 + 1.3 * gz ** 2).sqrt() - 0.6 + 0.05 * torch.randn(X, Y, Z, generator=g)
    chans = [sdf.clamp(-0.2, 0.2) / 0.2]
    if spec.use_tex:
        chans += [torch.rand(X, Y, Z, generator=g) for _ in range(spec.tex_channels)]
    return torch.stack(chans, dim=0)[None].contiguous()


# ---------------------------------------------------------------------------- parameter gradients (training row, groundwork)
GRAD_CASES = {
    "startx": dict(spec=dict(**SMALL), wseed=111, HWD=(12, 16, 10), B=2, T=1000, respacing="", seed=121, t=[3, 871]),
    "eps_odd": dict(spec=dict(**SMALL), wseed=112, HWD=(9, 14, 7), B=1, T=1000, respacing="", mean_type="epsilon", seed=122, t=[456]),
}


def make_grad_inputs(case):
    """-> (x_start in [-1, 1], q_sample noise, t)."""
    H, W, D = case["HWD"]
    shape = (case["B"], case["spec"]["in_channels"], H + D, W + D)
    g = torch.Generator().manual_seed(case["seed"])
    x0 = torch.rand(shape, generator=g) * 2 - 1
    nz = torch.randn(shape, generator=g)
    return x0, nz, torch.tensor(case["t"])


# ---------------------------------------------------------------------------- full-length BASELINE chains
# The BASELINE.json configurations end to end: the real reference's final latent for DDPM-1000 at cfg2, DDIM-100 at cfg3 (B=8,
# D = 138) and a 20-step DDPM chain at the cfg5 shape (B=8).  Fixtures hold sample 0 in full and every FULL_STRIDE-th element of
# the other samples (the composed tensor, dead corner included: both sides evolve it as pure noise) — see pack_full / full_errors.
C12 = dict(in_channels=12, model_channels=64, out_channels=12)
FULL_CASES = {
    "cfg2_ddpm1000": dict(spec=dict(**C12), wseed=1234, HWD=(92, 128, 92), B=1, T=1000, respacing="", ddim=False, nseed=201),
    "cfg3_ddim100": dict(spec=dict(**C12), wseed=1234, HWD=(92, 128, 138), B=8, T=1000, respacing="100", ddim=True, nseed=202),
    "cfg5_ddpm20": dict(spec=dict(**C12), wseed=1234, HWD=(92, 128, 92), B=8, T=1000, respacing="20", ddim=False, nseed=203),
}
FULL_STRIDE = 8


def pack_full(sample):
    """final composed latent [B, C, H+D, W+D] -> fixture arrays."""
    out = dict(sample0=sample[0].numpy())
    if sample.shape[0] > 1:
        out["rest_strided"] = sample[1:].reshape(sample.shape[0] - 1, -1)[:, ::FULL_STRIDE].contiguous().numpy()
    return out


def full_errors(got, fixture, H, W, D):
    """-> (rel-L2 over the three planes of sample 0, rel-L2 over the strided elements of the other samples or 0.0)."""
    from oracle.unet_ref import split_planes
    got = torch.as_tensor(got).cpu().double()
    w0 = torch.from_numpy(fixture["sample0"]).double()
    num = sum(((a - b) ** 2).sum() for a, b in zip(split_planes(got[0], H, W, D), split_planes(w0, H, W, D)))
    den = sum((b ** 2).sum() for b in split_planes(w0, H, W, D))
    rel0 = float((num / den).sqrt())
    rel_rest = 0.0
    if "rest_strided" in fixture:
        wr = torch.from_numpy(fixture["rest_strided"]).double()
        gr = got[1:].reshape(got.shape[0] - 1, -1)[:, ::FULL_STRIDE]
        rel_rest = float(((gr - wr) ** 2).sum().sqrt() / (wr ** 2).sum().sqrt())
    return rel0, rel_rest


# ---------------------------------------------------------------------------- guided sampling hooks
# cond_fn / denoised_fn loops (gaussian_diffusion.py:357-394, 233-327) and q_mean_variance (:172-187)
COND_CASES = {
    "ddpm_cond_mean": dict(spec=dict(**SMALL), wseed=131, HWD=(8, 12, 10), B=2, T=1000, respacing="12", ddim=False, nseed=141,
                           cond=True, denoise=False),
    "ddim_cond_score": dict(spec=dict(**SMALL), wseed=132, HWD=(10, 8, 8), B=1, T=1000, respacing="ddim10", ddim=True, nseed=142,
                            cond=True, denoise=True, eta=0.3, rescale_timesteps=True),
    "ddpm_denoised_fn_eps": dict(spec=dict(**SMALL), wseed=133, HWD=(8, 8, 12), B=1, T=1000, respacing="9", ddim=False, nseed=143,
                                 cond=False, denoise=True, mean_type="epsilon"),
}


def cond_fn(x, t, **kwargs):
    """deterministic stand-in for grad log p(y | x): smooth in x, depends on the (mapped) timestep"""
    return 0.3 * torch.tanh(x) * (1.0 + t.float().view(-1, 1, 1, 1) / 1000.0) - 0.05


def denoised_fn(x):
    return 0.9 * x + 0.05 * torch.sin(3.0 * x)
