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
