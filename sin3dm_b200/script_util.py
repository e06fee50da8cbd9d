"""Factories with the reference's names (src/diffusion/script_util.py:7-60) so sample.py / train.py
construct the B200 path by changing one import."""
from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps
from .unet_triplane import TriplaneUNetModelSmall, TriplaneUNetModelSmallRaw

_MODEL_KEYS = ("in_channels", "model_channels", "out_channels", "num_res_blocks", "dropout", "channel_mult",
               "use_checkpoint", "use_fp16", "use_scale_shift_norm")
_DIFFUSION_KEYS = ("learn_sigma", "steps", "noise_schedule", "timestep_respacing", "use_kl", "predict_xstart",
                   "rescale_timesteps", "rescale_learned_sigmas")


def create_model_and_diffusion_from_args(args):
    diffusion = create_gaussian_diffusion(**{k: getattr(args, k) for k in _DIFFUSION_KEYS})
    if isinstance(args.channel_mult, str):
        args.channel_mult = tuple(int(m) for m in args.channel_mult.split(","))
    kw = {k: getattr(args, k) for k in _MODEL_KEYS}
    if args.diff_net_type == "unet_small":
        model = TriplaneUNetModelSmall(**kw)
    elif args.diff_net_type == "unet_raw":
        model = TriplaneUNetModelSmallRaw(**kw)
    else:
        raise ValueError(f"unknown diff_net_type {args.diff_net_type}")
    return model, diffusion


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear", use_kl=False,
                              predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False,
                              timestep_respacing=""):
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if not timestep_respacing:
        timestep_respacing = [steps]
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
        model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
        model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps)
