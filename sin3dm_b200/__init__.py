"""sin3dm_b200 — B200-native (sm_100a) triplane denoising path with the Sin3DM class API.

    from sin3dm_b200.script_util import create_model_and_diffusion_from_args     # as in the reference
    from sin3dm_b200 import TriplaneUNetModelSmall, SpacedDiffusion, space_timesteps

All arithmetic runs in the hand-written CUDA kernels of ``lib/libsin3dm_b200.so`` (C ABI:
include/sin3dm_b200.h).  Importing the package does not load the library; the first call does, and
raises if it is missing — there is no CPU / torch fallback.
"""
from .gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,  # noqa: F401
                                 get_named_beta_schedule)
from .respace import SpacedDiffusion, space_timesteps  # noqa: F401
from .triplane_util import (compose_featmaps, decompose_featmaps, load_triplane_data, pad_composed_featmaps,  # noqa: F401
                            save_triplane_data)
from .unet_triplane import TriplaneUNetModelSmall, TriplaneUNetModelSmallRaw  # noqa: F401

__all__ = ["GaussianDiffusion", "SpacedDiffusion", "space_timesteps", "TriplaneUNetModelSmall",
           "TriplaneUNetModelSmallRaw", "ModelMeanType", "ModelVarType", "LossType", "get_named_beta_schedule",
           "compose_featmaps", "decompose_featmaps", "pad_composed_featmaps", "save_triplane_data", "load_triplane_data"]
