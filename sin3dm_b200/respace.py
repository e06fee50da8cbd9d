"""Timestep respacing — mirrors reference src/diffusion/respace.py (space_timesteps :7-60,
SpacedDiffusion :63-113, _WrappedModel :116-128).  The kept-step index set is integer / fp64 host
arithmetic and is reproduced bit-exactly (tests/test_host_logic.py against tests/golden/respace.json)."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            wanted = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == wanted:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    n_sections = len(section_counts)
    base, extra = divmod(num_timesteps, n_sections)
    first, chosen = 0, []
    for k, count in enumerate(section_counts):
        size = base + (1 if k < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            chosen.append(first + round(pos))      # fp64 accumulate, then round-half-even: keep this exact
            pos += stride
        first += size
    return set(chosen)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that keeps only ``use_timesteps`` of a base process (respace.py:63-113)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base_ac = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        prev, new_betas = 1.0, []
        for i, ac in enumerate(base_ac):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / prev)
                prev = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _model_timesteps(self, t):
        """What _WrappedModel hands to the network: original timestep, optionally rescaled (respace.py:123-128)."""
        tmap = th.tensor(self.timestep_map, device=t.device, dtype=th.long)
        new_ts = tmap[t.long()]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return new_ts

    def _scale_timesteps(self, t):
        return t      # scaling is applied together with the map above

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)


class _WrappedModel:
    """Callable kept for API parity (respace.py:116-128); the kernels consume the mapped timesteps directly."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model, self.timestep_map = model, timestep_map
        self.rescale_timesteps, self.original_num_steps = rescale_timesteps, original_num_steps

    def __call__(self, x, ts, **kwargs):
        new_ts = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
