"""Generate tests/golden/grads_*.npz from the REAL reference under torch.autograd (authoring container only).

    python -m oracle.make_golden_grads      # needs /root/reference/src (read-only import)

Groundwork for the training row: the unmodified ``TriplaneUNetModelSmall`` and
``SpacedDiffusion.training_losses`` (gaussian_diffusion.py:771-856) produce ``loss = mean(losses["loss"])`` as
``TrainLoop.forward_backward`` does with the uniform sampler (train_util.py:198-229), ``loss.backward()`` gives the gradient of
every parameter.  The oracle gradient — autograd through ``oracle.unet_ref.unet_forward`` / ``RefDiffusion.training_losses`` — is
asserted against it, and a compact fixture is frozen per parameter: L2 norm, sum, and a fixed strided sample of <= 256 entries
(the full gradient set is 28 MB).
"""
import os

import numpy as np
import torch

from oracle import diffusion_ref as dr
from oracle import unet_ref as ur
from oracle.cases import GRAD_CASES, make_grad_inputs
from oracle.make_golden import OUT, build_ref_diffusion, build_ref_unet, ref_modules


def sample_idx(n):
    """The fixed entries of a flattened gradient the fixture keeps."""
    return np.arange(0, n, max(1, n // 256))[:256]


def oracle_grads(case):
    """{name: grad} by autograd through the oracle forward (also used by tests/test_oracle_grads.py)."""
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    H, W, D = case["HWD"]
    x0, nz, t = make_grad_inputs(case)
    o = dr.RefDiffusion(case["T"], case["respacing"], "linear", case.get("mean_type", "start_x"))
    terms = o.training_losses(lambda xx, tt: ur.unet_forward(sdg, spec, xx, tt, H, W, D), x0, t, nz, H, W, D)
    loss = terms["loss"].mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in sdg.items()}


def main():
    gd, respace, ut = ref_modules()
    torch.set_num_threads(os.cpu_count())
    for name, case in GRAD_CASES.items():
        spec = ur.UNetSpec(**case["spec"])
        sd = ur.synthetic_state_dict(spec, case["wseed"])
        m = build_ref_unet(ut, spec, sd).train()
        d = build_ref_diffusion(gd, respace, case)
        H, W, D = case["HWD"]
        x0, nz, t = make_grad_inputs(case)
        losses = d.training_losses(m, x0, t, model_kwargs=dict(H=H, W=W, D=D), noise=nz)
        loss = losses["loss"].mean()
        loss.backward()
        want = {k: p.grad for k, p in m.named_parameters()}
        oloss, got = oracle_grads(case)
        assert abs(oloss - float(loss)) <= 1e-6 * abs(float(loss)), (oloss, float(loss))
        save = {"loss": np.float32(float(loss))}
        worst = 0.0
        for k, gw in want.items():
            go = got[k]
            err = float((gw - go).norm() / gw.norm().clamp(min=1e-20))
            worst = max(worst, err)
            assert err <= 1e-4, (name, k, err)
            flat = gw.reshape(-1).numpy()
            save[f"norm/{k}"] = np.float32(np.linalg.norm(flat.astype(np.float64)))
            save[f"sum/{k}"] = np.float32(flat.astype(np.float64).sum())
            save[f"sample/{k}"] = flat[sample_idx(flat.size)]
        np.savez_compressed(os.path.join(OUT, f"grads_{name}.npz"), **save)
        print("grads", name, "loss", float(loss), "tensors", len(want), "oracle-vs-ref worst rel-L2", worst)


if __name__ == "__main__":
    main()
