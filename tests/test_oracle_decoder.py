"""The CPU decoder oracle against the fixtures produced from the real reference (oracle/make_golden_decoder.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_ref as de
from oracle.cases import DECODER_CASES, make_decoder_inputs


@pytest.mark.parametrize("name", list(DECODER_CASES))
def test_decoder_oracle_matches_reference(golden_dir, name):
    case = DECODER_CASES[name]
    g = np.load(os.path.join(golden_dir, f"decoder_{name}.npz"))
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    maps, pts, aabb = make_decoder_inputs(case)
    assert np.array_equal(maps[0].numpy(), g["xy"]) and np.array_equal(aabb.numpy(), g["aabb"])
    if "grid" in case:
        out = de.decode_grid(sd, spec, maps, case["grid"], batch_size=1000, aabb=aabb)
        assert list(out.shape[:3]) == list(g["grid_shape"])
        out = out.reshape(-1, spec.out_channels).numpy()
    else:
        assert np.array_equal(pts.numpy(), g["pts"])
        out = de.decode_batch(sd, spec, maps, pts, batch_size=256, aabb=aabb).numpy()
    # same torch build => bit-exact; leave head-room for a different CPU kernel selection
    assert np.abs(out - g["out"]).max() <= 2e-5 * max(1.0, np.abs(g["out"]).max())
    assert out[:, 1:].min(initial=0.0) >= 0.0 and out[:, 1:].max(initial=0.0) <= 1.0
    planes = de.feature_planes(sd, spec, maps)
    for br, ps in planes.items():
        for pl, a in zip(de.PLANES, ps):
            w = g[f"planes/{br}/{pl}"]
            assert np.abs(a.numpy() - w).max() <= 2e-5 * max(1.0, np.abs(w).max())


def test_hoisted_planes_equal_per_chunk_recompute():
    """The reference recomputes the feature planes for every chunk (networks.py:203-213); hoisting changes nothing."""
    case = DECODER_CASES["tall"]
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    maps, pts, aabb = make_decoder_inputs(case)
    a = de.decode_batch(sd, spec, maps, pts, batch_size=128, aabb=aabb, hoist=True)
    b = de.decode_batch(sd, spec, maps, pts, batch_size=128, aabb=aabb, hoist=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["default", "odd", "sdf_only", "aligned", "custom", "pbr"])
def test_encoder_oracle_matches_reference(golden_dir, name):
    """oracle encode (networks.py:164-180 restated) against the planes the real reference produced."""
    from oracle.cases import ENCODER_CASES, make_encoder_inputs
    case = ENCODER_CASES[name]
    g = np.load(os.path.join(golden_dir, f"encoder_{name}.npz"))
    spec = de.DecoderSpec(**case["spec"])
    sd = de.synthetic_state_dict(spec, case["wseed"])
    planes = de.encode(sd, spec, make_encoder_inputs(case))
    for pl, a in zip(de.PLANES, planes):
        assert a.shape == g[pl].shape
        assert np.abs(a.numpy() - g[pl]).max() <= 1e-6
