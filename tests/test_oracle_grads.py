"""Groundwork for the training row: the oracle's parameter gradients (autograd through the oracle forward) against the compact
fixtures the real reference produced under torch.autograd (oracle/make_golden_grads.py)."""
import os

import numpy as np
import pytest

from oracle.cases import GRAD_CASES
from oracle.make_golden_grads import oracle_grads, sample_idx


@pytest.mark.parametrize("name", list(GRAD_CASES))
def test_oracle_gradients_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"grads_{name}.npz"))
    loss, grads = oracle_grads(GRAD_CASES[name])
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert len(grads) == 138
    for k, gr in grads.items():
        flat = gr.reshape(-1).numpy()
        n_ref = float(g[f"norm/{k}"])
        assert abs(np.linalg.norm(flat.astype(np.float64)) - n_ref) <= 1e-4 * max(n_ref, 1e-12), k
        want = g[f"sample/{k}"]
        assert np.abs(flat[sample_idx(flat.size)] - want).max() <= 1e-4 * max(np.abs(want).max(), n_ref / np.sqrt(flat.size), 1e-12), k
