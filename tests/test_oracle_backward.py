"""Groundwork for the training row: the folded-form backward of the rollout TriplaneConv (oracle/backward_ref.py) against
torch.autograd through the oracle forward (which is pinned to the real reference)."""
import pytest
import torch

from oracle import backward_ref as br
from oracle import unet_ref as ur


@pytest.mark.parametrize("C,Cout,HWD,B", [(4, 6, (7, 9, 5), 2), (3, 3, (1, 4, 2), 1), (5, 2, (6, 1, 3), 3)])
def test_folded_backward_equals_autograd(C, Cout, HWD, B):
    g = torch.Generator().manual_seed(C * 100 + Cout)
    H, W, D = HWD
    sd = {}
    for n in br.PLANES:
        sd[f"c.conv_{n}.weight"] = torch.randn(Cout, 3 * C, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
        sd[f"c.conv_{n}.bias"] = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True)
    planes = [torch.randn(B, C, r, c, generator=g, dtype=torch.float64, requires_grad=True) for r, c in ((H, W), (H, D), (W, D))]
    outs = ur.tri_conv(sd, "c", planes, 1, True)
    dys = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in outs]
    loss = sum((o * d).sum() for o, d in zip(outs, dys))
    loss.backward()
    dplanes, grads = br.tri_conv_backward_folded({k: v.detach() for k, v in sd.items()}, "c", [p.detach() for p in planes], dys)
    for p, dp in zip(planes, dplanes):
        assert torch.allclose(p.grad, dp, rtol=1e-10, atol=1e-10)
    for k, v in sd.items():
        assert torch.allclose(v.grad, grads[k], rtol=1e-10, atol=1e-10), k
