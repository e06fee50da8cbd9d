"""-m gpu: the fused AdamW + EMA step against torch.optim.AdamW + update_ema (train_util.py:82-84, 237-239; nn.py:53-63) on CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("weight_decay", [0.0, 0.05])
def test_fused_adamw_ema_matches_torch(weight_decay):
    from sin3dm_b200.optim import FusedAdamWEMA
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 12, 1, 1), (64,), (5, 7, 3), (1,), (129, 33)]       # sizes that are not multiples of 4 included
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    rates = [0.9999, 0.9]
    opt_ref = torch.optim.AdamW(ref, lr=2e-3, weight_decay=weight_decay)
    ema_ref = [[p.detach().clone() for p in ref] for _ in rates]
    opt = FusedAdamWEMA(mine, lr=2e-3, weight_decay=weight_decay, ema_rates=rates)
    for step in range(4):
        lr = 2e-3 * (1 - step / 10)                                      # TrainLoop._anneal_lr
        for pg in opt_ref.param_groups:
            pg["lr"] = lr
        for p, q in zip(ref, mine):
            gr = torch.randn(p.shape, generator=g) * (10.0 ** (step - 2))
            p.grad = gr.clone()
            q.grad.copy_(gr)                                             # autograd would accumulate into the flat gradient view
        opt_ref.step()
        for rate, params in zip(rates, ema_ref):
            for targ, src in zip(params, ref):
                targ.detach().mul_(rate).add_(src.detach(), alpha=1 - rate)
        opt.step(lr=lr)
    for p, q in zip(ref, mine):
        assert torch.allclose(q.detach().cpu(), p.detach(), rtol=2e-6, atol=1e-7), (p - q.cpu()).abs().max()
    for k in range(len(rates)):
        for e, q in zip(ema_ref[k], opt.ema_params(k)):
            assert torch.allclose(q.cpu(), e, rtol=2e-6, atol=1e-7)
    # the module's parameters are views of the flat buffer
    assert all(q.data_ptr() == opt.flat.data_ptr() + 4 * o for q, o in zip(mine, opt.offsets))
