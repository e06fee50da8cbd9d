#!/usr/bin/env python
"""Where the fixed cost of one public-API sampling call goes (cfg2, 20-step chain): host time to enqueue vs device time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import sin3dm_b200 as s3
    from sin3dm_b200.script_util import create_gaussian_diffusion
    from sin3dm_b200.synthetic import synthetic_state_dict_like
    dev = torch.device("cuda", 0)
    m = s3.TriplaneUNetModelSmall(12, 64, 12, 1, 0, (1, 2), use_scale_shift_norm=True)
    m.load_state_dict(synthetic_state_dict_like(m, 1234))
    m = m.to(dev).eval()
    H, W, D = 92, 128, 92
    x_host = torch.randn(1, 12, H + D, W + D).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    kd = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="20")

    def once(stamps=None):
        t0 = time.perf_counter()
        xin = x_host.to(dev, non_blocking=True)
        t1 = time.perf_counter()
        res = kd.p_sample_loop(m, list(x_host.shape), noise=xin, model_kwargs=dict(H=H, W=W, D=D), seed=1234)
        t2 = time.perf_counter()
        out_host.copy_(res, non_blocking=True)
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if stamps is not None:
            stamps.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
    with torch.no_grad():
        for _ in range(3):
            once()
        st = []
        for _ in range(20):
            once(st)
    import numpy as np
    a = np.array(st) * 1e6
    print("us: h2d enqueue %.0f | p_sample_loop host %.0f | d2h enqueue %.0f | final sync wait %.0f | total %.0f" % tuple(np.median(a, axis=0)))
    # finer: profile the host side of p_sample_loop
    import cProfile
    import pstats
    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        for _ in range(50):
            kd.p_sample_loop(m, list(x_host.shape), noise=x_host.to(dev), model_kwargs=dict(H=H, W=W, D=D), seed=1234)
        pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
