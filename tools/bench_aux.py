#!/usr/bin/env python
"""HBM-bound helper kernels against the measured copy bandwidth: fused AdamW+EMA, variational-bound terms, per-plane MSE, q_sample.

    python tools/bench_aux.py            # one JSON line per kernel: GB/s = algorithmic bytes / CUDA-event time, frac of MEASURED_PEAKS hbm
Buffers are sized like the real use (7.1 M parameters; the cfg3 batch-8 latent [8, 12, 230, 266]) and every timed call is preceded
by a write of a 256 MiB scratch buffer (L2 flush), so the numbers are HBM numbers."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench as B
    from sin3dm_b200 import _lib
    from sin3dm_b200.optim import FusedAdamWEMA
    from sin3dm_b200.script_util import create_gaussian_diffusion

    torch.cuda.set_device(0)
    peaks = B.load_peaks()
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            scratch.fill_(1)                       # flush L2
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    def report(name, ms, nbytes, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps(dict(kernel=name, ms=ms, algorithmic_bytes=nbytes, workload=what,
                              roofline=dict(bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                                            peak_source=f"MEASURED_PEAKS.json hbm_gbs ({peaks['src']})"))), flush=True)

    # ---- AdamW + one EMA over 7.1 M parameters (the UNet's size): 6 reads + 4 writes of 4 B per parameter
    n = 7_100_000
    p = [torch.nn.Parameter(torch.randn(n, device="cuda"))]
    opt = FusedAdamWEMA(p, lr=1e-4, weight_decay=0.0, ema_rates=[0.9999])
    opt.grad.normal_()
    report("k_adamw_ema", timed(lambda: opt.step()), opt.n * 4 * 10, "7.1 M parameters, 1 EMA copy")

    # ---- scheduler-side reductions on the cfg3 batch-8 latent
    shape = (8, 12, 230, 266)
    H, W, D = 92, 128, 138
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.rand(shape, device="cuda", generator=g) * 2 - 1
    nz = torch.randn(shape, device="cuda", generator=g)
    mo = torch.randn(shape, device="cuda", generator=g)
    t = torch.randint(0, 1000, (8,), device="cuda")
    d = create_gaussian_diffusion(predict_xstart=True)
    nb = x0.numel() * 4
    xt = d.q_sample(x0, t, nz)
    report("k_q_sample", timed(lambda: d.q_sample(x0, t, nz)), 3 * nb, "cfg3 latent, B=8: 2 reads + 1 write")
    model = lambda xx, tt, **k: mo
    report("k_vb_terms", timed(lambda: d._vb_device(model, x0, xt, t, True, None, noise=nz)), 5 * nb,
           "cfg3 latent, B=8: x_start, x_t, model_out, noise read, pred_xstart written")
    L = _lib.lib()
    ws = torch.empty(L.s3d_vb_workspace_bytes(8, x0[0].numel()), dtype=torch.uint8, device="cuda")
    mse = torch.empty(8, 3, device="cuda")
    report("k_plane_mse", timed(lambda: _lib.check(L.s3d_plane_mse(C.c_void_p(x0.data_ptr()), C.c_void_p(mo.data_ptr()), 8, 12, H, W, D,
                                                                   C.c_void_p(ws.data_ptr()), C.c_void_p(mse.data_ptr()),
                                                                   _lib.current_stream_ptr()))), 2 * nb, "cfg3 latent, B=8: 2 reads")


if __name__ == "__main__":
    main()
