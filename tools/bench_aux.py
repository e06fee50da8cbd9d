#!/usr/bin/env python
"""HBM-bound helper kernels against the measured copy bandwidth: fused AdamW+EMA, variational-bound terms, per-plane MSE, q_sample.

    python tools/bench_aux.py            # one JSON line per kernel: GB/s = algorithmic bytes / CUDA-event time, frac of MEASURED_PEAKS hbm
AdamW: 7.1 M parameters, every timed call preceded by a write of a 256 MiB scratch buffer (L2 flush).  Latent passes: a batch-32 cfg3
latent (94 MB per tensor, every call streams more than the L2 holds), called back to back through the C ABI."""
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

    # ---- scheduler-side passes on a batch-32 cfg3 latent [32, 12, 230, 266] (94 MB per tensor: every call streams 2-5 tensors, i.e.
    # more than the 126 MB L2 — no flush needed); the kernels are called through the C ABI back to back, so the time is device time
    shape = (32, 12, 230, 266)
    H, W, D = 92, 128, 138
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.rand(shape, device="cuda", generator=g) * 2 - 1
    nz = torch.randn(shape, device="cuda", generator=g)
    mo = torch.randn(shape, device="cuda", generator=g)
    xt = torch.randn(shape, device="cuda", generator=g)          # x_t: its own buffer (every stream of k_vb_terms is distinct memory)
    out = torch.empty_like(x0)
    t = torch.randint(1, 1000, (shape[0],), device="cuda").to(torch.int32)
    d = create_gaussian_diffusion(predict_xstart=True)
    coef = d.coef_table(x0.device)
    d._vb_device(lambda xx, tt, **k: mo[:1], x0[:1], x0[:1], t[:1], True, None)          # builds the log-variance table
    logvar = d._coef_cache[("logvar", str(x0.device))]
    nb, n = x0.numel() * 4, x0[0].numel()
    L = _lib.lib()
    st = _lib.current_stream_ptr()
    P = lambda v: C.c_void_p(v.data_ptr())

    def timed_dev(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    report("k_q_sample", timed_dev(lambda: _lib.check(L.s3d_q_sample(P(x0), P(nz), P(out), P(coef), P(t), shape[0], n, st))), 3 * nb,
           "cfg3 latent, B=32: 2 reads + 1 write")
    ws = torch.empty(L.s3d_vb_workspace_bytes(shape[0], n), dtype=torch.uint8, device="cuda")
    res = torch.empty(shape[0], 3, device="cuda")
    a = _lib.VbArgs()
    a.mean_type, a.clip_denoised, a.B, a.n_per_sample = _lib.START_X, 1, shape[0], n
    a.x_start, a.x_t, a.model_out, a.noise, a.pred_xstart = x0.data_ptr(), xt.data_ptr(), mo.data_ptr(), nz.data_ptr(), out.data_ptr()
    a.coef_dev, a.logvar_dev, a.t_idx_dev, a.workspace, a.out = coef.data_ptr(), logvar.data_ptr(), t.data_ptr(), ws.data_ptr(), res.data_ptr()
    report("k_vb_terms", timed_dev(lambda: _lib.check(L.s3d_vb_terms(C.byref(a), st))), 5 * nb,
           "cfg3 latent, B=32: x_start, x_t, model_out, noise read, pred_xstart written (+ k_vb_finalize)")
    report("k_plane_mse", timed_dev(lambda: _lib.check(L.s3d_plane_mse(P(x0), P(mo), shape[0], 12, H, W, D, P(ws), P(res), st))), 2 * nb,
           "cfg3 latent, B=32: 2 reads (+ k_plane_mse_finalize)")


if __name__ == "__main__":
    main()
