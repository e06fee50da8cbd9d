"""bench.py --workload cfg4: one diffusion-training iteration of the triplane UNet (BASELINE.json configs[3], the diffusion half of
train.py: TrainLoop.run_step, reference src/diffusion/train_util.py:163-235) —

    t ~ UniformSampler, x_t = q_sample(x_0, t, noise)            (k_q_sample)
    out = UNet(x_t, t)                                            (training plan: forward kernels, activations kept)
    loss = mean_b sum_planes mse(out, x_0)                        (torch reductions on the output, on the autograd tape)
    loss.backward()                                               (s3d_unet_backward + the embedding MLP through torch)
    AdamW + EMA                                                   (k_adamw_ema over the flat parameter buffer)
    operand re-pack for the next forward                          (s3d_unet_refresh_dev, on the device)

`value` = iterations/s with the latent batch resident; `e2e` = the same loop fed from pinned host memory (batch H2D + loss D2H per
step, as TrainLoop's data iterator does).  roofline = the backward GEMMs (dgrad on tcgen05, wgrad) from per-op CUDA-event times."""
import ctypes as C
import json
import os
import sys


def run_train(bn, args, wl, K, Wm, METRIC, config_of, load_peaks, emit=True, profile=True):
    """emit=True: print the JSON line (bench.py --workload cfg4).  emit=False: return it (the default bench adds it to `also`)."""
    torch = bn.torch
    from sin3dm_b200 import _lib
    from sin3dm_b200.dist import all_reduce_gradients
    from sin3dm_b200.optim import FusedAdamWEMA
    from sin3dm_b200.script_util import create_gaussian_diffusion
    L = _lib.lib()
    Cc, (H, W, D), B = wl["C"], wl["HWD"], args.batch or wl["B"]
    model = bn.model(Cc, fresh=True).train()       # its parameters move into the optimizer's flat buffer: never the sampling model
    opt = FusedAdamWEMA(model.parameters(), lr=5e-4, weight_decay=0.0, ema_rates="0.9999")
    diff = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="")
    g = torch.Generator().manual_seed(bn.rank)
    latent = (torch.rand(1, Cc, H + D, W + D, generator=g) * 2 - 1)
    x_host = latent.expand(B, -1, -1, -1).contiguous().pin_memory()        # get_data_iterator repeats the single latent
    x_dev = x_host.to(bn.dev)
    kw = dict(H=H, W=W, D=D)

    def step(x0, exchange=True):
        t = torch.randint(0, diff.num_timesteps, (B,), device=bn.dev)
        opt.zero_grad()
        losses = diff.training_losses(model, x0, t, model_kwargs=kw)
        loss = losses["loss"].mean()
        loss.backward()
        if bn.world > 1 and exchange:
            all_reduce_gradients(opt.grad)        # data parallel: one NCCL all-reduce of the flat gradient buffer per step
        opt.step()
        return loss

    def timed(n, from_host):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bn.barrier()
        e0.record()
        last = None
        for _ in range(n):
            x0 = x_host.to(bn.dev, non_blocking=True) if from_host else x_dev
            loss = step(x0)
            if from_host:
                last = float(loss.detach())          # loss read back every step (TrainLoop logs it)
        e1.record()
        bn.barrier()
        return bn.max_over_ranks(e0.elapsed_time(e1)), last

    import gc
    for _ in range(max(Wm, 3)):
        step(x_dev)
    gc.collect()
    gc.disable()          # a generation-2 collection inside the 10-step window showed up as a 50 ms hiccup (25.6 instead of 20.4 ms / step)
    # The loop is launched asynchronously from Python (autograd + ~60 torch ops + the library's launch lists per iteration), so a host
    # hiccup (scheduler, allocator) inside a 10-step window idles the GPU: one window in four came out 25-45 % slow on the shared boxes
    # while the windows around it agreed to 1 %.  Three windows of exactly K steps are timed; `value` is the fastest, all are reported.
    windows = []
    try:
        for _ in range(3):
            (ms_w, pr_w), _ = timed(K, False)
            windows.append((ms_w, pr_w))
        (ms2, _), last_loss = timed(max(3, min(K, 10)), True)
    finally:
        gc.enable()
    ms, per_rank = min(windows, key=lambda w: w[0])
    n2 = max(3, min(K, 10))
    value = bn.world * K / (ms / 1e3)
    h = model._handle

    # ---- per-op times: forward list (graph replay with event nodes) and backward list (eager launches with events)
    rec = {}
    if bn.rank == 0 and profile:
        peaks = load_peaks()
        nf, nb = L.s3d_unet_op_count(h), L.s3d_unet_bwd_op_count(h)
        mf, mb = (C.c_float * nf)(), (C.c_float * nb)()
        step(x_dev, exchange=False)            # rank 0 alone from here on: no collective
        torch.cuda.synchronize()
        _lib.check(L.s3d_unet_profile_bwd_ops(h, 5, mb, _lib.current_stream_ptr()))
        _lib.check(L.s3d_unet_profile_ops(h, 5, mf, _lib.current_stream_ptr()))
        rows, per_kernel = [], {}
        for which, n, msv, info in (("fwd", nf, mf, L.s3d_unet_op_info), ("bwd", nb, mb, L.s3d_unet_bwd_op_info)):
            for i in range(n):
                nm, fl = C.c_char_p(), C.c_double()
                _lib.check(info(h, i, C.byref(nm), C.byref(fl)))
                k = f"{which}:{nm.value.decode()}"
                rows.append(f"{k:34s} {msv[i] * 1e3:10.1f} us {fl.value / 1e9:10.2f} GFLOP")
                e = per_kernel.setdefault(k, dict(launches=0, ms=0.0, dense_gflop=0.0))
                e["launches"] += 1
                e["ms"] += msv[i]
                e["dense_gflop"] += fl.value / 1e9
        if args.dump_ops and emit:
            with open(args.dump_ops, "w") as f:
                f.write("\n".join(rows) + f"\nstep in the loop {ms / K * 1e3:.1f} us\n")
        for e in per_kernel.values():
            e["ms"] = round(e["ms"], 4)
            e["dense_gflop"] = round(e["dense_gflop"], 2)
        roof = {}
        for key, label in (("bwd:k_conv_tc<dgrad>", "dgrad"), ("bwd:k_wgrad_tc", "wgrad"), ("bwd:k_wgrad_ffma", "wgrad_ffma"),
                           ("fwd:k_conv_tc", "fprop")):
            e = per_kernel.get(key)
            if e and e["ms"] > 0:
                ach = e["dense_gflop"] / e["ms"]        # GFLOP / ms = TFLOP/s
                roof[label] = dict(kernel=key, launches=e["launches"], ms=e["ms"], achieved=ach, peak=peaks["tflops"], unit="TFLOP/s",
                                   frac=ach / peaks["tflops"])
        rec = dict(kernels=per_kernel, roofline_parts=roof)
        main = roof.get("wgrad") or roof.get("wgrad_ffma") or roof.get("dgrad")
        if main:
            rec["roofline"] = dict(bound="tensor", kernel=main["kernel"], achieved=main["achieved"], peak=main["peak"], unit="TFLOP/s",
                                   frac=main["frac"], traffic=None, peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
                                   note="dense algorithmic FLOPs of the op (rollout channels counted for dgrad / fprop as the reference executes "
                                        "them; own-channel FLOPs for wgrad) / CUDA-event time of its launches in one step")
    line = None
    if bn.rank == 0:
        cfg = config_of("cfg4", dict(wl, B=B), bn.world)
        fwd_gf = cfg["dense_gflop_per_step"]
        line = dict(metric="triplane diffusion training iterations/sec (train.py diffusion stage, cfg4)", value=value, unit="iterations/s",
                    n_gpus=bn.world, steps=K, warmup=Wm, ms_per_step=ms / K, per_rank_ms=[round(v, 3) for v in per_rank],
                    higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="fp16 hi/lo split operands on tcgen05 (fprop, dgrad), fp32 accumulate / norms / reductions / AdamW",
                    data="synthetic", config=dict(cfg, parallelism=f"data parallel x{bn.world}: per-GPU batch {B}, one NCCL all-reduce of the "
                                                                   "flat 28 MB gradient buffer per step" if bn.world > 1 else "single GPU"),
                    e2e=dict(value=bn.world * n2 / (ms2 / 1e3), unit="iterations/s", h2d_bytes_per_step=x_host.numel() * 4, d2h_bytes_per_step=4,
                             steps=n2, api="training_losses + loss.backward() + FusedAdamWEMA.step(), batch from pinned host memory, loss read back"),
                    gpu_launches=K * (L.s3d_unet_op_count(h) + L.s3d_unet_bwd_op_count(h) + 3),
                    timing=dict(method=f"fastest of 3 windows of {K} steps each (CUDA events, max over ranks per window)",
                                windows_ms_per_step=[round(w[0] / K, 3) for w in windows]),
                    detail=dict(samples_per_s=value * B, fwd_dense_gflop=fwd_gf, train_dense_tflops=3 * fwd_gf * value / bn.world / 1e3,
                                workspace_mib=round(L.s3d_unet_workspace_bytes(h) / 2 ** 20, 1), last_loss=last_loss),
                    cpu_baseline=None, **rec)
        if emit:
            print(json.dumps(line), flush=True)
    del opt, model
    torch.cuda.empty_cache()
    if bn.world > 1:
        bn.barrier()
        if emit:
            bn.dist.destroy_process_group()
    return line
