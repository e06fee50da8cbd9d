#!/usr/bin/env python
"""Benchmark of the triplane decoder: latent planes -> SDF + texture on a grid.

    python tools/bench_decoder.py [--reso 256] [--iters 5] [--no-cpu-baseline]

Workload: the cfg2 latent (C=12, (H,W,D)=(92,128,92)) decoded with decode_grid at `reso` over the towerruins-shaped aabb
(reso 256 -> 184 x 256 x 184 = 8.67 M points, what decode_texmesh does at sample.py time).  One JSON line:
  value       points/s of the fused decode kernel, feature planes resident (CUDA events, mean of `iters` launches)
  e2e         points/s from host latent planes to the host [nx,ny,nz,4] grid: H2D + the two feature-plane blocks + decode + D2H
  roofline    tensor: algorithmic MLP FLOPs per point x points / kernel time vs MEASURED_PEAKS.json
  cpu_baseline the oracle (bit-exact restatement of AutoEncoderGroupSkip.decode) on the host cores over 4 chunks of 16384
              points, feature planes recomputed per chunk exactly as the reference's decode_batch does
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

AABB = [-0.71875, -1.0, -0.71875, 0.71875, 1.0, 0.71875]      # 92 : 128 : 92


def mlp_flops_per_point(spec):
    up, hid, nh = spec.feat_channel_up, spec.mlp_hidden_channels, spec.mlp_hidden_layers
    per = lambda out: up * hid + (nh // 2) * hid * hid + (up + hid) * hid + max(nh // 2 - 1, 0) * hid * hid + hid * out
    macs = per(1) + (per(spec.tex_channels) if spec.use_tex else 0)
    return 2.0 * macs


def run(reso=256, iters=5, cpu_baseline=True, precision=3, device=0):
    import types
    import torch
    from sin3dm_b200.encoding import AutoEncoderGroupSkip, TriplaneDecoder, sample_grid_points_axes
    from sin3dm_b200.synthetic import synthetic_state_dict_like
    import bench as B

    torch.cuda.set_device(device)
    spec = types.SimpleNamespace(feat_channel_up=64, mlp_hidden_channels=256, mlp_hidden_layers=4, tex_channels=3, use_tex=True)
    net = AutoEncoderGroupSkip(4, 8, 64, 256, 4, use_tex=True, tex_channels=3)
    sd = synthetic_state_dict_like(net, 1234)       # the measured arm never imports oracle/ (only the cpu_baseline leg below does)
    sd["aabb"] = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    net.load_state_dict(sd)
    net.s3d_precision = precision
    net = net.cuda().eval()
    dec = TriplaneDecoder(net)
    H, W, D = 92, 128, 92
    g = torch.Generator().manual_seed(0)
    host_maps = [torch.tanh(torch.randn(1, 12, a, b, generator=g)).pin_memory() for a, b in ((H, W), (H, D), (W, D))]
    aabb = torch.tensor(AABB)
    xs, ys, zs = sample_grid_points_axes(aabb, reso)
    n = xs.numel() * ys.numel() * zs.numel()
    maps = [m.cuda() for m in host_maps]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- kernel only: planes resident
    out = dec.decode_grid(maps, reso, aabb=aabb)
    torch.cuda.synchronize()
    for _ in range(2):
        dec.decode_grid(maps, reso, aabb=aabb)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        out = dec.decode_grid(maps, reso, aabb=aabb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters

    # ---- feature planes alone (set_planes: 8 launches), forced by rebinding a fresh copy of the latent
    def rebind():
        fresh = [m.clone() for m in maps]
        net._bind_planes(fresh)
    rebind()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        rebind()
    e1.record()
    torch.cuda.synchronize()
    ms_planes = e0.elapsed_time(e1) / iters

    # ---- end to end: pinned host latent -> device -> planes + decode -> pinned host grid
    out_host = torch.empty(out.shape, dtype=torch.float32).pin_memory()

    def e2e_once():
        dm = [m.to("cuda", non_blocking=True) for m in host_maps]
        out_host.copy_(dec.decode_grid(dm, reso, aabb=aabb), non_blocking=True)
    e2e_once()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        e2e_once()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / iters

    peaks = B.load_peaks()
    fl = mlp_flops_per_point(spec)
    ach = fl * n / (ms * 1e-3) / 1e12
    line = dict(
        metric="triplane decode points/sec (decode_grid)", value=n / (ms * 1e-3), unit="points/s", n_gpus=1, steps=iters,
        ms_per_step=ms, higher_is_better=True, data="synthetic",
        dtype="fp16 hi/lo split operands (3 tcgen05 MMAs per tile-K, fp32-grade), fp32 accumulate" if precision == 3 else "fp16 operands, fp32 accumulate",
        config=dict(workload=f"decode_grid reso {reso}: cfg2 latent C=12 (92,128,92) -> {xs.numel()}x{ys.numel()}x{zs.numel()} = {n} points, "
                             "AutoEncoderGroupSkip defaults (fdim_up 64, hidden 256, 4 hidden layers, sdf + rgb)",
                    points=n, feature_planes_ms=ms_planes, l2="feature planes 16.4 MB + weights 2.4 MB stay L2-resident; output "
                    f"{n * 16 / 2**20:.0f} MiB streams to HBM"),
        e2e=dict(value=n / (ms_e2e * 1e-3), unit="points/s", ms=ms_e2e, h2d_bytes_per_step=sum(m.numel() for m in host_maps) * 4,
                 d2h_bytes_per_step=n * 16, api="TriplaneDecoder.decode_grid (host latent planes -> host [nx,ny,nz,4] grid)"),
        gpu_launches=iters, launches_per_step=1,
        roofline=dict(bound="tensor", kernel="k_dec_mlp_tc", achieved=ach, peak=peaks["tflops"], unit="TFLOP/s",
                      frac=ach / peaks["tflops"], traffic=None, flops_per_point=fl,
                      executed_flops_per_point=fl * (3 if precision == 3 else 1),
                      note="achieved = algorithmic MLP FLOPs (2 x MACs of the 12 Linear layers, 1.18 MFLOP/point) x points / "
                           "kernel time; the hi/lo split executes 3x that on the tensor cores",
                      peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})"))
    if cpu_baseline:
        from oracle import decoder_ref as de            # the checker / CPU baseline: the oracle on the same weights
        spec = de.DecoderSpec()
        pts = de.grid_points(aabb, reso).view(-1, 3)
        sel = pts[: 4 * 2 ** 14]
        cm = [m.clone() for m in host_maps]
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        with torch.no_grad():
            de.decode_batch(sd, spec, cm, sel[: 2 ** 14], aabb=aabb, hoist=False)
            t0 = time.perf_counter()
            want = de.decode_batch(sd, spec, cm, sel, aabb=aabb, hoist=False)
            dt = time.perf_counter() - t0
        got = out.view(-1, 4)[: sel.shape[0]].cpu()
        rel = float((got.double() - want.double()).norm() / want.double().norm())
        line["cpu_baseline"] = dict(value=sel.shape[0] / dt, unit="points/s", cores=torch.get_num_threads(), kind="port",
                                    sample=f"{sel.shape[0]} grid points in 4 chunks of 16384 ({dt:.2f} s), feature planes recomputed "
                                           "per chunk as the reference's decode_batch does (model.py:319-333)")
        line["parity_vs_oracle_rel_l2"] = rel
    return line


def run_encode(iters=10, cpu_baseline=True, device=0):
    """Encoder half (AutoEncoderGroupSkip.encode): the 184 x 256 x 184 sdf+rgb volume of the cfg2 latent -> three planes."""
    import torch
    from sin3dm_b200.encoding import AutoEncoderGroupSkip
    from sin3dm_b200.synthetic import synthetic_state_dict_like
    import bench as B

    torch.cuda.set_device(device)
    net = AutoEncoderGroupSkip(4, 8, 64, 256, 4, use_tex=True, tex_channels=3)
    sd = synthetic_state_dict_like(net, 1234)
    sd["aabb"] = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    X, Y, Z = 184, 256, 184
    g = torch.Generator().manual_seed(0)
    host_vol = (torch.rand(1, 4, X, Y, Z, generator=g) * 2 - 1).pin_memory()
    vol = host_vol.cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        planes = net.encode(vol)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        planes = net.encode(vol)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    host_out = [torch.empty(p.shape).pin_memory() for p in planes]

    def e2e_once():
        v = host_vol.to("cuda", non_blocking=True)
        for h, p in zip(host_out, net.encode(v)):
            h.copy_(p, non_blocking=True)
    e2e_once()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        e2e_once()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / iters
    peaks = B.load_peaks()
    H, W, D = X // 2, Y // 2, Z // 2
    nvox = H * W * D
    in_bytes = host_vol.numel() * 4
    out_bytes = sum(p.numel() for p in planes) * 4
    flops = 2.0 * nvox * (64 * 4 + 64 * 4 * 8)
    gbs = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9
    line = dict(metric="volume encode output voxels/sec (AutoEncoderGroupSkip.encode)", value=nvox / (ms * 1e-3), unit="voxels/s",
                n_gpus=1, steps=iters, ms_per_step=ms, higher_is_better=True, data="synthetic", dtype="f32",
                config=dict(workload=f"encode: sdf+rgb volume 4 x {X} x {Y} x {Z} -> planes 12 x ({H},{W},{D})", voxels=nvox,
                            gflop=flops / 1e9, fp32_tflops=flops / (ms * 1e-3) / 1e12),
                e2e=dict(value=nvox / (ms_e2e * 1e-3), unit="voxels/s", ms=ms_e2e, h2d_bytes_per_step=in_bytes,
                         d2h_bytes_per_step=out_bytes, api="AutoEncoderGroupSkip.encode (host volume -> host planes)"),
                gpu_launches=2 * iters, launches_per_step=2,
                roofline=dict(bound="hbm", kernel="k_enc_conv3d", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                              traffic=None, note="algorithmic bytes = the volume read once + the three planes written "
                              f"({(in_bytes + out_bytes) / 1e6:.1f} MB); the kernel also executes {flops / 1e9:.2f} GFLOP of fp32 FFMA "
                              "(2304 per output voxel), which is what bounds it on the CUDA cores",
                              peak_source=f"MEASURED_PEAKS.json hbm_gbs ({peaks['src']})"))
    if cpu_baseline:
        from oracle import decoder_ref as de            # the checker / CPU baseline
        spec = de.DecoderSpec()
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        with torch.no_grad():
            t0 = time.perf_counter()
            want = de.encode(sd, spec, host_vol)
            dt = time.perf_counter() - t0
        err = max(float((a.cpu() - b).abs().max()) for a, b in zip(planes, want))
        line["cpu_baseline"] = dict(value=nvox / dt, unit="voxels/s", cores=torch.get_num_threads(), kind="port",
                                    sample=f"one full encode of the same volume ({dt:.2f} s)")
        line["parity_vs_oracle_max_abs"] = err
    return line


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--encode", action="store_true", help="benchmark the encoder half instead of decode_grid")
    ap.add_argument("--reso", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.encode:
        print(json.dumps(run_encode(a.iters, not a.no_cpu_baseline)), flush=True)
        sys.exit(0)
    print(json.dumps(run(a.reso, a.iters, not a.no_cpu_baseline, a.precision)), flush=True)
