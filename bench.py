#!/usr/bin/env python
"""Benchmark of the triplane denoising path (BASELINE.json: "triplane denoising steps/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg1]

A "step" is one denoising iteration (UNet forward + scheduler update) of the local batch.  Default workload is
BASELINE.json configs[1]: DDPM-1000, default triplane C=12, (H,W,D)=(92,128,92), batch 1 per GPU.
  value  : steps/s with x_t resident in HBM (s3d_sample_loop, CUDA-graph replay, device timed with CUDA events)
  e2e    : the same K steps through the public API (SpacedDiffusion.p_sample_loop) with x_T coming from pinned host
           memory and the final sample copied back to the host inside the timed region
  roofline: the tcgen05 conv kernel (dominant): dense algorithmic FLOPs / CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline: the CPU oracle port (oracle/, torch CPU, all host threads) on a bounded sample of the same workload
`--impl reference` times that CPU path alone (the reference is pure Python/PyTorch and cannot travel to the GPU box;
the oracle is its bit-exact restatement, see oracle/make_golden.py).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (C, (H, W, D), per-GPU batch, sampler, T)
    "cfg1": dict(C=8, HWD=(32, 32, 32), B=1, sampler="ddim", T=1000, desc="DDIM, 3x8chx32x32 triplane, B=1"),
    "cfg2": dict(C=12, HWD=(92, 128, 92), B=1, sampler="ddpm", T=1000,
                 desc="DDPM-1000, default triplane C=12 (H,W,D)=(92,128,92), B=1 per GPU"),
    "cfg3": dict(C=12, HWD=(92, 128, 138), B=8, sampler="ddim", T=1000,
                 desc="DDIM-100 shape, --resize 1 1 1.5 -> (92,128,138), B=8 per GPU"),
    "cfg5": dict(C=12, HWD=(92, 128, 92), B=8, sampler="ddpm", T=1000, desc="DDPM-1000, cfg2 shape, B=8 per GPU"),
}
METRIC = "triplane denoising steps/sec (DDPM-1000, DDIM-100) at 1/2/4/8 B200"


def dense_gflop_per_step(Cc, H, W, D, B):
    """dense conv FLOPs as the reference executes them (rollout channels counted)."""
    a0 = H * W + H * D + W * D
    a1 = (H // 2) * (W // 2) + (H // 2) * (D // 2) + (W // 2) * (D // 2)
    m0 = 128 * Cc + 675840
    m1 = 1556480
    return 2.0 * B * (a0 * m0 + a1 * m1) / 1e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm=d.get("hbm_gbs"), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU path
def cpu_steps_per_s(wl, max_steps, warmup, budget_s):
    """Oracle port (bit-exact restatement of the reference's PyTorch path) on the host cores.
    -> (steps/s, steps timed, seconds, threads used).  The thread count is the fastest of {8,16,32,64,all cores}
    (oneDNN convolutions of this size slow down badly when oversubscribed), probed with one step each."""
    import torch
    from oracle import diffusion_ref as dr
    from oracle import unet_ref as ur
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    spec = ur.UNetSpec(in_channels=Cc, model_channels=64, out_channels=Cc)
    sd = ur.synthetic_state_dict(spec, 1234)
    o = dr.RefDiffusion(wl["T"], "")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cc, H + D, W + D, generator=g)
    model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
    fn = o.ddim_sample if wl["sampler"] == "ddim" else o.p_sample
    state = dict(x=x, i=o.num_timesteps - 1)

    def one_step():
        i = state["i"]
        t = torch.full((B,), i, dtype=torch.long)
        state["x"] = fn(model, state["x"], t, torch.randn(x.shape, generator=g))["sample"]
        state["i"] = i - 1 if i > 0 else o.num_timesteps - 1

    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], None
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            one_step()
            t0 = time.perf_counter()
            one_step()
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = c, dt
            elif dt > 1.5 * best_t:
                break
        torch.set_num_threads(best)
        for _ in range(warmup):
            one_step()
        done, t0 = 0, time.perf_counter()
        while done < max_steps:
            one_step()
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return done / dt, done, dt, best


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, done, dt, thr = cpu_steps_per_s(wl, args.steps, args.warmup, budget_s=200.0)
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    line = dict(metric=METRIC, value=sps, unit="steps/s", n_gpus=args.gpus, steps=done, warmup=args.warmup,
                ms_per_step=1e3 / sps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference",
                config=dict(workload=f"{args.workload}: {wl['desc']}", device="host CPU", truncated=done < args.steps),
                cpu_baseline=dict(value=sps, unit="steps/s", cores=thr, kind="port",
                                  sample=f"{done} consecutive {wl['sampler'].upper()} steps of {args.workload} after "
                                         f"{args.warmup} warm-up steps, torch CPU oracle port, {thr} threads "
                                         f"(fastest of 8/16/32/64/{os.cpu_count()})"),
                e2e=dict(value=sps, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gflops=dense_gflop_per_step(Cc, H, W, D, B) * sps)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU path
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from oracle import unet_ref as ur          # synthetic weight recipe only (shared with the CPU baseline)
    import sin3dm_b200 as s3
    from sin3dm_b200 import _lib
    from sin3dm_b200.script_util import create_gaussian_diffusion

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    model = s3.TriplaneUNetModelSmall(Cc, 64, Cc, 1, 0, (1, 2), use_scale_shift_norm=True)
    if rank == 0:
        model.load_state_dict(ur.synthetic_state_dict(ur.UNetSpec(in_channels=Cc, model_channels=64, out_channels=Cc), 1234))
    model = model.to(dev).eval()
    if world > 1:
        # one broadcast of the flattened checkpoint over NVLink, no collective in the step loop
        from sin3dm_b200.dist import broadcast_parameters
        broadcast_parameters(model, src=0)

    K, Wm = args.steps, max(args.warmup, 3)
    L = _lib.lib()
    kind = _lib.DDIM if wl["sampler"] == "ddim" else _lib.DDPM

    def chain_diffusion(nsteps):
        return create_gaussian_diffusion(predict_xstart=True, timestep_respacing="" if nsteps >= 1000 else str(nsteps))

    full = chain_diffusion(1000)
    coef = full.coef_table(dev)
    film = model.film_table(full._model_timesteps(torch.arange(1000, device=dev)).float())
    g = torch.Generator().manual_seed(rank)
    x_host = torch.randn(B, Cc, H + D, W + D, generator=g).pin_memory()
    x = x_host.to(dev)
    h = model.handle()

    def device_steps(n):
        """n steps of the DDPM-1000 chain starting from t = 999 (wraps every 1000), x resident."""
        left = n
        while left > 0:
            m = min(left, 1000)
            a = _lib.LoopArgs()
            a.kind, a.mean_type, a.clip_denoised, a.n_steps = kind, _lib.START_X, 1, m
            a.B, a.H, a.W, a.D = B, H, W, D
            a.x_dev = x.data_ptr()
            a.coef_dev = coef.data_ptr() + (1000 - m) * 12 * 4
            a.film_dev = film.data_ptr() + (1000 - m) * film.shape[1] * 4
            a.seed, a.sample_base, a.use_graph = 1234, rank * B, 1
            _lib.check(L.s3d_sample_loop(h, C.byref(a), _lib.current_stream_ptr()))
            left -= m

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (builds the plan, captures the graph)
    device_steps(Wm)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    device_steps(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches_per_step = L.s3d_unet_last_launches(h)
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * K / (ms / 1e3)

    # ---- end to end through the public API: pinned host x_T -> p_sample_loop -> host
    kd = chain_diffusion(min(K, 1000))
    reps = max(1, K // kd.num_timesteps)
    fn = kd.ddim_sample_loop if wl["sampler"] == "ddim" else kd.p_sample_loop
    out_host = torch.empty_like(x_host).pin_memory()

    def e2e_once():
        xin = x_host.to(dev, non_blocking=True)
        res = fn(model, list(x_host.shape), noise=xin, model_kwargs=dict(H=H, W=W, D=D), seed=1234, sample_base=rank * B)
        out_host.copy_(res, non_blocking=True)

    with torch.no_grad():
        e2e_once()          # warm-up: film table, coefficient table, graph for these pointers
        barrier()
        e0.record()
        for _ in range(reps):
            e2e_once()
        e1.record()
        barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_steps = reps * kd.num_timesteps
    e2e_value = world * e2e_steps / (float(ms2.item()) / 1e3)
    nbytes = x_host.numel() * 4

    # ---- roofline leg: CUDA events around every op of a step (same process, after the timed region)
    roof, per_kernel = None, {}
    if rank == 0:
        nops = L.s3d_unet_op_count(h)
        msv = (C.c_float * nops)()
        device_steps(3)
        torch.cuda.synchronize()
        _lib.check(L.s3d_unet_profile_ops(h, 20, msv, _lib.current_stream_ptr()))
        conv_ms = conv_fl = tot_ms = 0.0
        op_rows = []
        for i in range(nops):
            nm, fl = C.c_char_p(), C.c_double()
            _lib.check(L.s3d_unet_op_info(h, i, C.byref(nm), C.byref(fl)))
            k = nm.value.decode()
            op_rows.append(f"{i:3d} {k:24s} {msv[i] * 1e3:9.2f} us {fl.value / 1e9:9.3f} GFLOP")
            e = per_kernel.setdefault(k, dict(launches=0, ms=0.0, dense_gflop=0.0))
            e["launches"] += 1
            e["ms"] += msv[i]
            e["dense_gflop"] += fl.value / 1e9
            tot_ms += msv[i]
            if k == "k_conv_tc":
                conv_ms += msv[i]
                conv_fl += fl.value
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                f.write("\n".join(op_rows) + f"\nsum {tot_ms * 1e3:.2f} us; step in the loop {ms / K * 1e3:.2f} us\n")
        peaks = load_peaks()
        if conv_ms > 0:
            ach = conv_fl / (conv_ms * 1e-3) / 1e12
            traffic = None
            try:        # DRAM bytes per conv launch from the committed ncu --set full capture of this workload
                tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_conv_traffic.json")))
                if tj["workload"] == args.workload and B == 1:
                    traffic = tj["traffic_bytes_per_launch"]
            except (OSError, KeyError, ValueError):
                pass
            roof = dict(bound="tensor", kernel="k_conv_tc (8 launches/step, fp16x3 split => 3 tcgen05.mma per dense MAC tile)",
                        achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"], traffic=traffic,
                        traffic_unit="DRAM bytes per k_conv_tc launch (mean of the 8 launches of a step; profiles/r1d_step_ncu_full.md)",
                        peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
                        flops_per_step_dense=conv_fl, conv_ms_per_step=conv_ms, conv_share_of_event_timed_step=conv_ms / tot_ms,
                        timing="graph replay with event-record nodes" if L.s3d_unet_profile_mode(h) == 1 else "eager launches",
                        note="achieved = dense algorithmic conv FLOPs (rollout channels counted, SURVEY §8d) / summed CUDA-event "
                             "time of the conv launches of one step; executed MMA FLOPs are the same number (1/3 after the "
                             "exact rollout fold, x3 for the hi/lo split)")
        for e in per_kernel.values():
            e["ms"] = round(e["ms"], 5)
            e["dense_gflop"] = round(e["dense_gflop"], 3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, done, dt, thr = cpu_steps_per_s(wl, 200, 3, budget_s=20.0)
        cpu = dict(value=sps, unit="steps/s", cores=thr, kind="port",
                   sample=f"{done} consecutive {wl['sampler'].upper()} steps of {args.workload} ({dt:.1f} s) after 3 warm-up steps; "
                          f"oracle port = bit-exact torch-CPU restatement of the reference, {thr} threads "
                          f"(fastest of 8/16/32/64/{os.cpu_count()} probed)")
    ws_mb = L.s3d_unet_workspace_bytes(h) / 2 ** 20
    if rank == 0:
        gf = dense_gflop_per_step(Cc, H, W, D, B)
        line = dict(
            metric=METRIC, value=value, unit="steps/s", n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms / K,
            higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="fp16 hi/lo split operands (3 tcgen05 MMAs, fp32-grade), fp32 accumulate / norm / scheduler",
            data="synthetic",
            config=dict(workload=f"{args.workload}: {wl['desc']}", sampler=wl["sampler"], per_gpu_batch=B, global_batch=B * world,
                        parallelism=f"sample-sharded x{world} (independent chains, weights NCCL-broadcast once)",
                        sample_steps_per_s=value * B, dense_gflop_per_step=gf, dense_tflops=gf * value / world / 1e3,
                        l2="inputs larger than L2: one step streams %.0f MiB of plan workspace + 28 MiB weights through a "
                           "126 MB L2; steps run back to back exactly as in a sampling run (no flush)" % ws_mb,
                        workspace_mib=round(ws_mb, 1), cuda_graph=True),
            clocks=clk,
            e2e=dict(value=e2e_value, unit="steps/s", h2d_bytes_per_step=nbytes / kd.num_timesteps,
                     d2h_bytes_per_step=nbytes / kd.num_timesteps, steps=e2e_steps,
                     api=f"SpacedDiffusion.{'ddim' if wl['sampler'] == 'ddim' else 'p'}_sample_loop x{reps} "
                         f"({kd.num_timesteps}-step chain each; x_T from pinned host memory, final sample copied back)"),
            gpu_launches=K * launches_per_step, launches_per_step=launches_per_step,
            roofline=roof, kernels=per_kernel, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch steady-state times of one step to this file")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, wl)


if __name__ == "__main__":
    main()
