#!/usr/bin/env python
"""Benchmark of the triplane denoising path (BASELINE.json: "triplane denoising steps/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5|cfg1|cfg4]

A "step" is one denoising iteration (UNet forward + scheduler update) of the local batch.  Default workload is
BASELINE.json configs[1]: DDPM-1000, default triplane C=12, (H,W,D)=(92,128,92), batch 1 per GPU.
  value  : steps/s with x_t resident in HBM (s3d_sample_loop, CUDA-graph replay, device timed with CUDA events).  The warm-up
           runs the SAME call (same buffers, same graph key), and the line asserts that no graph was captured while timing.
  e2e    : the same K steps through the public API (SpacedDiffusion.p_sample_loop) with x_T coming from pinned host
           memory and the final sample copied back to the host inside the timed region
  roofline: the tcgen05 conv kernel (dominant): dense algorithmic FLOPs / CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline: the reference's CPU path (oracle/_ref = the staged, unmodified reference modules when present, else the oracle
           port) on a bounded sample of the same workload
  gpu_library_baseline: the reference's arithmetic run by torch eager on the same GPU (cuDNN / cuBLAS), tf32 off and on
  also   : the other halves of the metric measured the same way — cfg3 (DDIM-100, D=138, B=8) and cfg5 (DDPM, B=8 per GPU)
`--workload cfg4` times the training step instead (q_sample -> UNet forward -> per-plane MSE -> backward -> AdamW + EMA).
`--impl reference` times the CPU path alone.
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
    # name: C, (H, W, D), per-GPU batch, sampler, chain length of one full run
    "cfg1": dict(C=8, HWD=(32, 32, 32), B=1, sampler="ddim", chain=10, desc="DDIM-10, 3x8chx32x32 triplane, B=1"),
    "cfg2": dict(C=12, HWD=(92, 128, 92), B=1, sampler="ddpm", chain=1000,
                 desc="DDPM-1000, default triplane C=12 (H,W,D)=(92,128,92), B=1 per GPU"),
    "cfg3": dict(C=12, HWD=(92, 128, 138), B=8, sampler="ddim", chain=100,
                 desc="DDIM-100, --resize 1 1 1.5 -> (92,128,138), B=8 per GPU"),
    "cfg4": dict(C=12, HWD=(92, 128, 92), B=32, sampler="train", chain=1000,
                 desc="diffusion UNet training step (train.py: q_sample + forward + backward + AdamW + EMA), (92,128,92), B=32 per GPU"),
    "cfg5": dict(C=12, HWD=(92, 128, 92), B=8, sampler="ddpm", chain=1000, desc="DDPM-1000, cfg2 shape, B=8 per GPU"),
}
METRIC = "triplane denoising steps/sec (DDPM-1000, DDIM-100) at 1/2/4/8 B200"
DTYPE = {3: "fp16 hi/lo split operands (3 tcgen05 MMAs, fp32-grade), fp32 accumulate / norm / scheduler",
         2: "fp16 activations x fp16 hi/lo split weights (one N=128 tcgen05 MMA per tile step), fp32 accumulate / norm / scheduler",
         4: "fp16 hi/lo split activations x fp16 weights (2 tcgen05 MMAs), fp32 accumulate / norm / scheduler",
         5: "fp16 activations x fp16 hi/lo split weights (one N=128 tcgen05 MMA per tile step; rollout terms with the full 3-term split), "
            "fp32 accumulate / norm / scheduler",
         1: "fp16 operands (1 tcgen05 MMA), fp32 accumulate / norm / scheduler"}


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


def config_of(name, wl, world):
    """The `config` object both arms print (same keys, same workload string)."""
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    return dict(workload=f"{name}: {wl['desc']}", sampler=wl["sampler"], per_gpu_batch=B, global_batch=B * world,
                parallelism=f"sample-sharded x{world} (independent chains, weights NCCL-broadcast once)",
                dense_gflop_per_step=dense_gflop_per_step(Cc, H, W, D, B),
                l2="GPU arm: no flush — the timed region is the steady state of a sampling chain, steps back to back, each streaming "
                   "the plan workspace (66 MiB at B=1, 669-864 MiB at B=8; `detail.workspace_mib`) + 28 MiB of packed weights "
                   "through the 126 MB L2, exactly as a sampling run does")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
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
        time.sleep(0.1)
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
def _cpu_stepper(wl):
    """-> (one_step(), kind): one sampler iteration of the workload on the host.  kind "reference": the unmodified reference
    modules staged by oracle/stage_ref.py (SpacedDiffusion.p_sample / ddim_sample driving TriplaneUNetModelSmall); kind "port":
    the oracle's bit-exact restatement."""
    import contextlib
    import io
    import torch
    from oracle import unet_ref as ur
    from oracle.stage_ref import staged_path
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    spec = ur.UNetSpec(in_channels=Cc, model_channels=64, out_channels=Cc)
    sd = ur.synthetic_state_dict(spec, 1234)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cc, H + D, W + D, generator=g)
    resp = "" if wl["chain"] >= 1000 else str(wl["chain"])
    ref = staged_path()
    if ref:
        sys.path.insert(0, ref)
        from diffusion import gaussian_diffusion as gd, respace, unet_triplane as ut
        with contextlib.redirect_stdout(io.StringIO()):
            m = ut.TriplaneUNetModelSmall(in_channels=Cc, model_channels=64, out_channels=Cc, num_res_blocks=1, dropout=0,
                                          channel_mult=(1, 2), use_scale_shift_norm=True)
        m.load_state_dict(sd)
        m.eval()
        d = respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, resp or [1000]),
                                    betas=gd.get_named_beta_schedule("linear", 1000), model_mean_type=gd.ModelMeanType.START_X,
                                    model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE, rescale_timesteps=False)
        fn = d.ddim_sample if wl["sampler"] == "ddim" else d.p_sample
        nT = d.num_timesteps
        state = dict(x=x, i=nT - 1)

        def one_step():
            i = state["i"]
            t = torch.tensor([i] * B)
            state["x"] = fn(m, state["x"], t, model_kwargs=dict(H=H, W=W, D=D))["sample"]
            state["i"] = i - 1 if i > 0 else nT - 1
        return one_step, "reference"
    from oracle import diffusion_ref as dr
    o = dr.RefDiffusion(1000, resp)
    model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
    fn = o.ddim_sample if wl["sampler"] == "ddim" else o.p_sample
    state = dict(x=x, i=o.num_timesteps - 1)

    def one_step():
        i = state["i"]
        t = torch.full((B,), i, dtype=torch.long)
        state["x"] = fn(model, state["x"], t, torch.randn(x.shape, generator=g))["sample"]
        state["i"] = i - 1 if i > 0 else o.num_timesteps - 1
    return one_step, "port"


def cpu_steps_per_s(wl, max_steps, warmup, budget_s):
    """-> (steps/s, steps timed, seconds, threads used, kind).  The thread count is the fastest of {8,16,32,64,all cores}
    (oneDNN convolutions of this size slow down badly when oversubscribed), probed with one step each."""
    import torch
    one_step, kind = _cpu_stepper(wl)
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
    return done / dt, done, dt, best, kind


def cpu_sample_text(kind, done, wl, name, warm, thr, dt=None):
    what = ("unmodified reference modules (oracle/_ref: SpacedDiffusion + TriplaneUNetModelSmall, torch CPU)" if kind == "reference"
            else "oracle port = bit-exact torch-CPU restatement of the reference")
    return (f"{done} consecutive {wl['sampler'].upper()} steps of {name}" + (f" ({dt:.1f} s)" if dt else "") +
            f" after {warm} warm-up steps; {what}, {thr} threads (fastest of 8/16/32/64/{os.cpu_count()} probed)")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if wl["sampler"] == "train":
        print(json.dumps(dict(impl="reference", unavailable="cfg4 (training step) has no CPU reference arm in bench.py")), flush=True)
        return
    sps, done, dt, thr, kind = cpu_steps_per_s(wl, args.steps, args.warmup, budget_s=200.0)
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    cfg = config_of(args.workload, wl, world)
    line = dict(metric=METRIC, value=sps, unit="steps/s", n_gpus=args.gpus, steps=done, warmup=args.warmup,
                ms_per_step=1e3 / sps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference", config=cfg, device="host CPU", truncated=done < args.steps,
                cpu_baseline=dict(value=sps, unit="steps/s", cores=thr, kind=kind,
                                  sample=cpu_sample_text(kind, done, wl, args.workload, args.warmup, thr)),
                e2e=dict(value=sps, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gflops=dense_gflop_per_step(Cc, H, W, D, B) * sps)
    print(json.dumps(line), flush=True)


def torch_eager_steps_per_s(wl, steps, tf32):
    """The "library kernels to beat" (SURVEY §8d): the reference's arithmetic (oracle restatement, functional torch) run eagerly on
    cuda (cuDNN / cuBLAS / ATen), one sampler step per iteration.  Baseline leg only — not the product path."""
    import numpy as np
    import torch
    from oracle import diffusion_ref as dr
    from oracle import unet_ref as ur
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    dev = torch.device("cuda", torch.cuda.current_device())
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    spec = ur.UNetSpec(in_channels=Cc, model_channels=64, out_channels=Cc)
    sd = {k: v.to(dev) for k, v in ur.synthetic_state_dict(spec, 1234).items()}
    x = torch.randn(B, Cc, H + D, W + D, device=dev)
    o = dr.RefDiffusion(1000, "")
    o._x = lambda arr, t, like: torch.from_numpy(np.asarray(arr))[t.cpu()].float().to(like.device).view(-1, *([1] * (like.dim() - 1)))
    o.model_t = lambda t: t
    model = lambda xx, tt: ur.unet_forward(sd, spec, xx, tt, H, W, D)
    fn = o.ddim_sample if wl["sampler"] == "ddim" else o.p_sample
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(i, img):
        t = torch.full((B,), i, dtype=torch.long, device=dev)
        return fn(model, img, t, torch.randn_like(img))["sample"]
    with torch.no_grad():
        img = x
        for i in range(3):
            img = step(999 - i, img)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            img = step(996 - i, img)
        e1.record()
        torch.cuda.synchronize()
    return steps / (e0.elapsed_time(e1) / 1e3)


# ------------------------------------------------------------------------------------------------ GPU path
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.models = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        """-> (max, [per-rank])"""
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world == 1:
            return ms, [ms]
        allv = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(allv, t)
        v = [float(a.item()) for a in allv]
        return max(v), v

    def model(self, Cc, fresh=False):
        import sin3dm_b200 as s3
        from sin3dm_b200.synthetic import synthetic_state_dict_like
        key = ("fresh", Cc) if fresh else Cc        # fresh: a model of its own (the training leg moves its parameters around)
        if key not in self.models:
            m = s3.TriplaneUNetModelSmall(Cc, 64, Cc, 1, 0, (1, 2), use_scale_shift_norm=True)
            if self.rank == 0:
                m.load_state_dict(synthetic_state_dict_like(m, 1234))
            m = m.to(self.dev).eval()
            if self.world > 1:
                # one broadcast of the flattened checkpoint over NVLink, no collective in the step loop
                from sin3dm_b200.dist import broadcast_parameters
                broadcast_parameters(m, src=0)
            self.models[key] = m
        return self.models[key] if not fresh else self.models.pop(key)

    def sampling(self, name, wl, K, Wm, roofline=True):
        """device-timed value, e2e and (rank 0) the per-op roofline leg of one sampling workload"""
        torch = self.torch
        from sin3dm_b200 import _lib
        from sin3dm_b200.script_util import create_gaussian_diffusion
        L = _lib.lib()
        Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
        model = self.model(Cc)
        kind = _lib.DDIM if wl["sampler"] == "ddim" else _lib.DDPM
        chain = wl["chain"]

        def chain_diffusion(nsteps):
            return create_gaussian_diffusion(predict_xstart=True, timestep_respacing="" if nsteps >= 1000 else str(nsteps))

        full = chain_diffusion(chain)
        T = full.num_timesteps
        coef = full.coef_table(self.dev)
        film = model.film_table(full._model_timesteps(torch.arange(T)).float(), cache=True)
        g = torch.Generator().manual_seed(self.rank)
        x_host = torch.randn(B, Cc, H + D, W + D, generator=g).pin_memory()
        x = x_host.to(self.dev)
        h = model.handle()

        def device_steps(n):
            """n steps of the chain starting from t = T-1 (wraps every T), x resident; always the same buffers"""
            left = n
            while left > 0:
                m = min(left, T)
                a = _lib.LoopArgs()
                a.kind, a.mean_type, a.clip_denoised, a.n_steps, a.t_start = kind, _lib.START_X, 1, m, T - 1
                a.B, a.H, a.W, a.D, a.n_per_sample = B, H, W, D, x[0].numel()
                a.x_dev, a.coef_dev, a.film_dev = x.data_ptr(), coef.data_ptr(), film.data_ptr()
                a.seed, a.sample_base, a.use_graph = 1234, self.rank * B, 1
                _lib.check(L.s3d_sample_loop(h, C.byref(a), _lib.current_stream_ptr()))
                left -= m

        # ---- warm-up: builds the plan and captures the graph of exactly the call that is timed
        device_steps(max(Wm, 3))
        self.barrier()
        builds0 = L.s3d_unet_graph_builds(h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        device_steps(K)
        e1.record()
        self.barrier()
        assert L.s3d_unet_graph_builds(h) == builds0, "a CUDA graph was captured inside the timed region"
        ms, per_rank = self.max_over_ranks(e0.elapsed_time(e1))
        launches_per_step = L.s3d_unet_last_launches(h)
        value = self.world * K / (ms / 1e3)

        # ---- end to end through the public API: pinned host x_T -> *_sample_loop -> host
        kd = chain_diffusion(min(K, chain))
        reps = max(1, K // kd.num_timesteps)
        fn = kd.ddim_sample_loop if wl["sampler"] == "ddim" else kd.p_sample_loop
        out_host = torch.empty_like(x_host).pin_memory()

        def e2e_once():
            xin = x_host.to(self.dev, non_blocking=True)
            res = fn(model, list(x_host.shape), noise=xin, model_kwargs=dict(H=H, W=W, D=D), seed=1234, sample_base=self.rank * B)
            out_host.copy_(res, non_blocking=True)

        with torch.no_grad():
            e2e_once()          # warm-up: conditioning table, coefficient table, graph for these buffers
            self.barrier()
            builds1 = L.s3d_unet_graph_builds(h)
            e0.record()
            for _ in range(reps):
                e2e_once()
            e1.record()
            self.barrier()
        assert L.s3d_unet_graph_builds(h) == builds1, "a CUDA graph was captured inside the e2e timed region"
        ms2, _ = self.max_over_ranks(e0.elapsed_time(e1))
        e2e_steps = reps * kd.num_timesteps
        nbytes = x_host.numel() * 4
        rec = dict(value=value, ms_per_step=ms / K, per_rank_ms=[round(v, 4) for v in per_rank], steps=K,
                   sample_steps_per_s=value * B, launches_per_step=launches_per_step,
                   e2e=dict(value=self.world * e2e_steps / (ms2 / 1e3), unit="steps/s", h2d_bytes_per_step=nbytes / kd.num_timesteps,
                            d2h_bytes_per_step=nbytes / kd.num_timesteps, steps=e2e_steps,
                            api=f"SpacedDiffusion.{'ddim' if wl['sampler'] == 'ddim' else 'p'}_sample_loop x{reps} "
                                f"({kd.num_timesteps}-step chain each; x_T from pinned host memory, final sample copied back)"),
                   workspace_mib=round(L.s3d_unet_workspace_bytes(h) / 2 ** 20, 1))

        # ---- roofline leg: CUDA events around every op of a step (same process, after the timed region)
        if roofline and self.rank == 0:
            nops = L.s3d_unet_op_count(h)
            msv = (C.c_float * nops)()
            device_steps(3)
            torch.cuda.synchronize()
            _lib.check(L.s3d_unet_profile_ops(h, 20, msv, _lib.current_stream_ptr()))
            conv_ms = conv_fl = tot_ms = 0.0
            per_kernel, op_rows = {}, []
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
            if self.args.dump_ops and name == self.args.workload:
                with open(self.args.dump_ops, "w") as f:
                    f.write("\n".join(op_rows) + f"\nsum {tot_ms * 1e3:.2f} us; step in the loop {ms / K * 1e3:.2f} us\n")
            peaks = load_peaks()
            if conv_ms > 0:
                ach = conv_fl / (conv_ms * 1e-3) / 1e12
                traffic = None
                try:        # DRAM bytes per conv launch from the committed ncu --set full capture of this workload
                    tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
                    traffic = tj.get(name, {}).get("traffic_bytes_per_launch")
                except (OSError, KeyError, ValueError):
                    pass
                prec = model.s3d_precision
                rec["roofline"] = dict(
                    bound="tensor", kernel=f"k_conv_tc ({per_kernel['k_conv_tc']['launches']} launches/step, precision mode {prec})",
                    achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"], traffic=traffic,
                    traffic_unit="DRAM bytes per k_conv_tc launch (mean over the launches of a step; profiles/)",
                    peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
                    flops_per_step_dense=conv_fl, conv_ms_per_step=conv_ms, conv_share_of_event_timed_step=conv_ms / tot_ms,
                    timing="graph replay with event-record nodes" if L.s3d_unet_profile_mode(h) == 1 else "eager launches",
                    note="achieved = dense algorithmic conv FLOPs (rollout channels counted, SURVEY §8d) / summed CUDA-event "
                         "time of the conv launches of one step; executed MMA FLOPs = dense / 3 (exact rollout fold) x the "
                         "number of MMA terms of the precision mode")
            for e in per_kernel.values():
                e["ms"] = round(e["ms"], 5)
                e["dense_gflop"] = round(e["dense_gflop"], 3)
            rec["kernels"] = per_kernel
        return rec


def run_ours(args, wl):
    bn = Bench(args)
    torch = bn.torch
    K, Wm = args.steps, max(args.warmup, 3)
    if wl["sampler"] == "train":
        from tools.bench_train import run_train
        run_train(bn, args, wl, K, Wm, METRIC, config_of, load_peaks)
        return
    clocks = ClockSampler(bn.local)
    if bn.rank == 0:
        clocks.start()
    main = bn.sampling(args.workload, wl, K, Wm)
    clk = clocks.stop() if bn.rank == 0 else None
    also = {}
    if not args.no_also:
        for nm in ("cfg3", "cfg5"):
            if nm == args.workload:
                continue
            w2 = WORKLOADS[nm]
            K2 = 100 if nm == "cfg3" else 50
            r = bn.sampling(nm, w2, K2, 3)
            Cc, (H, W, D), B = w2["C"], w2["HWD"], w2["B"]
            entry = dict(config_of(nm, w2, bn.world), value=r["value"], unit="steps/s", ms_per_step=r["ms_per_step"], steps=K2,
                         sample_steps_per_s=r["sample_steps_per_s"], e2e=r["e2e"], per_rank_ms=r["per_rank_ms"],
                         workspace_mib=r["workspace_mib"])
            if "roofline" in r:
                entry["roofline"] = {k: r["roofline"][k] for k in ("kernel", "achieved", "peak", "unit", "frac", "conv_ms_per_step",
                                                                    "conv_share_of_event_timed_step")}
                entry["kernels"] = r["kernels"]
            also[f"{nm}_{w2['sampler']}{w2['chain']}"] = entry
    if not args.no_also:
        # (before the host-side baseline legs, which leave the GPU idle for ~30 s; 40 warm-up steps = 0.9 s: building the training plan
        # idles the GPU for seconds and the first ~10 iterations after it run 10-20 % slow)
        # the training step (BASELINE configs[3], diffusion half): forward + backward + AdamW/EMA at B=32 per GPU, data parallel at N>1
        try:
            from tools.bench_train import run_train
            tr = run_train(bn, args, WORKLOADS["cfg4"], 10, 40, METRIC, config_of, load_peaks, emit=False, profile=bn.world == 1)
            if tr is not None:
                also["cfg4_train_b32"] = {k: tr[k] for k in ("metric", "value", "unit", "ms_per_step", "per_rank_ms", "timing", "config", "e2e", "detail",
                                                             "roofline", "roofline_parts") if k in tr}
        except Exception as e:
            also["cfg4_train_b32"] = f"failed: {type(e).__name__}: {e}"
    cpu = lib = None
    if bn.rank == 0 and bn.world == 1 and not args.no_cpu_baseline:
        lib = dict(what="reference arithmetic (oracle restatement) run by torch eager on the same GPU: cuDNN / cuBLAS / ATen library "
                        "kernels, one sampler step per iteration, 20 steps after 3 warm-up", unit="steps/s", torch=torch.__version__)
        for tf32 in (0, 1):
            try:
                lib["tf32_on" if tf32 else "fp32"] = torch_eager_steps_per_s(wl, 20, tf32)
            except Exception as e:       # baseline leg only: never take the product line down with it
                lib["tf32_on" if tf32 else "fp32"] = f"failed: {type(e).__name__}: {e}"
        torch.backends.cudnn.allow_tf32 = True
        try:        # the step after the path: decode_grid of a sampled latent, with its CPU oracle beside it
            from tools.bench_decoder import run as run_decoder
            dl = run_decoder(reso=256, iters=3, cpu_baseline=True, device=bn.local)
            also["decoder_grid256"] = dict(metric=dl["metric"], value=dl["value"], unit=dl["unit"], ms_per_step=dl["ms_per_step"],
                                           workload=dl["config"]["workload"], e2e=dl["e2e"],
                                           roofline={k: dl["roofline"][k] for k in ("kernel", "achieved", "peak", "unit", "frac")},
                                           cpu_baseline=dl.get("cpu_baseline"), parity_vs_oracle_rel_l2=dl.get("parity_vs_oracle_rel_l2"))
        except Exception as e:
            also["decoder_grid256"] = f"failed: {type(e).__name__}: {e}"
        sps, done, dt, thr, kind = cpu_steps_per_s(wl, 200, 3, budget_s=20.0)
        cpu = dict(value=sps, unit="steps/s", cores=thr, kind=kind, sample=cpu_sample_text(kind, done, wl, args.workload, 3, thr, dt))
    if bn.rank == 0:
        Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
        cfg = config_of(args.workload, wl, bn.world)       # identical in both arms
        detail = dict(sample_steps_per_s=main["sample_steps_per_s"], dense_tflops=cfg["dense_gflop_per_step"] * main["value"] / bn.world / 1e3,
                      workspace_mib=main["workspace_mib"], cuda_graph=True, graph_captures_in_timed_region=0)
        line = dict(
            metric=METRIC, value=main["value"], unit="steps/s", n_gpus=bn.world, steps=K, warmup=Wm, ms_per_step=main["ms_per_step"],
            per_rank_ms=main["per_rank_ms"], higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype=DTYPE[bn.model(wl["C"]).s3d_precision], data="synthetic", config=cfg, clocks=clk, e2e=main["e2e"],
            gpu_launches=K * main["launches_per_step"], launches_per_step=main["launches_per_step"],
            roofline=main.get("roofline"), kernels=main.get("kernels"), cpu_baseline=cpu, gpu_library_baseline=lib,
            detail=detail, also=also)
        print(json.dumps(line), flush=True)
    if bn.world > 1:
        bn.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and torch-eager baseline legs")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary cfg3 / cfg5 records")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of the workload (cfg4)")
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
