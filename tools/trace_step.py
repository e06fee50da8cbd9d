"""Timeline of ONE denoising step from the opt-in device trace (s3d_unet_trace_enable): every kernel of the step stamps
%globaltimer at named phases, once per CTA.  Runs a short graph-replayed sampling loop, reads the stamps of the LAST
step and prints, per launch, when its CTAs entered / reached each phase / left, relative to the first entry of the step.

    python tools/trace_step.py [--workload cfg2] [--steps 30] > gpurun_out/trace_cfg2.txt
Times are microseconds; each cell is "min / median / max" over the CTAs that stamped that slot.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

SLOT_NAMES = {
    "k_conv_tc": {**{4 + 6 * lt + k: f"t{lt}.{n}" for lt in range(3)
                     for k, n in enumerate(("opnd", "mma_iss", "pre", "acc", "epi", "a_iss"))},
                  0: "entry", 1: "setup", 2: "roll.begin", 16: "roll.slot", 17: "roll.ldiss", 18: "roll.filled", 3: "roll.A_done",
                  19: "roll.staged", 20: "roll.stored", 21: "roll.fenced", 23: "roll_seen", 22: "exit",
                  24: "A.role", 25: "A.decoded", 26: "A.slot", 27: "B.role", 28: "B.first", 29: "B.ninth", 30: "epi.role", 31: "mma.role"},
    "k_gn_silu": {0: "entry", 1: "coef", 2: "stored", 3: "row_atom"},
    "k_boundary<in_conv>": {0: "entry", 1: "head", 2: "sched", 3: "in_conv"},
    "k_boundary<head>": {0: "entry", 1: "head", 2: "sched", 3: "in_conv"},
    "k_boundary<fused>": {0: "entry", 1: "head", 2: "sched", 3: "in_conv"},
    "k_avgpool2": {0: "entry", 1: "loop", 2: "exit"},
    "k_upcat": {0: "entry", 1: "loop", 2: "exit"},
    "k_sched_step": {0: "entry", 1: "loop"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--clock", action="store_true", help="also print per-CTA phase durations in SM cycles (clock64)")
    args = ap.parse_args()
    from bench import WORKLOADS
    from oracle import unet_ref as ur
    import sin3dm_b200 as s3
    from sin3dm_b200 import _lib
    from sin3dm_b200.script_util import create_gaussian_diffusion

    wl = WORKLOADS[args.workload]
    Cc, (H, W, D), B = wl["C"], wl["HWD"], wl["B"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model = s3.TriplaneUNetModelSmall(Cc, 64, Cc, 1, 0, (1, 2), use_scale_shift_norm=True)
    model.load_state_dict(ur.synthetic_state_dict(ur.UNetSpec(in_channels=Cc, model_channels=64, out_channels=Cc), 1234))
    model = model.to(dev).eval()
    L = _lib.lib()
    h = model.handle()
    _lib.check(L.s3d_unet_trace_enable(h, 1))
    full = create_gaussian_diffusion(predict_xstart=True, timestep_respacing="")
    coef = full.coef_table(dev)
    film = model.film_table(full._model_timesteps(torch.arange(1000, device=dev)).float())
    x = torch.randn(B, Cc, H + D, W + D, generator=torch.Generator().manual_seed(0)).to(dev)
    a = _lib.LoopArgs()
    a.kind = _lib.DDIM if wl["sampler"] == "ddim" else _lib.DDPM
    a.mean_type, a.clip_denoised, a.n_steps, a.t_start = _lib.START_X, 1, args.steps, 999
    a.B, a.H, a.W, a.D, a.n_per_sample = B, H, W, D, x[0].numel()
    a.x_dev, a.coef_dev, a.film_dev = x.data_ptr(), coef.data_ptr(), film.data_ptr()
    a.seed, a.sample_base, a.use_graph = 1234, 0, 1
    _lib.check(L.s3d_sample_loop(h, C.byref(a), _lib.current_stream_ptr()))
    torch.cuda.synchronize()

    nslots = L.s3d_trace_slots()
    nops = L.s3d_unet_op_count(h)
    maxc = 2048
    rows = []
    for i in list(range(nops)) + [-1, -2]:
        if i >= 0:
            nm = C.c_char_p()
            _lib.check(L.s3d_unet_op_info(h, i, C.byref(nm), None))
            name = nm.value.decode()
        else:
            name = "k_sched_step" if i == -1 else "k_boundary<fused>"
        buf = np.zeros((maxc, 2, nslots), dtype=np.uint64)
        if L.s3d_unet_trace_read(h, i, buf.ctypes.data_as(C.c_void_p), maxc) != 0:
            continue
        if not (buf[:, 0, 0] > 0).any():
            continue                      # op not part of the loop in this mode
        rows.append((i, name, buf))
    # in the fused loop the in_conv of the NEXT step runs inside the boundary kernel: order by first entry
    first_of = lambda r: int(r[2][:, 0, 0][r[2][:, 0, 0] > 0].min())
    latest = max(first_of(r) for r in rows)
    rows = [r for r in rows if first_of(r) > latest - 5_000_000]     # drop launches that are not part of the last step
    rows.sort(key=first_of)
    t0 = min(int(b[:, 0, 0][b[:, 0, 0] > 0].min()) for _, _, b in rows if (b[:, 0, 0] > 0).any())
    print(f"# workload {args.workload}: one step of a {args.steps}-step graph-replayed loop; us since the first CTA entry of the step")
    prev_end = None
    for i, name, buf in rows:
        g = buf[:, 0, :].astype(np.int64)
        used = g[:, 0] > 0
        n = int(used.sum())
        if n == 0:
            print(f"op {i:3d} {name}: no stamps")
            continue
        names = SLOT_NAMES.get(name, {})
        last = g[used].max()
        first = g[used, 0].min()
        gap = "" if prev_end is None else f"  gap after previous kernel's last stamp {(first - prev_end) / 1e3:6.2f}"
        print(f"op {i:3d} {name:22s} ctas {n:4d}  first entry {(first - t0) / 1e3:8.2f}  last stamp {(last - t0) / 1e3:8.2f}"
              f"  span {(last - first) / 1e3:7.2f}{gap}")
        for s in range(nslots):
            col = g[used, s]
            col = col[col > 0]
            if col.size == 0:
                continue
            rel = (col - t0) / 1e3
            print(f"      {names.get(s, 'slot%d' % s):10s} n={col.size:4d}  {rel.min():8.2f} / {np.median(rel):8.2f} / {rel.max():8.2f}")
        if args.clock:
            c = buf[:, 1, :].astype(np.int64)
            stamped = [s for s in range(nslots) if (g[used, s] > 0).all()]
            for s0, s1 in zip(stamped[:-1], stamped[1:]):
                d = c[used, s1] - c[used, s0]
                print(f"      cycles {names.get(s0, s0)} -> {names.get(s1, s1)}: {d.min()} / {int(np.median(d))} / {d.max()}")
        prev_end = last


if __name__ == "__main__":
    main()
