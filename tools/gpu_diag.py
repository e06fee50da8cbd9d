"""Bring-up diagnostics for a GPU box: runs the path stage by stage and prints where it first diverges
from the oracle.  Everything is wrapped so one failure does not hide the later stages.
    python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import unet_ref as ur
from oracle.cases import UNET_CASES, make_inputs
from tests.gpu_util import make_cuda_model, nhwc, plane_errors


def stage(name):
    def deco(fn):
        print(f"\n===== {name}", flush=True)
        t0 = time.time()
        try:
            fn()
            print(f"----- {name}: ok ({time.time() - t0:.1f}s)", flush=True)
        except Exception:
            traceback.print_exc()
            print(f"----- {name}: FAILED", flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("!! device error after stage:", e, flush=True)
        return fn
    return deco


print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def layer_report(case_name, impl, precision=3):
    case = UNET_CASES[case_name]
    spec = ur.UNetSpec(**case["spec"])
    sd = ur.synthetic_state_dict(spec, case["wseed"])
    x, t = make_inputs(case)
    H, W, D = case["HWD"]
    m = make_cuda_model(spec, sd, precision, impl)
    with torch.no_grad():
        out = m(x.cuda(), t.cuda(), H=H, W=W, D=D)
    torch.cuda.synchronize()
    trace = {}
    want = ur.unet_forward(sd, spec, x, t, H, W, D, trace=trace)
    acts = m.debug_activations()
    for k, v in acts.items():
        key = k[:-4] if k.endswith(".out") else k
        if key in trace:
            errs = ["%.2e" % float((a - nhwc(b)).abs().max() / (b.abs().max() + 1e-9)) for a, b in zip(v, trace[key])]
            print(f"  {k:34s} max-rel/plane {errs}")
        else:
            print(f"  {k:34s} absmax/plane {['%.3g' % float(a.abs().max()) for a in v]} nan={[bool(a.isnan().any()) for a in v]}")
    rel, mx = plane_errors(out.cpu(), want, H, W, D)
    print(f"  FINAL {case_name} impl={impl} prec={precision}: rel_l2={rel:.3e} max={mx:.3e}")


for impl in ("ffma", "tc"):
    for cname in ("small_even", "small_odd"):
        stage(f"unet {cname} impl={impl}")(lambda: layer_report(cname, impl))
stage("unet small_even tc precision=1")(lambda: layer_report("small_even", "tc", 1))
stage("unet raw tc")(lambda: layer_report("raw", "tc"))
stage("unet three_level tc")(lambda: layer_report("three_level", "tc"))


@stage("smoke")
def _():
    import __graft_entry__ as ge
    ge.smoke()
