"""Bring-up: one UNet forward at the cfg2 shape with batch B (argv[1], default 8); prints 'ok' after a device sync."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import unet_ref as ur
from tests.gpu_util import make_cuda_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W, D = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (92, 128, 92)
spec = ur.UNetSpec(in_channels=12, model_channels=64, out_channels=12)
m = make_cuda_model(spec, ur.synthetic_state_dict(spec, 3))
x = torch.randn(B, 12, H + D, W + D, device="cuda")
t = torch.full((B,), 500, device="cuda")
for i in range(3):
    with torch.no_grad():
        y = m(x, t, H=H, W=W, D=D)
    torch.cuda.synchronize()
    print("forward", i, "ok", float(y.abs().max()), flush=True)
