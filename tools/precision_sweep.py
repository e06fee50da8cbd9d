#!/usr/bin/env python
"""Full-chain error of every conv precision mode against the real-reference fixtures (tests/golden/full_*.npz), with the device
time per step of each: the measurement behind the default mode.  Writes gpurun_out/precision_sweep.json.

    python tools/precision_sweep.py [case ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from oracle.cases import FULL_CASES, full_errors
    from tests.test_gpu_full_chains import run_full_chain
    names = sys.argv[1:] or list(FULL_CASES)
    out = {}
    for name in names:
        fx = np.load(os.path.join(ROOT, "tests", "golden", f"full_{name}.npz"))
        for mode in [int(v) for v in os.environ.get('SWEEP_MODES', '3,2,4,1').split(',')]:
            t0 = time.time()
            got = run_full_chain(name, mode)
            rel0, rel_rest = full_errors(got, fx, *FULL_CASES[name]["HWD"])
            out[f"{name}/mode{mode}"] = dict(rel_l2_sample0=rel0, rel_l2_others=rel_rest, wall_s=round(time.time() - t0, 2))
            print(name, "mode", mode, out[f"{name}/mode{mode}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
