"""Builds libsin3dm_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library has a plain C ABI).  `python -m sin3dm_b200.build [--force] [-v]`.

Two translation units (the denoising path and the triplane decoder) are compiled to objects under build/ in
parallel and linked into one shared library; an object is rebuilt only when one of its sources is newer."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libsin3dm_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
HEADER = os.path.join(ROOT, "include", "sin3dm_b200.h")
UNITS = {
    "s3d.cu": ["s3d.cu", "kernels.cuh", "conv_tc.cuh", "boundary.cuh", "common.cuh", "ptx.cuh", "host_util.cuh", "train_kernels.cuh",
               "train_host.cuh", "wgrad_tc.cuh"],
    "dec.cu": ["dec.cu", "dec_kernels.cuh", "common.cuh", "ptx.cuh", "host_util.cuh"],
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def up_to_date():
    deps = [HEADER] + [os.path.join(CSRC, f) for fs in UNITS.values() for f in fs]
    return not _newer(OUT, deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_unit(item):
        src, deps = item
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _newer(obj, [HEADER] + [os.path.join(CSRC, f) for f in deps]):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS.items()))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
