"""Stage the reference's own hot-path modules under oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

    python -m oracle.stage_ref          # authoring container only: needs /root/reference (read-only)

The reference is a plain script tree (no packaging, pins torch==1.13.1), so there is nothing to pip-install: the unmodified files
that make up the denoising path — and only those — are copied verbatim, never into the tracked tree.  ``bench.py --impl
reference`` imports them from here when present (``cpu_baseline.kind = "reference"``) and falls back to the oracle port
otherwise.  Nothing under sin3dm_b200/ ever reads this directory."""
import os
import shutil
import sys

SRC = "/root/reference/src"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
FILES = ["diffusion/gaussian_diffusion.py", "diffusion/respace.py", "diffusion/unet_triplane.py", "diffusion/nn.py",
         "diffusion/losses.py", "diffusion/fp16_util.py", "diffusion/logger.py", "utils/triplane_util.py"]


def stage():
    if not os.path.isdir(SRC):
        return None
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
    for pkg in ("diffusion", "utils"):
        init = os.path.join(SRC, pkg, "__init__.py")
        if os.path.exists(init):
            shutil.copyfile(init, os.path.join(DST, pkg, "__init__.py"))
    return DST


def staged_path():
    """-> oracle/_ref/src when every file is there, else None."""
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


if __name__ == "__main__":
    print(stage() or "no /root/reference here", file=sys.stderr)
