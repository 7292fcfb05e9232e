"""Builds ml_qem_b200/lib/libbwq.so in-tree with nvcc for sm_100a (python -m ml_qem_b200.build)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "lowering.cpp", "sv_lowering.cpp", "variants.cpp")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "kernels_tma.cuh", "onchip.cuh", "sv_kernels.cuh", "program.h")] + [os.path.join(HERE, "..", "include", "bwq.h")]
OUT = os.path.join(HERE, "lib", "libbwq.so")


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_library(force=False, verbose=False, out=None, defines=()):
    """out/defines: alternative builds for kernel experiments (e.g. defines=["BWQ_KQ6_BLOCKS=3"])."""
    out = out or OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(f) for f in SRC + HDR):
        return out
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", out] + [f"-D{d}" for d in defines] + SRC
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv, out=outs[0] if outs else None,
                        defines=defs))
