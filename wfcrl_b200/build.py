"""Builds libwfcrl_b200.so (hand-written sm_100a CUDA kernels + C-ABI) IN-TREE with nvcc.

Usage: python -m wfcrl_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwfcrl_b200.so")
SOURCES = ["wf_api.cu", "wf_kernels.cu", "wf_fast.cu", "wf_fast64.cu"]
HEADERS = ["wf_device.cuh", "wf_host_const.h", "gen_baked.cu", os.path.join("..", "..", "include", "wfcrl_b200.h")]
BAKED = os.path.join(CSRC, "wf_fast_baked.inc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--resource-usage",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the wfcrl_b200 CUDA library cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    # step 1: bake the default model's kernel constants into a header (host program built and run at build time)
    gen_exe = os.path.join(HERE, "_gen_baked")
    subprocess.run([_nvcc(), "-O1", "-std=c++17", "-o", gen_exe, os.path.join(CSRC, "gen_baked.cu")], check=True,
                   capture_output=True, text=True)
    with open(BAKED, "w") as fp:
        fp.write(subprocess.run([gen_exe], check=True, capture_output=True, text=True).stdout)
    os.remove(gen_exe)
    # step 2: the library
    extra = os.environ.get("WFCRL_NVCC_EXTRA", "").split()  # e.g. -DWF_FAST_MINB=14 for tuning experiments
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fp:
        fp.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
