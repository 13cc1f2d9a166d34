"""Build recipe of libosmr_b200.so (hand-written CUDA for sm_100a; no torch, no JIT cache).

The library is built IN-TREE (osm_renderer_b200/libosmr_b200.so) so that it travels to the GPU box with the
repository snapshot.  -fmad=false is mandatory: the reference arithmetic has no fused multiply-adds.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libosmr_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "osmr_capi.cu")]
HEADERS = [
    os.path.join(HERE, "csrc", "osmr_kernels.cuh"),
    os.path.join(HERE, "csrc", "osmr_device.cuh"),
    os.path.join(HERE, "csrc", "osmr_auto.cuh"),
    os.path.join(HERE, "csrc", "osmr_png.cuh"),
    os.path.join(HERE, "csrc", "osmr_labels_host.hpp"),
    os.path.join(ROOT, "include", "osmr.h"),
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libosmr_b200.so cannot be built (there is no CPU fallback)")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [
        nvcc_path(),
        "-O3",
        "-std=c++17",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-fmad=false",
        "-lineinfo",
        "-Xcompiler",
        "-fPIC",
        "-shared",
        "-cudart",
        "static",
        "-I",
        os.path.join(ROOT, "include"),
        "-I",
        os.path.join(HERE, "csrc"),
        "-o",
        LIB_PATH,
    ] + SOURCES
    extra = os.environ.get("OSMR_EXTRA_NVCC_FLAGS", "").split()  # experiments only, e.g. -DOSMR_BW=32
    cmd[1:1] = extra
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_lib(force=True, verbose="-v" in sys.argv))
