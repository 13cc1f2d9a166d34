"""Kernel logic on the host SIMT emulator (tests/emu/): the CUDA sources of libosmr_b200.so, compiled by g++ against
a fiber-per-thread emulator of warps / blocks, must render fixture tiles bit-for-bit like the oracle.  This checks the
kernels' control flow and integer / f64 arithmetic in the GPU-less container (divergent or deadlocking warp
collectives abort the run); it says nothing about the GPU build itself -- the `-m gpu` tests do that on the B200.
The whole `-m gpu` suite can be run the same way:  OSMR_TEST_EMU=1 python -m pytest tests -m gpu
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [["17", "0", "7", "12"], ["14", "0"], ["18_2x", "5"]])
def test_emulated_kernels_equal_oracle(args):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    build_emu.build()
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_emu.py")] + args, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing pixels vs oracle: 0" in res.stdout
    assert "[emu]" not in res.stderr, res.stderr  # divergent collective / deadlock reports
