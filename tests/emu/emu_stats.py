"""Divergence / work statistics of the emulated kernels on a few tiles of the bench workload (OSMR_COUNT counters).

    OSMR_EMU_STATS=1 python tests/emu/emu_stats.py [n_tiles] [check]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import bench  # noqa: E402
from run_emu import emu_context  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    w = bench.build_workload("C2")
    sel = np.linspace(0, len(w["tiles"]) - 1, n).astype(int)
    ab = w["area_begin"]
    parts = [w["areas"][ab[i] : ab[i + 1]] for i in sel]
    b = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    a = np.concatenate(parts)
    ctx = emu_context(os.environ.get("OSMR_EMU_LIB"))
    ctx.set_geodata(w["bin"])
    ctx.set_table(w["table"])
    t0 = time.time()
    got = ctx.draw_tiles(w["tiles"][sel], b, a, w["canvas"], w["caps"])
    print(f"{n} tiles emulated in {time.time() - t0:.1f}s, visible ops {ctx.stats()['n_visible_ops']}")
    if "check" in sys.argv:
        import oracle

        want = np.stack(oracle.draw_tiles(w["bin"], w["table"], w["tiles"][sel], b, a, w["canvas"], w["caps"]))
        print("differing pixels vs oracle:", int((got != want).any(axis=-1).sum()))


if __name__ == "__main__":
    main()
