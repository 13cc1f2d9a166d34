"""Runs fixture tiles through the EMULATED kernels (tests/emu/_build/libosmr_emu.so) and compares with the oracle.

    python tests/emu/run_emu.py [config] [tile indices ...]     e.g.  python tests/emu/run_emu.py 17 0 7 12
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def emu_context(lib=None):
    from osm_renderer_b200 import _lib

    _lib._lib = None
    _lib.LIB_PATH = lib or os.path.join(HERE, "_build", "libosmr_emu.so")
    from osm_renderer_b200.drawer import GpuContext

    return GpuContext(0)


def main():
    import oracle
    from conftest import FixtureInputs

    cfg = sys.argv[1] if len(sys.argv) > 1 else "17"
    fx = FixtureInputs()
    tiles, begins, areas = fx.batches[cfg]
    sel = [int(a) for a in sys.argv[2:]] or [0]
    ctx = emu_context(os.environ.get("OSMR_EMU_LIB"))
    ctx.set_geodata(fx.bin)
    ctx.set_table(fx.table)
    parts = [areas[begins[i] : begins[i + 1]] for i in sel]
    b = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    a = np.concatenate(parts)
    t0 = time.time()
    got = ctx.draw_tiles(tiles[sel], b, a, fx.canvas_rgb, fx.use_caps_for_dashes)
    t1 = time.time()
    want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles[sel], b, a, fx.canvas_rgb, fx.use_caps_for_dashes))
    bad = (got != want).any(axis=-1)
    print(f"config {cfg} tiles {sel}: {ctx.stats()['n_visible_ops']} visible ops, emulated in {t1 - t0:.1f}s, "
          f"differing pixels vs oracle: {int(bad.sum())}")
    return int(bad.sum())


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
